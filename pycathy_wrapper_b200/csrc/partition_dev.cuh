// partition_dev.cuh -- row-block partition of one mesh over several GPUs, device side: peer-memory boxes, system-scope flags, halo send / receive, combination of the per-rank norms and step sums.
// Part of the single translation unit cathy_b200.cu (included in dependency order; shares its structs and helpers).
#pragma once

// ==========================================================================================
// Row-block partition of ONE mesh over several GPUs (BASELINE config 5).  Each rank holds a window of DEM rows (its owned
// node rows + DD_W ghost node rows per interior side) with the same layer-major numbering and the same 15-point DIA stencil.
// Ranks exchange data through peer memory over NVLink (CUDA IPC mapped "boxes"): halo rows are STORED straight into the
// neighbour's inbox by the kernel that produces them, all-reduces are slot writes + system-scope release/acquire flags,
// summed in rank order on every rank (bit-identical results on all ranks, hence identical control flow).
// ==========================================================================================
#define DD_W 2
#define DD_MAXW 8
#define DD_NRED 24
#define DD_TIMEOUT_CYCLES 12000000000LL   // ~6 s: a lost peer turns into an error, not a hang
struct DDBox {                               // lives in each rank's device memory, mapped by all peers
    unsigned int ar_flag[DD_MAXW];           // sequence number of the last all-reduce contribution of rank r
    unsigned int halo_flag[2];               // [0]: from the north neighbour, [1]: from the south neighbour
    int geom[4];                             // column-major layout of this rank: first owned row lo, end hi, local rows, halo rows (k_pcg_tma)
    unsigned int pad[2];
    double ar_slot[2][DD_MAXW][DD_NRED];     // [parity][rank][value]
};
struct DDCtx {
    int world, rank, north, south;           // neighbour ranks (-1: none)
    DDBox *me;
    DDBox *peer[DD_MAXW];                    // peer[rank] == me
    double *inbox_me;                        // [2 parities][2 sides][hcap]
    double *inbox_peer[DD_MAXW];
    long long hcap;
    int nc1, nlay, nnod, own_a, own_b;
    unsigned int *seq;                       // [0] all-reduce sequence, [1] halo sequence (device resident, advanced by the kernels)
    int *err;
};
__device__ __forceinline__ void dd_wait(const unsigned int *flag, unsigned int target, int *err, int site = 1)
{
    if (*(volatile int *)err) return;
    long long t0 = clock64();
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - target) >= 0) return;
        if (clock64() - t0 > DD_TIMEOUT_CYCLES) { *(volatile int *)err = site + 10 * (int)(target & 0xffffffu); return; }
    }
}
__device__ __forceinline__ void dd_release(unsigned int *flag, unsigned int v)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
// All-reduce (sum, rank order) of NV block-uniform values; called by every thread of every block of a kernel whose blocks
// all hold the same v[] (after a grid-wide local reduction).  sh: NV doubles of shared memory.
template <int NV>
__device__ __forceinline__ void dd_allreduce(const DDCtx &c, unsigned int &seq, double (&v)[NV], double *sh)
{
    ++seq;
    const int par = seq & 1u;
    if (blockIdx.x == 0 && (int)threadIdx.x < c.world) {
        DDBox *dst = c.peer[threadIdx.x];
#pragma unroll
        for (int q = 0; q < NV; ++q) dst->ar_slot[par][c.rank][q] = v[q];
        dd_release(&dst->ar_flag[c.rank], seq);
    }
    if ((int)threadIdx.x < c.world) dd_wait(&c.me->ar_flag[threadIdx.x], seq, c.err, 1);
    __syncthreads();
    if (threadIdx.x < NV) {
        double acc = 0.0;
        for (int r = 0; r < c.world; ++r) acc += *(volatile double *)&c.me->ar_slot[par][r][threadIdx.x];
        sh[threadIdx.x] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = sh[q];
    __syncthreads();
}
// element e of a halo message <-> (layer, row offset, column)
__device__ __forceinline__ long long dd_index(const DDCtx &c, long long e, int row0)
{
    int j = (int)(e % c.nc1);
    long long t = e / c.nc1;
    int w = (int)(t % DD_W), l = (int)(t / DD_W);
    return (long long)l * c.nnod + (long long)(row0 + w) * c.nc1 + j;
}
__device__ __forceinline__ void dd_send_rows(const DDCtx &c, unsigned int next_seq, const double *vec, long long tid, long long nthreads)
{   // my first / last DD_W owned rows -> the neighbours' south / north inboxes
    const long long E = (long long)DD_W * c.nlay * c.nc1;
    const int par = next_seq & 1u;
    bool any = false;
    if (c.north >= 0) { double *dst = c.inbox_peer[c.north] + ((size_t)par * 2 + 1) * c.hcap; for (long long e = tid; e < E; e += nthreads) { dst[e] = vec[dd_index(c, e, c.own_a)]; any = true; } }
    if (c.south >= 0) { double *dst = c.inbox_peer[c.south] + ((size_t)par * 2 + 0) * c.hcap; for (long long e = tid; e < E; e += nthreads) { dst[e] = vec[dd_index(c, e, c.own_b - DD_W)]; any = true; } }
    if (any) __threadfence_system();
}
// after a grid-wide barrier that follows the sends: publish, wait for the neighbours' rows, copy them into the ghost rows
__device__ __forceinline__ void dd_recv_rows(const DDCtx &c, unsigned int &seq, double *vec, long long tid, long long nthreads)
{
    ++seq;
    const int par = seq & 1u;
    if (blockIdx.x == 0 && threadIdx.x == 0 && c.north >= 0) dd_release(&c.peer[c.north]->halo_flag[1], seq);
    if (blockIdx.x == 0 && threadIdx.x == 1 && c.south >= 0) dd_release(&c.peer[c.south]->halo_flag[0], seq);
    if (threadIdx.x == 0 && c.north >= 0) dd_wait(&c.me->halo_flag[0], seq, c.err, 2);
    if (threadIdx.x == 1 && c.south >= 0) dd_wait(&c.me->halo_flag[1], seq, c.err, 3);
    __syncthreads();
    const long long E = (long long)DD_W * c.nlay * c.nc1;
    if (c.north >= 0) { const double *src = c.inbox_me + ((size_t)par * 2 + 0) * c.hcap; for (long long e = tid; e < E; e += nthreads) vec[dd_index(c, e, c.own_a - DD_W)] = *(volatile const double *)&src[e]; }
    if (c.south >= 0) { const double *src = c.inbox_me + ((size_t)par * 2 + 1) * c.hcap; for (long long e = tid; e < E; e += nthreads) vec[dd_index(c, e, c.own_b)] = *(volatile const double *)&src[e]; }
}
// stand-alone halo exchange of one N-vector between kernels of the nonlinear loop (two launches: the kernel boundary is the
// grid-wide barrier between "all rows stored" and "flag published")
__global__ void k_dd_send(DDCtx c, const double *__restrict__ vec)
{
    dd_send_rows(c, c.seq[1] + 1u, vec, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}
__global__ void k_dd_recv(DDCtx c, double *__restrict__ vec, unsigned int *counter)
{   // counter: arrival count so that the LAST block to finish advances the sequence number
    unsigned int seq = c.seq[1];
    dd_recv_rows(c, seq, vec, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(counter, 1u) == gridDim.x - 1) { c.seq[1] = seq; *counter = 0u; }
    }
}
// cross-rank combination of the per-iteration scalars (sums in rank order; max-norm with the reference's "last node wins" tie
// rule on GLOBAL node numbers) and of the per-step scalars; one block
__global__ void k_dd_combine_iter(DDCtx c, IterOut *__restrict__ io, int gnnod, int lo_shift)
{
    __shared__ double sh[DD_NRED];
    unsigned int seq = c.seq[0];
    // the two norms arrive squared-rooted from k_norms_final: square them back for the sum
    double v[7] = {io->pl2 * io->pl2, io->fl2 * io->fl2, io->dstore, io->adin, io->adout, io->anin, io->anout};
    dd_allreduce<7>(c, seq, v, sh);
    // max part: every rank publishes (pinf, global ik, pnew_ik, pold_ik, finf) in its slot, all pick the same winner
    int ikl = io->ikmax, lay = ikl / c.nnod;
    double gik = (double)((long long)lay * gnnod + (ikl - lay * c.nnod) + lo_shift);
    ++seq;
    const int par = seq & 1u;
    if ((int)threadIdx.x < c.world) {
        DDBox *dst = c.peer[threadIdx.x];
        double *sl = dst->ar_slot[par][c.rank];
        sl[0] = io->pinf; sl[1] = gik; sl[2] = io->pnew_ik; sl[3] = io->pold_ik; sl[4] = io->finf;
        dd_release(&dst->ar_flag[c.rank], seq);
    }
    if ((int)threadIdx.x < c.world) dd_wait(&c.me->ar_flag[threadIdx.x], seq, c.err, 4);
    __syncthreads();
    if (threadIdx.x == 0) {
        double pinf = -1.0, ik = -1.0, pn = 0.0, po = 0.0, finf = 0.0;
        for (int r = 0; r < c.world; ++r) {
            const volatile double *sl = c.me->ar_slot[par][r];
            if (sl[0] > pinf || (sl[0] == pinf && sl[1] > ik)) { pinf = sl[0]; ik = sl[1]; pn = sl[2]; po = sl[3]; }
            finf = fmax(finf, sl[4]);
        }
        io->pl2 = sqrt(v[0]); io->fl2 = sqrt(v[1]); io->dstore = v[2]; io->adin = v[3]; io->adout = v[4]; io->anin = v[5]; io->anout = v[6];
        io->pinf = pinf; io->ikmax = (int)ik; io->pnew_ik = pn; io->pold_ik = po; io->finf = finf;
        c.seq[0] = seq;
    }
}
__global__ void k_dd_combine_step(DDCtx c, StepOut *__restrict__ so, double *__restrict__ extra3)
{
    __shared__ double sh[DD_NRED];
    unsigned int seq = c.seq[0];
    double v[21];
    v[0] = so->store1; v[1] = so->apot; v[2] = so->aact; v[3] = so->ovflow; v[4] = so->reflow;
    v[5] = so->nhort; v[6] = so->ndunn; v[7] = so->npond; v[8] = so->nsat;
    for (int q = 0; q < 9; ++q) v[9 + q] = so->hgflag[q];
    for (int q = 0; q < 3; ++q) v[18 + q] = extra3 ? extra3[q] : 0.0;
    dd_allreduce<21>(c, seq, v, sh);
    if (threadIdx.x == 0) {
        so->store1 = v[0]; so->apot = v[1]; so->aact = v[2]; so->ovflow = v[3]; so->reflow = v[4];
        so->nhort = (int)v[5]; so->ndunn = (int)v[6]; so->npond = (int)v[7]; so->nsat = (int)v[8];
        for (int q = 0; q < 9; ++q) so->hgflag[q] = (int)v[9 + q];
        if (extra3) for (int q = 0; q < 3; ++q) extra3[q] = v[18 + q];
        c.seq[0] = seq;
    }
}
