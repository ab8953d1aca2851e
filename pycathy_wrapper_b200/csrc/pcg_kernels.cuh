// pcg_kernels.cuh -- SYMSLV on the device: the persistent PCG kernels k_pcg (streaming), k_pcg_res / k_pcg_res2 (CG vectors resident in shared memory), the cluster kernels (pcg_cluster.cuh) and k_pcg2 (single-reduction CG).
// Part of the single translation unit cathy_b200.cu (included in dependency order; shares its structs and helpers).
#pragma once

// ------------------------------------------------------------------------------------------
// K5-K7: the whole SYMSLV (SRC/solscal-extended.f:4669-4699) as ONE persistent cooperative
// kernel: preconditioner set-up, x0 = M^-1 b, and the GRADDP recurrence (:1260-1380) with two
// grid-wide barriers per iteration.  Reductions are fixed-order (block partials, then every block
// adds the partials in the same order), so results are bit-reproducible run to run.
//   phase A: p = z + beta p_old (recomputed on the fly for the neighbours), B = A p, (p.r), (p.B)
//   phase B: r -= alfa B, x += alfa p, z = M^-1 r, (B.z), ||r_free||^2
// Residual norm excludes Dirichlet rows exactly like GRADDP (:1286-1297, :1356-1371).
// ------------------------------------------------------------------------------------------
struct PcgArgs {
    int n, nnod, itmax;
    double tol;
    Diag A;
    const double *diag;      // main diagonal with the Dirichlet penalty
    const double *rhs;
    double *x, *r, *z, *p0, *p1, *bv;
    const int *ifatm;
    const unsigned char *contp_flag;
    double *partial;         // [3][gridDim.x]
    unsigned int *counter;   // grid barrier counter (monotonic)
    unsigned int epoch0;     // its value at launch
    IterOut *out;
    int prefetch;               // 1: software prefetch of the next grid-stride row into L2
    const unsigned char *own;   // row-block partition: bit0 = owned row, bit1 / bit2 = row is sent to the north / south neighbour
    DDCtx dd;
    int rows_cta;               // k_pcg_res: rows owned by one CTA (multiple of 32)
    int xres;                   // k_pcg_res: 1 = the solution vector lives in shared memory too
    int cm;                     // k_pcg: 1 = the arrays are in the column-major permutation (Dirichlet rows are recognised by their penalty diagonal)
};

// Grid-wide barrier for the persistent kernel: one arrival per block on a monotonically increasing counter
// (release), then a spin on an acquire load.  All blocks are co-resident (cooperative launch).
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - epoch) < 0);
    }
    __syncthreads();
}
// three sums at once: block partials (one shared-memory round), one grid barrier, then every block adds the
// partials in the same fixed order -> bit-reproducible and identical in all blocks
template <int BLOCK, bool CUSTOM>
__device__ __forceinline__ void grid_reduce3(cg::grid_group &grid, unsigned int *counter, unsigned int &epoch, double a, double b, double c,
                                             double *partial, double (*sh)[3], double &ra, double &rb, double &rc)
{
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { sh[w][0] = a; sh[w][1] = b; sh[w][2] = c; }
    __syncthreads();
    if (w == 0) {
        double t0 = lane < BLOCK / 32 ? sh[lane][0] : 0.0, t1 = lane < BLOCK / 32 ? sh[lane][1] : 0.0, t2 = lane < BLOCK / 32 ? sh[lane][2] : 0.0;
        t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
        if (lane == 0) { partial[blockIdx.x] = t0; partial[nb + blockIdx.x] = t1; partial[2 * nb + blockIdx.x] = t2; }
    }
    if (CUSTOM) grid_barrier(counter, epoch); else grid.sync();
    if (w < 3) {
        double s0 = 0.0, s1 = 0.0;
        const volatile double *pp = partial + w * nb;
        int i = lane;
        for (; i + 32 < nb; i += 64) { s0 += pp[i]; s1 += pp[i + 32]; }
        if (i < nb) s0 += pp[i];
        double t = warp_sum(s0 + s1);
        if (lane == 0) sh[0][w] = t;
    }
    __syncthreads();
    ra = sh[0][0]; rb = sh[0][1]; rc = sh[0][2];
    __syncthreads();
}

// two sums at once, second generation (k_pcg_res, k_pcg_res2): half-warp butterflies, double-buffered partials, see k_pcg_res2
#define FULLMASK 0xffffffffu
template <int BLOCK>
__device__ __forceinline__ void grid_reduce2(unsigned int *counter, unsigned int &epoch, unsigned int &par, double a, double b, double *partial,
                                             double (*sh)[2], double (*res)[2], double &ra, double &rb)
{
    static_assert(BLOCK == 1024, "32 warps: the second level is one half-warp butterfly");
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool hi = lane >= 16;
    double keep = hi ? b : a, send = hi ? a : b;
    keep += __shfl_xor_sync(FULLMASK, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(FULLMASK, keep, o);
    if ((lane & 15) == 0) sh[w][hi] = keep;
    __syncthreads();
    double *pp = partial + (size_t)par * 2 * nb;   // [2][nb], buffer of this reduction
    if (w == 0) {
        double v = hi ? sh[lane - 16][1] + sh[lane][1] : sh[lane][0] + sh[lane + 16][0];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
        if ((lane & 15) == 0) pp[(hi ? nb : 0) + blockIdx.x] = v;
        __syncwarp();
        epoch += nb;
        // only thread 0 spins and nobody of its warp waits at a __syncwarp meanwhile: a lane spinning next to parked lanes of
        // the same warp costs +1.5 us per reduction on B200 (tools/bench_barrier4.cu)
        if (lane == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            unsigned int c;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(counter) : "memory"); } while ((int)(c - epoch) < 0);
        }
    } else
        epoch += nb;
    __syncthreads();
    if (w < 2) {      // warp 0 sums the first quantity, warp 1 the second: independent loads, fixed order
        constexpr int MAXJ = 5;    // up to 160 CTAs (B200: 148)
        double v[MAXJ];
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int i = lane + 32 * j;
            v[j] = 0.0;
            if (i < nb) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v[j]) : "l"(pp + w * nb + i) : "memory");
        }
        double t = (((v[0] + v[1]) + v[2]) + v[3]) + v[4];
        for (int i = lane + 32 * MAXJ; i < nb; i += 32) t += ((volatile double *)pp)[w * nb + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULLMASK, t, o);
        if (lane == 0) res[par][w] = t;
    }
    __syncthreads();
    ra = res[par][0]; rb = res[par][1];
    par ^= 1u;
}
// The same two sums when the whole solve runs in ONE thread-block cluster (small meshes): every CTA pushes its pair of partial sums
// into the slot it owns in every CTA's shared memory (st.shared::cluster), one hardware cluster barrier (release / acquire at
// cluster scope, ~0.2 us instead of the ~2 us of the global-memory barrier above), then every thread adds the slots in rank order
// -> the same value in all CTAs, bit-reproducible.  Double-buffered like grid_reduce2: one barrier per reduction suffices.
constexpr int PCG_CL_MAX = 16;
template <int BLOCK>
__device__ __forceinline__ void cluster_reduce2(cg::cluster_group &cl, unsigned int &par, double a, double b, double (*sh)[2],
                                                double (*cp)[PCG_CL_MAX][2], double &ra, double &rb)
{
    static_assert(BLOCK == 1024, "32 warps: the second level is one half-warp butterfly");
    const int nc = (int)cl.num_blocks(), lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool hi = lane >= 16;
    double keep = hi ? b : a, send = hi ? a : b;
    keep += __shfl_xor_sync(FULLMASK, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(FULLMASK, keep, o);
    if ((lane & 15) == 0) sh[w][hi] = keep;
    __syncthreads();
    if (w == 0) {
        double v = hi ? sh[lane - 16][1] + sh[lane][1] : sh[lane][0] + sh[lane + 16][0];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
        const double va = __shfl_sync(FULLMASK, v, 0), vb = __shfl_sync(FULLMASK, v, 16);
        if (lane < nc) {
            double *dst = cl.map_shared_rank(&cp[par][cl.block_rank()][0], lane);
            *reinterpret_cast<double2 *>(dst) = make_double2(va, vb);
        }
    }
    cl.sync();
    double s0 = 0.0, s1 = 0.0;
    for (int c = 0; c < nc; ++c) { const double2 q = *reinterpret_cast<const double2 *>(&cp[par][c][0]); s0 += q.x; s1 += q.y; }
    ra = s0; rb = s1;
    par ^= 1u;
}
__device__ __forceinline__ void l2_prefetch(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int BLOCK, bool CUSTOM, bool DD, int MINB = 1024 / BLOCK>
__global__ void __launch_bounds__(BLOCK, MINB) k_pcg(PcgArgs a)
{
    const bool PF = a.prefetch != 0;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[BLOCK / 32][3];
    __shared__ double shdd[4];
    unsigned int epoch = a.epoch0;
    unsigned int seq_ar = 0, seq_h = 0;
    if (DD) { seq_ar = a.dd.seq[0]; seq_h = a.dd.seq[1]; }
    const int n = a.n, stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const double *__restrict__ dg = a.diag;
    const unsigned char *__restrict__ own = a.own;
    // x0 = M^-1 b ; xlung = ||b_free||^2   (PRODDP call at :4686, XLUNG at :1286-1297)
    double xl = 0.0;
    for (int k = t0; k < n; k += stride) {
        double b = a.rhs[k];
        a.x[k] = b / dg[k];
        if ((!DD || (own[k] & 1)) && !(a.cm ? dg[k] > 1.0e80 : is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag))) xl += b * b;
    }
    double xlung, d1, d2;
    grid_reduce3<BLOCK, CUSTOM>(grid, a.counter, epoch, xl, 0.0, 0.0, a.partial, sh, xlung, d1, d2);
    if (DD) {   // global ||b||^2, and x0 on the ghost rows from their owners
        double v[1] = {xlung};
        dd_allreduce<1>(a.dd, seq_ar, v, shdd);
        xlung = v[0];
        dd_send_rows(a.dd, seq_h + 1u, a.x, t0, stride);
        grid_barrier(a.counter, epoch);
        dd_recv_rows(a.dd, seq_h, a.x, t0, stride);
        grid_barrier(a.counter, epoch);
    }
    // r = b - A x0 ; z = M^-1 r ; p_old = 0 so that p = z in the first phase A
    for (int k = t0; k < n; k += stride) {
        double r = a.rhs[k] - dia_row(a.A, dg, a.x, k, n);
        a.r[k] = r;
        a.z[k] = r / dg[k];
        a.p0[k] = 0.0;
    }
    if (CUSTOM) grid_barrier(a.counter, epoch); else grid.sync();
    if (DD) {
        dd_send_rows(a.dd, seq_h + 1u, a.z, t0, stride);
        grid_barrier(a.counter, epoch);
        dd_recv_rows(a.dd, seq_h, a.z, t0, stride);
        grid_barrier(a.counter, epoch);
    }
    double beta = 0.0, err = 0.0;
    double *pold = a.p0, *pnew = a.p1;
    int niter = 1;
    for (;;) {
        // ---- phase A
        double s_pr = 0.0, s_pb = 0.0;
        {
            const double *z = a.z;      // NOT __restrict__/read-only: rewritten every iteration by other SMs
            const double *po = pold;
            for (int k = t0; k < n; k += stride) {
                if (PF && k + stride < n) {   // pull the next row's DRAM-bound streams into L2 while this row's FMA chain runs
                    const int kn = k + stride;
#pragma unroll
                    for (int d = 1; d < NDIAG; ++d) l2_prefetch(&a.A.d[d][kn]);
                    l2_prefetch(&dg[kn]); l2_prefetch(&z[kn]); l2_prefetch(&po[kn]); l2_prefetch(&a.r[kn]);
                }
                double pk = z[k] + beta * po[k];
                double acc = dg[k] * pk;
#pragma unroll
                for (int d = 1; d < NDIAG; ++d) {
                    const int o = a.A.off[d];
                    acc += a.A.d[d][k] * (z[k + o] + beta * po[k + o]);
                }
#pragma unroll
                for (int d = 1; d < NDIAG; ++d) {
                    const int o = a.A.off[d];
                    acc += a.A.d[d][k - o] * (z[k - o] + beta * po[k - o]);
                }
                pnew[k] = pk;
                a.bv[k] = acc;
                if (!DD || (own[k] & 1)) { s_pr += pk * a.r[k]; s_pb += pk * acc; }
            }
        }
        double pr, pb;
        grid_reduce3<BLOCK, CUSTOM>(grid, a.counter, epoch, s_pr, s_pb, 0.0, a.partial, sh, pr, pb, d1);
        if (DD) { double v[2] = {pr, pb}; dd_allreduce<2>(a.dd, seq_ar, v, shdd); pr = v[0]; pb = v[1]; }
        double alfa = pr / pb;
        // ---- phase B (row-block partition: the new z of my boundary rows goes straight into the neighbours' inboxes)
        double s_bz = 0.0, s_rr = 0.0;
        const int hpar = (seq_h + 1u) & 1u;
        bool sent = false;
        for (int k = t0; k < n; k += stride) {
            if (PF && k + stride < n) {
                const int kn = k + stride;
                l2_prefetch(&a.bv[kn]); l2_prefetch(&a.r[kn]); l2_prefetch(&a.x[kn]); l2_prefetch(&pnew[kn]); l2_prefetch(&dg[kn]);
            }
            double bk = a.bv[k];
            double r = a.r[k] - alfa * bk;
            a.r[k] = r;
            a.x[k] = a.x[k] + alfa * pnew[k];
            double zz = r / dg[k];
            a.z[k] = zz;
            if (!DD || (own[k] & 1)) {
                s_bz += bk * zz;
                if (!(a.cm ? dg[k] > 1.0e80 : is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag))) s_rr += r * r;
            }
            if (DD && (own[k] & 6)) {
                const DDCtx &c = a.dd;
                int l = k / c.nnod, sidx = k - l * c.nnod, row = sidx / c.nc1, j = sidx - row * c.nc1;
                if ((own[k] & 2) && c.north >= 0) c.inbox_peer[c.north][((size_t)hpar * 2 + 1) * c.hcap + ((size_t)l * DD_W + (row - c.own_a)) * c.nc1 + j] = zz;
                if ((own[k] & 4) && c.south >= 0) c.inbox_peer[c.south][((size_t)hpar * 2 + 0) * c.hcap + ((size_t)l * DD_W + (row - (c.own_b - DD_W))) * c.nc1 + j] = zz;
                sent = true;
            }
        }
        if (DD && sent) __threadfence_system();
        double bz, rr;
        grid_reduce3<BLOCK, CUSTOM>(grid, a.counter, epoch, s_bz, s_rr, 0.0, a.partial, sh, bz, rr, d1);
        if (DD) {
            double v[2] = {bz, rr};
            dd_allreduce<2>(a.dd, seq_ar, v, shdd);
            bz = v[0]; rr = v[1];
            dd_recv_rows(a.dd, seq_h, a.z, t0, stride);     // publish my rows (stored in phase B), fetch the neighbours'
            grid_barrier(a.counter, epoch);
        }
        beta = -bz / pb;
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / n);
        double *t = pold; pold = pnew; pnew = t;
        if (err > a.tol && niter < a.itmax && !(DD && *(volatile int *)a.dd.err)) { ++niter; continue; }
        break;
    }
    if (t0 == 0) {
        a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)epoch;
        if (DD) { a.dd.seq[0] = seq_ar; a.dd.seq[1] = seq_h; }
    }
}



// ------------------------------------------------------------------------------------------
// SYMSLV with the CG vectors RESIDENT IN SHARED MEMORY (default whenever they fit: n <= #CTAs x ~9.6k rows, i.e. up to
// ~1.4 M nodes on one B200).  Every CTA owns a contiguous block of rows_cta rows for the whole solve and keeps r, p and
// B = A p (and x when there is room) of its rows in its 227 KB of shared memory; only z = M^-1 r, which the neighbours'
// stencils need, goes through global memory (L2).  The search direction is never gathered: by linearity
//     p = z + beta p_old   =>   B = A p = A z + beta B_old,
// so phase A is ONE stencil product on z (15 gathered operands per row instead of 29) and two shared-memory recurrences.
// Same recurrence otherwise (GRADDP, SRC/solscal-extended.f:1260-1380): x0 = M^-1 b, alfa = (p.r)/(p.B),
// beta = -(B.z)/(p.B), residual test on the non-Dirichlet rows; two grid barriers per iteration, fixed-order reductions.
// Per row and iteration the kernel moves 8 diagonals + z (read, write) + the diagonal again in phase B = 88 B
// (+16 B for x when it is not resident) instead of 168 B.
// ------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1) k_pcg_res(PcgArgs a)
{
    extern __shared__ __align__(16) double smv[];
    __shared__ double sh[BLOCK / 32][2];
    __shared__ double res[2][2];
    unsigned int epoch = a.epoch0, par = 0;
    const bool PF = a.prefetch != 0;
    const int R = a.rows_cta, row0 = blockIdx.x * R, cnt = max(0, min(R, a.n - row0)), tid = threadIdx.x;
    double *rs = smv, *ps = smv + R, *bs = smv + 2 * (size_t)R, *xs = a.xres ? smv + 3 * (size_t)R : a.x + row0;
    const double *__restrict__ dg = a.diag;
    // x0 = M^-1 b ; xlung = ||b_free||^2 ; Dirichlet rows of this thread as a bit mask (row j*BLOCK + tid -> bit j)
    unsigned int dmask = 0;
    double xl = 0.0;
    for (int i = tid, j = 0; i < cnt; i += BLOCK, ++j) {
        const int k = row0 + i;
        double b = a.rhs[k];
        a.x[k] = b / dg[k];
        if (is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag)) dmask |= 1u << j; else xl += b * b;
    }
    double xlung, d1;
    grid_reduce2<BLOCK>(a.counter, epoch, par, xl, 0.0, a.partial, sh, res, xlung, d1);
    // r = b - A x0 ; z = M^-1 r ; p = B = 0
    for (int i = tid; i < cnt; i += BLOCK) {
        const int k = row0 + i;
        double r = a.rhs[k] - dia_row(a.A, dg, a.x, k, a.n);
        rs[i] = r;
        a.z[k] = r / dg[k];
        ps[i] = 0.0;
        bs[i] = 0.0;
        if (a.xres) xs[i] = a.x[k];
    }
    grid_barrier(a.counter, epoch);
    double beta = 0.0, err = 0.0;
    int niter = 1;
    const double *z = a.z;      // NOT __restrict__/read-only: rewritten every iteration by other SMs
    for (;;) {
        // ---- phase A: B = A z + beta B, p = z + beta p, (p.r), (p.B)
        double s_pr = 0.0, s_pb = 0.0;
        for (int i = tid; i < cnt; i += BLOCK) {
            const int k = row0 + i;
            if (PF && i + BLOCK < cnt) {
#pragma unroll
                for (int d = 1; d < NDIAG; ++d) l2_prefetch(&a.A.d[d][k + BLOCK]);
                l2_prefetch(&dg[k + BLOCK]);
            }
            const double zk = z[k];
            double acc = dg[k] * zk;
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) acc += a.A.d[d][k] * z[k + a.A.off[d]];
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) acc += a.A.d[d][k - a.A.off[d]] * z[k - a.A.off[d]];
            const double pk = zk + beta * ps[i], bk = acc + beta * bs[i];
            ps[i] = pk;
            bs[i] = bk;
            s_pr += pk * rs[i];
            s_pb += pk * bk;
        }
        double pr, pb;
        grid_reduce2<BLOCK>(a.counter, epoch, par, s_pr, s_pb, a.partial, sh, res, pr, pb);
        const double alfa = pr / pb;
        // ---- phase B: r -= alfa B, x += alfa p, z = M^-1 r, (B.z), ||r_free||^2
        double s_bz = 0.0, s_rr = 0.0;
        for (int i = tid, j = 0; i < cnt; i += BLOCK, ++j) {
            const int k = row0 + i;
            const double bk = bs[i], r = rs[i] - alfa * bk;
            rs[i] = r;
            xs[i] += alfa * ps[i];
            const double zz = r / dg[k];
            a.z[k] = zz;
            s_bz += bk * zz;
            if (!((dmask >> j) & 1u)) s_rr += r * r;
        }
        double bz, rr;
        grid_reduce2<BLOCK>(a.counter, epoch, par, s_bz, s_rr, a.partial, sh, res, bz, rr);
        beta = -bz / pb;
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / a.n);
        if (err > a.tol && niter < a.itmax) { ++niter; continue; }
        break;
    }
    if (a.xres) for (int i = tid; i < cnt; i += BLOCK) a.x[row0 + i] = xs[i];
    if (blockIdx.x == 0 && tid == 0) { a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)epoch; }
}

// ------------------------------------------------------------------------------------------
// k_pcg_res2 (default, CATHY_PCG_ALGO=4): the resident-vector PCG above, re-cut after ncu showed k_pcg_res bound by the L1/LSU
// data pipe (l1tex__data_pipe_lsu_wavefronts 55 % of peak over the whole launch, ~90 % inside the phases) and by the grid
// reduction (tools/bench_barrier*.cu: 2.9 us each = 1400 cycles of fp64 shuffles + 870 fence + 1480 arrive/poll + 1000 re-read):
//  * every thread owns TWO consecutive rows (k0 even, k0+1) and loads 16-byte aligned pairs; the element that a misaligned
//    window lacks comes from the neighbouring lane by shuffle (edge lanes fetch it themselves).  The stencil offsets come in
//    pairs (o, o+1) -- {-1,0,1}, {NC1,NC1+1}, {NNOD-NC1-1,NNOD-NC1}, {NNOD-1,NNOD} -- so one 3-element z window serves two
//    diagonals of both rows: ~60 instead of 81 LSU wavefronts per 32 rows;
//  * M^-1 is applied as a multiplication with the reciprocal diagonal computed once per solve (no fp64 division per row);
//  * the grid reduction sums two quantities in ONE half-warp butterfly (a in lanes 0-15, b in lanes 16-31), the partials are
//    double-buffered (a fast CTA can no longer overwrite what a slow one still reads) and fetched with independent loads.
// Same recurrence and stopping test as k_pcg_res; inside a row the products are summed pair of diagonals by pair of diagonals.
// ------------------------------------------------------------------------------------------
// paired-row loads: this thread needs p[0..1] (pair) or p[0..2] (win3); lanes own consecutive pairs of rows, so lane+1 needs
// p[2..], lane-1 p[-2..].  ODD (compile time, uniform): p is 8 but not 16 bytes aligned.  The 16-byte aligned pair is loaded, the
// missing element comes from the neighbouring lane by shuffle; edge_lo / edge_hi: the lane below / above does not hold the
// continuation (lane 0 / lane 31 or the last active pair) and the element is fetched directly.  Loads (`*_ld`) and shuffles
// (`*_fin`) are separate calls so that all loads of a group are in flight before the first shuffle waits for one of them.
struct PairLd { double2 q; double e; };
template <bool ODD> __device__ __forceinline__ PairLd pair_ld(const double *p, bool edge_hi)
{
    PairLd r; r.e = 0.0;
    if (!ODD) r.q = *reinterpret_cast<const double2 *>(p);
    else { r.q = *reinterpret_cast<const double2 *>(p - 1); if (edge_hi) r.e = p[1]; }
    return r;
}
template <bool ODD> __device__ __forceinline__ void pair_fin(const PairLd &r, bool edge_hi, double &v0, double &v1)
{
    if (!ODD) { v0 = r.q.x; v1 = r.q.y; }
    else { v0 = r.q.y; const double t = __shfl_down_sync(FULLMASK, r.q.x, 1); v1 = edge_hi ? r.e : t; }
}
template <bool ODD> __device__ __forceinline__ PairLd win3_ld(const double *p, bool edge_lo, bool edge_hi)
{
    PairLd r; r.e = 0.0;
    if (!ODD) { r.q = *reinterpret_cast<const double2 *>(p); if (edge_hi) r.e = p[2]; }
    else { r.q = *reinterpret_cast<const double2 *>(p + 1); if (edge_lo) r.e = p[0]; }
    return r;
}
template <bool ODD> __device__ __forceinline__ void win3_fin(const PairLd &r, bool edge_lo, bool edge_hi, double &v0, double &v1, double &v2)
{
    if (!ODD) { v0 = r.q.x; v1 = r.q.y; const double t = __shfl_down_sync(FULLMASK, r.q.x, 1); v2 = edge_hi ? r.e : t; }
    else { v1 = r.q.x; v2 = r.q.y; const double t = __shfl_up_sync(FULLMASK, r.q.y, 1); v0 = edge_lo ? r.e : t; }
}
// one pair of diagonals (o, o+1) = (da, da+1): upper and lower products of rows k, k+1
template <bool ODD>
__device__ __forceinline__ void pair_group(const Diag &A, const double *z, int da, int o, int k, bool elo, bool ehi, double &a0, double &a1)
{
    const double2 ua = *reinterpret_cast<const double2 *>(A.d[da] + k), ub = *reinterpret_cast<const double2 *>(A.d[da + 1] + k);
    const PairLd rw = win3_ld<ODD>(z + k + o, elo, ehi), rm = win3_ld<!ODD>(z + k - o - 1, elo, ehi);
    const PairLd ra = pair_ld<ODD>(A.d[da] + k - o, ehi), rb = pair_ld<!ODD>(A.d[da + 1] + k - o - 1, ehi);   // L_d = (A_d[k - off_d], A_d[k + 1 - off_d])
    double w0, w1, w2, m0, m1, m2, la0, la1, lb0, lb1;
    win3_fin<ODD>(rw, elo, ehi, w0, w1, w2);
    win3_fin<!ODD>(rm, elo, ehi, m0, m1, m2);
    pair_fin<ODD>(ra, ehi, la0, la1);
    pair_fin<!ODD>(rb, ehi, lb0, lb1);
    a0 += ua.x * w0;  a1 += ua.y * w1;
    a0 += ub.x * w1;  a1 += ub.y * w2;
    a0 += la0 * m1;   a1 += la1 * m2;
    a0 += lb0 * m0;   a1 += lb1 * m1;
}
// CL = true: the grid is ONE thread-block cluster (meshes of a few thousand to a few ten thousand rows, e.g. BASELINE config 1 and the
// members of small-catchment ensembles): reductions and barriers are cluster-scope (cluster_reduce2), so an iteration costs ~2 us
// instead of ~7 us, and a solve occupies only its cluster's SMs -- other members' solves run beside it.
template <int BLOCK, int PAR, bool CL = false>     // PAR: parities of the offsets off[2], off[4], off[6] (bits 0, 1, 2)
__global__ void __launch_bounds__(BLOCK, 1) k_pcg_res2(PcgArgs a)
{
    extern __shared__ __align__(16) double smv[];
    __shared__ double sh[BLOCK / 32][2];
    __shared__ double res[2][2];
    __shared__ __align__(16) double cpart[2][PCG_CL_MAX][2];
    cg::cluster_group cl = cg::this_cluster();
    unsigned int epoch = a.epoch0, par = 0;
    const int R = a.rows_cta, row0 = blockIdx.x * R, cnt = max(0, min(R, a.n - row0)), tid = threadIdx.x, lane = tid & 31;
    double *rs = smv, *ps = smv + R, *bs = smv + 2 * (size_t)R, *xs = a.xres ? smv + 3 * (size_t)R : a.x + row0;
    const double *__restrict__ dg = a.diag;
    double *dinv = a.p0;          // k_pcg's search-direction buffer is free here: reciprocal diagonal
    // x0 = M^-1 b ; xlung = ||b_free||^2 ; Dirichlet rows of this thread as a bit mask (pass j: rows 2 tid + 2 BLOCK j + {0,1} -> bits 2j, 2j+1)
    unsigned int dmask = 0;
    double xl = 0.0;
    for (int i = 2 * tid, j = 0; i < cnt; i += 2 * BLOCK, ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (i + q < cnt) {
                const int k = row0 + i + q;
                const double b = a.rhs[k], dv = 1.0 / dg[k];
                dinv[k] = dv;
                a.x[k] = b * dv;
                if (is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag)) dmask |= 1u << (2 * j + q); else xl += b * b;
            }
    double xlung, d1;
    if (CL) cluster_reduce2<BLOCK>(cl, par, xl, 0.0, sh, cpart, xlung, d1);
    else grid_reduce2<BLOCK>(a.counter, epoch, par, xl, 0.0, a.partial, sh, res, xlung, d1);
    // r = b - A x0 ; z = M^-1 r ; p = B = 0
    for (int i = tid; i < cnt; i += BLOCK) {
        const int k = row0 + i;
        const double r = a.rhs[k] - dia_row(a.A, dg, a.x, k, a.n);
        rs[i] = r;
        a.z[k] = r * dinv[k];
        ps[i] = 0.0;
        bs[i] = 0.0;
        if (a.xres) xs[i] = a.x[k];
    }
    if (CL) cl.sync(); else grid_barrier(a.counter, epoch);
    double beta = 0.0, err = 0.0;
    int niter = 1;
    const double *z = a.z;      // NOT __restrict__/read-only: rewritten every iteration by other SMs
    const int o2 = a.A.off[2], o4 = a.A.off[4], o6 = a.A.off[6];      // off[1] = 1, off[3] = o2 + 1, off[5] = o4 + 1, off[7] = o6 + 1 (checked by the host)
    const int last = (cnt - 1) & ~1;                                   // first row of the last pair
    const int iwarp_end = cnt;                                         // a warp runs a pass while its first pair exists
    for (;;) {
        // ---- phase A: B = A z + beta B, p = z + beta p, (p.r), (p.B)
        double s_pr = 0.0, s_pb = 0.0;
        for (int iw = 2 * (tid - lane); iw < iwarp_end; iw += 2 * BLOCK) {
            const int i_own = iw + 2 * lane;
            const bool act = i_own < cnt, ok1 = i_own + 1 < cnt;
            const int i = act ? i_own : last;                          // idle lanes of the last warp shadow the last pair (their shuffles feed nobody)
            const bool ehi = lane == 31 || i_own + 2 >= cnt, elo = lane == 0;
            const int k = row0 + i;
            // centre window z[k-1..k+2]: (z[k], z[k+1]) is the aligned pair
            const double2 zc = *reinterpret_cast<const double2 *>(z + k);
            double zm = __shfl_up_sync(FULLMASK, zc.y, 1), zp = __shfl_down_sync(FULLMASK, zc.x, 1);
            if (elo) zm = z[k - 1];
            if (ehi) zp = z[k + 2];
            const double2 dd = *reinterpret_cast<const double2 *>(dg + k);
            const double2 u1 = *reinterpret_cast<const double2 *>(a.A.d[1] + k);
            // lower part of diagonal 1: A1[k-1] (from the lane below), A1[k] = u1.x
            double l1 = __shfl_up_sync(FULLMASK, u1.y, 1);
            if (elo) l1 = a.A.d[1][k - 1];
            double a0 = dd.x * zc.x, a1 = dd.y * zc.y;
            a0 += u1.x * zc.y;  a1 += u1.y * zp;
            a0 += l1 * zm;      a1 += u1.x * zc.x;
            // the three offset pairs (o, o+1), one after the other (keeps the live registers under the 64 a 1024-thread CTA gets)
            pair_group<(PAR & 1) != 0>(a.A, z, 2, o2, k, elo, ehi, a0, a1);
            pair_group<(PAR & 2) != 0>(a.A, z, 4, o4, k, elo, ehi, a0, a1);
            pair_group<(PAR & 4) != 0>(a.A, z, 6, o6, k, elo, ehi, a0, a1);
            if (act) {
                double2 pv = *reinterpret_cast<double2 *>(ps + i), bv = *reinterpret_cast<double2 *>(bs + i);
                const double2 rv = *reinterpret_cast<const double2 *>(rs + i);
                pv.x = zc.x + beta * pv.x; pv.y = zc.y + beta * pv.y;
                bv.x = a0 + beta * bv.x;   bv.y = a1 + beta * bv.y;
                s_pr += pv.x * rv.x; s_pb += pv.x * bv.x;
                if (ok1) {
                    s_pr += pv.y * rv.y; s_pb += pv.y * bv.y;
                    *reinterpret_cast<double2 *>(ps + i) = pv; *reinterpret_cast<double2 *>(bs + i) = bv;
                } else { ps[i] = pv.x; bs[i] = bv.x; }
            }
        }
        double pr, pb;
        if (CL) cluster_reduce2<BLOCK>(cl, par, s_pr, s_pb, sh, cpart, pr, pb);
        else grid_reduce2<BLOCK>(a.counter, epoch, par, s_pr, s_pb, a.partial, sh, res, pr, pb);
        const double alfa = pr / pb;
        // ---- phase B: r -= alfa B, x += alfa p, z = M^-1 r, (B.z), ||r_free||^2
        double s_bz = 0.0, s_rr = 0.0;
        for (int i = 2 * tid, j = 0; i < cnt; i += 2 * BLOCK, ++j) {
            const int k = row0 + i;
            const bool ok1 = i + 1 < cnt;
            const double2 bv = *reinterpret_cast<const double2 *>(bs + i), pv = *reinterpret_cast<const double2 *>(ps + i);
            double2 rv = *reinterpret_cast<double2 *>(rs + i);
            const double2 dv = *reinterpret_cast<const double2 *>(dinv + k);
            rv.x -= alfa * bv.x; rv.y -= alfa * bv.y;
            double2 zz; zz.x = rv.x * dv.x; zz.y = rv.y * dv.y;
            s_bz += bv.x * zz.x;
            if (!((dmask >> (2 * j)) & 1u)) s_rr += rv.x * rv.x;
            if (ok1) {
                double2 xv = *reinterpret_cast<double2 *>(xs + i);
                xv.x += alfa * pv.x; xv.y += alfa * pv.y;
                *reinterpret_cast<double2 *>(xs + i) = xv;
                *reinterpret_cast<double2 *>(rs + i) = rv;
                *reinterpret_cast<double2 *>(a.z + k) = zz;
                s_bz += bv.y * zz.y;
                if (!((dmask >> (2 * j + 1)) & 1u)) s_rr += rv.y * rv.y;
            } else { xs[i] += alfa * pv.x; rs[i] = rv.x; a.z[k] = zz.x; }
        }
        double bz, rr;
        if (CL) cluster_reduce2<BLOCK>(cl, par, s_bz, s_rr, sh, cpart, bz, rr);
        else grid_reduce2<BLOCK>(a.counter, epoch, par, s_bz, s_rr, a.partial, sh, res, bz, rr);
        beta = -bz / pb;
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / a.n);
        if (err > a.tol && niter < a.itmax) { ++niter; continue; }
        break;
    }
    if (a.xres) for (int i = tid; i < cnt; i += BLOCK) a.x[row0 + i] = xs[i];
    if (blockIdx.x == 0 && tid == 0) { a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)epoch; }
}

#include "pcg_cluster.cuh"

// ------------------------------------------------------------------------------------------
// SYMSLV, second formulation (opt-in, CATHY_PCG_ALGO=2; measured slower than k_pcg on B200 except on tiny meshes, see
// profiles/r1_pcg_experiments.md): the system is scaled symmetrically, As = D^-1/2 A D^-1/2 (unit
// diagonal, y = D^1/2 x), so that the Jacobi-preconditioned CG of k_pcg becomes plain CG without the z vector and without the
// diagonal; and the recurrence is the single-reduction form of CG (Chronopoulos & Gear): with w = As r,
//     gamma = (r,r), delta = (w,r);  beta = gamma/gamma_old;  alpha = gamma / (delta - beta*gamma/alpha_old)
//     p = r + beta p;  s = w + beta s;  y += alpha p;  r -= alpha s;  w = As r
// The new w needs the new r of the 14 neighbours, which every thread recomputes on the fly from the OLD r, w, s
// (r_j - alpha (w_j + beta s_j)); r, w, s are double-buffered.  ONE grid-wide barrier per iteration (inside the reduction)
// instead of two, 144 instead of 168 bytes per row and iteration.  Same iterates as SYMSLV/GRADDP in exact arithmetic (same
// x0 = M^-1 b, same stopping test on the unscaled residual, Dirichlet rows excluded).
// ------------------------------------------------------------------------------------------
__global__ void k_sym_scale(int n, Diag A, const double *__restrict__ diag_bc, double *__restrict__ dis)
{   // pass 1: dis = 1/sqrt(diag)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) dis[k] = 1.0 / sqrt(diag_bc[k]);
}
__global__ void k_sym_scale2(int n, Diag A, const double *__restrict__ dis)
{   // pass 2: off-diagonals in place (dis carries a halo; the entries that reach into it are structurally zero)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double dk = dis[k];
#pragma unroll
        for (int d = 1; d < NDIAG; ++d) A.d[d][k] = (A.d[d][k] * dk) * dis[k + A.off[d]];
    }
}
struct Pcg2Args {
    int n, nnod, itmax, prefetch;
    double tol;
    Diag A;                  // scaled off-diagonals in d[1..7]
    const double *dis;       // 1/sqrt(diagonal with the Dirichlet penalty)
    const double *rhs;
    double *y, *p, *r0, *r1, *w0, *w1, *s0, *s1;
    const int *ifatm;
    const unsigned char *contp_flag;
    double *partial;
    unsigned int *counter;
    unsigned int epoch0;
    IterOut *out;
};
__device__ __forceinline__ double dia_offrow(const Diag &A, const double *x, int k)
{   // sum over the 14 off-diagonal entries of row k (unit diagonal not included)
    double acc = 0.0;
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) acc += A.d[d][k] * x[k + A.off[d]];
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) acc += A.d[d][k - A.off[d]] * x[k - A.off[d]];
    return acc;
}
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) k_pcg2(Pcg2Args a)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[BLOCK / 32][3];
    unsigned int epoch = a.epoch0;
    const int n = a.n, stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const double *__restrict__ dis = a.dis;
    const bool PF = a.prefetch != 0;
    // y0 = D^1/2 x0 = b/sqrt(d) (x0 = M^-1 b, :4686);  xlung = ||b_free||^2 (:1286-1297)
    double xl = 0.0;
    for (int k = t0; k < n; k += stride) {
        double b = a.rhs[k];
        a.y[k] = b * dis[k];
        a.p[k] = 0.0; a.s0[k] = 0.0;
        if (!is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag)) xl += b * b;
    }
    double xlung, g0, d0;
    grid_reduce3<BLOCK, true>(grid, a.counter, epoch, xl, 0.0, 0.0, a.partial, sh, xlung, g0, d0);
    // r = b~ - As y0
    for (int k = t0; k < n; k += stride) a.r0[k] = a.rhs[k] * dis[k] - (a.y[k] + dia_offrow(a.A, a.y, k));
    grid_barrier(a.counter, epoch);
    // w = As r ; gamma = (r,r) ; delta = (w,r)
    double sg = 0.0, sd = 0.0;
    for (int k = t0; k < n; k += stride) {
        double r = a.r0[k], w = r + dia_offrow(a.A, a.r0, k);
        a.w0[k] = w;
        sg += r * r; sd += w * r;
    }
    double gamma, delta, rr;
    grid_reduce3<BLOCK, true>(grid, a.counter, epoch, sg, sd, 0.0, a.partial, sh, gamma, delta, rr);
    double alpha = gamma / delta, beta = 0.0, err = 0.0;
    double *rc = a.r0, *rn = a.r1, *wc = a.w0, *wn = a.w1, *sc = a.s0, *sn = a.s1;
    int niter = 1;
    for (;;) {
        const double ab = alpha * beta;
        double s_g = 0.0, s_d = 0.0, s_rr = 0.0;
        for (int k = t0; k < n; k += stride) {
            if (PF && k + stride < n) {
                const int kn = k + stride;
#pragma unroll
                for (int d = 1; d < NDIAG; ++d) l2_prefetch(&a.A.d[d][kn]);
                l2_prefetch(&rc[kn]); l2_prefetch(&wc[kn]); l2_prefetch(&sc[kn]); l2_prefetch(&a.p[kn]); l2_prefetch(&a.y[kn]); l2_prefetch(&dis[kn]);
            }
            const double r = rc[k], w = wc[k], so = sc[k];
            const double s = w + beta * so;
            const double p = r + beta * a.p[k];
            const double r2 = (r - alpha * w) - ab * so;      // = r - alpha s, in the very form the neighbours use below
            double acc = r2;                                 // unit diagonal
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) {
                const int j = k + a.A.off[d];
                acc += a.A.d[d][k] * ((rc[j] - alpha * wc[j]) - ab * sc[j]);
            }
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) {
                const int j = k - a.A.off[d];
                acc += a.A.d[d][j] * ((rc[j] - alpha * wc[j]) - ab * sc[j]);
            }
            sn[k] = s; a.p[k] = p; a.y[k] = a.y[k] + alpha * p; rn[k] = r2; wn[k] = acc;
            s_g += r2 * r2; s_d += acc * r2;
            if (!is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag)) { double di = dis[k]; s_rr += (r2 * r2) / (di * di); }   // unscaled residual
        }
        double g1, d1;
        grid_reduce3<BLOCK, true>(grid, a.counter, epoch, s_g, s_d, s_rr, a.partial, sh, g1, d1, rr);
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / n);
        double *t;
        t = rc; rc = rn; rn = t; t = wc; wc = wn; wn = t; t = sc; sc = sn; sn = t;
        if (err > a.tol && niter < a.itmax) {
            beta = g1 / gamma;
            alpha = g1 / (d1 - beta * g1 / alpha);
            gamma = g1;
            ++niter;
            continue;
        }
        break;
    }
    // x = D^-1/2 y
    for (int k = t0; k < n; k += stride) a.y[k] = a.y[k] * dis[k];
    if (t0 == 0) { a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)epoch; }
}
