// cathy_prepro.cu -- the CATHY pre-processor (terrain analysis: CSORT, DEPIT, CCA, SMEAN, DSF, HG) on sm_100a.
// C ABI: include/cathy_prepro.h.  PRE = /root/reference/examples/SSHydro/weill_exemple/prepro/src (HAP v11.7).
//
// What is parallel and what is not.  The reference is a chain of sequential sweeps whose RESULTS depend on their order; the work
// here is finding the parallelism that leaves every result bit-identical:
//  * CSORT: the reference's unstable quicksort is REPLAYED, not replaced -- the permutation it leaves among equal elevations reaches
//    qoi_a, the routing order of SRC/route.f:47-56.  Hoare's partition loop is a function of the incoming values (flag scans + pairing
//    + independent exchanges, one 1024-thread CTA) and sub-arrays are finished side by side in shared memory (k_pp_qsplit / k_pp_qsmall).
//  * DEPIT: in-place Gauss-Seidel raising of pits; its fixed point depends on the visiting order, so ONE CTA visits in the reference's
//    order, with eight lanes per visit, a heap-driven first sweep that skips the visits that cannot change anything, and the parallel
//    quicksort replay between sweeps.  A parallel check first proves the common case "no pit at all".
//  * DSF: the reference sweeps the cells in descending elevation and lets every cell push area and deviation sums to its receivers
//    (PRE/dsf.f90:71-529).  Here a dependency WAVEFRONT: a cell is ready when all neighbours that precede it in the elevation order are
//    finished, then GATHERS the contributions of its donors in that same order, so the floating-point sums are the reference's while
//    all cells of a wavefront run side by side (persistent cooperative grid, one grid barrier per wavefront, O(cells) work).
//  * window analysis (facets, curvature), hydraulic geometry: one thread per cell; SMEAN: one warp adding in the reference's order.
// All arithmetic that decides something goes through __dmul_rn / __dadd_rn / __ddiv_rn: nvcc would contract a*b+c into FMAs, the
// reference binary (gfortran -O, x86-64) has none.  There is no host fallback for any of it.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
#include "../../include/cathy_prepro.h"

namespace cg = cooperative_groups;

static thread_local char pp_err[512] = "";
#define PFAIL(code, ...) do { snprintf(pp_err, sizeof pp_err, __VA_ARGS__); rc = (code); goto done; } while (0)
#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) PFAIL(-100, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// round-to-nearest operations that nvcc may not contract into FMAs: the reference binary (gfortran -O, x86-64) has none
__device__ __forceinline__ double rmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double radd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double rsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double rdiv(double a, double b) { return __ddiv_rn(a, b); }

#define EPS64 2.220446049250313e-16
#define EPS32 1.1920929e-07f

struct PP {
    CathyPreproParams h;
    int N, M, nb, nc;
    double *q;               // [nb] elevation, -1 = no cell (dtm_quota, PRE/mbbio.f90:766-776)
    unsigned char *pres;     // [nb]
    int *order, *rank;       // [nc+1] i_basin by descending elevation; [nb] position of a cell in it
    double *key;             // [nc+1] sort keys
    int *lst1, *lst2, *stamp;
    // window analysis
    double *smax, *dev1, *dev2, *de1, *de2;
    signed char *po1, *po2;
    float *Kp;
    int *dm0;
    // sweep
    double *Ain, *sdn, *Aout, *sd1, *sd2;
    float *w1, *w2, *ls1, *ls2, *epl1, *epl2, *ASk;
    int *p1, *p2, *hc, *dm, *hc_after;
    unsigned char *sflag;    // bit 0: direction 1 carries the deviation sum, bit 1: direction 2 does
    int *dep, *front[2];
    int *cnt;                // [16] counters: 0..2 frontier sizes, 3 processed, 4 waves, 5 pits, 6 modifications, 7 error, 8 passes
    double *scal;            // [4] mean_s_max
    // hydraulic geometry
    float *Ws1, *Ws2, *b1, *kS1, *kS2, *y1, *nrc;
};

// ------------------------------------------------------------------ QSORT (PRE/qsort.f90:9-125), arr/brr 1-based
__device__ void pp_qsort(int n, double *arr, int *brr)
{
    const int MM = 7;
    int istack[64];
    int jstack = 0, l = 1, ir = n;
    for (;;) {
        if (ir - l < MM) {
            for (int j = l + 1; j <= ir; ++j) {
                double a = arr[j];
                int b = brr[j];
                int i = j - 1;
                for (; i >= 1; --i) {
                    if (arr[i] <= a) break;
                    arr[i + 1] = arr[i];
                    brr[i + 1] = brr[i];
                }
                arr[i + 1] = a;
                brr[i + 1] = b;
            }
            if (jstack == 0) return;
            ir = istack[jstack];
            l = istack[jstack - 1];
            jstack -= 2;
        } else {
            int k = (l + ir) / 2;
            double t; int u;
#define PP_SWAP(x, y) do { t = arr[x]; arr[x] = arr[y]; arr[y] = t; u = brr[x]; brr[x] = brr[y]; brr[y] = u; } while (0)
            PP_SWAP(k, l + 1);
            if (arr[l + 1] > arr[ir]) PP_SWAP(l + 1, ir);
            if (arr[l] > arr[ir]) PP_SWAP(l, ir);
            if (arr[l + 1] > arr[l]) PP_SWAP(l + 1, l);
            int i = l + 1, j = ir;
            double a = arr[l];
            int b = brr[l];
            for (;;) {
                do { ++i; } while (arr[i] < a);
                do { --j; } while (arr[j] > a);
                if (j < i) break;
                PP_SWAP(i, j);
            }
#undef PP_SWAP
            arr[l] = arr[j]; arr[j] = a;
            brr[l] = brr[j]; brr[j] = b;
            jstack += 2;
            if (ir - i + 1 >= j - l) { istack[jstack] = ir; istack[jstack - 1] = i; ir = j - 1; }
            else { istack[jstack] = j - 1; istack[jstack - 1] = l; l = i; }
        }
    }
}

__device__ __forceinline__ void pp_ij(const PP &S, int ib, int &i, int &j)
{
    int jr = ib % S.M;
    if (jr != 0) { j = jr; i = (ib - jr) / S.M + 1; } else { j = S.M; i = ib / S.M; }
}

// ------------------------------------------------------------------ load + CSORT (PRE/csort.f90:32-62)
__global__ void k_pp_load(PP S, const double *qin, const unsigned char *present)
{
    for (int ib = blockIdx.x * blockDim.x + threadIdx.x + 1; ib < S.nb; ib += gridDim.x * blockDim.x) {
        bool p = present[ib - 1] != 0;
        double v = p ? qin[ib - 1] : -1.0;
        if (p && !(v > 0.0)) atomicExch(&S.cnt[7], 4);
        S.q[ib] = v;
        S.pres[ib] = p;
        S.rank[ib] = 0;
        S.stamp[ib] = 0;
    }
}

// ------------------------------------------------------------------ WBB boundary channel (PRE/wbb_sr.f90:95-160)
// rim cells (a neighbour outside the raster or outside the catchment) are lowered to quota_min*cqm, the last of them in
// the reference's scan (north row first, west to east) to quota_min*cqm*cqg: the outlet
__global__ void k_pp_gronda_mark(PP S)
{
    for (int ib = blockIdx.x * blockDim.x + threadIdx.x + 1; ib < S.nb; ib += gridDim.x * blockDim.x) {
        if (!S.pres[ib]) continue;
        atomicMin((unsigned long long *)&S.scal[1], (unsigned long long)__double_as_longlong(S.q[ib]));   // positive doubles order like their bits
        int i, j;
        pp_ij(S, ib, i, j);
        bool rim = false;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj) {
                int ii = i + di, jj = j + dj;
                if (ii < 1 || ii > S.N || jj < 1 || jj > S.M) { rim = true; continue; }
                if (!S.pres[ib + S.M * di + dj]) rim = true;
            }
        if (rim) {
            S.stamp[ib] = -1;
            atomicAdd(&S.cnt[9], 1);
            atomicMax(&S.cnt[10], (S.M - j) * S.N + (i - 1));
        }
    }
}

__global__ void k_pp_gronda_apply(PP S)
{
    const double quota_min = S.scal[1];
    const double qg = rmul(quota_min, (double)S.h.cqm0), qc = rmul(qg, (double)S.h.cqg0);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double rise = rmul(rmul(rmul((double)S.cnt[9], S.h.delta_x0), (double)sqrtf(2.0f)), S.h.pt0);
        if (radd(qg, rise) >= quota_min) atomicExch(&S.cnt[7], 5);                            // wbb_sr.f90:144-158
    }
    for (int ib = blockIdx.x * blockDim.x + threadIdx.x + 1; ib < S.nb; ib += gridDim.x * blockDim.x) {
        if (S.stamp[ib] != -1) continue;
        int i, j;
        pp_ij(S, ib, i, j);
        S.q[ib] = ((S.M - j) * S.N + (i - 1) == S.cnt[10]) ? qc : qg;
        S.stamp[ib] = 0;
    }
}

// ------------------------------------------------------------------ the reference's quicksort, in parallel, exactly
// QSORT (PRE/qsort.f90) is unstable and the permutation it leaves among equal elevations is part of the output, so the
// device must replay THAT algorithm, not merely sort.  Two facts make the replay parallel:
//  (1) Hoare's partition loop (qsort.f90:79-96) is a function of the INCOMING values: with pivot a, the k-th element
//      >= a from the left (u_k) is exchanged with the k-th element <= a from the right (v_k) for as long as u_k < v_k,
//      because each pointer only ever reads positions no exchange has touched yet, except the last one written by the
//      other pointer, which stops it.  So one partition = two flag scans, a pairing, and independent exchanges.
//  (2) Sorting a sub-array [l, ir] is self-contained -- the insertion scan of qsort.f90:29 runs down to index 1 but
//      stops at l-1, where every element is <= those of the sub-array -- so sub-arrays can be sorted in any order and
//      side by side without changing the permutation.
// k_pp_qsplit (one 1024-thread CTA) partitions every range longer than PP_CAP that way and lists the shorter ranges;
// k_pp_qsmall sorts each of those in shared memory, one CTA per range, replaying qsort.f90 literally (pp_qsort).
#define PP_CAP 12288
#define PP_QT 1024
#define PP_STK 4096
#define PP_DSORT 8                                   /* concurrent sorters inside k_pp_depit */
#define PP_DCAP ((PP_CAP + 1) / PP_DSORT - 1)        /* their range length: the staging buffer split PP_DSORT ways */

__device__ __forceinline__ int pp_block_exscan(int v, int *ws, int &total)
{
    // exclusive prefix sum over the PP_QT threads of the CTA; ws = 33 ints of shared memory
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();
    if (lane == 31) ws[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int w = ws[lane], z = w;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
        ws[lane] = z - w;
        if (lane == 31) ws[32] = z;
    }
    __syncthreads();
    total = ws[32];
    return ws[wid] + x - v;
}

// One CTA of PP_QT threads: partitions every range of arr[1..n] / brr[1..n] longer than cap, lists the others as tasks
// (l, len) and returns their number.  U, V: scratch of n ints each.
__device__ int pp_qsplit_block(double *arr, int *brr, int *U, int *V, int *task, int n, int cap, int *err)
{
    __shared__ int stk[PP_STK];
    __shared__ int ws[33];
    __shared__ int sh_l, sh_ir, sh_sp, sh_nt;
    __shared__ double sh_a;
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) {
        sh_sp = 0; sh_nt = 0;
        if (n > cap) { stk[0] = 1; stk[1] = n; sh_sp = 2; }
        else if (n >= 2) { task[0] = 1; task[1] = n; sh_nt = 1; }
    }
    __syncthreads();
    for (;;) {
        if (sh_sp == 0) break;
        __syncthreads();
        if (tid == 0) {
            sh_sp -= 2;
            const int l = stk[sh_sp], ir = stk[sh_sp + 1];
            // median of three (qsort.f90:49-76)
            const int k = (l + ir) / 2;
            double t; int u;
#define PP_SWAP(x, y) do { t = arr[x]; arr[x] = arr[y]; arr[y] = t; u = brr[x]; brr[x] = brr[y]; brr[y] = u; } while (0)
            PP_SWAP(k, l + 1);
            if (arr[l + 1] > arr[ir]) PP_SWAP(l + 1, ir);
            if (arr[l] > arr[ir]) PP_SWAP(l, ir);
            if (arr[l + 1] > arr[l]) PP_SWAP(l + 1, l);
#undef PP_SWAP
            sh_l = l; sh_ir = ir; sh_a = arr[l];
        }
        __syncthreads();
        const int l = sh_l, ir = sh_ir;
        const double a = sh_a;
        const int lo = l + 2, hi = ir - 1, cnt = hi - lo + 1;
        const int chunk = (cnt + PP_QT - 1) / PP_QT;
        const int p0 = min(lo + tid * chunk, hi + 1), p1 = min(p0 + chunk, hi + 1);
        int cL = 0, cR = 0;
        for (int p = p0; p < p1; ++p) { double v = arr[p]; cL += (v >= a); cR += (v <= a); }
        int nu, nv;
        const int exL = pp_block_exscan(cL, ws, nu);
        const int exR = pp_block_exscan(cR, ws, nv);
        int kL = exL;
        for (int p = p0; p < p1; ++p) if (arr[p] >= a) U[++kL] = p;
        int kR = nv - exR - cR;                                     // candidates <= a to the right of my chunk
        for (int p = p1 - 1; p >= p0; --p) if (arr[p] <= a) V[++kR] = p;
        __syncthreads();
        const int Kmin = min(nu, nv);
        int c = 0;
        for (int k = 1 + tid; k <= Kmin; k += PP_QT) c += (U[k] < V[k]);
        int K;
        pp_block_exscan(c, ws, K);
        for (int k = 1 + tid; k <= K; k += PP_QT) {
            const int x = U[k], y = V[k];
            double t = arr[x]; arr[x] = arr[y]; arr[y] = t;
            int u = brr[x]; brr[x] = brr[y]; brr[y] = u;
        }
        __syncthreads();
        if (tid == 0) {
            const int uK = K >= 1 ? U[K] : l + 1, vK = K >= 1 ? V[K] : ir;
            const int E = (K + 1 <= Kmin && U[K + 1] == V[K + 1]) ? 1 : 0;
            const int kk = K + 1 + E;
            const int i = min(kk <= nu ? U[kk] : 0x7fffffff, vK);
            const int j = max(kk <= nv ? V[kk] : -1, uK);
            arr[l] = arr[j]; arr[j] = a;                           // qsort.f90:97-100
            { int b = brr[l]; brr[l] = brr[j]; brr[j] = b; }
            const int rl[2] = {l, i}, rr[2] = {j - 1, ir};
            for (int h = 0; h < 2; ++h) {
                const int len = rr[h] - rl[h] + 1;
                if (len > cap) {
                    if (sh_sp + 2 > PP_STK) { atomicExch(err, 8); continue; }
                    stk[sh_sp] = rl[h]; stk[sh_sp + 1] = rr[h]; sh_sp += 2;
                } else if (len >= 2) { task[2 * sh_nt] = rl[h]; task[2 * sh_nt + 1] = len; ++sh_nt; }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    return sh_nt;
}

// keys S.key[1..n] / ids S.lst1[1..n]; scratch S.lst2, S.front[0]; tasks -> S.front[1], their number -> cnt[12]
__global__ void __launch_bounds__(PP_QT) k_pp_qsplit(PP S, int n)
{
    const int nt = pp_qsplit_block(S.key, S.lst1, S.lst2, S.front[0], S.front[1], n, PP_CAP, &S.cnt[7]);
    if (threadIdx.x == 0) S.cnt[12] = nt;
}

__global__ void __launch_bounds__(256) k_pp_qsmall(PP S)
{
    extern __shared__ double pp_sm[];
    double *sk = pp_sm;                                  // [PP_CAP + 1], 1-based
    int *si = (int *)(pp_sm + PP_CAP + 1);
    const int *task = S.front[1];
    const int nt = S.cnt[12];
    for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const int l = task[2 * t], len = task[2 * t + 1];
        for (int k = threadIdx.x; k < len; k += blockDim.x) { sk[1 + k] = S.key[l + k]; si[1 + k] = S.lst1[l + k]; }
        __syncthreads();
        if (threadIdx.x == 0) pp_qsort(len, sk, si);
        __syncthreads();
        for (int k = threadIdx.x; k < len; k += blockDim.x) { S.key[l + k] = sk[1 + k]; S.lst1[l + k] = si[1 + k]; }
        __syncthreads();
    }
}

// CSORT's records in ascending cell number (csort.f90:39-48): ordered compaction, one CTA
__global__ void __launch_bounds__(PP_QT) k_pp_records(PP S)
{
    __shared__ int ws[33];
    int n = 0;
    for (int base = 1; base < S.nb; base += PP_QT) {
        const int ib = base + threadIdx.x;
        const int p = (ib < S.nb && S.pres[ib]) ? 1 : 0;
        int tot;
        const int ex = pp_block_exscan(p, ws, tot);
        if (p) { S.key[n + ex + 1] = S.q[ib]; S.lst1[n + ex + 1] = ib; }
        n += tot;
    }
}

__global__ void k_pp_order(PP S)
{
    const int n = S.nc;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x + 1; l <= n; l += gridDim.x * blockDim.x) {      // descending (csort.f90:57-61)
        int ib = S.lst1[n - l + 1];
        S.order[l] = ib;
        S.rank[ib] = l;
    }
}

// ------------------------------------------------------------------ DEPIT
// parallel proof that DEPIT has nothing to do: no cell but the outlet lacks a strictly lower neighbour
__global__ void k_pp_pitcheck(PP S)
{
    int outlet = S.order[S.nc];
    if (blockIdx.x == 0 && threadIdx.x == 0 && S.nc > 1) {
        double ql = S.q[outlet], qsl = S.q[S.order[S.nc - 1]];
        if (rsub(qsl, ql) < rmul(EPS64, S.h.delta_x0)) atomicExch(&S.cnt[7], 2);       // depit.f90:56
    }
    for (int ib = blockIdx.x * blockDim.x + threadIdx.x + 1; ib < S.nb; ib += gridDim.x * blockDim.x) {
        if (!S.pres[ib] || ib == outlet) continue;
        int i, j;
        pp_ij(S, ib, i, j);
        double qc = S.q[ib];
        bool lower = false;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj) {
                int ii = i + di, jj = j + dj;
                if ((di == 0 && dj == 0) || ii < 1 || ii > S.N || jj < 1 || jj > S.M) continue;
                double qn = S.q[ib + S.M * di + dj];
                if (qn >= 0.0 && qn < qc) lower = true;
            }
        if (!lower) { S.dep[ib] = 1; S.front[0][atomicAdd(&S.cnt[5], 1)] = ib; }
    }
}

// the reference's sweeps (PRE/depit.f90:63-140) by ONE CTA.  The order of the visits is the reference's: the fixed point of
// this in-place Gauss-Seidel raising depends on it.  What is parallel is what does not change it:
//  * one visit: the eight neighbours are read by eight lanes of warp 0 at once ("some neighbour is lower" and "the lowest
//    neighbour" do not depend on the order they are read in); the list of the next sweep is appended in the reference's
//    neighbour order through a ballot;
//  * the first sweep visits ALL cells in ascending elevation, but a cell can only be raised if it is a pit on the incoming
//    elevations or if a neighbour was raised earlier in the same sweep -- every other visit is a no-op.  So it pops a
//    min-heap of list positions seeded with the pits k_pp_pitcheck found and fed with the later neighbours of every raised
//    cell: same visits that matter, same order;
//  * the re-sort of a long list between two sweeps is the parallel quicksort replay (pp_qsplit_block + staged pp_qsort).
__device__ __forceinline__ void pp_heap_push(int *h, int &n, int v)
{
    int c = n++;
    while (c > 0) { int p = (c - 1) >> 1; if (h[p] <= v) break; h[c] = h[p]; c = p; }
    h[c] = v;
}

__device__ __forceinline__ int pp_heap_pop(int *h, int &n)
{
    int top = h[0], v = h[--n], c = 0;
    for (;;) {
        int l = 2 * c + 1;
        if (l >= n) break;
        if (l + 1 < n && h[l + 1] < h[l]) ++l;
        if (h[l] >= v) break;
        h[c] = h[l]; c = l;
    }
    if (n > 0) h[c] = v;
    return top;
}

__global__ void __launch_bounds__(PP_QT) k_pp_depit(PP S)
{
    extern __shared__ double pp_sm[];
    double *sk = pp_sm;                                  // [PP_CAP + 1], staging of the short ranges
    int *si = (int *)(pp_sm + PP_CAP + 1);
    __shared__ int sh_cmd, sh_n;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nc = S.nc, N = S.N, M = S.M;
    const double eps = rmul(S.h.pt0, S.h.delta_x0);
    const int outlet = S.order[nc];
    int *heap = S.front[1];
    int hn = 0, n_pits = nc, total = 0, pass = 0, max_list = 0, cur = 0;     // cur: which of lst1 / lst2 holds the list being swept
    long long sort_cycles = 0, t_start = clock64();
    if (tid == 0) for (int t = 0; t < S.cnt[5]; ++t) pp_heap_push(heap, hn, nc - S.rank[S.front[0][t]] + 1);
    // neighbour of lane k < 8 in the reference's enumeration (ii outer, jj inner, the cell itself skipped)
    const int kk = lane < 4 ? lane : lane + 1, ndi = kk / 3 - 1, ndj = kk % 3 - 1;
    for (;;) {
        ++pass;
        int *pit1 = cur ? S.lst2 : S.lst1, *pit2 = cur ? S.lst1 : S.lst2;
        if (wid == 0) {
            int nn_mod = 0, nn_pit = 0, n = 0;
            for (;;) {
                int ib = -1, pos = 0;
                if (lane == 0) {
                    if (pass == 1) { if (hn > 0) { pos = pp_heap_pop(heap, hn); ib = S.order[nc - pos + 1]; } }
                    else if (++n <= n_pits) ib = pit1[n];
                }
                int nib = -1;                                    // the visit after this one, as far as it is known now
                if (lane == 0) {
                    if (pass == 1) { if (hn > 0) nib = S.order[nc - heap[0] + 1]; }
                    else if (n + 1 <= n_pits) nib = pit1[n + 1];
                }
                ib = __shfl_sync(0xffffffffu, ib, 0);
                pos = __shfl_sync(0xffffffffu, pos, 0);
                nib = __shfl_sync(0xffffffffu, nib, 0);
                if (ib < 0) break;
                if (nib > 0 && lane >= 8 && lane < 17) {         // idle lanes pull its 3x3 window towards the SM: a hint, not a read
                    int i2, j2;
                    pp_ij(S, nib, i2, j2);
                    const int k2 = lane - 8, d2i = k2 / 3 - 1, d2j = k2 % 3 - 1;
                    if (i2 + d2i >= 1 && i2 + d2i <= N && j2 + d2j >= 1 && j2 + d2j <= M)
                        { double sink; asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(sink) : "l"(&S.q[nib + M * d2i + d2j])); (void)sink; }
                }
                if (ib == outlet) continue;
                const double qc = S.q[ib];
                int i, j;
                pp_ij(S, ib, i, j);
                const int ii = i + ndi, jj = j + ndj;
                const bool valid = lane < 8 && ii >= 1 && ii <= N && jj >= 1 && jj <= M;
                const int nbr = valid ? ib + M * ndi + ndj : 0;
                const double qn = valid ? S.q[nbr] : -1.0;
                const bool hasq = valid && qn >= 0.0;
                if (__any_sync(0xffffffffu, hasq && qn < qc)) continue;
                double qmin = hasq ? qn : DBL_MAX;
                for (int o = 4; o >= 1; o >>= 1) qmin = fmin(qmin, __shfl_xor_sync(0xffffffffu, qmin, o));
                qmin = __shfl_sync(0xffffffffu, qmin, 0);
                if (qc <= qmin) {
                    if (lane == 0) S.q[ib] = radd(qmin, eps);
                    ++total; ++nn_mod;
                    const bool pr = hasq;                        // a catchment cell <=> a positive elevation (k_pp_load refuses the rest)
                    const int st = pr ? S.stamp[nbr] : pass, dp = (pr && pass == 1) ? S.dep[nbr] : 1, rk = (pr && pass == 1) ? S.rank[nbr] : 0;
                    const unsigned m = __ballot_sync(0xffffffffu, pr && st != pass);
                    if (pr && st != pass) { S.stamp[nbr] = pass; pit2[nn_pit + 1 + __popc(m & ((1u << lane) - 1))] = nbr; }
                    nn_pit += __popc(m);
                    if (pass == 1) {
                        const int pn = nc - rk + 1;
                        const bool push = pr && dp == 0 && pn > pos;
                        unsigned pm = __ballot_sync(0xffffffffu, push);
                        if (push) S.dep[nbr] = 1;
                        while (pm) {
                            const int src = __ffs(pm) - 1;
                            pm &= pm - 1;
                            const int v = __shfl_sync(0xffffffffu, pn, src);
                            if (lane == 0) pp_heap_push(heap, hn, v);
                        }
                    }
                }
                __syncwarp();
            }
            if (lane == 0) { sh_cmd = nn_mod > 0 ? 1 : 0; sh_n = nn_pit; }
        }
        __syncthreads();
        if (sh_cmd == 0) break;
        const int nn = sh_n;
        long long t0 = clock64();
        for (int k = 1 + tid; k <= nn; k += PP_QT) S.key[k] = S.q[pit2[k]];
        __syncthreads();
        if (nn <= 64) {
            if (tid == 0) pp_qsort(nn, S.key, pit2);
        } else {
            // scratch: the swept list's buffer (pit1), the pit list of k_pp_pitcheck (front[0], consumed) and -- after the first
            // sweep -- nothing else is live; the heap (front[1]) is empty whenever a sweep has ended
            int *task = S.front[1];
            const int nt = pp_qsplit_block(S.key, pit2, pit1, S.front[0], task, nn, PP_DCAP, &S.cnt[7]);
            // the short ranges: PP_DSORT warps at a time, each with its own slice of the staging buffer
            if (wid < PP_DSORT) {
                double *wk = sk + wid * (PP_DCAP + 1);
                int *wi = si + wid * (PP_DCAP + 1);
                for (int t = wid; t < nt; t += PP_DSORT) {
                    const int l = task[2 * t], len = task[2 * t + 1];
                    for (int k = lane; k < len; k += 32) { wk[1 + k] = S.key[l + k]; wi[1 + k] = pit2[l + k]; }
                    __syncwarp();
                    if (lane == 0) pp_qsort(len, wk, wi);
                    __syncwarp();
                    for (int k = lane; k < len; k += 32) { S.key[l + k] = wk[1 + k]; pit2[l + k] = wi[1 + k]; }
                    __syncwarp();
                }
            }
        }
        sort_cycles += clock64() - t0;
        __syncthreads();
        if (nn > max_list) max_list = nn;
        n_pits = nn;
        cur ^= 1;
    }
    if (tid == 0) {
        S.cnt[6] = total;
        S.cnt[8] = pass;
        S.cnt[13] = max_list;
        S.scal[2] = (double)sort_cycles;
        S.scal[3] = (double)(clock64() - t_start);
    }
}

// ------------------------------------------------------------------ FACET (PRE/facet.f90:10-47)
__device__ __forceinline__ void pp_facet(double e0, double e1, double e2, double dx, double dy, double &r, double &s)
{
    const double pi = 4.0 * atan(1.0);
    double s1 = rdiv(rsub(e0, e1), dx), s2 = rdiv(rsub(e1, e2), dx);
    if (fabs(s1) < EPS64) r = (s2 >= 0.0) ? pi / 2.0 : -pi / 2.0;
    else r = atan(rdiv(s2, s1));
    double sp = sqrt(radd(rmul(s1, s1), rmul(s2, s2)));
    double sd = rdiv(rsub(e0, e2), sqrt(radd(rmul(dx, dx), rmul(dy, dy))));
    if (r >= 0.0 && r <= pi / 4.0 && s1 >= 0.0) { s = sp; return; }
    if (s1 > sd) { s = s1; r = 0.0; } else { s = sd; r = pi / 4.0; }
}

__constant__ signed char PP_FA[8] = {2, 2, 6, 6, 8, 8, 4, 4};      // e1 of the eight facets in DSF's order (dsf.f90:103-245)
__constant__ signed char PP_FB[8] = {1, 3, 3, 9, 9, 7, 7, 1};      // e2
__constant__ signed char PP_DI[10] = {0, -1, -1, -1, 0, 0, 0, 1, 1, 1};   // p_outflow = 3*di + dj + 5 (dsf.f90:368-375)
__constant__ signed char PP_DJ[10] = {0, -1, 0, 1, -1, 0, 1, -1, 0, 1};

// window analysis of every cell: CCA (PRE/cca.f90:43-85), the facet search of SMEAN / DSF and the deviations (dsf.f90:247-268)
__global__ void k_pp_local(PP S)
{
    const double dx = S.h.delta_x, dy = S.h.delta_y;
    const double pi = 4.0 * atan(1.0);
    for (int ib = blockIdx.x * blockDim.x + threadIdx.x + 1; ib < S.nb; ib += gridDim.x * blockDim.x) {
        if (!S.pres[ib]) continue;
        int i, j;
        pp_ij(S, ib, i, j);
        double e[10];
        bool full = true;
        int l = 0;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj) {
                ++l;
                e[l] = 0.0;
                int ii = i + di, jj = j + dj;
                if (ii < 1 || ii > S.N || jj < 1 || jj > S.M) { full = false; continue; }
                double qn = S.q[ib + S.M * di + dj];
                if (qn < 0.0) { full = false; continue; }
                e[l] = qn;
            }
        // CCA: default method 2 (D8) on the rim of the catchment, curvature decides inside
        int dm = 2;
        float Kp = 0.f;
        if (full) {
            double dx2 = rmul(dx, dx);
            double zx = rdiv(rsub(e[6], e[4]), rmul(2.0, dx)), zy = rdiv(rsub(e[2], e[8]), rmul(2.0, dx));
            double zxx = rdiv(radd(rsub(e[6], rmul(2.0, e[5])), e[4]), dx2);
            double zyy = rdiv(radd(rsub(e[2], rmul(2.0, e[5])), e[8]), dx2);
            double zxy = rdiv(rsub(radd(radd(-e[1], e[3]), e[7]), e[9]), rmul(4.0, dx2));
            double p, qq;
            if (fabs(zx) > EPS64 || fabs(zy) > EPS64) { p = radd(rmul(zx, zx), rmul(zy, zy)); qq = radd(p, 1.0); }
            else { p = (double)1.0e-9f; qq = radd(p, 1.0); }
            double num = radd(rsub(rmul(rmul(zxx, zy), zy), rmul(rmul(rmul(2.0, zxy), zx), zy)), rmul(rmul(zyy, zx), zx));
            double Kc = rdiv(num, pow(p, 1.5));
            if (fabs(Kc) < EPS64) Kc = 0.0;
            double nump = radd(radd(rmul(rmul(zxx, zx), zx), rmul(rmul(rmul(2.0, zxy), zx), zy)), rmul(rmul(zyy, zy), zy));
            Kp = (float)rdiv(nump, rmul(p, pow(qq, 1.5)));
            dm = (Kc < (double)S.h.CC_threshold) ? 1 : 2;
        }
        S.dm0[ib] = dm;
        S.Kp[ib] = Kp;
        // steepest facet
        double e0 = e[5], smax = 0.0, rmax = 0.0, e1f = 0.0, e2f = 0.0, sigma = 1.0;
        int po1 = 0, po2 = 0;
        for (int f = 0; f < 8; ++f) {
            double e1 = e[PP_FA[f]], e2 = e[PP_FB[f]];
            if (e1 * e2 != 0.0) {
                double r, s;
                pp_facet(e0, e1, e2, dx, dy, r, s);
                if (s > smax) { e1f = e1; e2f = e2; rmax = r; smax = s; po1 = PP_FA[f]; po2 = PP_FB[f]; sigma = (f & 1) ? -1.0 : 1.0; }
            }
        }
        S.smax[ib] = smax;
        if (smax > 0.0) {
            double d1, d2;
            if (S.h.imethod == 1) { d1 = rmax; d2 = rsub(pi / 4, rmax); }
            else { d1 = rmul(dx, sin(rmax)); d2 = rmul(rmul(dx, sqrt(2.0)), sin(rsub(pi / 4.0, rmax))); }
            if (sigma == 1.0) d2 = -d2; else d1 = -d1;
            S.dev1[ib] = d1; S.dev2[ib] = d2;
            S.de1[ib] = rsub(e0, e1f); S.de2[ib] = rsub(e0, e2f);
            S.po1[ib] = (signed char)po1; S.po2[ib] = (signed char)po2;
        } else {
            // no facet with both corners inside the catchment: the lowest neighbour takes everything (dsf.f90:447-460)
            double emin = DBL_MAX;
            int pL = 0;
            for (int k = 1; k <= 9; ++k)
                if (e[k] != 0.0 && k != 5 && e[k] < emin) { emin = e[k]; pL = k; }
            S.dev1[ib] = 0.0; S.dev2[ib] = 0.0;
            S.de1[ib] = rsub(e0, emin); S.de2[ib] = rsub(e0, emin);
            S.po1[ib] = (signed char)pL; S.po2[ib] = (signed char)pL;
            if (pL == 0 && ib != S.order[S.nc]) atomicExch(&S.cnt[7], 6);
        }
    }
}

// SMEAN (PRE/smean.f90:33-164): the mean of the steepest slopes, summed in the reference's order by one warp
__global__ void k_pp_smean(PP S)
{
    const int lane = threadIdx.x;
    double sum = 0.0;
    const int outlet = S.order[S.nc];
    for (int base = 1; base <= S.nc; base += 32) {
        const int n = base + lane;
        const int ib = (n <= S.nc) ? S.order[n] : 0;
        const double v = (ib && ib != outlet) ? S.smax[ib] : 0.0;      // s_max >= 0: a skipped cell adds +0, which changes nothing
        double w[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) w[k] = __shfl_sync(0xffffffffu, v, k);
#pragma unroll
        for (int k = 0; k < 32; ++k) sum = radd(sum, w[k]);
    }
    if (lane == 0) S.scal[0] = rdiv(sum, (double)(S.nc - 1));
}

// ------------------------------------------------------------------ DSF as a dependency wavefront
__global__ void k_pp_deps(PP S)
{
    for (int ib = blockIdx.x * blockDim.x + threadIdx.x + 1; ib < S.nb; ib += gridDim.x * blockDim.x) {
        if (!S.pres[ib]) continue;
        int i, j;
        pp_ij(S, ib, i, j);
        int r = S.rank[ib], d = 0;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj) {
                int ii = i + di, jj = j + dj;
                if ((di == 0 && dj == 0) || ii < 1 || ii > S.N || jj < 1 || jj > S.M) continue;
                int nbr = ib + S.M * di + dj;
                if (S.pres[nbr] && S.rank[nbr] < r) ++d;
            }
        // a cell without a usable facet reads the channel flag the reference's loop carries over from the cell before
        // it in the order (local hcID of dsf.f90 is never reset): one more dependency
        if (S.smax[ib] == 0.0 && r > 1 && r < S.nc) ++d;
        S.dep[ib] = d;
        if (d == 0) S.front[0][atomicAdd(&S.cnt[0], 1)] = ib;
    }
}

__device__ __forceinline__ int pp_channel(const PP &S, double A_out, float ASk)
{
    if (S.h.nchc == 1) return (A_out <= S.h.A_threshold) ? 0 : 1;
    return (ASk <= S.h.ASk_threshold) ? 0 : 1;
}

__device__ void pp_cell(const PP &S, int ib)
{
    const int N = S.N, M = S.M;
    const double dx = S.h.delta_x;
    const double A_cell = rmul(dx, dx);
    const int r = S.rank[ib];
    int i, j;
    pp_ij(S, ib, i, j);
    // donors, in the order the reference visits them
    int dn[8], dr[8], dk[8], nd = 0;
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
            int ii = i + di, jj = j + dj;
            if ((di == 0 && dj == 0) || ii < 1 || ii > N || jj < 1 || jj > M) continue;
            int nbr = ib + M * di + dj;
            if (!S.pres[nbr]) continue;
            int rn = S.rank[nbr];
            if (rn > r) continue;
            int p = 3 * (-di) + (-dj) + 5;      // direction from the neighbour to this cell
            int which = (S.p1[nbr] == p) ? 1 : (S.p2[nbr] == p) ? 2 : 0;
            if (!which) continue;
            int k = nd++;
            while (k > 0 && dr[k - 1] > rn) { dr[k] = dr[k - 1]; dn[k] = dn[k - 1]; dk[k] = dk[k - 1]; --k; }
            dr[k] = rn; dn[k] = nbr; dk[k] = which;
        }
    double A = 0.0, Sd = 0.0;
    int hcin = 0;
    for (int k = 0; k < nd; ++k) {
        int n = dn[k];
        double w = (double)(dk[k] == 1 ? S.w1[n] : S.w2[n]);
        double aw = rmul(S.Aout[n], w);
        A = radd(A, aw);
        if (S.sflag[n] & dk[k]) Sd = radd(Sd, rmul(aw, dk[k] == 1 ? S.sd1[n] : S.sd2[n]));
        if (hcin == 0 && fabsf((float)w) > EPS32) hcin = S.hc[n];
    }
    S.Ain[ib] = A;
    S.sdn[ib] = Sd;
    if (r == S.nc) { S.hc[ib] = hcin; S.dm[ib] = S.dm0[ib]; return; }              // outlet: features assigned apart (dsf.f90:81)
    const double A_out = radd(A, A_cell);
    S.Aout[ib] = A_out;
    double sumdev = (A == 0.0) ? 0.0 : rdiv(Sd, A);
    const double smax = S.smax[ib], lam = S.h.lambda;
    if (smax > 0.0) {
        const double d1 = S.dev1[ib], d2 = S.dev2[ib];
        if (fabs(d1) <= EPS64 || fabs(d2) <= EPS64) sumdev = 0.0;
        int dm = S.dm0[ib], hc = hcin;
        float epl1 = (float)dx, epl2 = (float)rmul(sqrt(2.0), dx);
        float ls1 = (float)rdiv(S.de1[ib], (double)epl1), ls2 = (float)rdiv(S.de2[ib], (double)epl2);
        float ASk = (float)rmul(A_out, pow(smax, (double)S.h.kas));
        double sd1 = radd(rmul(lam, sumdev), d1), sd2 = radd(rmul(lam, sumdev), d2);
        if (hc == 0) hc = pp_channel(S, A_out, ASk);
        if (hc == 1 && S.h.ndcf == 1) dm = 2;
        double a1 = fabs(sd1), a2 = fabs(sd2);
        float w1, w2;
        if (dm == 1) {
            if (fabs(d1) <= EPS64) { w1 = 1.f; w2 = 0.f; }
            else if (fabs(d2) <= EPS64) { w1 = 0.f; w2 = 1.f; }
            else {
                w1 = (float)rdiv(a2, radd(a1, a2));
                w2 = (float)rdiv(a1, radd(a1, a2));
                if (w1 < 1.0e-6f) { w1 = 0.f; w2 = 1.f; }
                if (w2 < 1.0e-6f) { w1 = 1.f; w2 = 0.f; }
            }
        } else {
            if (rdiv(fabs(rsub(a1, a2)), dx) < 10e-14 && S.de1[ib] > 0.0) { w1 = 1.f; w2 = 0.f; epl2 = 0.f; ls2 = 0.f; }
            else if (a1 < a2 && S.de1[ib] > 0.0) { w1 = 1.f; w2 = 0.f; epl2 = 0.f; ls2 = 0.f; }
            else if (a1 > a2 || S.de2[ib] > 0.0) { w1 = 0.f; w2 = 1.f; epl1 = 0.f; ls1 = 0.f; }
            else { w1 = 0.f; w2 = 0.f; atomicExch(&S.cnt[7], 7); }
        }
        S.w1[ib] = w1; S.w2[ib] = w2; S.epl1[ib] = epl1; S.epl2[ib] = epl2; S.ls1[ib] = ls1; S.ls2[ib] = ls2;
        S.sd1[ib] = sd1; S.sd2[ib] = sd2; S.sflag[ib] = 3;
        S.dm[ib] = dm; S.hc[ib] = hc; S.hc_after[ib] = hc; S.ASk[ib] = ASk;
        __threadfence();
        S.p1[ib] = S.po1[ib]; S.p2[ib] = S.po2[ib];
    } else {
        int hc = (r > 1) ? S.hc_after[S.order[r - 1]] : 0;        // carried over, see k_pp_deps
        const int pL = S.po1[ib];
        S.dm[ib] = S.dm0[ib];
        if ((pL & 1) == 0) {
            float epl1 = (float)dx, ls1 = (float)rdiv(S.de1[ib], (double)epl1);
            float ASk = (float)rmul(A_out, pow((double)ls1, (double)S.h.kas));
            if (hc == 0) hc = pp_channel(S, A_out, ASk);
            S.w1[ib] = 1.f; S.epl1[ib] = epl1; S.ls1[ib] = ls1; S.sflag[ib] = 0; S.sd1[ib] = rmul(lam, sumdev);
            S.ASk[ib] = ASk; S.hc[ib] = hc; S.hc_after[ib] = hc;
            __threadfence();
            S.p1[ib] = pL;
        } else {
            float epl2 = (float)rmul(sqrt(2.0), dx), ls2 = (float)rdiv(S.de2[ib], (double)epl2);
            float ASk = (float)rmul(A_out, pow((double)ls2, (double)S.h.kas));
            S.w2[ib] = 1.f; S.epl2[ib] = epl2; S.ls2[ib] = ls2; S.sflag[ib] = 2; S.sd2[ib] = rmul(lam, sumdev);
            S.ASk[ib] = ASk; S.hc[ib] = hc; S.hc_after[ib] = hc;
            __threadfence();
            S.p2[ib] = pL;
        }
    }
}

__global__ void __launch_bounds__(256) k_pp_sweep(PP S)
{
    cg::grid_group grid = cg::this_grid();
    // cell t of a wavefront goes to CTA t mod gridDim.x: a wavefront holds a few hundred cells, each a chain of dependent loads,
    // so they are spread over all SMs instead of filling the first CTA
    const int gtid = blockIdx.x + gridDim.x * threadIdx.x, gsz = gridDim.x * blockDim.x;
    int wave = 0;
    for (;;) {
        const int ncur = *((volatile int *)&S.cnt[wave % 3]);
        if (ncur == 0) break;
        int *cur = S.front[wave & 1], *nxt = S.front[(wave + 1) & 1];
        int *nnext = &S.cnt[(wave + 1) % 3];
        if (gtid == 0) { S.cnt[(wave + 2) % 3] = 0; S.cnt[3] += ncur; S.cnt[4] = wave + 1; }
        for (int t = gtid; t < ncur; t += gsz) {
            const int ib = cur[t];
            pp_cell(S, ib);
            __threadfence();
            const int r = S.rank[ib];
            int i, j;
            pp_ij(S, ib, i, j);
            for (int di = -1; di <= 1; ++di)
                for (int dj = -1; dj <= 1; ++dj) {
                    int ii = i + di, jj = j + dj;
                    if ((di == 0 && dj == 0) || ii < 1 || ii > S.N || jj < 1 || jj > S.M) continue;
                    int nbr = ib + S.M * di + dj;
                    if (S.pres[nbr] && S.rank[nbr] > r && atomicSub(&S.dep[nbr], 1) == 1) nxt[atomicAdd(nnext, 1)] = nbr;
                }
            if (r + 1 < S.nc) {
                int nx = S.order[r + 1];
                if (S.smax[nx] == 0.0 && atomicSub(&S.dep[nx], 1) == 1) nxt[atomicAdd(nnext, 1)] = nx;
            }
        }
        grid.sync();
        ++wave;
    }
}

// phantom channel end of the outlet cell (PRE/dsf.f90:531-600)
__global__ void k_pp_outlet(PP S)
{
    if (blockIdx.x || threadIdx.x) return;
    const int N = S.N, M = S.M, ib = S.order[S.nc];
    const double dx = S.h.delta_x, A_cell = rmul(dx, dx);
    int i, j;
    pp_ij(S, ib, i, j);
    int nvo = 0, p_out = 0;
    double A_max = 0.0;
    float ls_out = 0.f;
    for (int ii = i - 1; ii <= i + 1; ++ii)
        for (int jj = j - 1; jj <= j + 1; ++jj) {
            if (ii < 1 || ii > N || jj < 1 || jj > M) continue;
            int nbr = (ii - 1) * M + jj;
            if (!S.pres[nbr]) continue;
            int p_in = 3 * (ii - i) + (jj - j) + 5;
            for (int k = 1; k <= 2; ++k) {
                int pk = (k == 1) ? S.p1[nbr] : S.p2[nbr];
                if (p_in + pk != 10) continue;
                double Ao = rmul(radd(S.Ain[nbr], A_cell), (double)((k == 1) ? S.w1[nbr] : S.w2[nbr]));
                if (Ao > A_max) {
                    A_max = Ao;
                    p_out = pk;
                    ls_out = (k == 1) ? S.ls1[nbr] : S.ls2[nbr];
                    int ivo = i + PP_DI[p_out], jvo = j + PP_DJ[p_out];
                    nvo = (ivo < 1 || ivo > N || jvo < 1 || jvo > M || !S.pres[(ivo - 1) * M + jvo]) ? 1 : 0;
                }
            }
        }
    if (nvo == 0) p_out = S.h.p_outflow_vo;
    if ((p_out & 1) == 0) { S.p1[ib] = p_out; S.w1[ib] = 1.f; S.epl1[ib] = (float)dx; S.ls1[ib] = ls_out; }
    else { S.p2[ib] = p_out; S.w2[ib] = 1.f; S.epl2[ib] = (float)rmul(sqrt(2.0), dx); S.ls2[ib] = ls_out; }
}

// ------------------------------------------------------------------ HG (PRE/hg.f90:39-113)
__device__ __forceinline__ float pp_law(float c, float Q, float ex1, float ex2, float wexp, double RA, float w)
{
    // c * Q**(-ex1) in single precision, (RA*w)**(wexp*(ex2-ex1)) in double, product stored in single precision
    float lead = __fmul_rn(c, (float)pow((double)Q, (double)(-ex1)));
    float ex = __fmul_rn(wexp, __fsub_rn(ex2, ex1));
    return (float)rmul((double)lead, pow(rmul(RA, (double)w), (double)ex));
}

__global__ void k_pp_hg(PP S)
{
    const CathyPreproParams &h = S.h;
    const double A_cell = rmul(h.delta_x, h.delta_y);
    for (int ib = blockIdx.x * blockDim.x + threadIdx.x + 1; ib < S.nb; ib += gridDim.x * blockDim.x) {
        if (!S.pres[ib]) continue;
        const double A_out = radd(S.Ain[ib], A_cell);
        const float w1 = S.w1[ib], w2 = S.w2[ib];
        const bool rill = S.hc[ib] == 0;
        const double RA = rdiv(A_out, rill ? h.As_rf : h.As_cf);
        const float Q = rill ? h.Qsf_rf : h.Qsf_cf, we = rill ? h.w_rf : h.w_cf, Wsf = rill ? h.Wsf_rf : h.Wsf_cf;
        const float b1 = rill ? h.b1_rf : h.b1_cf, b2 = rill ? h.b2_rf : h.b2_cf;
        const float kS = rill ? h.kSsf_rf : h.kSsf_cf, y1 = rill ? h.y1_rf : h.y1_cf, y2 = rill ? h.y2_rf : h.y2_cf;
        S.b1[ib] = b1;
        S.y1[ib] = y1;
        S.Ws1[ib] = (fabsf(w1) > EPS32) ? pp_law(Wsf, Q, b1, b2, we, RA, w1) : 0.f;
        S.Ws2[ib] = (fabsf(w2) > EPS32) ? pp_law(Wsf, Q, b1, b2, we, RA, w2) : 0.f;
        S.kS1[ib] = (fabsf(w1) > EPS32) ? pp_law(kS, Q, y1, y2, we, RA, w1) : 0.f;
        S.kS2[ib] = (fabsf(w2) > EPS32) ? pp_law(kS, Q, y1, y2, we, RA, w2) : 0.f;
        S.nrc[ib] = rill ? (float)rdiv(h.delta_x, h.dr) : 1.f;
    }
}

// ------------------------------------------------------------------ host driver
template <class T>
static cudaError_t pp_alloc(T **p, size_t n, bool zero = true)
{
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess && zero) e = cudaMemset(*p, 0, n * sizeof(T));
    return e;
}

// CSORT on the device (PRE/csort.f90:32-62)
static cudaError_t pp_csort(PP &S, int grid, int tpb, int nsm, int &launches)
{
    k_pp_records<<<1, PP_QT>>>(S); ++launches;
    k_pp_qsplit<<<1, PP_QT>>>(S, S.nc); ++launches;
    k_pp_qsmall<<<nsm, 256, (PP_CAP + 1) * (sizeof(double) + sizeof(int))>>>(S); ++launches;
    k_pp_order<<<grid, tpb>>>(S); ++launches;
    return cudaGetLastError();
}

extern "C" const char *cathy_prepro_last_error(void) { return pp_err; }

extern "C" int32_t cathy_prepro_run(const CathyPreproParams *p, const double *quota_in, const uint8_t *present, int32_t device,
                                    CathyPreproOut *out)
{
    int32_t rc = 0;
    PP S;
    memset(&S, 0, sizeof S);
    void *all[64];
    int nall = 0;
    double *d_qin = nullptr;
    unsigned char *d_present = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evs[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int cnt[16];
    int launches = 0;
    pp_err[0] = 0;
    if (!p || !quota_in || !present || !out) PFAIL(-1, "cathy_prepro_run: null argument");
    if (p->abi_version != CATHY_PREPRO_ABI_VERSION) PFAIL(-1, "cathy_prepro_run: ABI version %d, library has %d", p->abi_version, CATHY_PREPRO_ABI_VERSION);
    if (p->N < 1 || p->M < 1 || (long long)p->N * p->M > 2000000000LL) PFAIL(-1, "cathy_prepro_run: bad raster size %d x %d", p->N, p->M);
    if (p->N < 2 || p->M < 2) PFAIL(-1, "a raster one cell wide has no facet: the reference then carries an uninitialised channel flag from cell to cell (PRE/dsf.f90:64,470), its result is undefined");
    if (p->imethod != 1 && p->imethod != 2) PFAIL(-1, "unespected imethod!");
    if (p->nchc == 3) PFAIL(-1, "channel initiation by normalised divergence (nchc = 3) is not built: the reference reads an uninitialised curvature on the rim cells (PRE/cca.f90:24,87)");
    if (p->nchc != 1 && p->nchc != 2) PFAIL(-1, "nchc out of range!");
    if (!(p->delta_x > 0.0) || !(p->delta_x0 > 0.0)) PFAIL(-1, "cathy_prepro_run: grid spacing must be positive");
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) PFAIL(-100, "cathy_prepro_run: no CUDA device (there is no CPU path)");
    }
    PCK(cudaSetDevice(device));
    PCK(cudaFuncSetAttribute(k_pp_depit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((PP_CAP + 1) * (sizeof(double) + sizeof(int)))));
    PCK(cudaFuncSetAttribute(k_pp_qsmall, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((PP_CAP + 1) * (sizeof(double) + sizeof(int)))));
    {
        S.h = *p;
        S.N = p->N; S.M = p->M; S.nb = p->N * p->M + 1;
        long long nc = 0;
        for (long long k = 0; k < (long long)p->N * p->M; ++k) nc += present[k] != 0;
        if (nc < 1) PFAIL(-1, "cathy_prepro_run: no catchment cell");
        S.nc = (int)nc;
        const size_t nb = (size_t)S.nb, n1 = (size_t)S.nc + 1;
#define PA(field, n) do { PCK(pp_alloc(&S.field, (n))); all[nall++] = (void *)S.field; } while (0)
        PA(q, nb); PA(pres, nb); PA(order, n1); PA(rank, nb); PA(key, n1); PA(lst1, n1); PA(lst2, n1); PA(stamp, nb);
        PA(smax, nb); PA(dev1, nb); PA(dev2, nb); PA(de1, nb); PA(de2, nb); PA(po1, nb); PA(po2, nb); PA(Kp, nb); PA(dm0, nb);
        PA(Ain, nb); PA(sdn, nb); PA(Aout, nb); PA(sd1, nb); PA(sd2, nb);
        PA(w1, nb); PA(w2, nb); PA(ls1, nb); PA(ls2, nb); PA(epl1, nb); PA(epl2, nb); PA(ASk, nb);
        PA(p1, nb); PA(p2, nb); PA(hc, nb); PA(dm, nb); PA(hc_after, nb); PA(sflag, nb);
        PA(dep, nb); PA(front[0], n1); PA(front[1], n1); PA(cnt, 16); PA(scal, 4);
        PA(Ws1, nb); PA(Ws2, nb); PA(b1, nb); PA(kS1, nb); PA(kS2, nb); PA(y1, nb); PA(nrc, nb);
#undef PA
        PCK(pp_alloc(&d_qin, nb - 1, false));
        PCK(pp_alloc(&d_present, nb - 1, false));
        PCK(cudaMemcpy(d_qin, quota_in, (nb - 1) * sizeof(double), cudaMemcpyHostToDevice));
        PCK(cudaMemcpy(d_present, present, nb - 1, cudaMemcpyHostToDevice));
        PCK(cudaEventCreate(&ev0));
        PCK(cudaEventCreate(&ev1));
        for (int k = 0; k < 8; ++k) PCK(cudaEventCreate(&evs[k]));
        PCK(cudaEventRecord(ev0, 0));
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
        const int TPB = 256;
        const int GRID = nsm * 8;
        k_pp_load<<<GRID, TPB>>>(S, d_qin, d_present); ++launches;
        if (p->bcc != 0) {
            const double everest = (double)8844.43f;                            // wbb_sr.f90:185, a single-precision literal
            PCK(cudaMemcpy(&S.scal[1], &everest, sizeof(double), cudaMemcpyHostToDevice));
            k_pp_gronda_mark<<<GRID, TPB>>>(S); ++launches;
            k_pp_gronda_apply<<<GRID, TPB>>>(S); ++launches;
        }
        PCK(cudaEventRecord(evs[0], 0));
        PCK(pp_csort(S, GRID, TPB, nsm, launches));
        PCK(cudaEventRecord(evs[1], 0));
        k_pp_pitcheck<<<GRID, TPB>>>(S); ++launches;
        PCK(cudaMemcpy(cnt, S.cnt, sizeof cnt, cudaMemcpyDeviceToHost));
        if (cnt[7] == 4) PFAIL(-4, "non-positive elevation inside the catchment: the reference uses 0 and negative values as 'no cell' marks");
        if (cnt[7] == 5) PFAIL(-5, "boundary channel: after the depitting the highest boundary channel cell would be above the lowest dem cell; a smaller coefficient for boundary channel elevation definition is needed");
        if (cnt[7] == 2) PFAIL(-2, "catchment with more than one outlet cell!");
        PCK(cudaEventRecord(evs[2], 0));
        if (cnt[5] > 0) {
            k_pp_depit<<<1, PP_QT, (PP_CAP + 1) * (sizeof(double) + sizeof(int))>>>(S); ++launches;
            PCK(cudaEventRecord(evs[3], 0));
            PCK(pp_csort(S, GRID, TPB, nsm, launches));
        } else PCK(cudaEventRecord(evs[3], 0));
        PCK(cudaEventRecord(evs[4], 0));
        k_pp_local<<<GRID, TPB>>>(S); ++launches;
        k_pp_smean<<<1, 32>>>(S); ++launches;
        PCK(cudaEventRecord(evs[5], 0));
        k_pp_deps<<<GRID, TPB>>>(S); ++launches;
        {
            int per_sm = 0;
            PCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pp_sweep, TPB, 0));
            if (per_sm < 1) PFAIL(-100, "k_pp_sweep does not fit on an SM");
            if (per_sm > 1) per_sm = 1;      // a wavefront holds a few hundred cells: one CTA per SM keeps the grid barrier short
            void *args[] = {(void *)&S};
            PCK(cudaLaunchCooperativeKernel((void *)k_pp_sweep, dim3(nsm * per_sm), dim3(TPB), args, 0, 0)); ++launches;
        }
        PCK(cudaEventRecord(evs[6], 0));
        k_pp_outlet<<<1, 32>>>(S); ++launches;
        k_pp_hg<<<GRID, TPB>>>(S); ++launches;
        PCK(cudaGetLastError());
        PCK(cudaEventRecord(ev1, 0));
        PCK(cudaDeviceSynchronize());
        PCK(cudaMemcpy(cnt, S.cnt, sizeof cnt, cudaMemcpyDeviceToHost));
        if (cnt[7] == 8) PFAIL(-1, "quicksort range stack overflow (raster too large for PP_STK)");
        if (cnt[7] == 6) PFAIL(-1, "s_max = 0, unexpected case! (a cell without any neighbour)");
        if (cnt[7] == 7) PFAIL(-1, "s_max < 0, unexpected case!");
        if (cnt[3] != S.nc) PFAIL(-1, "drainage sweep finished %d of %d cells (dependency cycle?)", cnt[3], S.nc);
        float ms = 0.f;
        PCK(cudaEventElapsedTime(&ms, ev0, ev1));
        const size_t ncell = nb - 1;
#define PD(dst, src, T) do { if (out->dst) PCK(cudaMemcpy(out->dst, S.src + 1, ncell * sizeof(T), cudaMemcpyDeviceToHost)); } while (0)
        PD(quota, q, double); PD(A_inflow, Ain, double);
        PD(w_1, w1, float); PD(w_2, w2, float); PD(local_slope_1, ls1, float); PD(local_slope_2, ls2, float);
        PD(epl_1, epl1, float); PD(epl_2, epl2, float);
        PD(Ws1_sf_1, Ws1, float); PD(Ws1_sf_2, Ws2, float); PD(b1_sf, b1, float); PD(kSs1_sf_1, kS1, float); PD(kSs1_sf_2, kS2, float);
        PD(y1_sf, y1, float); PD(nrc, nrc, float);
        PD(p_outflow_1, p1, int32_t); PD(p_outflow_2, p2, int32_t); PD(hcID, hc, int32_t); PD(dmID, dm, int32_t);
#undef PD
        if (out->order) PCK(cudaMemcpy(out->order, S.order + 1, (size_t)S.nc * sizeof(int32_t), cudaMemcpyDeviceToHost));
        double scal[4];
        PCK(cudaMemcpy(scal, S.scal, sizeof scal, cudaMemcpyDeviceToHost));
        out->n_cells = S.nc;
        out->n_modifications = cnt[6];
        out->n_waves = cnt[4];
        out->n_launches = launches;
        out->mean_s_max = scal[0];
        if (getenv("CATHY_PREPRO_DEBUG"))
            fprintf(stderr, "[prepro] DEPIT: %d sweeps, %d raises, longest list %d, sort share of the kernel %.2f\n", cnt[8], cnt[6], cnt[13],
                    scal[3] > 0 ? scal[2] / scal[3] : 0.0);
        out->device_ms = ms;
        {
            // csort, pit check, depit, second csort, window analysis + smean, drainage sweep, outlet + hg
            const int a[7] = {0, 1, 2, 3, 4, 5, 6};
            for (int k = 0; k < 7; ++k) {
                float t = 0.f;
                PCK(cudaEventElapsedTime(&t, evs[a[k]], k == 6 ? ev1 : evs[a[k] + 1]));
                out->stage_ms[k] = t;
            }
            out->stage_ms[7] = cnt[8];
        }
    }
done:
    for (int k = 0; k < nall; ++k) cudaFree(all[k]);
    if (d_qin) cudaFree(d_qin);
    if (d_present) cudaFree(d_present);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (int k = 0; k < 8; ++k) if (evs[k]) cudaEventDestroy(evs[k]);
    return rc;
}

// ------------------------------------------------------------------ host side: the fixed-width text of RBB (PRE/mrbb_sr.f90:445-470)
// A raster file of a million cells is 21 MB of Fortran-formatted numbers; formatting is exact decimal conversion
// (snprintf), done row-parallel on host threads.  Every row is ncols fields of width w plus a newline, so rows are
// written at fixed offsets.
static bool pp_fmt_e(double x, int w, int d, char *dst)
{
    char buf[64], body[64];
    if (!(x == x) || x - x != 0.0) return false;
    int len;
    if (x == 0.0) {
        len = snprintf(body, sizeof body, "0.%0*dE+00", d, 0);
    } else {
        int n = snprintf(buf, sizeof buf, "%.*E", d - 1, x < 0 ? -x : x);      // D.DDDDE+XX
        char *e = strchr(buf, 'E');
        if (!e || n <= 0) return false;
        int ex = atoi(e + 1) + 1;
        if (ex > 99 || ex < -99) return false;
        char *o = body;
        if (x < 0) *o++ = '-';
        *o++ = '0'; *o++ = '.';
        *o++ = buf[0];
        for (char *c = buf + 2; c < e; ++c) *o++ = *c;
        o += snprintf(o, 8, "E%c%02d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
        len = (int)(o - body);
    }
    char *b = body;
    if (len > w) {                                    // gfortran drops the optional leading zero before giving up
        if (b[0] == '0') { ++b; --len; }
        else if (b[0] == '-' && b[1] == '0') { b[1] = '-'; ++b; --len; }
    }
    if (len > w) { memset(dst, '*', w); return true; }
    memset(dst, ' ', w - len);
    memcpy(dst + (w - len), b, len);
    return true;
}

static bool pp_fmt_f(double x, int w, int d, char *dst)
{
    char body[400];
    if (!(x == x) || x - x != 0.0) return false;
    int len = snprintf(body, sizeof body, "%.*f", d, x);
    if (len <= 0 || len >= (int)sizeof body) return false;
    char *b = body;
    if (len > w) {
        if (b[0] == '0' && b[1] == '.') { ++b; --len; }
        else if (b[0] == '-' && b[1] == '0' && b[2] == '.') { b[1] = '-'; ++b; --len; }
    }
    if (len > w) { memset(dst, '*', w); return true; }
    memset(dst, ' ', w - len);
    memcpy(dst + (w - len), b, len);
    return true;
}

template <class F>
static int64_t pp_rows_parallel(int64_t nrows, int32_t nthreads, F row)
{
    int hw = (int)std::thread::hardware_concurrency();
    int nt = nthreads > 0 ? nthreads : (hw > 32 ? 32 : (hw < 1 ? 1 : hw));
    if ((int64_t)nt > nrows) nt = (int)(nrows < 1 ? 1 : nrows);
    std::atomic<int> bad(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            for (int64_t r = t; r < nrows; r += nt) if (!row(r)) bad = 1;
        });
    for (auto &x : th) x.join();
    return bad ? -1 : 0;
}

// kind 0: Ew.d, 1: Fw.d.  out holds nrows * (ncols * w + 1) bytes.  Returns the byte count, -1 if a value needs a
// form this routine does not write (NaN, infinity, three-digit exponent): the caller formats that file itself.
extern "C" int64_t cathy_prepro_format_real(const double *v, int64_t nrows, int64_t ncols, int32_t w, int32_t d, int32_t kind,
                                            char *out, int32_t nthreads)
{
    if (!v || !out || nrows < 0 || ncols < 0 || w < 1 || w > 60 || d < 0 || d > 40) return -1;
    const int64_t stride = ncols * w + 1;
    int64_t rc = pp_rows_parallel(nrows, nthreads, [&](int64_t r) {
        char *o = out + r * stride;
        for (int64_t c = 0; c < ncols; ++c)
            if (!(kind == 0 ? pp_fmt_e(v[r * ncols + c], w, d, o + c * w) : pp_fmt_f(v[r * ncols + c], w, d, o + c * w))) return false;
        o[ncols * w] = '\n';
        return true;
    });
    return rc < 0 ? -1 : nrows * stride;
}

extern "C" int64_t cathy_prepro_format_int(const int32_t *v, int64_t nrows, int64_t ncols, int32_t w, char *out, int32_t nthreads)
{
    if (!v || !out || nrows < 0 || ncols < 0 || w < 1 || w > 30) return -1;
    const int64_t stride = ncols * w + 1;
    pp_rows_parallel(nrows, nthreads, [&](int64_t r) {
        char *o = out + r * stride, body[32];
        for (int64_t c = 0; c < ncols; ++c) {
            int len = snprintf(body, sizeof body, "%d", v[r * ncols + c]);
            if (len > w) memset(o + c * w, '*', w);
            else { memset(o + c * w, ' ', w - len); memcpy(o + c * w + (w - len), body, len); }
        }
        o[ncols * w] = '\n';
        return true;
    });
    return nrows * stride;
}
