// cathy_enkf.cu -- dense EnKF analysis update on B200 (fp64 tensor cores, DMMA m8n8k4).
//
// Reference arithmetic: pyCATHY/DA/enkf.py:82-205 (enkf_analysis) and :282-324
// (enkf_analysis_localized_with_inflation):
//     S = HX - mean_j(HX)            D = y - HX
//     C = S S^T/(Ne-1) + R^T         B = C^-1 D          (Sakov: B = D / diag(R))
//     P = X' S^T/(Ne-1)  (o L)       Xa = X + P B        (+ inflation about the analysis mean)
// Layout: X is [n][ne] row-major, i.e. state index major / member minor, the same orientation the reference
// uses (ensemble.shape = (N_state, N_ens)); members are contiguous so ensemble reductions are coalesced.
// Multi-GPU: members (columns) are sharded over ranks; the caller all-reduces the row sums and the partial
// cross-covariance P (N x m) between the stages below (NCCL through torch.distributed, see da.py).
// All entry points take DEVICE pointers except cathy_enkf_gain, whose operands are tiny (m x Ne).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/cathy_b200.h"

static thread_local char e_err[512] = "";
extern "C" const char *cathy_enkf_last_error(void) { return e_err; }
#define EFAIL(code, ...) do { snprintf(e_err, sizeof e_err, __VA_ARGS__); return (code); } while (0)
#define ECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) EFAIL(-100, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{   // D(8x8) += A(8x4, row) * B(4x8, col)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// D(16x8) += A(16x16, row) * B(16x8, col): the largest fp64 tensor-core shape (PTX ISA 8.0, sm_90+).  Fragments (g = lane / 4, t = lane % 4):
// a[i]: row g + 8 (i % 2), column t + 4 (i / 2);  b[v]: row (k) t + 4 v, column g;  d[0..3]: (g, 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---- stage 1 (tiny): S, D, C, B ------------------------------------------------------------
__global__ void k_enkf_obs(int m, int ne, const double *__restrict__ hx, const double *__restrict__ y, int y_ld,
                           double *__restrict__ S, double *__restrict__ D)
{   // one block per observation row: obs_avg = sum/Ne (enkf.py:139-142), S = HX - avg, D = y - HX
    __shared__ double sh[32];
    int i = blockIdx.x;
    double acc = 0.0;
    for (int j = threadIdx.x; j < ne; j += blockDim.x) acc += hx[(size_t)i * ne + j];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w]; sh[0] = (1.0 / ne) * t; }
    __syncthreads();
    double avg = sh[0];
    for (int j = threadIdx.x; j < ne; j += blockDim.x) {
        double h = hx[(size_t)i * ne + j];
        S[(size_t)i * ne + j] = h - avg;
        D[(size_t)i * ne + j] = (y_ld ? y[(size_t)i * y_ld + j] : y[i]) - h;
    }
}
__global__ void k_enkf_cov(int m, int ne, const double *__restrict__ S, const double *__restrict__ R, double *__restrict__ C)
{   // C = S S^T/(Ne-1) + R^T (enkf.py:166)
    int i = blockIdx.x, j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double acc = 0.0;
    for (int k = 0; k < ne; ++k) acc += S[(size_t)i * ne + k] * S[(size_t)j * ne + k];
    C[(size_t)i * m + j] = (1.0 / (ne - 1)) * acc + R[(size_t)j * m + i];
}
// B = C^-1 D by Gaussian elimination with partial pivoting (what numpy.linalg.solve's LAPACK gesv does), one CTA.
__global__ void k_enkf_solve(int m, int ne, double *__restrict__ C, double *__restrict__ D, int *__restrict__ info)
{
    __shared__ double s_val[32];
    __shared__ int s_idx[32], s_piv;
    for (int c = 0; c < m; ++c) {
        double best = -1.0; int bi = c;
        for (int r = c + threadIdx.x; r < m; r += blockDim.x) { double v = fabs(C[(size_t)r * m + c]); if (v > best) { best = v; bi = r; } }
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, best, o); int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
            s_piv = bi;
            if (best == 0.0) *info = c + 1;
        }
        __syncthreads();
        int p = s_piv;
        if (p != c) {
            for (int k = threadIdx.x; k < m; k += blockDim.x) { double t = C[(size_t)c * m + k]; C[(size_t)c * m + k] = C[(size_t)p * m + k]; C[(size_t)p * m + k] = t; }
            for (int k = threadIdx.x; k < ne; k += blockDim.x) { double t = D[(size_t)c * ne + k]; D[(size_t)c * ne + k] = D[(size_t)p * ne + k]; D[(size_t)p * ne + k] = t; }
        }
        __syncthreads();
        double piv = C[(size_t)c * m + c];
        // eliminate below: rows r > c, columns of C (k > c) and of D, flattened over threads
        int ncol = (m - c - 1) + ne;
        long long work = (long long)(m - c - 1) * ncol;
        for (long long w = threadIdx.x; w < work; w += blockDim.x) {
            int r = c + 1 + (int)(w / ncol), k = (int)(w % ncol);
            double f = C[(size_t)r * m + c] / piv;
            if (k < m - c - 1) C[(size_t)r * m + c + 1 + k] -= f * C[(size_t)c * m + c + 1 + k];
            else D[(size_t)r * ne + (k - (m - c - 1))] -= f * D[(size_t)c * ne + (k - (m - c - 1))];
        }
        __syncthreads();
    }
    // back substitution, one column of D per thread
    for (int k = threadIdx.x; k < ne; k += blockDim.x)
        for (int r = m - 1; r >= 0; --r) {
            double v = D[(size_t)r * ne + k];
            for (int q = r + 1; q < m; ++q) v -= C[(size_t)r * m + q] * D[(size_t)q * ne + k];
            D[(size_t)r * ne + k] = v / C[(size_t)r * m + r];
        }
}
__global__ void k_enkf_sakov(int m, int ne, const double *__restrict__ R, double *__restrict__ D)
{   // B = D / diag(R^T) (enkf.py:161-164)
    int i = blockIdx.x;
    double d = R[(size_t)i * m + i];
    for (int j = threadIdx.x; j < ne; j += blockDim.x) D[(size_t)i * ne + j] = D[(size_t)i * ne + j] / d;
}

// ---- stage 2: row sums over the local members ------------------------------------------------
__global__ void k_enkf_rowsum(long long n, int ne, const double *__restrict__ X, double *__restrict__ rowsum)
{   // one warp per state row; members are contiguous -> coalesced
    long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; row < n; row += nw) {
        double acc = 0.0;
        for (int j = lane; j < ne; j += 32) acc += X[row * ne + j];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) rowsum[row] = acc;
    }
}

// ---- stage 3: partial cross covariance P = (X - mean) S^T / (Ne-1), fp64 tensor cores ----------
// CTA = 4 warps, tile = 32 state rows (8 per warp) x all m observations; K (members) in chunks of 32.
#define EK 32
#define EPAD 36   // padded leading dimension of the shared tiles (avoids bank conflicts of the fragment loads)
template <int MT>   // MT = number of 8-wide observation tiles held per warp (m <= 8*MT)
__global__ void __launch_bounds__(128) k_enkf_crosscov(long long n, int ne, int m, const double *__restrict__ X, const double *__restrict__ mean,
                                                       const double *__restrict__ S, double scale, double *__restrict__ P)
{
    extern __shared__ double smem[];
    double *Xs = smem;                 // [32][EPAD]
    double *Ss = smem + 32 * EPAD;     // [8*MT][EPAD]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (long long row0 = (long long)blockIdx.x * 32; row0 < n; row0 += (long long)gridDim.x * 32) {
        double acc[MT][2];
#pragma unroll
        for (int q = 0; q < MT; ++q) acc[q][0] = acc[q][1] = 0.0;
        for (int k0 = 0; k0 < ne; k0 += EK) {
            __syncthreads();
            for (int e = threadIdx.x; e < 32 * EK; e += 128) {
                int r = e / EK, c = e % EK;
                long long row = row0 + r;
                double v = 0.0;
                if (row < n && k0 + c < ne) v = X[row * ne + k0 + c] - mean[row];
                Xs[r * EPAD + c] = v;
            }
            for (int e = threadIdx.x; e < 8 * MT * EK; e += 128) {
                int r = e / EK, c = e % EK;
                Ss[r * EPAD + c] = (r < m && k0 + c < ne) ? S[(size_t)r * ne + k0 + c] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < EK; kk += 4) {
                double a = Xs[(warp * 8 + g) * EPAD + kk + t];
#pragma unroll
                for (int q = 0; q < MT; ++q) {
                    double b = Ss[(q * 8 + g) * EPAD + kk + t];     // B(k=t, n=g) = S[n][k]
                    dmma884(acc[q][0], acc[q][1], a, b);
                }
            }
        }
        long long row = row0 + warp * 8 + g;
        if (row < n)
#pragma unroll
            for (int q = 0; q < MT; ++q) {
                int c = q * 8 + t * 2;
                if (c < m) P[row * m + c] = scale * acc[q][0];
                if (c + 1 < m) P[row * m + c + 1] = scale * acc[q][1];
            }
    }
}

// ---- stage 4: Xa = X + (P o L) B, optional inflation about the analysis mean ---------------------
// CTA = 4 warps, tile = 32 state rows x 64 members per pass; K = m observations in chunks of 32.
// X and Xa may alias (in-place update): every element is read and written by the same thread, so neither is __restrict__.
__global__ void __launch_bounds__(128) k_enkf_update(long long n, int ne, int m, const double *X, const double *__restrict__ P,
                                                     const double *__restrict__ L, long long n_loc, const double *__restrict__ B,
                                                     const double *__restrict__ mean, const double *__restrict__ bbar, double inflate,
                                                     long long n_infl, double inflate2, double *Xa)
{
    extern __shared__ double smem[];
    double *Ps = smem;                 // [32][EPAD]   (rows x k)
    double *Bs = smem + 32 * EPAD;     // [64][EPAD]   (member x k)  -> B(k, n) fragments read as Bs[n][k]
    __shared__ double s_ma[32];        // analysis mean of the tile rows (for inflation)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bool infl_any = inflate != 1.0 || inflate2 != 1.0;
    for (long long row0 = (long long)blockIdx.x * 32; row0 < n; row0 += (long long)gridDim.x * 32) {
        if (infl_any) {   // mean_a = mean + (P o L) bbar, one thread per tile row
            __syncthreads();
            if (threadIdx.x < 32) {
                long long row = row0 + threadIdx.x;
                double v = 0.0;
                if (row < n) {
                    v = mean[row];
                    for (int k = 0; k < m; ++k) { double p = P[row * m + k]; if (L && row < n_loc) p *= L[row * m + k]; v += p * bbar[k]; }
                }
                s_ma[threadIdx.x] = v;
            }
        }
        for (int c0 = 0; c0 < ne; c0 += 64) {
            double acc[8][2];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = 0.0;
            for (int k0 = 0; k0 < m; k0 += EK) {
                __syncthreads();
                for (int e = threadIdx.x; e < 32 * EK; e += 128) {
                    int r = e / EK, c = e % EK;
                    long long row = row0 + r;
                    double v = 0.0;
                    if (row < n && k0 + c < m) { v = P[row * m + k0 + c]; if (L && row < n_loc) v *= L[row * m + k0 + c]; }
                    Ps[r * EPAD + c] = v;
                }
                for (int e = threadIdx.x; e < 64 * EK; e += 128) {
                    int c = e / 64, r = e % 64;          // r = member within the pass (contiguous in B rows), c = k
                    Bs[r * EPAD + c] = (k0 + c < m && c0 + r < ne) ? B[(size_t)(k0 + c) * ne + c0 + r] : 0.0;
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < EK; kk += 4) {
                    double a = Ps[(warp * 8 + g) * EPAD + kk + t];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        double b = Bs[(q * 8 + g) * EPAD + kk + t];
                        dmma884(acc[q][0], acc[q][1], a, b);
                    }
                }
            }
            long long row = row0 + warp * 8 + g;
            if (row < n) {
                const double fac = row < n_infl ? inflate : inflate2;
                const double ma = infl_any ? s_ma[warp * 8 + g] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    int c = c0 + q * 8 + t * 2;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (c + h < ne) {
                            double v = X[row * ne + c + h] + acc[q][h];
                            if (fac != 1.0) v = ma + fac * (v - ma);
                            Xa[row * ne + c + h] = v;
                        }
                }
            }
        }
    }
}


// ---- the same two stages with the SMALL operand resident in shared memory (default whenever it fits: m x Ne x 8 B <= ~200 KB) -------
// k_enkf_crosscov / k_enkf_update above restage the m x Ne obs-perturbation matrix S (resp. the gain columns B) for every 32-row
// tile of the state: at N = 163,216, Ne = 256, m = 64 that is 2 x the traffic of X itself through shared memory, with two CTA barriers
// per 32 members -- 0.15 of the pair's roof (profiles/r2a_ncu_full_summary.json).  Here a persistent 256-thread CTA stages S (B) ONCE,
// padded so that the DMMA B-fragment loads of a half-warp hit 16 different banks, and then streams the state: the A fragments
// (X - mean, resp. P o L) go from global memory straight into registers (each lane one double; a warp-wide load covers 8 rows x 32
// contiguous bytes = whole sectors), so the main loop has no barrier at all: 1 LDG + MT x (LDS + DMMA) per 4 members.
template <int MT>
__global__ void __launch_bounds__(512) k_enkf_crosscov_res(long long n, int ne, int m, int ld, const double *__restrict__ X, const double *__restrict__ mean,
                                                           const double *__restrict__ S, double scale, double *__restrict__ P)
{
    extern __shared__ double Ss[];      // [8*MT][ld], zero beyond (m, ne); ld >= ne rounded up to 16
    for (int e = threadIdx.x; e < 8 * MT * ld; e += blockDim.x) {
        const int r = e / ld, c = e - r * ld;
        Ss[e] = (r < m && c < ne) ? S[(size_t)r * ne + c] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, ne16 = (ne + 15) & ~15;
    for (long long row0 = (long long)blockIdx.x * (blockDim.x >> 1); row0 < n; row0 += (long long)gridDim.x * (blockDim.x >> 1)) {
        const long long ra = row0 + warp * 16 + g, rb = ra + 8;
        const bool va = ra < n, vb = rb < n;
        const double mua = va ? mean[ra] : 0.0, mub = vb ? mean[rb] : 0.0;
        const double *xa = X + (va ? ra : 0) * ne, *xb = X + (vb ? rb : 0) * ne;
        double acc[MT][4];
#pragma unroll
        for (int q = 0; q < MT; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
#pragma unroll 2
        for (int k0 = 0; k0 < ne16; k0 += 16) {
            double a[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = k0 + t + 4 * j;
                a[2 * j] = (va && c < ne) ? xa[c] - mua : 0.0;
                a[2 * j + 1] = (vb && c < ne) ? xb[c] - mub : 0.0;
            }
            const double *sb = Ss + (size_t)g * ld + k0 + t;
#pragma unroll
            for (int q = 0; q < MT; ++q) {
                const double *sq = sb + (size_t)q * 8 * ld;
                const double b[4] = {sq[0], sq[4], sq[8], sq[12]};
                dmma16816(acc[q], a, b);
            }
        }
#pragma unroll
        for (int q = 0; q < MT; ++q) {
            const int c = q * 8 + t * 2;
            if (va) { if (c < m) P[ra * m + c] = scale * acc[q][0]; if (c + 1 < m) P[ra * m + c + 1] = scale * acc[q][1]; }
            if (vb) { if (c < m) P[rb * m + c] = scale * acc[q][2]; if (c + 1 < m) P[rb * m + c + 1] = scale * acc[q][3]; }
        }
    }
}
// X and Xa may alias: every element is read and written by the same thread
__global__ void __launch_bounds__(512) k_enkf_update_res(long long n, int ne, int m, int ld, const double *X, const double *__restrict__ P,
                                                         const double *__restrict__ L, long long n_loc, const double *__restrict__ B,
                                                         const double *__restrict__ mean, const double *__restrict__ bbar, double inflate,
                                                         long long n_infl, double inflate2, double *Xa)
{
    extern __shared__ double Bs[];      // [m16][ld]: B(k, member), zero beyond (m, ne)
    const int m16 = (m + 15) & ~15;
    for (int e = threadIdx.x; e < m16 * ld; e += blockDim.x) {
        const int k = e / ld, c = e - k * ld;
        Bs[e] = (k < m && c < ne) ? B[(size_t)k * ne + c] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bool infl_any = inflate != 1.0 || inflate2 != 1.0;
    for (long long row0 = (long long)blockIdx.x * (blockDim.x >> 1); row0 < n; row0 += (long long)gridDim.x * (blockDim.x >> 1)) {
        const long long rr[2] = {row0 + warp * 16 + g, row0 + warp * 16 + g + 8};
        bool valid[2], loc[2];
        const double *pr[2], *lr[2];
        double ma[2] = {0.0, 0.0}, fac[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            valid[h] = rr[h] < n; loc[h] = L != nullptr && rr[h] < n_loc;
            pr[h] = P + (valid[h] ? rr[h] : 0) * m; lr[h] = L + (loc[h] ? rr[h] : 0) * m;
            fac[h] = rr[h] < n_infl ? inflate : inflate2;
            if (infl_any) {   // mean_a = mean + (P o L) bbar, summed over the observations in order by one lane of the row's group
                if (valid[h] && t == 0) {
                    double v = mean[rr[h]];
                    for (int k = 0; k < m; ++k) { double p = pr[h][k]; if (loc[h]) p *= lr[h][k]; v += p * bbar[k]; }
                    ma[h] = v;
                }
                ma[h] = __shfl_sync(0xffffffffu, ma[h], lane & ~3);
            }
        }
        for (int c0 = 0; c0 < ne; c0 += 64) {
            double acc[8][4];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
            for (int k0 = 0; k0 < m16; k0 += 16) {
                double a[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k0 + t + 4 * j;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double v = 0.0;
                        if (valid[h] && k < m) { v = pr[h][k]; if (loc[h]) v *= lr[h][k]; }
                        a[2 * j + h] = v;
                    }
                }
                const double *bb = Bs + (size_t)(k0 + t) * ld + c0 + g;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double b[4] = {bb[q * 8], bb[q * 8 + 4 * (size_t)ld], bb[q * 8 + 8 * (size_t)ld], bb[q * 8 + 12 * (size_t)ld]};
                    dmma16816(acc[q], a, b);      // columns beyond ne are zero padding (ld >= c0 + 64)
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (valid[h])
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int c = c0 + q * 8 + t * 2;
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                            if (c + e < ne) {
                                double v = X[rr[h] * ne + c + e] + acc[q][2 * h + e];
                                if (fac[h] != 1.0) v = ma[h] + fac[h] * (v - ma[h]);
                                Xa[rr[h] * ne + c + e] = v;
                            }
                    }
        }
    }
}

// The same with the m8n8k4 shape (what the hardware executes: SASS DMMA.8x8x4), 8 rows per warp and 1024 threads per CTA: the loop is
// bound by the latency of the LDS -> DMMA chains, so more, thinner warps win (tools/bench_dmma.cu: 37 TFLOP/s needs 32 warps per SM)
template <int MT>
__global__ void __launch_bounds__(1024) k_enkf_crosscov_r8(long long n, int ne, int m, int ld, const double *__restrict__ X, const double *__restrict__ mean,
                                                          const double *__restrict__ S, double scale, double *__restrict__ P)
{
    extern __shared__ double Ss[];      // [8*MT][ld]
    for (int e = threadIdx.x; e < 8 * MT * ld; e += blockDim.x) {
        const int r = e / ld, c = e - r * ld;
        Ss[e] = (r < m && c < ne) ? S[(size_t)r * ne + c] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, ne4 = (ne + 3) & ~3, rows_cta = blockDim.x >> 2;
    for (long long row0 = (long long)blockIdx.x * rows_cta; row0 < n; row0 += (long long)gridDim.x * rows_cta) {
        const long long row = row0 + warp * 8 + g;
        const bool valid = row < n;
        const double mu = valid ? mean[row] : 0.0;
        const double *xr = X + (valid ? row : 0) * ne;
        double acc[MT][2];
#pragma unroll
        for (int q = 0; q < MT; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll 4
        for (int k0 = 0; k0 < ne4; k0 += 4) {
            const double a = (valid && k0 + t < ne) ? xr[k0 + t] - mu : 0.0;
            const double *sb = Ss + (size_t)g * ld + k0 + t;
#pragma unroll
            for (int q = 0; q < MT; ++q) dmma884(acc[q][0], acc[q][1], a, sb[(size_t)q * 8 * ld]);
        }
        if (valid)
#pragma unroll
            for (int q = 0; q < MT; ++q) {
                const int c = q * 8 + t * 2;
                if (c < m) P[row * m + c] = scale * acc[q][0];
                if (c + 1 < m) P[row * m + c + 1] = scale * acc[q][1];
            }
    }
}
__global__ void __launch_bounds__(1024) k_enkf_update_r8(long long n, int ne, int m, int ld, const double *X, const double *__restrict__ P,
                                                        const double *__restrict__ L, long long n_loc, const double *__restrict__ B,
                                                        const double *__restrict__ mean, const double *__restrict__ bbar, double inflate,
                                                        long long n_infl, double inflate2, double *Xa)
{
    extern __shared__ double Bs[];      // [m16][ld]: B(k, member)
    const int m16 = (m + 15) & ~15, m4 = (m + 3) & ~3;
    for (int e = threadIdx.x; e < m16 * ld; e += blockDim.x) {
        const int k = e / ld, c = e - k * ld;
        Bs[e] = (k < m && c < ne) ? B[(size_t)k * ne + c] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, rows_cta = blockDim.x >> 2;
    const bool infl_any = inflate != 1.0 || inflate2 != 1.0;
    for (long long row0 = (long long)blockIdx.x * rows_cta; row0 < n; row0 += (long long)gridDim.x * rows_cta) {
        const long long row = row0 + warp * 8 + g;
        const bool valid = row < n, loc = L != nullptr && row < n_loc;
        const double *pr = P + (valid ? row : 0) * m, *lr = L + (loc ? row : 0) * m;
        double ma = 0.0;
        if (infl_any) {
            if (valid && t == 0) {
                double v = mean[row];
                for (int k = 0; k < m; ++k) { double p = pr[k]; if (loc) p *= lr[k]; v += p * bbar[k]; }
                ma = v;
            }
            ma = __shfl_sync(0xffffffffu, ma, lane & ~3);
        }
        const double fac = row < n_infl ? inflate : inflate2;
        for (int c0 = 0; c0 < ne; c0 += 64) {
            double acc[8][2];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll 2
            for (int k0 = 0; k0 < m4; k0 += 4) {
                double a = 0.0;
                if (valid && k0 + t < m) { a = pr[k0 + t]; if (loc) a *= lr[k0 + t]; }
                const double *bb = Bs + (size_t)(k0 + t) * ld + c0 + g;
#pragma unroll
                for (int q = 0; q < 8; ++q) dmma884(acc[q][0], acc[q][1], a, bb[q * 8]);
            }
            if (valid)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int c = c0 + q * 8 + t * 2;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (c + h < ne) {
                            double v = X[row * ne + c + h] + acc[q][h];
                            if (fac != 1.0) v = ma + fac * (v - ma);
                            Xa[row * ne + c + h] = v;
                        }
                }
        }
    }
}

__global__ void k_enkf_scale(long long n, double a, double *__restrict__ v)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] *= a;
}
// Particle weights (pf.py:60-82): log w_i = -0.5 sum_k ((y_k - HX_ki)/sigma_k)^2, shifted by the max, normalised;
// w[ne] receives n_eff = 1/sum w^2.  One CTA, sums in member order like the reference's loops.
__global__ void k_pf_weights(int m, int ne, const double *__restrict__ hx, const double *__restrict__ y, const double *__restrict__ sd,
                             double *__restrict__ w)
{
    __shared__ double s_red[256];
    double mx = -1.0e300;
    for (int i = threadIdx.x; i < ne; i += blockDim.x) {
        double acc = 0.0;
        for (int k = 0; k < m; ++k) { double d = (y[k] - hx[(size_t)k * ne + i]) / sd[k]; acc += d * d; }
        double lw = -0.5 * acc;
        w[i] = lw;
        mx = fmax(mx, lw);
    }
    s_red[threadIdx.x] = mx;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) { if ((int)threadIdx.x < o) s_red[threadIdx.x] = fmax(s_red[threadIdx.x], s_red[threadIdx.x + o]); __syncthreads(); }
    mx = s_red[0];
    __syncthreads();
    for (int i = threadIdx.x; i < ne; i += blockDim.x) w[i] = exp(w[i] - mx);
    __syncthreads();
    if (threadIdx.x == 0) {   // ne is a few hundred: sequential sums keep the reference's summation order
        double s = 0.0;
        for (int i = 0; i < ne; ++i) s += w[i];
        double s2 = 0.0;
        for (int i = 0; i < ne; ++i) { double v = w[i] / s; w[i] = v; s2 += v * v; }
        w[ne] = 1.0 / s2;
    }
}
__global__ void k_pf_gather(long long n, int ne, const double *__restrict__ X, const int *__restrict__ idx, double *__restrict__ Xo)
{
    long long tot = n * ne;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        long long r = e / ne; int j = (int)(e - r * ne);
        Xo[e] = X[r * ne + idx[j]];
    }
}

// Gaspari-Cohn localisation matrix (pyCATHY/DA/localisation.py:136-188): L[i][k] = gc(|grid_i - obs_k| / radius), 2-D distances
__device__ __forceinline__ double gaspari_cohn(double d, double radius)
{
    double r = fabs(d) / radius;
    if (r <= 1.0) return (((-0.25 * r + 0.5) * r + 0.625) * r - 5.0 / 3.0) * (r * r) + 1.0;
    if (r <= 2.0) return ((((r / 12.0 - 0.5) * r + 0.625) * r + 5.0 / 3.0) * r - 5.0) * r + 4.0 - 2.0 / (3.0 * r);
    return 0.0;
}
__global__ void k_enkf_localization(long long n, int m, const double *__restrict__ gxy, const double *__restrict__ oxy, double radius,
                                    double *__restrict__ L)
{
    long long tot = n * m;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        long long i = e / m; int k = (int)(e - i * m);
        double dx = oxy[2 * k] - gxy[2 * i], dy = oxy[2 * k + 1] - gxy[2 * i + 1];
        L[e] = gaspari_cohn(sqrt(dx * dx + dy * dy), radius);
    }
}

extern "C" {

// S, D=B on output.  hx [m][ne], y [m] (y_ld = 0) or [m][ne] (y_ld = ne), R [m][m]; host pointers.
int32_t cathy_enkf_gain(const double *hx, const double *y, int32_t y_is_matrix, const double *R, int32_t m, int32_t ne,
                        int32_t sakov, double *S_out, double *B_out)
{
    if (m < 1 || ne < 2) EFAIL(-1, "cathy_enkf_gain: need m >= 1 and ne >= 2");
    double *d_hx = nullptr, *d_y = nullptr, *d_R = nullptr, *d_S = nullptr, *d_D = nullptr, *d_C = nullptr;
    int *d_info = nullptr;
    size_t bmn = (size_t)m * ne * sizeof(double), bmm = (size_t)m * m * sizeof(double), by = y_is_matrix ? bmn : (size_t)m * sizeof(double);
    ECK(cudaMalloc(&d_hx, bmn)); ECK(cudaMalloc(&d_y, by)); ECK(cudaMalloc(&d_R, bmm)); ECK(cudaMalloc(&d_S, bmn));
    ECK(cudaMalloc(&d_D, bmn)); ECK(cudaMalloc(&d_C, bmm)); ECK(cudaMalloc(&d_info, sizeof(int)));
    ECK(cudaMemcpy(d_hx, hx, bmn, cudaMemcpyHostToDevice)); ECK(cudaMemcpy(d_y, y, by, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(d_R, R, bmm, cudaMemcpyHostToDevice)); ECK(cudaMemset(d_info, 0, sizeof(int)));
    k_enkf_obs<<<m, 128>>>(m, ne, d_hx, d_y, y_is_matrix ? ne : 0, d_S, d_D);
    if (sakov) k_enkf_sakov<<<m, 128>>>(m, ne, d_R, d_D);
    else {
        k_enkf_cov<<<dim3(m, (m + 127) / 128), 128>>>(m, ne, d_S, d_R, d_C);
        k_enkf_solve<<<1, 256>>>(m, ne, d_C, d_D, d_info);
    }
    int info = 0;
    ECK(cudaMemcpy(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost));
    if (S_out) ECK(cudaMemcpy(S_out, d_S, bmn, cudaMemcpyDeviceToHost));
    if (B_out) ECK(cudaMemcpy(B_out, d_D, bmn, cudaMemcpyDeviceToHost));
    cudaFree(d_hx); cudaFree(d_y); cudaFree(d_R); cudaFree(d_S); cudaFree(d_D); cudaFree(d_C); cudaFree(d_info);
    if (info) EFAIL(-3, "cathy_enkf_gain: singular observation covariance (pivot %d)", info);
    return 0;
}

// DEVICE pointers below; `stream` is a cudaStream_t passed as an integer (0 = default stream).
int32_t cathy_enkf_rowsum(const double *dX, int64_t n, int32_t ne, double *d_rowsum, uint64_t stream)
{
    int blocks = (int)std::min<int64_t>((n * 32 + 255) / 256, 148 * 16);
    k_enkf_rowsum<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(n, ne, dX, d_rowsum);
    ECK(cudaGetLastError());
    return 0;
}
int32_t cathy_enkf_crosscov(const double *dX, const double *d_mean, const double *dS_local, int64_t n, int32_t ne_local, int32_t m,
                            int32_t ne_total, double *dP, uint64_t stream)
{
    if (m > 256) EFAIL(-2, "cathy_enkf_crosscov: m = %d observations > 256 not supported by this build", m);
    int mt = (m + 7) / 8;
    size_t sm = (size_t)(32 + 8 * (mt <= 8 ? 8 : mt <= 16 ? 16 : 32)) * EPAD * sizeof(double);
    int blocks = (int)std::min<int64_t>((n + 31) / 32, 148 * 8);
    double scale = 1.0 / (ne_total - 1);
    cudaStream_t st = (cudaStream_t)stream;
    {   // S resident in shared memory (k_enkf_crosscov_res) when it fits; leading dimension = 4 mod 16 words: conflict-free B fragments
        const int ne16 = (ne_local + 15) & ~15, ld = ne16 + 4, mtt = mt <= 8 ? 8 : mt <= 16 ? 16 : 32;
        const size_t smr = (size_t)8 * mtt * ld * sizeof(double);
        if (smr <= 200 * 1024 && !getenv("CATHY_ENKF_TILED")) {
            const int blk = getenv("CATHY_ENKF_BLOCK") ? atoi(getenv("CATHY_ENKF_BLOCK")) : 1024, rows_cta = blk > 512 ? blk / 4 : blk / 2;
            const int gridr = (int)std::max<int64_t>(1, std::min<int64_t>((n + rows_cta - 1) / rows_cta, 148));
            if (blk > 512 && mtt <= 16) {      // 1024 threads x 8 rows per warp (m8n8k4): <= 64 registers per thread needs <= 16 observation tiles
                if (mtt == 8) { ECK(cudaFuncSetAttribute(k_enkf_crosscov_r8<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
                                k_enkf_crosscov_r8<8><<<gridr, 1024, smr, st>>>(n, ne_local, m, ld, dX, d_mean, dS_local, scale, dP); }
                else { ECK(cudaFuncSetAttribute(k_enkf_crosscov_r8<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
                       k_enkf_crosscov_r8<16><<<gridr, 1024, smr, st>>>(n, ne_local, m, ld, dX, d_mean, dS_local, scale, dP); }
                ECK(cudaGetLastError());
                return 0;
            }
            if (mtt == 8) { ECK(cudaFuncSetAttribute(k_enkf_crosscov_res<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
                            k_enkf_crosscov_res<8><<<gridr, blk, smr, st>>>(n, ne_local, m, ld, dX, d_mean, dS_local, scale, dP); }
            else if (mtt == 16) { ECK(cudaFuncSetAttribute(k_enkf_crosscov_res<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
                                  k_enkf_crosscov_res<16><<<gridr, blk, smr, st>>>(n, ne_local, m, ld, dX, d_mean, dS_local, scale, dP); }
            else { ECK(cudaFuncSetAttribute(k_enkf_crosscov_res<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
                   k_enkf_crosscov_res<32><<<gridr, blk, smr, st>>>(n, ne_local, m, ld, dX, d_mean, dS_local, scale, dP); }
            ECK(cudaGetLastError());
            return 0;
        }
    }
    if (mt <= 8) k_enkf_crosscov<8><<<blocks, 128, sm, st>>>(n, ne_local, m, dX, d_mean, dS_local, scale, dP);
    else if (mt <= 16) k_enkf_crosscov<16><<<blocks, 128, sm, st>>>(n, ne_local, m, dX, d_mean, dS_local, scale, dP);
    else {
        ECK(cudaFuncSetAttribute(k_enkf_crosscov<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_enkf_crosscov<32><<<blocks, 128, sm, st>>>(n, ne_local, m, dX, d_mean, dS_local, scale, dP);
    }
    ECK(cudaGetLastError());
    return 0;
}
int32_t cathy_enkf_update(const double *dX, const double *dP, const double *dL, int64_t n_loc, const double *dB_local, const double *d_mean,
                          const double *d_bbar, double inflate, int64_t n_infl, double inflate2, int64_t n, int32_t ne_local, int32_t m, double *dXa,
                          uint64_t stream)
{
    if ((inflate != 1.0 || inflate2 != 1.0) && (!d_mean || !d_bbar)) EFAIL(-1, "cathy_enkf_update: inflation needs the ensemble mean and the mean gain column");
    {   // B resident in shared memory (k_enkf_update_res) when it fits; columns padded to whole 64-member passes
        const int ne64 = (ne_local + 63) & ~63, ld = ne64 + 4, m16 = (m + 15) & ~15;
        const size_t smr = (size_t)m16 * ld * sizeof(double);
        if (smr <= 200 * 1024 && !getenv("CATHY_ENKF_TILED")) {
            const int blk = getenv("CATHY_ENKF_BLOCK") ? atoi(getenv("CATHY_ENKF_BLOCK")) : 1024, rows_cta = blk > 512 ? blk / 4 : blk / 2;
            const int gridr = (int)std::max<int64_t>(1, std::min<int64_t>((n + rows_cta - 1) / rows_cta, 148));
            if (blk > 512) {
                ECK(cudaFuncSetAttribute(k_enkf_update_r8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
                k_enkf_update_r8<<<gridr, 1024, smr, (cudaStream_t)stream>>>(n, ne_local, m, ld, dX, dP, dL, n_loc, dB_local, d_mean, d_bbar, inflate, n_infl, inflate2, dXa);
                ECK(cudaGetLastError());
                return 0;
            }
            ECK(cudaFuncSetAttribute(k_enkf_update_res, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
            k_enkf_update_res<<<gridr, blk, smr, (cudaStream_t)stream>>>(n, ne_local, m, ld, dX, dP, dL, n_loc, dB_local, d_mean, d_bbar, inflate, n_infl, inflate2, dXa);
            ECK(cudaGetLastError());
            return 0;
        }
    }
    size_t sm = (size_t)(32 + 64) * EPAD * sizeof(double);
    int blocks = (int)std::min<int64_t>((n + 31) / 32, 148 * 8);
    k_enkf_update<<<blocks, 128, sm, (cudaStream_t)stream>>>(n, ne_local, m, dX, dP, dL, n_loc, dB_local, d_mean, d_bbar, inflate, n_infl, inflate2, dXa);
    ECK(cudaGetLastError());
    return 0;
}


int32_t cathy_enkf_scale(double *d_v, int64_t n, double a, uint64_t stream)
{
    int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    k_enkf_scale<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(n, a, d_v);
    ECK(cudaGetLastError());
    return 0;
}

// Whole analysis, HOST buffers in and out: H2D of X (and L), the three device stages, D2H of Xa.
int32_t cathy_enkf_analysis_host(const double *X, int64_t n, int32_t ne, const double *HX, const double *y, int32_t y_is_matrix,
                                 const double *R, int32_t m, int32_t sakov, const double *L, int64_t n_loc, double inflate, int64_t n_infl,
                                 double inflate2, double *Xa, double *B, double *P, int32_t device, double *device_ms)
{
    if (n < 1 || ne < 2 || m < 1) EFAIL(-1, "cathy_enkf_analysis_host: need n >= 1, ne >= 2, m >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) EFAIL(-102, "no CUDA device available: the EnKF analysis has no CPU fallback");
    ECK(cudaSetDevice(device));
    std::vector<double> hS((size_t)m * ne), hB((size_t)m * ne), hbbar(m, 0.0);
    int rc = cathy_enkf_gain(HX, y, y_is_matrix, R, m, ne, sakov, hS.data(), hB.data());
    if (rc) return rc;
    for (int i = 0; i < m; ++i) { double t = 0.0; for (int j = 0; j < ne; ++j) t += hB[(size_t)i * ne + j]; hbbar[i] = t / ne; }
    double *dX = nullptr, *dS = nullptr, *dB = nullptr, *dP = nullptr, *dL = nullptr, *dmean = nullptr, *dbbar = nullptr;
    size_t bX = (size_t)n * ne * sizeof(double), bmn = (size_t)m * ne * sizeof(double), bP = (size_t)n * m * sizeof(double);
    cudaEvent_t e0, e1;
    ECK(cudaEventCreate(&e0)); ECK(cudaEventCreate(&e1));
    ECK(cudaMalloc(&dX, bX)); ECK(cudaMalloc(&dS, bmn)); ECK(cudaMalloc(&dB, bmn)); ECK(cudaMalloc(&dP, bP));
    ECK(cudaMalloc(&dmean, (size_t)n * sizeof(double))); ECK(cudaMalloc(&dbbar, (size_t)m * sizeof(double)));
    if (L && n_loc > 0) { ECK(cudaMalloc(&dL, (size_t)n_loc * m * sizeof(double))); ECK(cudaMemcpy(dL, L, (size_t)n_loc * m * sizeof(double), cudaMemcpyHostToDevice)); }
    ECK(cudaMemcpy(dX, X, bX, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(dS, hS.data(), bmn, cudaMemcpyHostToDevice)); ECK(cudaMemcpy(dB, hB.data(), bmn, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(dbbar, hbbar.data(), (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
    ECK(cudaEventRecord(e0, 0));
    rc = cathy_enkf_rowsum(dX, n, ne, dmean, 0);
    if (!rc) rc = cathy_enkf_scale(dmean, n, 1.0 / ne, 0);
    if (!rc) rc = cathy_enkf_crosscov(dX, dmean, dS, n, ne, m, ne, dP, 0);
    if (!rc) rc = cathy_enkf_update(dX, dP, dL, dL ? n_loc : 0, dB, dmean, dbbar, inflate, n_infl, inflate2, n, ne, m, dX, 0);
    ECK(cudaEventRecord(e1, 0));
    ECK(cudaDeviceSynchronize());
    if (!rc) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (device_ms) *device_ms = ms;
        if (Xa) ECK(cudaMemcpy(Xa, dX, bX, cudaMemcpyDeviceToHost));
        if (P) ECK(cudaMemcpy(P, dP, bP, cudaMemcpyDeviceToHost));
        if (B) memcpy(B, hB.data(), bmn);
    }
    cudaFree(dX); cudaFree(dS); cudaFree(dB); cudaFree(dP); cudaFree(dmean); cudaFree(dbbar); if (dL) cudaFree(dL);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

// ---- particle filter ---------------------------------------------------------------------------
int32_t cathy_pf_weights(const double *hx, const double *y, const double *obs_std, int32_t m, int32_t ne, double *weights, double *n_eff)
{
    if (m < 1 || ne < 1) EFAIL(-1, "cathy_pf_weights: need m >= 1 and ne >= 1");
    double *d_hx = nullptr, *d_y = nullptr, *d_s = nullptr, *d_w = nullptr;
    ECK(cudaMalloc(&d_hx, (size_t)m * ne * sizeof(double))); ECK(cudaMalloc(&d_y, (size_t)m * sizeof(double)));
    ECK(cudaMalloc(&d_s, (size_t)m * sizeof(double))); ECK(cudaMalloc(&d_w, ((size_t)ne + 1) * sizeof(double)));
    ECK(cudaMemcpy(d_hx, hx, (size_t)m * ne * sizeof(double), cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(d_y, y, (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(d_s, obs_std, (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
    k_pf_weights<<<1, 256>>>(m, ne, d_hx, d_y, d_s, d_w);
    ECK(cudaGetLastError());
    std::vector<double> h((size_t)ne + 1);
    ECK(cudaMemcpy(h.data(), d_w, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    if (weights) memcpy(weights, h.data(), (size_t)ne * sizeof(double));
    if (n_eff) *n_eff = h[ne];
    cudaFree(d_hx); cudaFree(d_y); cudaFree(d_s); cudaFree(d_w);
    return 0;
}
int32_t cathy_pf_systematic_resample(const double *weights, int32_t ne, double u, int32_t *indices)
{   // ne-element cumulative sum + searchsorted(side='left'): control logic, done where the weights live (host)
    if (ne < 1 || !(u >= 0.0 && u < 1.0)) EFAIL(-1, "cathy_pf_systematic_resample: need ne >= 1 and 0 <= u < 1");
    std::vector<double> cs(ne);
    double acc = 0.0;
    for (int i = 0; i < ne; ++i) { acc += weights[i]; cs[i] = acc; }
    for (int j = 0; j < ne; ++j) {
        double pos = (j + u) / ne;
        indices[j] = (int32_t)(std::lower_bound(cs.begin(), cs.end(), pos) - cs.begin());
    }
    return 0;
}
int32_t cathy_pf_gather_members(const double *dX, int64_t n, int32_t ne, const int32_t *d_idx, double *dXout, uint64_t stream)
{
    if (dX == dXout) EFAIL(-1, "cathy_pf_gather_members: output must not alias the input");
    int blocks = (int)std::min<int64_t>((n * ne + 255) / 256, 148 * 16);
    k_pf_gather<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(n, ne, dX, d_idx, dXout);
    ECK(cudaGetLastError());
    return 0;
}


// DEVICE pointers: grid_xy [n][2], obs_xy [m][2] -> L [n][m]
int32_t cathy_enkf_localization(const double *d_grid_xy, int64_t n, const double *d_obs_xy, int32_t m, double radius, double *dL, uint64_t stream)
{
    if (!(radius > 0.0)) EFAIL(-1, "cathy_enkf_localization: the localisation radius must be positive");
    int blocks = (int)std::min<int64_t>((n * m + 255) / 256, 148 * 16);
    k_enkf_localization<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(n, m, d_grid_xy, d_obs_xy, radius, dL);
    ECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
