// Micro-benchmark of the grid-wide reduction used by the persistent PCG kernels (k_pcg / k_pcg_res):
// how many microseconds does ONE "three sums over 148 CTAs" cost, and which part of it is the barrier?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bench_barrier tools/bench_barrier.cu
// run (GPU box): ./bench_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_xor(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- variant 0: the shipped scheme (fence + atomicAdd + acquire spin, partials re-read by 3 warps)
__device__ __forceinline__ void grid_barrier0(unsigned int *counter, unsigned int &epoch)
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
template <int BLOCK>
__device__ __forceinline__ void reduce3_v0(unsigned int *counter, unsigned int &epoch, double a, double b, double c, double *partial, double (*sh)[3],
                                           double &ra, double &rb, double &rc)
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
    grid_barrier0(counter, epoch);
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

// ---- variant 1: warp 0 does everything between two __syncthreads: it writes the partials, arrives with ONE red.release,
// spins, and then its lanes fetch all partials with independent (unrolled, predicated) loads; fixed summation order.
// Partials are double-buffered by the parity of the reduction count so that a fast CTA cannot overwrite what a slow one
// still reads.
template <int BLOCK, int MAXNB>
__device__ __forceinline__ void reduce3_v1(unsigned int *counter, unsigned int &epoch, unsigned int &par, double a, double b, double c, double *partial,
                                           double (*sh)[3], double &ra, double &rb, double &rc)
{
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { sh[w][0] = a; sh[w][1] = b; sh[w][2] = c; }
    __syncthreads();
    if (w == 0) {
        double t0 = lane < BLOCK / 32 ? sh[lane][0] : 0.0, t1 = lane < BLOCK / 32 ? sh[lane][1] : 0.0, t2 = lane < BLOCK / 32 ? sh[lane][2] : 0.0;
        t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
        double *pp = partial + (size_t)par * 3 * MAXNB;
        epoch += nb;
        if (lane == 0) {
            pp[blockIdx.x] = t0; pp[MAXNB + blockIdx.x] = t1; pp[2 * MAXNB + blockIdx.x] = t2;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            unsigned int v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - epoch) < 0);
        }
        __syncwarp();
        double s[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            double v[MAXNB / 32];
#pragma unroll
            for (int j = 0; j < MAXNB / 32; ++j) {
                const int i = lane + 32 * j;
                v[j] = 0.0;
                if (i < nb) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v[j]) : "l"(pp + q * MAXNB + i) : "memory");
            }
#pragma unroll
            for (int j = 0; j < MAXNB / 32; ++j) s[q] += v[j];
        }
        s[0] = warp_sum_xor(s[0]); s[1] = warp_sum_xor(s[1]); s[2] = warp_sum_xor(s[2]);
        if (lane == 0) { sh[0][0] = s[0]; sh[0][1] = s[1]; sh[0][2] = s[2]; }
        par ^= 1u;
    }
    __syncthreads();
    ra = sh[0][0]; rb = sh[0][1]; rc = sh[0][2];
    __syncthreads();
}

// ---- variant 2: as variant 1, but the block-level sum runs on one warp per quantity in parallel (warps 0,1,2), and
// the second __syncthreads after the read is replaced by a double-buffered result slot.
template <int BLOCK, int MAXNB>
__device__ __forceinline__ void reduce3_v2(unsigned int *counter, unsigned int &epoch, unsigned int &par, double a, double b, double c, double *partial,
                                           double (*sh)[3], double (*res)[3], double &ra, double &rb, double &rc)
{
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { sh[w][0] = a; sh[w][1] = b; sh[w][2] = c; }
    __syncthreads();
    if (w == 0) {
        double *pp = partial + (size_t)par * 3 * MAXNB;
        // lanes 0..31 hold warp partials; three quantities interleaved over the shuffle tree
        double t0 = lane < BLOCK / 32 ? sh[lane][0] : 0.0, t1 = lane < BLOCK / 32 ? sh[lane][1] : 0.0, t2 = lane < BLOCK / 32 ? sh[lane][2] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t0 += __shfl_down_sync(0xffffffffu, t0, o); t1 += __shfl_down_sync(0xffffffffu, t1, o); t2 += __shfl_down_sync(0xffffffffu, t2, o);
        }
        epoch += nb;
        if (lane == 0) {
            pp[blockIdx.x] = t0; pp[MAXNB + blockIdx.x] = t1; pp[2 * MAXNB + blockIdx.x] = t2;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            unsigned int v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - epoch) < 0);
        }
        __syncwarp();
        double v[3][MAXNB / 32];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int j = 0; j < MAXNB / 32; ++j) {
                const int i = lane + 32 * j;
                v[q][j] = 0.0;
                if (i < nb) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v[q][j]) : "l"(pp + q * MAXNB + i) : "memory");
            }
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < MAXNB / 32; ++j) { s0 += v[0][j]; s1 += v[1][j]; s2 += v[2][j]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) { res[par][0] = s0; res[par][1] = s1; res[par][2] = s2; }
    }
    __syncthreads();
    ra = res[par][0]; rb = res[par][1]; rc = res[par][2];
    par ^= 1u;
}

// ---- variant 3: no partials at all: fp64 atomics into a rotating accumulator (NOT bit-reproducible; lower bound probe)
template <int BLOCK>
__device__ __forceinline__ void reduce3_v3(unsigned int *counter, unsigned int &epoch, unsigned int &slot, double a, double b, double c, double *acc,
                                           double (*sh)[3], double &ra, double &rb, double &rc)
{
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { sh[w][0] = a; sh[w][1] = b; sh[w][2] = c; }
    __syncthreads();
    if (w == 0) {
        double t0 = lane < BLOCK / 32 ? sh[lane][0] : 0.0, t1 = lane < BLOCK / 32 ? sh[lane][1] : 0.0, t2 = lane < BLOCK / 32 ? sh[lane][2] : 0.0;
        t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
        epoch += nb;
        if (lane == 0) {
            double *A = acc + 4 * slot, *Z = acc + 4 * ((slot + 1) % 3);
            atomicAdd(A, t0); atomicAdd(A + 1, t1); atomicAdd(A + 2, t2);
            if (blockIdx.x == 0) { Z[0] = 0.0; Z[1] = 0.0; Z[2] = 0.0; }    // the next slot: last read before barrier i-1, first added to after barrier i
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            unsigned int v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - epoch) < 0);
            double r0, r1, r2;
            asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(r0) : "l"(A) : "memory");
            asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(r1) : "l"(A + 1) : "memory");
            asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(r2) : "l"(A + 2) : "memory");
            sh[0][0] = r0; sh[0][1] = r1; sh[0][2] = r2;
        }
        slot = (slot + 1) % 3;
    }
    __syncthreads();
    ra = sh[0][0]; rb = sh[0][1]; rc = sh[0][2];
    __syncthreads();
}

struct Args { unsigned int *counter; double *partial; double *acc; double *z; double *out; int nit; int variant; int work; int n; int off; };

// work = 0: reductions only.  work = 1: between reductions every thread stores one double to z (its row) and, after the
// barrier, reads 15 neighbours (a tiny-mesh PCG phase).  work = 2: phase B-like store only.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1) k_bench(Args a)
{
    __shared__ double sh[BLOCK / 32][3];
    __shared__ double res[2][3];
    unsigned int epoch = *((volatile unsigned int *)a.counter) / gridDim.x * gridDim.x;   // all CTAs read the same idle value
    // make sure every CTA has read the counter before anyone bumps it
    cg::this_grid().sync();
    unsigned int par = 0, slot = 0;
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    double x = 1.0 + 1e-3 * k, s = 0.0;
    for (int it = 0; it < a.nit; ++it) {
        double ra, rb, rc;
        if (a.work) {
            a.z[a.off + k] = x;
        }
        switch (a.variant) {
        case 0: reduce3_v0<BLOCK>(a.counter, epoch, x, 2.0 * x, 0.0, a.partial, sh, ra, rb, rc); break;
        case 1: reduce3_v1<BLOCK, 160>(a.counter, epoch, par, x, 2.0 * x, 0.0, a.partial, sh, ra, rb, rc); break;
        case 2: reduce3_v2<BLOCK, 160>(a.counter, epoch, par, x, 2.0 * x, 0.0, a.partial, sh, res, ra, rb, rc); break;
        default: reduce3_v3<BLOCK>(a.counter, epoch, slot, x, 2.0 * x, 0.0, a.acc, sh, ra, rb, rc); break;
        }
        if (a.work == 1) {
            const volatile double *z = a.z + a.off;
            double acc = 0.0;
#pragma unroll
            for (int d = -7; d <= 7; ++d) acc += z[k + d * 37];
            x = 1.0 + 1e-9 * acc + 1e-12 * ra;
        } else
            x = 1.0 + 1e-12 * (ra + rb);
        s += ra + rb + rc;
    }
    if (threadIdx.x == 0) a.out[blockIdx.x] = s + x;
}

int main()
{
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    printf("SMs %d\n", sms);
    const int BLOCK = 1024;
    Args a{};
    const int n = sms * BLOCK, off = 1024;
    CK(cudaMalloc(&a.counter, 64)); CK(cudaMemset(a.counter, 0, 64));
    CK(cudaMalloc(&a.partial, 2 * 3 * 160 * sizeof(double) + 3 * 1024 * sizeof(double))); CK(cudaMemset(a.partial, 0, 2 * 3 * 160 * sizeof(double) + 3 * 1024 * sizeof(double)));
    CK(cudaMalloc(&a.acc, 16 * sizeof(double))); CK(cudaMemset(a.acc, 0, 16 * sizeof(double)));
    CK(cudaMalloc(&a.z, (n + 2 * off) * sizeof(double))); CK(cudaMemset(a.z, 0, (n + 2 * off) * sizeof(double)));
    CK(cudaMalloc(&a.out, sms * sizeof(double)));
    a.n = n; a.off = off;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int grid : {sms, 37, 8}) {
        for (int work = 0; work < 2; ++work)
            for (int variant = 0; variant < 4; ++variant) {
                a.variant = variant; a.work = work;
                float best = 1e30f;
                for (int rep = 0; rep < 4; ++rep) {
                    a.nit = rep == 0 ? 50 : 2000;
                    CK(cudaMemset(a.counter, 0, 64)); CK(cudaMemset(a.acc, 0, 16 * sizeof(double)));
                    void *args[] = {&a};
                    CK(cudaEventRecord(e0));
                    CK(cudaLaunchCooperativeKernel((void *)k_bench<BLOCK>, dim3(grid), dim3(BLOCK), args, 0, 0));
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep > 0 && ms < best) best = ms;
                }
                double h[1];
                CK(cudaMemcpy(h, a.out, sizeof(double), cudaMemcpyDeviceToHost));
                printf("grid %3d work %d variant %d: %.3f us per reduction (check %.6e)\n", grid, work, variant, best * 1e3 / 2000, h[0]);
            }
    }
    return 0;
}
