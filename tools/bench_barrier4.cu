// Why is grid_reduce2 (k_pcg_res2) slower in the kernel than the pieces measured in bench_barrier2/3?  Variants of the spin:
//  A: lane 0 spins, lanes 1-31 parked at __syncwarp (as first written)   B: all 32 lanes spin on the counter (no divergence)
//  C: thread 0 spins between two __syncthreads, then warps 0 and 1 read one quantity each
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_barrier4 tools/bench_barrier4.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
#define FULLMASK 0xffffffffu

template <int BLOCK, int VAR>
__device__ __forceinline__ void grid_reduce2(unsigned int *counter, unsigned int &epoch, unsigned int &par, double a, double b, double *partial,
                                             double (*sh)[2], double (*res)[2], double &ra, double &rb)
{
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool hi = lane >= 16;
    double keep = hi ? b : a, send = hi ? a : b;
    keep += __shfl_xor_sync(FULLMASK, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(FULLMASK, keep, o);
    if ((lane & 15) == 0) sh[w][hi] = keep;
    __syncthreads();
    double *pp = partial + (size_t)par * 2 * nb;
    if (w == 0) {
        double v = hi ? sh[lane - 16][1] + sh[lane][1] : sh[lane][0] + sh[lane + 16][0];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
        if ((lane & 15) == 0) pp[(hi ? nb : 0) + blockIdx.x] = v;
        __syncwarp();
        epoch += nb;
        if (VAR == 0) {
            if (lane == 0) {
                __threadfence();
                atomicAdd(counter, 1u);
                unsigned int c;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(counter) : "memory"); } while ((int)(c - epoch) < 0);
            }
            __syncwarp();
        } else if (VAR == 1) {
            if (lane == 0) { __threadfence(); atomicAdd(counter, 1u); }
            unsigned int c;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(counter) : "memory"); } while ((int)(c - epoch) < 0);
        } else {
            if (lane == 0) {
                __threadfence();
                atomicAdd(counter, 1u);
                unsigned int c;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(counter) : "memory"); } while ((int)(c - epoch) < 0);
            }
        }
    }
    if (VAR == 2) {
        __syncthreads();
        if (w < 2) {
            double v[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int i = lane + 32 * j;
                v[j] = 0.0;
                if (i < nb) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v[j]) : "l"(pp + w * nb + i) : "memory");
            }
            double t = (((v[0] + v[1]) + v[2]) + v[3]) + v[4];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULLMASK, t, o);
            if (lane == 0) res[par][w] = t;
        }
    } else if (w == 0) {
        constexpr int MAXJ = 5;
        double va[MAXJ], vb[MAXJ];
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int i = lane + 32 * j;
            va[j] = 0.0; vb[j] = 0.0;
            if (i < nb) {
                asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(va[j]) : "l"(pp + i) : "memory");
                asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(vb[j]) : "l"(pp + nb + i) : "memory");
            }
        }
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) { sa += va[j]; sb += vb[j]; }
        double k2 = hi ? sb : sa, s2 = hi ? sa : sb;
        k2 += __shfl_xor_sync(FULLMASK, s2, 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) k2 += __shfl_xor_sync(FULLMASK, k2, o);
        if ((lane & 15) == 0) res[par][hi] = k2;
    }
    __syncthreads();
    ra = res[par][0]; rb = res[par][1];
    par ^= 1u;
}

struct Args { unsigned int *counter; double *partial; double *z; double *out; int nit; int work; int off; };

template <int BLOCK, int VAR>
__global__ void __launch_bounds__(BLOCK, 1) k_bench(Args a)
{
    __shared__ double sh[BLOCK / 32][2];
    __shared__ double res[2][2];
    unsigned int epoch = 0, par = 0;
    cg::this_grid().sync();
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    double x = 1.0 + 1e-3 * k, s = 0.0;
    for (int it = 0; it < a.nit; ++it) {
        if (a.work) a.z[a.off + k] = x;
        double ra, rb;
        grid_reduce2<BLOCK, VAR>(a.counter, epoch, par, x, 2.0 * x, a.partial, sh, res, ra, rb);
        if (a.work == 1) {
            const double *z = a.z + a.off;
            double acc = 0.0;
#pragma unroll
            for (int d = -7; d <= 7; ++d) acc += z[k + d * 37];
            x = 1.0 + 1e-9 * acc + 1e-12 * ra;
        } else
            x = 1.0 + 1e-12 * (ra + rb);
        s += ra + rb;
    }
    if (threadIdx.x == 0) a.out[blockIdx.x] = s + x;
}

template <int VAR>
int run(Args a, int grid, int work, const char *name)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    a.work = work;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        a.nit = rep == 0 ? 50 : 2000;
        CK(cudaMemset(a.counter, 0, 64));
        void *args[] = {&a};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void *)k_bench<1024, VAR>, dim3(grid), dim3(1024), args, 0, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    double o[1];
    CK(cudaMemcpy(o, a.out, sizeof(o), cudaMemcpyDeviceToHost));
    printf("grid %3d work %d %-44s %.3f us/red (check %.6e)\n", grid, work, name, best * 1e3 / 2000, o[0]);
    return 0;
}

int main()
{
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    Args a{};
    const int n = sms * 1024, off = 1024;
    CK(cudaMalloc(&a.counter, 64));
    CK(cudaMalloc(&a.partial, 4 * 160 * sizeof(double))); CK(cudaMemset(a.partial, 0, 4 * 160 * sizeof(double)));
    CK(cudaMalloc(&a.z, (n + 2 * off) * sizeof(double))); CK(cudaMemset(a.z, 0, (n + 2 * off) * sizeof(double)));
    CK(cudaMalloc(&a.out, sms * sizeof(double)));
    a.off = off;
    for (int grid : {sms, 8})
        for (int work = 0; work < 2; ++work) {
            if (run<0>(a, grid, work, "A lane 0 spins, others parked at syncwarp")) return 1;
            if (run<1>(a, grid, work, "B all lanes of warp 0 spin")) return 1;
            if (run<2>(a, grid, work, "C thread 0 spins, syncthreads, 2 warps read")) return 1;
        }
    return 0;
}
