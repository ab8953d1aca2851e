// Instrumented grid reduction: where do the ~2.9 us of one "three sums over all CTAs" go?  (clock64 around each segment,
// thread 0 of CTA 0), for a few fence / spin flavours.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_barrier2 tools/bench_barrier2.cu
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
struct Args { unsigned int *counter; double *partial; double *z; double *out; long long *clk; int nit; int fence; int spin; int work; int off; };

template <int BLOCK, int FENCE, int SPIN>
__global__ void __launch_bounds__(BLOCK, 1) k_bench(Args a)
{
    __shared__ double sh[BLOCK / 32][3];
    unsigned int epoch = 0;
    cg::this_grid().sync();
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    double x = 1.0 + 1e-3 * k, s = 0.0;
    long long T[7] = {0, 0, 0, 0, 0, 0, 0};
    unsigned int par = 0;
    for (int it = 0; it < a.nit; ++it) {
        if (a.work) a.z[a.off + k] = x;
        long long c0 = clock64();
        double va = warp_sum(x), vb = warp_sum(2.0 * x), vc = warp_sum(0.5 * x);
        if (lane == 0) { sh[w][0] = va; sh[w][1] = vb; sh[w][2] = vc; }
        __syncthreads();
        double *pp = a.partial + par * 3 * 160;
        if (w == 0) {
            double t0 = sh[lane][0], t1 = sh[lane][1], t2 = sh[lane][2];
            t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
            if (lane == 0) { pp[blockIdx.x] = t0; pp[160 + blockIdx.x] = t1; pp[320 + blockIdx.x] = t2; }
        }
        __syncthreads();
        long long c1 = clock64(), c2 = c1, c3 = c1, c4 = c1;
        if (threadIdx.x == 0) {
            epoch += nb;
            if (FENCE == 0) __threadfence();
            else if (FENCE == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            c2 = clock64();
            if (FENCE == 3) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.counter) : "memory");
            else asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(a.counter) : "memory");
            c3 = clock64();
            unsigned int v;
            if (SPIN == 0) {
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.counter) : "memory"); } while ((int)(v - epoch) < 0);
            } else {
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.counter) : "memory"); } while ((int)(v - epoch) < 0);
                if (SPIN == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
            c4 = clock64();
        }
        __syncthreads();
        long long c5 = clock64();
        if (w < 3) {
            double v[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int i = lane + 32 * j;
                v[j] = 0.0;
                if (i < nb) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v[j]) : "l"(pp + w * 160 + i) : "memory");
            }
            double t = warp_sum((((v[0] + v[1]) + v[2]) + v[3]) + v[4]);
            if (lane == 0) sh[0][w] = t;
        }
        __syncthreads();
        double ra = sh[0][0], rb = sh[0][1], rc = sh[0][2];
        __syncthreads();
        long long c6 = clock64();
        par ^= 1u;
        if (a.work == 1) {
            const double *z = a.z + a.off;
            double acc = 0.0;
#pragma unroll
            for (int d = -7; d <= 7; ++d) acc += z[k + d * 37];
            x = 1.0 + 1e-9 * acc + 1e-12 * ra;
        } else
            x = 1.0 + 1e-12 * (ra + rb);
        s += ra + rb + rc;
        long long c7 = clock64();
        T[0] += c1 - c0; T[1] += c2 - c1; T[2] += c3 - c2; T[3] += c4 - c3; T[4] += c5 - c4; T[5] += c6 - c5; T[6] += c7 - c6;
    }
    if (threadIdx.x == 0) a.out[blockIdx.x] = s + x;
    if (threadIdx.x == 0 && blockIdx.x == 0) for (int i = 0; i < 7; ++i) a.clk[i] = T[i];
}

template <int F, int S>
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
        CK(cudaLaunchCooperativeKernel((void *)k_bench<1024, F, S>, dim3(grid), dim3(1024), args, 0, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    long long h[7];
    CK(cudaMemcpy(h, a.clk, sizeof(h), cudaMemcpyDeviceToHost));
    printf("grid %3d work %d %-34s %.3f us/red | cycles: blocksum %lld fence %lld red %lld spin %lld sync %lld read+sum %lld phase %lld\n", grid, work, name,
           best * 1e3 / 2000, h[0] / 2000, h[1] / 2000, h[2] / 2000, h[3] / 2000, h[4] / 2000, h[5] / 2000, h[6] / 2000);
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
    CK(cudaMalloc(&a.partial, 2 * 3 * 160 * sizeof(double))); CK(cudaMemset(a.partial, 0, 2 * 3 * 160 * sizeof(double)));
    CK(cudaMalloc(&a.z, (n + 2 * off) * sizeof(double))); CK(cudaMemset(a.z, 0, (n + 2 * off) * sizeof(double)));
    CK(cudaMalloc(&a.out, sms * sizeof(double)));
    CK(cudaMalloc(&a.clk, 64));
    a.off = off;
    for (int grid : {sms, 8})
        for (int work = 0; work < 2; ++work) {
            if (run<0, 0>(a, grid, work, "threadfence + ld.acquire spin")) return 1;
            if (run<1, 0>(a, grid, work, "fence.acq_rel + ld.acquire spin")) return 1;
            if (run<3, 0>(a, grid, work, "red.release + ld.acquire spin")) return 1;
            if (run<0, 1>(a, grid, work, "threadfence + relaxed spin + fence")) return 1;
            if (run<1, 1>(a, grid, work, "fence.acq_rel + relaxed spin + fence")) return 1;
            if (run<2, 2>(a, grid, work, "NO fences (timing only)")) return 1;
        }
    return 0;
}
