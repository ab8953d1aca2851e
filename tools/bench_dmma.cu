// Peak rate of the fp64 tensor-core instruction (mma.sync m8n8k4.f64 = SASS DMMA.8x8x4; the m16n8k16 PTX shape lowers to eight of
// them on sm_100a) and of the plain fp64 FMA pipe, registers only: the roof of k_enkf_crosscov / k_enkf_update.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bench_dmma tools/bench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k_dmma(int iters, double *out)
{
    double acc[NACC][2];
    for (int q = 0; q < NACC; ++q) acc[q][0] = acc[q][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int q = 0; q < NACC; ++q)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[q][0]), "+d"(acc[q][1]) : "d"(a), "d"(b));
    double s = 0.0;
    for (int q = 0; q < NACC; ++q) s += acc[q][0] + acc[q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma(int iters, double *out)
{
    double acc[NACC];
    for (int q = 0; q < NACC; ++q) acc[q] = q;
    double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 2e-3;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[q] = fma(acc[q], a, b);
    double s = 0.0;
    for (int q = 0; q < NACC; ++q) s += acc[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *d; cudaMalloc(&d, 148 * 1024 * 8 * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_dmma<8><<<148 * (1024 / threads), threads>>>(iters, d);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flop = 148.0 * 1024 / 32 * iters * 8 * 512.0;
        printf("DMMA.8x8x4, %4d threads/CTA, 1024 threads/SM, 8 independent accumulators per warp: %.1f TFLOP/s\n", threads, flop / ms / 1e9);
    }
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_dfma<8><<<148, 1024>>>(iters, d); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA, 1024 threads/SM, 8 independent chains per thread: %.1f TFLOP/s\n", 148.0 * 1024 * iters * 8 * 2.0 / ms / 1e9);
    return 0;
}
