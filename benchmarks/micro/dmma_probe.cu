// dmma_probe.cu -- latency / throughput of mma.sync.m8n8k4.f64 (DMMA) and DFMA on sm_100a as a function of
// independent chains per warp and resident warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k_mma(int iters, double* sink, long long* cyc)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[CH][2];
#pragma unroll
    for (int k = 0; k < CH; ++k) { c[k][0] = k; c[k][1] = -k; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CH; ++k) dmma(c[k][0], c[k][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += c[k][0] + c[k][1];
    if (s == 1.2345) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CH>
__global__ void k_fma(int iters, double* sink, long long* cyc)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    double c[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) c[k] = k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CH; ++k) c[k] = c[k] * a + b;
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += c[k];
    if (s == 1.2345) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// shared-memory LDS.64 in the fragment pattern (pitch 132) feeding dmma: the inner loop of k_pool_mma
__global__ void k_lds_mma(int iters, double* sink, long long* cyc)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 42 * 132; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
    __syncthreads();
    const int lane = threadIdx.x & 31, k4 = lane & 3, g = lane >> 2;
    const double* pa = sm + g * 132 + k4;
    const double* pb = sm + (12 + g) * 132 + k4;
    double c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 40; k += 4) {
            double b = pb[k];
            dmma(c0, c1, pa[k], b);
            dmma(d0, d1, pa[k + 8 * 132 > 0 ? k : k], b);
        }
    }
    long long t1 = clock64();
    if (c0 + c1 + d0 + d1 == 1.2345) sink[0] = c0;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <class K>
void run(const char* name, K kern, int ch, int warps, int ctas_per_sm, int sms, double flop_per_iter_warp, size_t smem = 0)
{
    double* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (smem) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<sms * ctas_per_sm, warps * 32, smem>>>(iters, sink, cyc);
    cudaEventRecord(e0);
    kern<<<sms * ctas_per_sm, warps * 32, smem>>>(iters, sink, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double tf = (double)sms * ctas_per_sm * warps * iters * flop_per_iter_warp / (ms * 1e-3) / 1e12;
    printf("%-10s chains %2d warps/CTA %2d CTAs/SM %2d : %8.1f cycles/iter/warp  %7.2f TFLOP/s  (%s)\n", name, ch, warps, ctas_per_sm,
           (double)h / iters, tf, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink); cudaFree(cyc);
}

int main()
{
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("== DMMA m8n8k4 (512 flop per warp instruction)\n");
    run("dmma", k_mma<1>, 1, 1, 1, sms, 1 * 512.0);
    run("dmma", k_mma<2>, 2, 1, 1, sms, 2 * 512.0);
    run("dmma", k_mma<4>, 4, 1, 1, sms, 4 * 512.0);
    run("dmma", k_mma<8>, 8, 1, 1, sms, 8 * 512.0);
    run("dmma", k_mma<1>, 1, 4, 1, sms, 1 * 512.0);
    run("dmma", k_mma<1>, 1, 4, 2, sms, 1 * 512.0);
    run("dmma", k_mma<1>, 1, 4, 5, sms, 1 * 512.0);
    run("dmma", k_mma<2>, 2, 4, 5, sms, 2 * 512.0);
    run("dmma", k_mma<4>, 4, 4, 2, sms, 4 * 512.0);
    run("dmma", k_mma<4>, 4, 4, 5, sms, 4 * 512.0);
    run("dmma", k_mma<8>, 8, 8, 8, sms, 8 * 512.0);
    printf("== DFMA (64 flop per warp instruction)\n");
    run("dfma", k_fma<1>, 1, 1, 1, sms, 1 * 64.0);
    run("dfma", k_fma<4>, 4, 1, 1, sms, 4 * 64.0);
    run("dfma", k_fma<8>, 8, 1, 1, sms, 8 * 64.0);
    run("dfma", k_fma<2>, 2, 4, 2, sms, 2 * 64.0);
    run("dfma", k_fma<2>, 2, 4, 5, sms, 2 * 64.0);
    run("dfma", k_fma<4>, 4, 4, 5, sms, 4 * 64.0);
    run("dfma", k_fma<8>, 8, 8, 8, sms, 8 * 64.0);
    printf("== LDS.64 fragment loads + DMMA (k_pool_mma inner loop: 10 k-steps x 2 mma = 20 x 512 flop per iter)\n");
    run("lds+dmma", k_lds_mma, 2, 4, 1, sms, 20 * 512.0, 42 * 132 * 8);
    run("lds+dmma", k_lds_mma, 2, 4, 5, sms, 20 * 512.0, 42 * 132 * 8);
    return 0;
}
