// cuda_emu.h -- TEST-ONLY stand-in for the CUDA runtime and execution model.
//
// The build container has no GPU.  To debug the index logic of the kernels in
// ace_jl_b200/csrc before spending GPU minutes, tests/emu/build_emu.sh compiles the very same
// sources with g++ -DACEB200_EMU against this header: every CUDA thread of a block becomes one
// std::thread, __syncthreads() a pthread barrier, device memory the host heap.  Blocks run one
// after the other.  The resulting tests/emu/libaceb200_emu.so is loaded only by
// tests/test_emu_*.py; the product (ace_jl_b200/_lib.py) only ever loads csrc/libaceb200.so and has
// no CPU path.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
inline int __double2hiint(double v) { union { double d; unsigned long long u; } w; w.d = v; return (int)(w.u >> 32); }
inline int __double2loint(double v) { union { double d; unsigned long long u; } w; w.d = v; return (int)(w.u & 0xffffffffu); }
inline double __hiloint2double(int hi, int lo) { union { double d; unsigned long long u; } w; w.u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; return w.d; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace emu {
struct BlockCtx { pthread_barrier_t bar; pthread_barrier_t wbar[32]; unsigned char* smem; double frag_a[32][32]; double frag_b[32][32]; };
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local dim3 t_blockDim, t_gridDim;
inline thread_local BlockCtx* t_ctx;
inline unsigned char* smem_base() { return t_ctx->smem; }

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F fn)
{
    unsigned nt = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        BlockCtx ctx;
        pthread_barrier_init(&ctx.bar, nullptr, nt);
        for (unsigned w = 0; w * 32 < nt; ++w) pthread_barrier_init(&ctx.wbar[w], nullptr, std::min(32u, nt - w * 32));
        ctx.smem = (unsigned char*)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([=, &ctx]() {
                t_threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                t_blockIdx = uint3{bx, by, bz};
                t_blockDim = block;
                t_gridDim = grid;
                t_ctx = &ctx;
                fn();
            });
        for (auto& x : th) x.join();
        free(ctx.smem);
        pthread_barrier_destroy(&ctx.bar);
        for (unsigned w = 0; w * 32 < nt; ++w) pthread_barrier_destroy(&ctx.wbar[w]);
    }
}
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::t_blockDim
#define gridDim emu::t_gridDim
inline void __syncthreads() { pthread_barrier_wait(&emu::t_ctx->bar); }
inline void __syncwarp()
{
    unsigned t = emu::t_threadIdx.x + emu::t_blockDim.x * (emu::t_threadIdx.y + emu::t_blockDim.y * emu::t_threadIdx.z);
    pthread_barrier_wait(&emu::t_ctx->wbar[t / 32]);
}
template <class T> inline T __ldg(const T* p) { return *p; }
namespace emu {
// mma.sync.aligned.m8n8k4.row.col.f64: D = A(8x4, row) * B(4x8, col) + C.  Fragments (PTX ISA): lane i holds
// A[i / 4][i % 4], B[i % 4][i / 4], C[i / 4][2 (i % 4) + {0, 1}].
inline void mma_m8n8k4(double& c0, double& c1, double a, double b)
{
    unsigned t = t_threadIdx.x + t_blockDim.x * (t_threadIdx.y + t_blockDim.y * t_threadIdx.z);
    unsigned w = t / 32, lane = t % 32;
    t_ctx->frag_a[w][lane] = a; t_ctx->frag_b[w][lane] = b;
    pthread_barrier_wait(&t_ctx->wbar[w]);
    unsigned row = lane / 4, col = 2 * (lane % 4);
    for (unsigned k = 0; k < 4; ++k) {
        double av = t_ctx->frag_a[w][row * 4 + k];
        c0 = std::fma(av, t_ctx->frag_b[w][col * 4 + k], c0);
        c1 = std::fma(av, t_ctx->frag_b[w][(col + 1) * 4 + k], c1);
    }
    pthread_barrier_wait(&t_ctx->wbar[w]);
}
inline double shfl_xor(double v, int mask)
{
    unsigned t = t_threadIdx.x + t_blockDim.x * (t_threadIdx.y + t_blockDim.y * t_threadIdx.z);
    unsigned w = t / 32, lane = t % 32;
    t_ctx->frag_a[w][lane] = v;
    pthread_barrier_wait(&t_ctx->wbar[w]);
    double r = t_ctx->frag_a[w][lane ^ (unsigned)mask];
    pthread_barrier_wait(&t_ctx->wbar[w]);
    return r;
}
}  // namespace emu
inline void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMax(int* p, int v) { int o = *p; while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }

inline double atomicAdd(double* p, double v)
{
    double o = *p, n;
    do { n = o + v; } while (!__atomic_compare_exchange(p, &o, &n, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return o;
}

// ---- runtime API ------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct emuEvent { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaDevAttrMultiProcessorCount = 16,
       cudaDevAttrMaxSharedMemoryPerBlockOptin = 97, cudaEventDefault = 0, cudaHostAllocDefault = 0 };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 2; return 0; }   // two pretend devices: exercises aceb200_set_devices
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
    *v = (attr == cudaDevAttrMultiProcessorCount) ? 2 : 227 * 1024; return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, ((n + 255) / 256 + 1) * 256); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
enum { cudaStreamNonBlocking = 1 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return 0; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{0}; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return 0; }
