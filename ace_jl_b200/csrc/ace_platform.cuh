// ace_platform.cuh -- the one place that knows whether the source is being compiled by nvcc for
// sm_100a (the product) or by g++ against tests/emu/cuda_emu.h (a TEST-ONLY thread-per-CUDA-thread
// emulation used to debug index logic in a container that has no GPU; never built into or loaded by
// the product library).
#pragma once

#ifdef ACEB200_EMU
#include "cuda_emu.h"   // found through -I tests/emu
#define ACE_HD
#define ACE_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(grid, block, smem, [=]() { kernel(__VA_ARGS__); })
#define ACE_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::smem_base())
#else
#include <cuda_runtime.h>
#define ACE_HD __host__ __device__
#define ACE_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define ACE_DYN_SMEM(type, name) \
    extern __shared__ __align__(16) unsigned char ace_dyn_smem_raw[]; \
    type* name = reinterpret_cast<type*>(ace_dyn_smem_raw)
#endif

#include <cmath>
#include <cstdint>
