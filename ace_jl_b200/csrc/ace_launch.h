// ace_launch.h -- the two large kernel families (k_pool, k_forces) are instantiated in one translation unit per
// radial bound NMAX (inst_NN.cu, generated from inst_template.cuh), and k_adjoint_stream in one per number of leaf
// factors NF (stream_nfN.cu, from stream_template.cuh), so that the library builds in parallel.
#pragma once

#include <stdexcept>
#include <string>

#include "../../include/aceb200.h"
#include "ace_kernels.cuh"
#include "ace_tables.h"

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            throw ModelError(err__ == cudaErrorMemoryAllocation ? ACEB200_ENOMEM : ACEB200_ECUDA, \
                             std::string(#call) + ": " + cudaGetErrorString(err__));              \
    } while (0)

namespace aceb200 {

#define ACE_INST_DECL(N)                                                                                                    \
    void forces_inst_##N(int pb, bool species, bool staticL, const ForceParams& p, unsigned grid, size_t smem, cudaStream_t st); \
    void pool_inst_##N(bool species, bool staticL, const PoolParams& p, unsigned grid, size_t smem, cudaStream_t st);               \
    void pool_mma_inst_##N(bool staticL, const PoolMmaParams& p, unsigned grid, size_t smem, cudaStream_t st);                     \
    void forces_mma_inst_##N(bool staticL, const ForceMmaParams& p, unsigned grid, size_t smem, cudaStream_t st);
ACE_INST_DECL(4) ACE_INST_DECL(8) ACE_INST_DECL(12) ACE_INST_DECL(16) ACE_INST_DECL(20) ACE_INST_DECL(24) ACE_INST_DECL(32)
#undef ACE_INST_DECL

void stream_inst_nf2(int pb, bool cw, int epl, const StreamParams& p, int grid, size_t smem, cudaStream_t st);
void stream_inst_nf3(int pb, bool cw, int epl, const StreamParams& p, int grid, size_t smem, cudaStream_t st);
void stream_inst_nf4(int pb, bool cw, int epl, const StreamParams& p, int grid, size_t smem, cudaStream_t st);

bool launch_basis_inst(int nfac, int nch, bool cw, int epl, const BasisParams& p, int grid, size_t smem, cudaStream_t st);
bool basis_geom(int nfac, int nch, bool cw, int& LB, int& LPC, int& HDR, int& W);
int basis_blocks_per_sm(int nfac, int nch, bool cw, int epl, int threads, size_t smem);
bool basis_epl2(int nch, bool cw);

inline void launch_stream_inst(int nf, int pb, bool cw, int epl, const StreamParams& p, int grid, size_t smem, cudaStream_t st)
{
    if (nf == 2) stream_inst_nf2(pb, cw, epl, p, grid, smem, st);
    else if (nf == 3) stream_inst_nf3(pb, cw, epl, p, grid, smem, st);
    else stream_inst_nf4(pb, cw, epl, p, grid, smem, st);
}

inline void launch_forces_inst(int nmax, int pb, bool species, bool staticL, const ForceParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    switch (nmax) {
    case 4: forces_inst_4(pb, species, staticL, p, grid, smem, st); break;
    case 8: forces_inst_8(pb, species, staticL, p, grid, smem, st); break;
    case 12: forces_inst_12(pb, species, staticL, p, grid, smem, st); break;
    case 16: forces_inst_16(pb, species, staticL, p, grid, smem, st); break;
    case 20: forces_inst_20(pb, species, staticL, p, grid, smem, st); break;
    case 24: forces_inst_24(pb, species, staticL, p, grid, smem, st); break;
    default: forces_inst_32(pb, species, staticL, p, grid, smem, st); break;
    }
}

inline void launch_pool_inst(int nmax, bool species, bool staticL, const PoolParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    switch (nmax) {
    case 4: pool_inst_4(species, staticL, p, grid, smem, st); break;
    case 8: pool_inst_8(species, staticL, p, grid, smem, st); break;
    case 12: pool_inst_12(species, staticL, p, grid, smem, st); break;
    case 16: pool_inst_16(species, staticL, p, grid, smem, st); break;
    case 20: pool_inst_20(species, staticL, p, grid, smem, st); break;
    case 24: pool_inst_24(species, staticL, p, grid, smem, st); break;
    default: pool_inst_32(species, staticL, p, grid, smem, st); break;
    }
}

inline void launch_forces_mma_inst(int nmax, bool staticL, const ForceMmaParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    switch (nmax) {
    case 4: forces_mma_inst_4(staticL, p, grid, smem, st); break;
    case 8: forces_mma_inst_8(staticL, p, grid, smem, st); break;
    case 12: forces_mma_inst_12(staticL, p, grid, smem, st); break;
    case 16: forces_mma_inst_16(staticL, p, grid, smem, st); break;
    case 20: forces_mma_inst_20(staticL, p, grid, smem, st); break;
    case 24: forces_mma_inst_24(staticL, p, grid, smem, st); break;
    default: forces_mma_inst_32(staticL, p, grid, smem, st); break;
    }
}

inline void launch_pool_mma_inst(int nmax, bool staticL, const PoolMmaParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    switch (nmax) {
    case 4: pool_mma_inst_4(staticL, p, grid, smem, st); break;
    case 8: pool_mma_inst_8(staticL, p, grid, smem, st); break;
    case 12: pool_mma_inst_12(staticL, p, grid, smem, st); break;
    case 16: pool_mma_inst_16(staticL, p, grid, smem, st); break;
    case 20: pool_mma_inst_20(staticL, p, grid, smem, st); break;
    case 24: pool_mma_inst_24(staticL, p, grid, smem, st); break;
    default: pool_mma_inst_32(staticL, p, grid, smem, st); break;
    }
}

}  // namespace aceb200
