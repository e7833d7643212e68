// inst_template.cuh -- included by inst_NN.cu with ACE_INST_NMAX defined: every k_pool / k_forces instantiation for
// that radial bound (species x harmonics walk x channels per pass), behind two plain functions (ace_launch.h).
#include "ace_launch.h"

#define ACE_CAT2(a, b) a##b
#define ACE_CAT(a, b) ACE_CAT2(a, b)

namespace aceb200 {

template <int PB, bool SPECIES, int WALK>
static void forces_go(const ForceParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    auto kfn = k_forces<ACE_INST_NMAX, PB, SPECIES, WALK>;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACE_LAUNCH(kfn, dim3(grid), dim3(kForceThreads), smem, st, p);
}

template <int PB>
static void forces_pb(bool species, bool staticL, const ForceParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    if (species) { if (staticL) forces_go<PB, true, kWalkStatic>(p, grid, smem, st); else forces_go<PB, true, kWalkRolled>(p, grid, smem, st); }
    else { if (staticL) forces_go<PB, false, kWalkStatic>(p, grid, smem, st); else forces_go<PB, false, kWalkRolled>(p, grid, smem, st); }
}

void ACE_CAT(forces_inst_, ACE_INST_NMAX)(int pb, bool species, bool staticL, const ForceParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    if (pb == 1) forces_pb<1>(species, staticL, p, grid, smem, st);
    else if (pb == 3) forces_pb<3>(species, staticL, p, grid, smem, st);
    else forces_pb<2>(species, staticL, p, grid, smem, st);
}

template <bool SPECIES, int WALK>
static void pool_go(const PoolParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    auto kfn = k_pool<ACE_INST_NMAX, SPECIES, WALK>;
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACE_LAUNCH(kfn, dim3(grid), dim3(kPoolThreads), smem, st, p);
}

void ACE_CAT(pool_inst_, ACE_INST_NMAX)(bool species, bool staticL, const PoolParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    if (species) { if (staticL) pool_go<true, kWalkStatic>(p, grid, smem, st); else pool_go<true, kWalkRolled>(p, grid, smem, st); }
    else { if (staticL) pool_go<false, kWalkStatic>(p, grid, smem, st); else pool_go<false, kWalkRolled>(p, grid, smem, st); }
}

template <int WALK>
static void pool_mma_go(const PoolMmaParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    auto kfn = k_pool_mma<ACE_INST_NMAX, WALK>;
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACE_LAUNCH(kfn, dim3(grid), dim3(kPoolMmaThreads), smem, st, p);
}

void ACE_CAT(pool_mma_inst_, ACE_INST_NMAX)(bool staticL, const PoolMmaParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    if (staticL) pool_mma_go<kWalkStatic>(p, grid, smem, st); else pool_mma_go<kWalkRolled>(p, grid, smem, st);
}

template <int WALK>
static void forces_mma_go(const ForceMmaParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    auto kfn = k_forces_mma<ACE_INST_NMAX, WALK>;
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACE_LAUNCH(kfn, dim3(grid), dim3(kFmmaThreads), smem, st, p);
}

void ACE_CAT(forces_mma_inst_, ACE_INST_NMAX)(bool staticL, const ForceMmaParams& p, unsigned grid, size_t smem, cudaStream_t st)
{
    if (staticL) forces_mma_go<kWalkStatic>(p, grid, smem, st); else forces_mma_go<kWalkRolled>(p, grid, smem, st);
}

}  // namespace aceb200
