// basis_inst.cu -- every k_basis_stream instantiation (factors per leaf x real output channels x real / complex weights),
// behind one plain function (ace_launch.h).
#include "ace_launch.h"

namespace aceb200 {

template <int NFAC, int NCH, bool CW, int EPL>
static void basis_go(const BasisParams& p, int grid, size_t smem, cudaStream_t st)
{
    auto kfn = k_basis_stream<NFAC, NCH, CW, EPL>;
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACE_LAUNCH(kfn, dim3(grid), dim3(32 * p.nw), smem, st, p);
}

template <int NFAC>
static bool basis_nfac(int nch, bool cw, int epl, const BasisParams& p, int grid, size_t smem, cudaStream_t st)
{
    // two environments per lane up to 9 channels (18 accumulators per lane); 18 channels keep one
#define ACE_B(N, C) if (nch == N && cw == C) { constexpr bool E2 = (N <= 9);                                            \
                                               if (epl == 2 && E2) basis_go<NFAC, N, C, (E2 ? 2 : 1)>(p, grid, smem, st);  \
                                               else basis_go<NFAC, N, C, 1>(p, grid, smem, st); return true; }
    ACE_B(1, false) ACE_B(2, false)
    ACE_B(1, true) ACE_B(2, true) ACE_B(3, true) ACE_B(6, true) ACE_B(9, true) ACE_B(18, true)
#undef ACE_B
    return false;
}

bool launch_basis_inst(int nfac, int nch, bool cw, int epl, const BasisParams& p, int grid, size_t smem, cudaStream_t st)
{
    switch (nfac) {
    case 1: case 2: return basis_nfac<2>(nch, cw, epl, p, grid, smem, st);
    case 3: return basis_nfac<3>(nch, cw, epl, p, grid, smem, st);
    case 4: return basis_nfac<4>(nch, cw, epl, p, grid, smem, st);
    default: return false;
    }
}

// resident CTAs per SM of an instantiation at a given geometry (registers and shared memory both count)
template <int NFAC, int NCH, bool CW, int EPL>
static int basis_occ(int threads, size_t smem)
{
#ifdef ACEB200_EMU
    (void)threads; (void)smem;
    return 2;
#else
    auto kfn = k_basis_stream<NFAC, NCH, CW, EPL>;
    int nb = 0;
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kfn, threads, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return nb;
#endif
}

template <int NFAC>
static int basis_occ_nfac(int nch, bool cw, int epl, int threads, size_t smem)
{
#define ACE_B(N, C) if (nch == N && cw == C) { constexpr bool E2 = (N <= 9);                                   \
                                               if (epl == 2 && E2) return basis_occ<NFAC, N, C, (E2 ? 2 : 1)>(threads, smem);  \
                                               return basis_occ<NFAC, N, C, 1>(threads, smem); }
    ACE_B(1, false) ACE_B(2, false)
    ACE_B(1, true) ACE_B(2, true) ACE_B(3, true) ACE_B(6, true) ACE_B(9, true) ACE_B(18, true)
#undef ACE_B
    return 0;
}

int basis_blocks_per_sm(int nfac, int nch, bool cw, int epl, int threads, size_t smem)
{
    switch (nfac) {
    case 1: case 2: return basis_occ_nfac<2>(nch, cw, epl, threads, smem);
    case 3: return basis_occ_nfac<3>(nch, cw, epl, threads, smem);
    case 4: return basis_occ_nfac<4>(nch, cw, epl, threads, smem);
    default: return 0;
    }
}

// host mirror of BasisGeom for the supported combinations
bool basis_geom(int nfac, int nch, bool cw, int& LB, int& LPC, int& HDR, int& W)
{
    const int cs = cw ? 2 : 1, nwd = nch * cs;
    if (!((nch == 1 || nch == 2) || (cw && (nch == 3 || nch == 6 || nch == 9 || nch == 18)))) return false;
    if (nfac < 1 || nfac > 4) return false;
    HDR = (nwd == 1) ? 8 : 16;
    LB = (HDR + 8 * nwd + 15) / 16 * 16;
    LPC = (LB <= 32) ? 512 / LB : std::max(1, 1024 / LB);
    W = (nch >= 9 ? 1 : nch <= 2 ? 4 / nch : (16 + nch - 1) / nch) * nch;
    return true;
}

// measured (config 4b, 9 complex-weight channels): one environment per lane x 16 warps beats two x 12 (3.07 vs 2.71 x 10^7 env/s)
bool basis_epl2(int nch, bool cw) { (void)cw; return nch <= 6; }

}  // namespace aceb200
