// stream_template.cuh -- included by stream_nfN.cu with ACE_STREAM_NF defined: every k_adjoint_stream instantiation for
// that number of leaf factors (channels per pass x real / complex weights x environments per lane), behind one plain
// function (ace_launch.h).
#include "ace_launch.h"

#define ACE_CAT2(a, b) a##b
#define ACE_CAT(a, b) ACE_CAT2(a, b)

namespace aceb200 {

template <int PB, bool CW, int EPL = 1>
static void stream_go(const StreamParams& p, int grid, size_t smem, cudaStream_t st)
{
    auto kfn = k_adjoint_stream<ACE_STREAM_NF, PB, CW, EPL>;
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACE_LAUNCH(kfn, dim3(grid), dim3(32 * StreamGeom<ACE_STREAM_NF, PB, CW>::NW), smem, st, p);
}

void ACE_CAT(stream_inst_nf, ACE_STREAM_NF)(int pb, bool cw, int epl, const StreamParams& p, int grid, size_t smem, cudaStream_t st)
{
#define ACE_S(PBV) { if (cw) stream_go<PBV, true>(p, grid, smem, st); else stream_go<PBV, false>(p, grid, smem, st); }
    switch (pb) {
    case 1:
        if (!cw && epl == 2) stream_go<1, false, 2>(p, grid, smem, st);
        else ACE_S(1)
        break;
    case 2: ACE_S(2) break;
    case 4: ACE_S(4) break;
    default: ACE_S(8) break;
    }
#undef ACE_S
}

}  // namespace aceb200
