// k_group.cu -- instantiations of the generic on-chip group kernel (ldpc_kernels.cuh: ldpc_ms_group_kernel) and the
// choice among them for a plan's shape.
#include "ldpc_launch.h"
#include "ldpc_kernels.cuh"

namespace ldpc_b200 {
namespace {

template <int G, int DMAX, bool TAB, int MAXT, bool YS, class PROF = GenericProfile, bool T16 = false>
int launch_group_t(const GroupParams& q, int grid, int threads, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ldpc_ms_group_kernel<G, DMAX, TAB, MAXT, YS, PROF, T16>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ldpc_ms_group_kernel<G, DMAX, TAB, MAXT, YS, PROF, T16><<<grid, threads, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

template <class P>
bool profile_matches(const GroupSel& pl, const GroupParams& q) {
    if (pl.CS != P::CS || pl.VS != P::VS) return false;
    for (int i = 0; i < P::CS; ++i) if (q.cdeg[i] != P::cdeg(i)) return false;
    for (int i = 0; i < P::VS; ++i) if (q.vdeg[i] != P::vdeg(i)) return false;
    return true;
}

}  // namespace

int k_launch_group(const GroupSel& pl, const GroupParams& q, int grid, size_t sm, cudaStream_t stream) {
    const int th = pl.threads;
    if (pl.dmax > 16) return kNoKernel;  // no instantiation unrolls more than 16 edges per check
    if (pl.G == 16 && pl.tab_smem) {
        if (pl.dmax == 8) return pl.y_smem ? launch_group_t<16, 8, true, 1024, true>(q, grid, th, sm, stream)
                                           : launch_group_t<16, 8, true, 1024, false>(q, grid, th, sm, stream);
        if (pl.y_smem) return th <= 768 ? launch_group_t<16, 16, true, 768, true>(q, grid, th, sm, stream)
                                        : launch_group_t<16, 16, true, 1024, true>(q, grid, th, sm, stream);
        return th <= 768 ? launch_group_t<16, 16, true, 768, false>(q, grid, th, sm, stream)
                         : launch_group_t<16, 16, true, 1024, false>(q, grid, th, sm, stream);
    }
    if (pl.G == 4 && pl.tab_smem && th <= 288) {
        if (!pl.y_smem && pl.allow_profile && profile_matches<ProfileWimax34B576L72>(pl, q))
            return launch_group_t<4, 16, true, 288, false, ProfileWimax34B576L72>(q, grid, th, sm, stream);
        return pl.y_smem ? launch_group_t<4, 16, true, 288, true>(q, grid, th, sm, stream)
                         : launch_group_t<4, 16, true, 288, false>(q, grid, th, sm, stream);
    }
    if (pl.G == 8 && pl.tab_smem && th <= 384) {
        if (pl.t16) return launch_group_t<8, 16, true, 384, false, ProfileWimax34B576, true>(q, grid, th, sm, stream);
        if (!pl.y_smem && pl.allow_profile && profile_matches<ProfileWimax34B576>(pl, q))
            return launch_group_t<8, 16, true, 384, false, ProfileWimax34B576>(q, grid, th, sm, stream);
        if (pl.dmax == 8) return pl.y_smem ? launch_group_t<8, 8, true, 384, true>(q, grid, th, sm, stream)
                                           : launch_group_t<8, 8, true, 384, false>(q, grid, th, sm, stream);
        return pl.y_smem ? launch_group_t<8, 16, true, 384, true>(q, grid, th, sm, stream)
                         : launch_group_t<8, 16, true, 384, false>(q, grid, th, sm, stream);
    }
    if (pl.G == 1 && !pl.tab_smem && pl.dmax == 8 && pl.allow_profile && profile_matches<ProfileRegular36N8192>(pl, q))
        return pl.t16 ? launch_group_t<1, 8, false, 1024, false, ProfileRegular36N8192, true>(q, grid, th, sm, stream)
                      : launch_group_t<1, 8, false, 1024, false, ProfileRegular36N8192>(q, grid, th, sm, stream);
    if (pl.G == 1 && pl.y_smem) {
        if (pl.tab_smem) return pl.dmax == 8 ? launch_group_t<1, 8, true, 1024, true>(q, grid, th, sm, stream)
                                             : launch_group_t<1, 16, true, 1024, true>(q, grid, th, sm, stream);
        return pl.dmax == 8 ? launch_group_t<1, 8, false, 1024, true>(q, grid, th, sm, stream)
                            : launch_group_t<1, 16, false, 1024, true>(q, grid, th, sm, stream);
    }
    if (pl.G == 1) {
        if (pl.tab_smem) return pl.dmax == 8 ? launch_group_t<1, 8, true, 1024, false>(q, grid, th, sm, stream)
                                             : launch_group_t<1, 16, true, 1024, false>(q, grid, th, sm, stream);
        return pl.dmax == 8 ? launch_group_t<1, 8, false, 1024, false>(q, grid, th, sm, stream)
                            : launch_group_t<1, 16, false, 1024, false>(q, grid, th, sm, stream);
    }
    return kNoKernel;
}
}  // namespace ldpc_b200
