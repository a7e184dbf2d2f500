// k_qcw.cu -- the warp-per-codeword quasi-cyclic kernel (ldpc_qcw.cuh): one instantiation per 802.16e rate at z = 24
// and z = 32 (N = 576, 768), and the host-side check that a handle's H is the compiled code.
#include <cstring>

#include "ldpc_launch.h"
#define LDPC_QCW_DEVICE
#include "ldpc_qcw.cuh"

namespace ldpc_b200 {
namespace {

// rows[br] = the circulants (block column, shift) of block row br: row r meets column (r + s) mod z.  true = H is
// exactly the code compiled into P; fills the per-lane table of the syndrome rounds.
template <class P>
bool qcw_build(const HostTables& t, const std::vector<std::vector<QcBlk>>& rows, std::vector<uint32_t>* syn_tab) {
    constexpr int z = P::Z, NB = P::NB, MB = P::MB;
    if (t.N != NB * z || t.M != MB * z || (int)rows.size() != MB) return false;
    for (int br = 0; br < MB; ++br) {
        if ((int)rows[br].size() != P::cdeg(br)) return false;
        for (int j = 0; j < P::cdeg(br); ++j)
            if (rows[br][j].bc != P::cbc(P::e0(br) + j) || rows[br][j].s != P::csh(P::e0(br) + j)) return false;
    }
    syn_tab->assign((size_t)P::ROUNDS * 32, 0xffffffffu);
    for (int r = 0; r < P::ROUNDS; ++r)
        for (int lane = 0; lane < 32; ++lane) {
            const int br = r * (32 / P::SEG) + lane / P::SEG, j = lane % P::SEG;
            if (br < MB && j < P::cdeg(br)) (*syn_tab)[(size_t)r * 32 + lane] = ((uint32_t)rows[br][j].bc << 8) | (uint32_t)rows[br][j].s;
        }
    return true;
}

template <class P>
int launch_qcw_t(const QcwParams& q, int grid, int warps, cudaStream_t stream) {
    const size_t smem = (size_t)warps * P::WARP_BYTES;
    auto kernel = q.fmt == 0 ? ldpc_ms_qcw_kernel<P, false> : ldpc_ms_qcw_kernel<P, true>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kernel<<<grid, warps * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

#define QCW_PROFILE(C) {C::Z, (int)QcwProfile<C>::WARP_BYTES, &qcw_build<QcwProfile<C>>, &launch_qcw_t<QcwProfile<C>>}
const QcwProfileEntry kTable[] = {
    QCW_PROFILE(QcwCode34B_24), QCW_PROFILE(QcwCode34A_24), QCW_PROFILE(QcwCode23B_24),
    QCW_PROFILE(QcwCode23A_24), QCW_PROFILE(QcwCode12_24),  QCW_PROFILE(QcwCode56_24),
    QCW_PROFILE(QcwCode34B_32), QCW_PROFILE(QcwCode34A_32), QCW_PROFILE(QcwCode23B_32),
    QCW_PROFILE(QcwCode23A_32), QCW_PROFILE(QcwCode12_32),  QCW_PROFILE(QcwCode56_32),
};
#undef QCW_PROFILE

}  // namespace

const QcwProfileEntry* qcw_profiles(int* n) {
    *n = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace ldpc_b200
