// k_qcw.cu -- the warp-per-codeword quasi-cyclic kernel (ldpc_qcw.cuh): one instantiation per 802.16e rate at z = 24,
// the host-side table builder, and this unit's __constant__ table bank.
#include <cstring>

#include "ldpc_launch.h"
#define LDPC_QCW_DEVICE
#include "ldpc_qcw.cuh"

namespace ldpc_b200 {
namespace {

// rows[br] = the circulants (block column, shift) of block row br: row r meets column (r + s) mod z.
template <class P>
bool qcw_build(const HostTables& t, const std::vector<std::vector<QcBlk>>& rows, std::vector<unsigned char>* tab_bytes,
               std::vector<uint32_t>* syn_tab) {
    constexpr int z = P::Z, NB = P::NB, MB = P::MB;
    if (t.N != NB * z || t.M != MB * z || (int)rows.size() != MB) return false;
    struct Col { int br, j, s; };
    std::vector<std::vector<Col>> cols(NB);
    for (int br = 0; br < MB; ++br) {
        if ((int)rows[br].size() != P::cdeg(br)) return false;
        for (int j = 0; j < (int)rows[br].size(); ++j) cols[rows[br][j].bc].push_back({br, j, rows[br][j].s});
    }
    for (int b = 0; b < NB; ++b)
        if ((int)cols[b].size() != P::vdeg(b)) return false;
    tab_bytes->assign(sizeof(QcwTab<P>), 0);
    QcwTab<P>& tab = *reinterpret_cast<QcwTab<P>*>(tab_bytes->data());
    for (int br = 0; br < MB; ++br)
        for (int j = 0; j < P::cdeg(br); ++j)
            tab.cn_t[P::coff(br) + j] = (uint32_t)(rows[br][j].bc * 2 * z + rows[br][j].s) * 4u;
    for (int b = 0; b < NB; ++b)
        for (int k = 0; k < P::vdeg(b); ++k) {   // ascending block row = ascending row: the summation order
            const Col& c = cols[b][k];
            tab.vn[P::voff(b) + k] = make_uint2(P::T_BYTES + (uint32_t)(P::e0(c.br) + c.j) * z * 4u - (uint32_t)c.s * 4u, (uint32_t)c.s);
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
    cudaError_t e = cudaFuncSetAttribute(ldpc_ms_qcw_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ldpc_ms_qcw_kernel<P><<<grid, warps * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

int upload_qcw_bank(int slot, const void* tab, size_t bytes) {
    if (slot < 0 || slot >= kQcTabSlots || bytes > (size_t)kQcwBankBytes) return (int)cudaErrorInvalidValue;
    return (int)cudaMemcpyToSymbol(g_qcw_bank, tab, bytes, (size_t)slot * kQcwBankBytes, cudaMemcpyHostToDevice);
}

#define QCW_PROFILE(R, Z) {Z, (int)QcwProfile<R, Z>::WARP_BYTES, &qcw_build<QcwProfile<R, Z>>, &launch_qcw_t<QcwProfile<R, Z>>, &upload_qcw_bank}
const QcwProfileEntry kTable[] = {
    QCW_PROFILE(QcwRate34B, 24), QCW_PROFILE(QcwRate34A, 24), QCW_PROFILE(QcwRate23B, 24),
    QCW_PROFILE(QcwRate23A, 24), QCW_PROFILE(QcwRate12, 24),  QCW_PROFILE(QcwRate56, 24),
};
#undef QCW_PROFILE

}  // namespace

const QcwProfileEntry* qcw_profiles(int* n) {
    *n = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace ldpc_b200
