// k_qcm.cu -- the group-of-warps-per-codeword quasi-cyclic kernel (ldpc_qcm.cuh): one instantiation per 802.16e degree
// profile (any z, any shifts), the host-side table builder, and this unit's __constant__ table bank.
#include <cstring>

#include "ldpc_launch.h"
#define LDPC_QCM_DEVICE
#include "ldpc_qcm.cuh"

namespace ldpc_b200 {
namespace {

// rows[br] = the circulants (block column, shift) of block row br in ascending column order: row r meets column
// (r + s) mod z.  true = the code has R's degree sequences; fills the table block and the launch geometry.
template <class R>
bool qcm_build(const HostTables& t, int z, const std::vector<std::vector<QcBlk>>& rows, size_t smem_limit, QcmParams* out,
               std::vector<unsigned char>* tab_bytes, int* groups_out) {
    constexpr int NB = R::NB, MB = R::MB;
    if (z < 1 || z > 96 || t.N != NB * z || t.M != MB * z || (int)rows.size() != MB) return false;
    struct Col { int br, j, s; };
    std::vector<std::vector<Col>> cols(NB);
    for (int br = 0; br < MB; ++br) {
        if ((int)rows[br].size() != R::cdeg(br)) return false;
        for (int j = 0; j < (int)rows[br].size(); ++j) cols[rows[br][j].bc].push_back({br, j, rows[br][j].s});
    }
    for (int b = 0; b < NB; ++b)
        if ((int)cols[b].size() != R::vdeg(b)) return false;
    const uint32_t zb = (uint32_t)z * 4u, t_bytes = (uint32_t)NB * 2u * zb;
    tab_bytes->assign(sizeof(QcmTab<R>), 0);
    QcmTab<R>& tab = *reinterpret_cast<QcmTab<R>*>(tab_bytes->data());
    for (int br = 0; br < MB; ++br)
        for (int j = 0; j < R::cdeg(br); ++j)
            tab.cn_t[QcmLayout<R>::coff(br) + j] = (uint32_t)(rows[br][j].bc * 2 * z + rows[br][j].s) * 4u;
    for (int b = 0; b < NB; ++b)
        for (int k = 0; k < R::vdeg(b); ++k) {   // ascending block row = ascending row: the summation order
            const Col& c = cols[b][k];
            tab.vn[R::v0(b) + k] = make_uint2(t_bytes + (uint32_t)(R::e0(c.br) + c.j) * zb - (uint32_t)c.s * 4u, (uint32_t)c.s);
        }
    QcmParams& q = *out;
    std::memset(&q, 0, sizeof(q));
    q.z = z; q.NW = (z + 31) / 32; q.RW = (z + q.NW - 1) / q.NW;
    q.zb = zb; q.t_bytes = t_bytes;
    q.N = t.N;
    // [T | R | 128 B slack (idle lanes read past the last circulant) | bit buffer]
    q.bits_off = t_bytes + (uint32_t)R::E * zb + 128u;
    q.word_bytes = (q.bits_off + (uint32_t)(((t.N + 7) / 8 + 3) & ~3) + 4u + 15u) & ~15u;
    int groups = (int)(smem_limit / q.word_bytes);
    groups = std::min(groups, kQcmMaxWarps / q.NW);
    if (q.NW > 1) groups = std::min(groups, kQcmMaxGroups);
    if (groups < 2) return false;
    *groups_out = groups;
    q.PK = 1;
    return true;
}

// Geometry of ldpc_ms_qcm_multi_kernel for `pack` codewords per group (same tables, same slice layout).
// false = that many do not fit.
bool qcm_multi_geometry(const QcmParams& single, int pack, size_t smem_limit, QcmParams* out, int* groups_out) {
    if (pack < 2 || pack > kQcmMaxPack) return false;
    QcmParams q = single;
    q.PK = pack;
    q.NW = (pack * q.z + 31) / 32;
    q.RW = (pack * q.z + q.NW - 1) / q.NW;
    if (q.NW < 2 || q.NW > kQcmMaxWarps) return false;
    // consecutive slices z words apart modulo the 32 banks: the lanes of a warp that straddles two codewords
    // (rows z-1-i .. z-1 of one, 0 .. j of the next) then hit consecutive banks as the lanes of one codeword do
    if (q.z % 4 == 0) q.word_bytes += ((uint32_t)(q.z * 4) % 128u + 128u - q.word_bytes % 128u) % 128u;
    if (smem_limit < 1024) return false;
    int groups = (int)((smem_limit - 1024) / ((size_t)pack * q.word_bytes));   // (the kernel's static shared memory counts too)
    groups = std::min(groups, kQcmMaxWarps / q.NW);
    groups = std::min(groups, kQcmMaxGroups);
    if (groups < 1) return false;
    *out = q;
    *groups_out = groups;
    return true;
}

template <class R>
int launch_qcm_multi_t(const QcmParams& q, int grid, int groups, cudaStream_t stream) {
    const size_t smem = (size_t)groups * q.PK * q.word_bytes;
    auto kernel = ldpc_ms_qcm_multi_kernel<R>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kernel<<<grid, groups * q.NW * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

template <class R>
int launch_qcm_t(const QcmParams& q, int grid, int groups, cudaStream_t stream) {
    const size_t smem = (size_t)groups * q.word_bytes;
    auto kernel = q.fmt == 0 ? ldpc_ms_qcm_kernel<R, false> : ldpc_ms_qcm_kernel<R, true>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kernel<<<grid, groups * q.NW * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

int upload_qcm_bank(int slot, const void* tab, size_t bytes) {
    if (slot < 0 || slot >= kQcTabSlots || bytes > (size_t)kQcmBankBytes) return (int)cudaErrorInvalidValue;
    return (int)cudaMemcpyToSymbol(g_qcm_bank, tab, bytes, (size_t)slot * kQcmBankBytes, cudaMemcpyHostToDevice);
}

// (the degree sequences of a rate do not depend on z: the z = 24 code tables carry them)
#define QCM_PROFILE(C) {&qcm_build<C>, &launch_qcm_t<C>, &upload_qcm_bank, &qcm_multi_geometry, &launch_qcm_multi_t<C>}
const QcmProfileEntry kTable[] = {
    QCM_PROFILE(QcwCode34B_24), QCM_PROFILE(QcwCode34A_24), QCM_PROFILE(QcwCode23B_24),
    QCM_PROFILE(QcwCode23A_24), QCM_PROFILE(QcwCode12_24),  QCM_PROFILE(QcwCode56_24),
};
#undef QCM_PROFILE

}  // namespace

const QcmProfileEntry* qcm_profiles(int* n) {
    *n = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace ldpc_b200
