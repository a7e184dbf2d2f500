// ldpc_qc_host.h -- host-side table builder of the compiled quasi-cyclic profiles (ldpc_qc.cuh).  A template over
// the profile, instantiated next to the kernel it feeds (k_qc.cu).
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "ldpc_launch.h"
#include "ldpc_qc.cuh"

namespace ldpc_b200 {

template <class P>
bool qc_build(const HostTables& t, const std::vector<std::vector<QcBlk>>& rows, QcParams* out, std::vector<unsigned char>* tab_bytes,
              size_t* smem_out) {
    constexpr int z = P::Z, G = P::G, SUB = 32 / G, W = P::W;
    constexpr uint32_t ROWB = (uint32_t)G * 4u, RS = (uint32_t)(z + SUB) * ROWB;
    static_assert(z % SUB == 0, "a group of node lanes must not straddle blocks");
    if (t.M % z || t.N % z) return false;
    const int MB = t.M / z, NB = t.N / z, gpb = z / SUB;
    if (MB * gpb != P::CS * W || NB * gpb != P::VS * W || (int)rows.size() != MB) return false;
    struct Col { int br, j, s; };
    std::vector<std::vector<Col>> cols(NB);
    for (int br = 0; br < MB; ++br)
        for (int j = 0; j < (int)rows[br].size(); ++j) cols[rows[br][j].bc].push_back({br, j, rows[br][j].s});
    std::vector<int> border(MB), corder(NB);
    for (int i = 0; i < MB; ++i) border[i] = i;
    for (int i = 0; i < NB; ++i) corder[i] = i;
    std::stable_sort(border.begin(), border.end(), [&](int a, int b) { return rows[a].size() > rows[b].size(); });
    std::stable_sort(corder.begin(), corder.end(), [&](int a, int b) { return cols[a].size() > cols[b].size(); });
    // R blocks of a block row: as many as the degree of the slot(s) its groups are processed with (>= its own)
    std::vector<int> dpad(MB, 0), eb0(MB + 1, 0);
    for (int p = 0; p < MB * gpb; ++p) {
        if ((int)rows[border[p / gpb]].size() > P::cdeg(p / W)) return false;
        dpad[border[p / gpb]] = std::max(dpad[border[p / gpb]], P::cdeg(p / W));
    }
    for (int br = 0; br < MB; ++br) eb0[br + 1] = eb0[br] + dpad[br];
    const uint32_t t_bytes = (uint32_t)NB * (z + SUB) * ROWB, r_bytes = (uint32_t)eb0[MB] * RS;
    const uint32_t zero_row = t_bytes + r_bytes, inf_row = zero_row + 128u;
    QcParams& q = *out;
    std::memset(&q, 0, sizeof(q));
    tab_bytes->assign(sizeof(QcWarpTab<P>) * W, 0);
    QcWarpTab<P>* tabs = reinterpret_cast<QcWarpTab<P>*>(tab_bytes->data());
    for (int p = 0; p < MB * gpb; ++p) {
        const int br = border[p / gpb], g = p % gpb, slot = p / W, w = p % W, r0 = g * SUB;
        QcWarpTab<P>& tb = tabs[w];
        tb.cn_r[slot] = t_bytes + (uint32_t)eb0[br] * RS + (uint32_t)(SUB + r0) * ROWB;
        if (g == gpb - 1) tb.cdup |= 1u << slot;
        for (int j = 0; j < P::cdeg(slot); ++j)
            tb.cn_t[QcLayout<P>::coff(slot) + j] = j < (int)rows[br].size()
                ? (uint32_t)(rows[br][j].bc * (z + SUB) + (r0 + rows[br][j].s) % z) * ROWB
                : inf_row;  // padded edge: T = -inf is neutral for the minima, the sign parity and the syndrome
    }
    for (int p = 0; p < NB * gpb; ++p) {
        const int bc = corder[p / gpb], g = p % gpb, slot = p / W, w = p % W, i0 = g * SUB;
        const int d = (int)cols[bc].size();
        if (d > P::vdeg(slot)) return false;
        QcWarpTab<P>& tb = tabs[w];
        tb.vn_t[slot] = (uint32_t)(bc * (z + SUB) + i0) * ROWB;
        tb.var0[slot] = (uint32_t)(bc * z + i0);
        if (g == 0) tb.vdup |= 1u << slot;
        for (int k = 0; k < P::vdeg(slot); ++k) {
            if (k >= d) { tb.vn_r[QcLayout<P>::voff(slot) + k] = zero_row; continue; }
            const Col& cd = cols[bc][k];  // ascending block row = ascending row: the summation order
            int m = ((i0 - cd.s) % z + z) % z;
            if (m > z - SUB) m -= z;      // the group wraps: its first rows are read through the leading pad
            tb.vn_r[QcLayout<P>::voff(slot) + k] = t_bytes + (uint32_t)(eb0[cd.br] + cd.j) * RS + (uint32_t)(SUB + m) * ROWB;
        }
    }
    // the kernel branches once per pass on "this warp owns wrapped rows": all of a warp's groups or none
    for (int w = 0; w < W; ++w) {
        if (tabs[w].cdup != 0u && tabs[w].cdup != (1u << P::CS) - 1u) return false;
        if (tabs[w].vdup != 0u && tabs[w].vdup != (1u << P::VS) - 1u) return false;
    }
    q.N = t.N; q.NB = NB;
    q.t_bytes = t_bytes; q.r_bytes = r_bytes;
    *smem_out = (size_t)t_bytes + r_bytes + 256;
    return true;
}

}  // namespace ldpc_b200
