// ldpc_spq.cuh -- probability-domain sum-product (decodeType DecodeSP) for quasi-cyclic codes of ANY block size
// z <= 96: the layout of ldpc_qcm.cuh (a codeword per group of ceil(z / 32) warps, lane = row in the check pass /
// column in the variable pass, warp-uniform offsets from __constant__ memory) with the arithmetic of ldpc_sp.cuh.
//
// The on-chip sum-product kernel (ldpc_sp.cuh) needs the group layout with 8 or 16 codewords per CTA, which only the
// shortest codes of the reference's family fit (N = 576 .. 768); every other block size ran the any-size kernel
// (ldpc_big.cuh: messages in a global workspace, correctness first) at 0.2 Gbit/s.  Here a codeword's slice of shared
// memory holds
//   HB[24][2z]  its hard decisions, one BYTE per variable (a word each cost a codeword in flight per SM from z = 56 on),
//               every block column stored twice (the cyclic wrap of the check pass is a plain offset) -- the place of
//               T in ldpc_qcm.cuh, the same table entries divided by four
//   E [E ][z]   one float per edge that alternates meaning as in ldpc_sp.cuh: after a variable pass E_e = q0_e - q1_e,
//               after a check pass E_e = d_e = product over the row's OTHER edges of (q0 - q1), in row-list order
// and the tables, the geometry and the work queue are those of the group-of-warps min-sum kernel (QcmTab / QcmParams).
// Restates decodeInit / refreshR / hardDecision / checkResult / refreshQ (decodeCL.c:3-108) under Coder::decodeOnceSP's
// loop (MyLdpc.cpp:977-1059): every product in the reference's list order (ascending column in a row, ascending row in
// a column) with one IEEE fp32 rounding per operation, IEEE divisions, exp(8y) by sp_expf -- operation for operation
// what ldpc_sp.cuh does, so the two kernels and the pinned oracle agree bit for bit.
#pragma once
#include "ldpc_qcm.cuh"
#include "ldpc_sp.cuh"

namespace ldpc_b200 {

#ifdef LDPC_SPQ_DEVICE   // the kernel and its table bank: only the unit that instantiates them (k_spq.cu)
static __constant__ uint4 g_spq_bank[kQcTabSlots][kQcmBankBytes / 16];

__device__ __forceinline__ uint32_t spq_lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void spq_sts_u8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void spq_atoms_or(uint32_t a, uint32_t v) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// One block row: lane = row.  refreshR (decodeCL.c:25-41): d_j = ((((x_0 x_1) ...) x_{j-1}) x_{j+1}) ... x_{D-1}, the
// reference's left-to-right product with edge j skipped -- its first j factors are the running prefix shared by all
// later edges.  Returns the row's syndrome bit over the hard decisions of the previous variable pass (checkResult).
template <class R, int I>
__device__ __forceinline__ uint32_t spq_check(const QcmTab<R>& tb, uint32_t la, uint32_t lb, uint32_t eb, uint32_t zb, bool act) {
    constexpr int D = R::cdeg(I);
    uint32_t syn = 0u;
#pragma unroll
    for (int j = 0; j < D; j += 2) {  // two warp-uniform bases per LDCU.64 (byte offsets of the float layout: / 4 here)
        const uint2 e = *reinterpret_cast<const uint2*>(tb.cn_t + QcmLayout<R>::coff(I) + j);
        syn ^= spq_lds_u8(lb + (e.x >> 2));
        if (j + 1 < D) syn ^= spq_lds_u8(lb + (e.y >> 2));
    }
    const uint32_t r0 = la + eb + (uint32_t)R::e0(I) * zb;   // this lane's row of the block row's first circulant
    float x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = lds_f32(r0 + (uint32_t)j * zb);
    float pre = 1.0f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        float d = pre;
#pragma unroll
        for (int k = j + 1; k < D; ++k) d = __fmul_rn(d, x[k]);
        if (act) sts_f32(r0 + (uint32_t)j * zb, d);
        pre = __fmul_rn(pre, x[j]);  // 1 * x_0 = x_0 exactly
    }
    return syn & 1u;
}

template <class R, int I>
__device__ __forceinline__ uint32_t spq_cn(const QcmTab<R>& tb, uint32_t la, uint32_t lb, uint32_t eb, uint32_t zb, bool act) {
    if constexpr (I < R::MB) {
        const uint32_t u = spq_check<R, I>(tb, la, lb, eb, zb, act);
        return u | spq_cn<R, I + 1>(tb, la, lb, eb, zb, act);
    } else {
        return 0u;
    }
}

// One block column: lane = column c, channel term t = exp(8y).  hardDecision (decodeCL.c:64-86) and refreshQ
// (decodeCL.c:43-62); edge k's message sits at row (c - s) mod z of its circulant.  `bits` bit B = this column's hard
// decision of block column B.
template <class R, int B>
__device__ __forceinline__ void spq_vn(const QcmTab<R>& tb, uint32_t la, uint32_t laz, uint32_t lb, uint32_t c, uint32_t z, const float* tt,
                                       uint32_t& bits, bool act) {
    if constexpr (B < R::NB) {
        constexpr int D = R::vdeg(B), V0 = R::v0(B);
        uint32_t a[D];
        float r0[D], r1[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const uint2 u = tb.vn[V0 + k];
            a[k] = (c < u.y ? laz : la) + u.x;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float d = lds_f32(a[k]);
            r0[k] = __fmul_rn(__fadd_rn(1.0f, d), 0.5f);  // x / 2 == x * 0.5 for every float
            r1[k] = __fmul_rn(__fsub_rn(1.0f, d), 0.5f);
        }
        const float t = tt[B];
        const float one_t = __fadd_rn(1.0f, t);
        const float p0 = __fdiv_rn(t, one_t), p1 = __fdiv_rn(1.0f, one_t);  // priorP0 / priorP1, decodeCL.c:13-18
        float pu0 = p0, pu1 = p1;  // running prefix p * r[0] ... r[k-1]
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float u0 = pu0, u1 = pu1;
#pragma unroll
            for (int m = k + 1; m < D; ++m) { u0 = __fmul_rn(u0, r0[m]); u1 = __fmul_rn(u1, r1[m]); }
            const float s = __fadd_rn(u0, u1);
            const float xq = __fsub_rn(__fdiv_rn(u0, s), __fdiv_rn(u1, s));
            if (act) sts_f32(a[k], xq);
            pu0 = __fmul_rn(pu0, r0[k]);
            pu1 = __fmul_rn(pu1, r1[k]);
        }
        // after the last edge the prefixes ARE hardDecision's products over the whole column (same factors, same order)
        const uint32_t prev = (bits >> B) & 1u;
        const uint32_t bit = (pu0 > pu1) ? 0u : ((pu0 < pu1) ? 1u : prev);
        if (act) {
            bits = (bits & ~(1u << B)) | (bit << B);
            spq_sts_u8(lb + (uint32_t)(2 * B) * z, bit);
            spq_sts_u8(lb + (uint32_t)(2 * B + 1) * z, bit);
        }
        spq_vn<R, B + 1>(tb, la, laz, lb, c, z, tt, bits, act);
    }
}

template <class R>
__global__ void __launch_bounds__(kQcmMaxWarps * 32, 1) ldpc_sp_qcm_kernel(const __grid_constant__ QcmParams p) {
    constexpr int NB = R::NB;
    static_assert(sizeof(QcmTab<R>) <= kQcmBankBytes, "profile tables exceed a bank slot");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_word[kQcmMaxWarps];
    __shared__ uint32_t s_flag[kQcmMaxWarps][2];

    const uint32_t lane = threadIdx.x & 31u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int ws = warp / p.NW, sw = warp - ws * p.NW;
    const QcmTab<R>& tb = *reinterpret_cast<const QcmTab<R>*>(&g_spq_bank[p.tab_slot][0]);
    const uint32_t wb = smem_u32(smem_raw) + (uint32_t)ws * p.word_bytes;
    const uint32_t c = (uint32_t)(sw * p.RW) + lane;
    const bool act = lane < (uint32_t)p.RW && c < (uint32_t)p.z;
    const uint32_t zb = p.zb;
    const uint32_t la = wb + c * 4u;                      // this lane's row of a float array that starts at the slice base
    const uint32_t lb = wb + c;                           // ... of the byte array HB
    // the table entries of the variable pass are offsets in the min-sum layout (messages behind p.t_bytes of posteriors):
    // here E starts behind p.hb_bytes of hard decisions
    const uint32_t lav = la + p.hb_bytes - p.t_bytes, laz = lav + zb;
    const uint32_t bitbuf = wb + p.bits_off;
    const int KB = (p.K + 7) >> 3, NB8 = (p.N + 7) >> 3;
    const int gl = sw * 32 + (int)lane, gn = p.NW * 32;
    const bool single = p.NW == 1;
    auto gsync = [&]() {
        if (single) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(ws + 1), "r"(gn) : "memory");
    };
    auto claim = [&]() -> long long { return (sw == 0 && lane == 0) ? (long long)atomicAdd(p.counter64, 1ull) : 0ll; };
    auto prefetch_y = [&](long long w) {
        const char* src = reinterpret_cast<const char*>(p.llr + (size_t)w * p.N);
        for (int o0 = 0; o0 < p.N * 4; o0 += 32 * 128) {
            const int o = o0 + (int)lane * 128;
            if (o < p.N * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
        }
    };
    auto landed = [&](long long w) -> bool {
        return w < p.ncw && (!p.avail || qc_wait_input(p.avail, w, true, p.status, p.wait_ns));
    };

    float tt[NB];
    long long wn = 0, tick = 0;
    if (sw == 0) {
        if (lane == 0) { s_flag[ws][0] = 0u; s_flag[ws][1] = 0u; }
        wn = __shfl_sync(0xffffffffu, claim(), 0);
        tick = claim();
        if (!landed(wn)) wn = p.ncw;
        if (wn < p.ncw) prefetch_y(wn);
    }
    for (;;) {
        if (sw == 0) {
            if (lane == 0) s_word[ws] = wn;
            for (int o0 = 0; o0 < NB8 + 4; o0 += 128) {
                const int o = o0 + (int)lane * 4;
                if (o < NB8 + 4) qc_sts_u32(bitbuf + (uint32_t)o, 0u);
            }
        }
        gsync();
        const long long w = s_word[ws];
        if (w >= p.ncw) break;
        // ---- start word w (decodeInit, decodeCL.c:3-22): t = exp(8y), E_e = q0 - q1 on every edge of the column, bits 0
        {
            const float* src = p.llr + (size_t)w * p.N + c;
#pragma unroll
            for (int b = 0; b < NB; ++b) tt[b] = act ? __ldg(src + b * p.z) : 0.0f;
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) tt[b] = sp_expf(__fmul_rn(8.0f, tt[b]));
        if (act) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                spq_sts_u8(lb + (uint32_t)(2 * b * p.z), 0u);
                spq_sts_u8(lb + (uint32_t)((2 * b + 1) * p.z), 0u);
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float one_t = __fadd_rn(1.0f, tt[b]);
            const float x0 = __fsub_rn(__fdiv_rn(tt[b], one_t), __fdiv_rn(1.0f, one_t));
            for (int k = 0; k < R::vdeg(b); ++k) {
                const uint2 u = tb.vn[R::v0(b) + k];
                if (act) sts_f32((c < u.y ? laz : lav) + u.x, x0);
            }
        }
        uint32_t bits = 0u;
        if (sw == 0) {   // the next word's values travel while this one is decoded
            wn = __shfl_sync(0xffffffffu, tick, 0);
            if (!landed(wn)) wn = p.ncw;
            if (wn < p.ncw) prefetch_y(wn);
            tick = claim();
        }
        gsync();

        int it = 0;
        uint32_t ph = 0u;
        for (;;) {
            // refreshR + checkResult of the previous hard decision
            const uint32_t unsat = spq_cn<R, 0>(tb, la, lb, p.hb_bytes, zb, act);
            const bool check = p.early_term && it >= 1;
            if (check && __any_sync(0xffffffffu, act && unsat != 0u) && lane == 0) s_flag[ws][ph] = 1u;  // same-value race, benign
            gsync();
            if (check && s_flag[ws][ph] == 0u) break;            // the hard decisions of iteration `it` satisfy every check
            if (sw == 0 && lane == 0) s_flag[ws][ph ^ 1u] = 0u;
            ph ^= 1u;
            spq_vn<R, 0>(tb, lav, laz, lb, c, (uint32_t)p.z, tt, bits, act);     // hardDecision + refreshQ
            gsync();
            ++it;
            if (it >= p.max_iter) break;
        }

        // ---- word w leaves (toChar, decodeCL.c:188-199): a warp ballots its columns of every block column into the
        // word's bit buffer, then the group copies the bytes out
        for (int b = 0; b < NB; ++b) {
            const uint32_t bal = __ballot_sync(0xffffffffu, act && ((bits >> b) & 1u) != 0u);
            if (lane == 0) {
                const uint32_t g = (uint32_t)(b * p.z + sw * p.RW), sh = g & 31u;
                spq_atoms_or(bitbuf + (g >> 5) * 4u, bal << sh);
                if (sh) spq_atoms_or(bitbuf + (g >> 5) * 4u + 4u, bal >> (32u - sh));
            }
        }
        gsync();
        {
            const int nby = p.hard ? NB8 : KB;
            for (int b0 = 0; b0 < nby; b0 += gn) {
                const int b = b0 + gl;
                if (b < nby) {
                    const uint32_t v = (qc_lds_u32(bitbuf + (uint32_t)(b & ~3)) >> (8 * (b & 3))) & 0xffu;
                    if (p.hard) p.hard[(size_t)w * NB8 + b] = (uint8_t)v;
                    if (p.info && b < KB) {
                        const uint32_t keep = (b == KB - 1 && (p.K & 7)) ? ((1u << (p.K & 7)) - 1u) : 0xffu;
                        p.info[(size_t)w * KB + b] = (uint8_t)(v & keep);
                    }
                }
            }
        }
        if (p.iters && gl == 0) p.iters[w] = it;
        gsync();
    }
}

#endif  // LDPC_SPQ_DEVICE

}  // namespace ldpc_b200
