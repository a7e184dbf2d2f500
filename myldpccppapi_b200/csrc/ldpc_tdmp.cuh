// ldpc_tdmp.cuh -- layered (TDMP) min-sum decoder for sm_100a, on the group layout of ldpc_kernels.cuh.
//
// What the reference's DecodeTDMP path intends (host loop Coder::decodeOnceTDMP, MyLdpc.cpp:889-976; kernels
// decodeInitTDMP / refreshQTDMP / refreshRTDMP / refreshPostPTDMP / hardDecisionTDMP, decodeCL.c:203-292):
// the rows are swept in layers of z consecutive rows (one block row of the quasi-cyclic matrix).  Inside a
// layer, for every edge:  Q = P[col] - R_old;  R_new = +-min(1000, min of the row's other |Q|), sign = xor of
// the other (Q < 0);  P[col] = Q + R_new.  After the last layer: hard decision (P > 0 -> 0, P < 0 -> 1,
// P == 0 keeps the previous bit), syndrome check, ++time, stop when clean or time == cap.
//
// Mapping.  A CTA decodes G codewords; a warp instruction covers SUB = 32/G rows x G codewords and the CTA's
// W = z/SUB warps cover exactly one layer, so a layer is one straight-line check per thread followed by one
// __syncthreads.  No two rows of a layer share a column (checked on the host), so the posterior update is
// a plain scatter store: there is no variable-node pass at all.  State in shared memory:
//   T[v][G]     negated posterior -P (natural column order: consecutive rows of a circulant touch consecutive
//               columns, so the gathers are bank-conflict free by structure)
//   R[row][32]  the message of every edge, rows = (warp, layer, j), private to the thread that owns the check
//   tables      T byte offsets, quads [layer][j/4][h][4], padded with per-h dummy rows holding -inf
// Per edge and iteration: T gather, R load, R store, T scatter (+ one T gather in the syndrome sweep).
//
// Zeros.  S = T + R_old = -Q; (Q < 0) is taken as !signbit(S).  That is wrong only when Q is exactly zero,
// and then every other message of the row has magnitude zero, so only signs of zeros can differ from the
// reference; all non-zero values are identical and the hard decision tests P == 0 explicitly (keeping the
// previous bit, which the owning thread holds in a register mask).  After the decision a zero posterior is
// rewritten as +-0 with the sign that encodes its bit, so that syndrome and output read sign bits only.
#pragma once

#include "ldpc_kernels.cuh"

namespace ldpc_b200 {

constexpr int kTdmpMaxLayers = 16;

struct TdmpParams {
    const uint32_t* cn_tab;  // [warp][layer][quad][h][4] byte offsets into T
    int M, N, K, W, L, VS;
    int cn_stride;           // table words per warp
    int r_rows_per_warp;     // sum of the layers' slot degrees
    int max_iter, early_term;
    const float* llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;
    unsigned long long* counter64;
    uint8_t ldeg[kTdmpMaxLayers];  // slot degree (largest row degree) of each layer
};

// One row of exact slot degree D: the check-node arithmetic of grp_check plus the posterior scatter.
template <int D, int SUB>
__device__ __forceinline__ void tdmp_check(uint32_t tab, uint32_t rrow, uint32_t c4, int h, bool act) {
    constexpr int NQ = (D + 3) / 4;
    uint32_t ent[NQ * 4];
#pragma unroll
    for (int jq = 0; jq < NQ; ++jq) {
        const uint4 o = lds_u128(tab + (uint32_t)(jq * SUB + h) * 16u);
        ent[jq * 4 + 0] = o.x + c4; ent[jq * 4 + 1] = o.y + c4; ent[jq * 4 + 2] = o.z + c4; ent[jq * 4 + 3] = o.w + c4;
    }
    float S[D];
    {
        float tv[D];
#pragma unroll
        for (int j = 0; j < D; ++j) tv[j] = lds_f32(ent[j]);
#pragma unroll
        for (int j = 0; j < D; ++j) S[j] = lds_f32(rrow + (uint32_t)j * 128u);
#pragma unroll
        for (int j = 0; j < D; ++j) S[j] = __fadd_rn(tv[j], S[j]);  // = -Q_j = -(P - R_old)
    }
    uint32_t px = 0u;
#pragma unroll
    for (int j = 0; j + 1 < D; j += 2) px = px ^ __float_as_uint(S[j]) ^ __float_as_uint(S[j + 1]);
    if (D & 1) px ^= __float_as_uint(S[D - 1]);
    float rn[D];
    ms_new_messages<D>(S, px, rn);   // (Q_j < 0) = !signbit(S_j); sign(R_j) = xor of the others = total ^ own
#pragma unroll
    for (int j = 0; j < D; ++j) {
        if (act) {
            sts_f32(rrow + (uint32_t)j * 128u, rn[j]);
            sts_f32(ent[j], __fsub_rn(S[j], rn[j]));  // -(Q + R_new)
        }
    }
}

template <int G, int MAX_THREADS>
__global__ void __launch_bounds__(MAX_THREADS, 1) ldpc_tdmp_group_kernel(const __grid_constant__ TdmpParams p) {
    constexpr int SUB = 32 / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_flag[2][32];
    __shared__ long long s_cw[32];

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int c = lane & (G - 1), h = lane / G;
    const int W = p.W, L = p.L, VS = p.VS;
    const int NL = W * SUB;                 // = layer height z
    const int RD = W * p.r_rows_per_warp;

    const uint32_t t_base = smem_u32(smem_raw);
    const uint32_t t_bytes = ((uint32_t)(p.N + SUB) * G * 4 + 127u) & ~127u;  // N columns + SUB dummy rows (-inf)
    const uint32_t r_base = t_base + t_bytes;
    const uint32_t cn_base = r_base + (uint32_t)RD * 128;
    {
        uint32_t* cn_s = reinterpret_cast<uint32_t*>(smem_raw + (cn_base - t_base));
        for (int i = threadIdx.x; i < W * p.cn_stride; i += blockDim.x) cn_s[i] = __ldg(p.cn_tab + i) + t_base;
    }
    const uint32_t c4 = (uint32_t)c * 4u;
    if (warp == 0) {
        sts_f32(t_base + (uint32_t)p.N * G * 4 + lane * 4, -INFINITY);  // SUB * G = 32 dummy elements
        s_flag[0][lane] = 0u; s_flag[1][lane] = 0u;
    }
    const uint32_t t_own = t_base + (uint32_t)threadIdx.x * 4u;  // column (s*NL + warp*SUB + h), codeword c
    const uint32_t t_stride = (uint32_t)NL * G * 4u;
    const uint32_t r_own = r_base + (uint32_t)warp * p.r_rows_per_warp * 128u + (uint32_t)lane * 4u;
    const uint32_t cn_w = cn_base + (uint32_t)warp * p.cn_stride * 4u;

    uint32_t bits = 0u;  // hard decisions of this thread's VS <= 32 columns (needed for the P == 0 rule)
    long long cw = -1;
    bool live = false, done = false, loading = false;
    int it = 0, my_iters = 0;

    auto layer_sweep = [&](bool act) {
        uint32_t tab = cn_w, rrow = r_own;
        for (int l = 0; l < L; ++l) {
            const int d = p.ldeg[l];
#define TDMP_CASE(D) case D: tdmp_check<D, SUB>(tab, rrow, c4, h, act); break;
            switch (d) {
                TDMP_CASE(1) TDMP_CASE(2) TDMP_CASE(3) TDMP_CASE(4) TDMP_CASE(5) TDMP_CASE(6) TDMP_CASE(7)
                TDMP_CASE(8) TDMP_CASE(9) TDMP_CASE(10) TDMP_CASE(11) TDMP_CASE(12) TDMP_CASE(13) TDMP_CASE(14)
                TDMP_CASE(15) TDMP_CASE(16) TDMP_CASE(17) TDMP_CASE(18) TDMP_CASE(19) TDMP_CASE(20)
                default: break;
            }
#undef TDMP_CASE
            rrow += (uint32_t)d * 128u;
            tab += 16u * (uint32_t)(((d + 3) >> 2) * SUB);
            __syncthreads();  // the next layer reads the posteriors this one wrote
        }
    };
    auto decide = [&](bool act) {  // hardDecisionTDMP, decodeCL.c:261-281
        for (int s = 0; s < VS; ++s) {
            const uint32_t a = t_own + (uint32_t)s * t_stride;
            const float t = lds_f32(a);
            const uint32_t prev = (bits >> s) & 1u;
            const uint32_t nb = t < 0.0f ? 0u : (t > 0.0f ? 1u : prev);  // T = -P
            if (act) {
                bits = (bits & ~(1u << s)) | (nb << s);
                if (t == 0.0f) sts_f32(a, nb ? 0.0f : -0.0f);  // bit = !signbit(T) from here on
            }
        }
    };
    auto syndrome = [&]() -> uint32_t {  // checkResult over this thread's L rows
        uint32_t unsat = 0u, tab = cn_w + (uint32_t)h * 16u;
        for (int l = 0; l < L; ++l) {
            const int nq = (p.ldeg[l] + 3) >> 2;
            uint32_t sx = 0u;
            for (int jq = 0; jq < nq; ++jq) {
                const uint4 o = lds_u128(tab);
                tab += (uint32_t)SUB * 16u;
                // padded entries read -inf (sign bit set = bit 0); a quad has an even number of entries
                sx ^= __float_as_uint(lds_f32(o.x + c4)) ^ __float_as_uint(lds_f32(o.y + c4));
                sx ^= __float_as_uint(lds_f32(o.z + c4)) ^ __float_as_uint(lds_f32(o.w + c4));
            }
            unsat |= sx >> 31;
        }
        return unsat;
    };
    auto emit = [&](bool sel) {  // toChar, decodeCL.c:188-199
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = warp * SUB + h; b < KB; b += NL) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.K) v |= ((~__float_as_uint(lds_f32(t_base + (uint32_t)n * (G * 4) + c4))) >> 31) << t;
                }
                if (sel) p.info[(size_t)cw * KB + b] = (uint8_t)v;
            }
        }
        if (p.hard) {
            const int NB = (p.N + 7) >> 3;
            for (int b = warp * SUB + h; b < NB; b += NL) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.N) v |= ((~__float_as_uint(lds_f32(t_base + (uint32_t)n * (G * 4) + c4))) >> 31) << t;
                }
                if (sel) p.hard[(size_t)cw * NB + b] = (uint8_t)v;
            }
        }
        if (p.post && sel) {
            for (int s = 0; s < VS; ++s)
                p.post[(size_t)cw * p.N + s * NL + warp * SUB + h] = -lds_f32(t_own + (uint32_t)s * t_stride);
        }
        if (p.iters && warp == 0 && h == 0 && sel) p.iters[cw] = my_iters;
    };
    auto fetch = [&](bool want) {
        if (warp == 0 && h == 0 && want) s_cw[c] = (long long)atomicAdd(p.counter64, 1ull);
        __syncthreads();
        if (want) {
            live = false; done = false;
            cw = s_cw[c];
            if (cw < p.ncw) {
                loading = true;
                const float* src = p.llr + (size_t)cw * p.N + warp * SUB + h;
                for (int s = 0; s < VS; ++s)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(t_own + (uint32_t)s * t_stride), "l"(src + s * NL) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto start_loaded = [&]() {  // decodeInitTDMP: P = y, R = 0
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (loading) {
            for (int s = 0; s < VS; ++s) {
                const uint32_t a = t_own + (uint32_t)s * t_stride;
                sts_f32(a, -lds_f32(a));
            }
            for (int r = 0; r < p.r_rows_per_warp; ++r) sts_f32(r_own + (uint32_t)r * 128u, 0.0f);
            bits = 0u;
            loading = false; live = true; done = false; it = 0;
        }
        __syncthreads();
    };

    __syncthreads();
    fetch(true);
    uint32_t ph = 0;
    for (;;) {
        if (__any_sync(0xffffffffu, loading)) start_loaded();
        const bool retire = live && done;
        if (__any_sync(0xffffffffu, retire)) {
            emit(retire);
            fetch(retire);
            if (__any_sync(0xffffffffu, loading)) start_loaded();
        }
        if (!__any_sync(0xffffffffu, live || loading)) break;

        const bool act = live && !done;
        layer_sweep(act);
        decide(act);
        __syncthreads();
        if (act) ++it;
        if (p.early_term) {
            const uint32_t unsat = syndrome();
            if (act && unsat) s_flag[ph & 1][c] = 1u;
            __syncthreads();
            if (act && s_flag[ph & 1][c] == 0u) { done = true; my_iters = it; }
            if (warp == 0) s_flag[(ph + 1) & 1][lane] = 0u;
            ++ph;
        }
        if (act && !done && it == p.max_iter) { done = true; my_iters = it; }
    }
}

}  // namespace ldpc_b200
