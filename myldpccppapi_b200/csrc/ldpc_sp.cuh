// ldpc_sp.cuh -- probability-domain sum-product decoder (decodeType DecodeSP) on the group layout.
//
// Restates on sm_100a what the reference does with its OpenCL kernels decodeInit / refreshR /
// hardDecision / checkResult / refreshQ (decodeCL.c:3-108) under Coder::decodeOnceSP's host loop
// (MyLdpc.cpp:977-1059).  Same data layout and control flow as ldpc_ms_group_kernel (G codewords per
// CTA, lane = (node lane h, codeword c), lane refill), with one per-edge array E in shared memory that
// alternates meaning:  after a variable pass  E_e = q0_e - q1_e,  after a check pass  E_e = d_e =
// product over the row's OTHER edges of (q0 - q1), in row-list order.  From d: r0 = (1+d)/2, r1 = (1-d)/2.
//   check pass     d_e for every edge of the row (O(deg^2) products, the reference's own order) and the
//                  syndrome of the hard bits of the previous iteration (checkResult folded in)
//   variable pass  hardDecision: t0 = prior0 * prod r0, t1 = prior1 * prod r1 over the column list,
//                  bit = 0 if t0 > t1, 1 if t0 < t1, unchanged on a tie;  refreshQ: per edge the same
//                  products without that edge, q0 = u0/(u0+u1), q1 = u1/(u0+u1), E_e = q0 - q1
// Every product is taken in the reference's list order with one IEEE fp32 rounding per operation
// (--fmad=false), divisions are IEEE; exp(8y) is sp_expf -- the reference's exp is an OpenCL built-in
// with implementation-defined rounding, so oracle and kernel share one fixed fp32 routine
// (the test checker restates it operation for operation) and are compared bit for bit.
#pragma once
#include "ldpc_kernels.cuh"

namespace ldpc_b200 {

// fixed operation sequence; the test checker mirrors it
__device__ __forceinline__ float sp_expf(float x) {
    if (x > 88.72283f) return INFINITY;
    if (x < -103.0f) return 0.0f;
    const float kf = __fsub_rn(__fadd_rn(__fmul_rn(x, 1.44269504f), 12582912.0f), 12582912.0f);
    float r = __fsub_rn(x, __fmul_rn(kf, 0.693359375f));
    r = __fsub_rn(r, __fmul_rn(kf, -2.12194440e-4f));
    float p = 1.38888889e-3f;
    p = __fadd_rn(__fmul_rn(p, r), 8.33333377e-3f);
    p = __fadd_rn(__fmul_rn(p, r), 4.16666679e-2f);
    p = __fadd_rn(__fmul_rn(p, r), 1.66666672e-1f);
    p = __fadd_rn(__fmul_rn(p, r), 0.5f);
    p = __fadd_rn(__fmul_rn(p, r), 1.0f);
    p = __fadd_rn(__fmul_rn(p, r), 1.0f);
    const int k = (int)kf;
    const int k1 = k / 2, k2 = k - k1;
    const float s1 = __uint_as_float((uint32_t)(k1 + 127) << 23), s2 = __uint_as_float((uint32_t)(k2 + 127) << 23);
    return __fmul_rn(__fmul_rn(p, s1), s2);
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// One check of slot degree D: d_j = prod_{k != j} x_k in ascending k; padded edges act as x = 1.
template <int D, int SUB>
__device__ __forceinline__ uint32_t sp_check(uint32_t tab, uint32_t dummy_hb, uint32_t rrow, uint32_t c4, int h) {
    constexpr int NQ = (D + 3) / 4;
    uint32_t ent[NQ * 4];
#pragma unroll
    for (int jq = 0; jq < NQ; ++jq) {
        const uint4 o = lds_u128(tab + (uint32_t)(jq * SUB + h) * 16u);
        ent[jq * 4 + 0] = o.x; ent[jq * 4 + 1] = o.y; ent[jq * 4 + 2] = o.z; ent[jq * 4 + 3] = o.w;
    }
    float x[D];
    uint32_t syn = 0u;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float v = lds_f32(rrow + (uint32_t)j * 128u);
        x[j] = (ent[j] != dummy_hb) ? v : 1.0f;
        syn ^= lds_u32(ent[j] + c4);  // hard bit of the row's j-th variable (dummy entry holds 0)
    }
    // d_j = ((((x_0 x_1) ...) x_{j-1}) x_{j+1}) ... x_{D-1}: the reference's left-to-right product with edge j skipped
    // (decodeCL.c:31-37).  Its first j factors are the running prefix, shared by all later edges -- same operations in
    // the same order, D(D-1)/2 + D - 1 multiplies instead of D(D-1).
    float pre = 1.0f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        float d = pre;
#pragma unroll
        for (int k = j + 1; k < D; ++k) d = __fmul_rn(d, x[k]);
        sts_f32(rrow + (uint32_t)j * 128u, d);
        pre = __fmul_rn(pre, x[j]);  // 1 * x_0 = x_0 exactly
    }
    return syn & 1u;
}

// One variable of slot degree D (channel term t = exp(8y) in a register).  Returns the new hard bit.
template <int D, int SUB>
__device__ __forceinline__ uint32_t sp_variable(uint32_t q, uint32_t dummy_e, float t, uint32_t prev_bit, uint32_t c4, bool frozen) {
    constexpr int NQ = (D + 3) / 4;
    uint32_t e[NQ * 4];
#pragma unroll
    for (int jq = 0; jq < NQ; ++jq) {
        const uint4 o = lds_u128(q + (uint32_t)(jq * SUB) * 16u);
        e[jq * 4 + 0] = o.x; e[jq * 4 + 1] = o.y; e[jq * 4 + 2] = o.z; e[jq * 4 + 3] = o.w;
    }
    const float one_t = __fadd_rn(1.0f, t);
    const float p0 = __fdiv_rn(t, one_t), p1 = __fdiv_rn(1.0f, one_t);  // priorP0 / priorP1, decodeCL.c:13-18
    float r0[D], r1[D];
    bool valid[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        valid[k] = e[k] != dummy_e;
        const float d = lds_f32(e[k] + c4);
        r0[k] = valid[k] ? __fmul_rn(__fadd_rn(1.0f, d), 0.5f) : 1.0f;  // x / 2 == x * 0.5 for every float
        r1[k] = valid[k] ? __fmul_rn(__fsub_rn(1.0f, d), 0.5f) : 1.0f;
    }
    float t0 = p0, t1 = p1;  // hardDecision, decodeCL.c:64-86
#pragma unroll
    for (int k = 0; k < D; ++k) { t0 = __fmul_rn(t0, r0[k]); t1 = __fmul_rn(t1, r1[k]); }
    const uint32_t bit = (t0 > t1) ? 0u : ((t0 < t1) ? 1u : prev_bit);
    float pu0 = p0, pu1 = p1;  // running prefix p * r[0] ... r[k-1], shared like the check products
#pragma unroll
    for (int k = 0; k < D; ++k) {  // refreshQ, decodeCL.c:43-62
        float u0 = pu0, u1 = pu1;
#pragma unroll
        for (int m = k + 1; m < D; ++m) { u0 = __fmul_rn(u0, r0[m]); u1 = __fmul_rn(u1, r1[m]); }
        const float s = __fadd_rn(u0, u1);
        const float xq = __fsub_rn(__fdiv_rn(u0, s), __fdiv_rn(u1, s));
        if (valid[k] && !frozen) sts_f32(e[k] + c4, xq);
        pu0 = __fmul_rn(pu0, r0[k]);
        pu1 = __fmul_rn(pu1, r1[k]);
    }
    return bit;
}

// decodeInit for one variable: E_e = q0 - q1 on every edge of the column (decodeCL.c:3-12)
template <int SUB>
__device__ __forceinline__ void sp_init_variable(uint32_t q, int d, uint32_t dummy_e, float t, uint32_t c4) {
    const float one_t = __fadd_rn(1.0f, t);
    const float x0 = __fsub_rn(__fdiv_rn(t, one_t), __fdiv_rn(1.0f, one_t));
    for (int k = 0; k < d; ++k) {
        const uint32_t e = lds_u32(q + (uint32_t)((k >> 2) * SUB) * 16u + (uint32_t)(k & 3) * 4u);
        if (e != dummy_e) sts_f32(e + c4, x0);
    }
}

template <int G, int DMAX, int MAX_THREADS>
__global__ void __launch_bounds__(MAX_THREADS, (MAX_THREADS <= 384 ? 2 : 1)) ldpc_sp_group_kernel(const __grid_constant__ GroupParams p) {
    constexpr int SUB = 32 / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_flag[2][32];
    __shared__ long long s_cw[32];

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int c = lane & (G - 1), h = lane / G;
    const int W = p.W, CS = p.CS, VS = p.VS;
    const int NL = W * SUB;
    const int PD = VS * NL;
    const int RD = W * p.r_rows_per_warp;

    // same carve-up as the min-sum kernel: HB (hard bits, shape of T) | E (shape of R) | tables
    const uint32_t hb_base = smem_u32(smem_raw);
    const uint32_t hb_bytes = ((uint32_t)(PD + 1) * G * 4 + 127u) & ~127u;
    const uint32_t e_base = hb_base + hb_bytes;
    const uint32_t cn_base = e_base + (uint32_t)(RD + 1) * 128;
    const uint32_t vn_base = cn_base + (uint32_t)W * p.cn_stride * 4;
    {
        uint32_t* cn_s = reinterpret_cast<uint32_t*>(smem_raw + (cn_base - hb_base));
        uint32_t* vn_s = reinterpret_cast<uint32_t*>(smem_raw + (vn_base - hb_base));
        for (int i = threadIdx.x; i < W * p.cn_stride; i += blockDim.x) cn_s[i] = __ldg(p.cn_tab + i) + hb_base;
        for (int i = threadIdx.x; i < W * p.vn_stride; i += blockDim.x) vn_s[i] = __ldg(p.vn_tab + i) + e_base;
    }
    const uint32_t c4 = (uint32_t)c * 4u;
    const uint32_t dummy_hb = hb_base + (uint32_t)PD * G * 4, dummy_e = e_base + (uint32_t)RD * 128;
    if (warp == 0) {
        if (lane < G) sts_u32(dummy_hb + lane * 4, 0u);
        sts_f32(dummy_e + lane * 4, 1.0f);
        s_flag[0][lane] = 0u; s_flag[1][lane] = 0u;
    }
    const uint32_t hb_own = hb_base + (uint32_t)threadIdx.x * 4u;
    const uint32_t t_stride = (uint32_t)NL * G * 4u;
    const uint32_t e_own = e_base + (uint32_t)warp * p.r_rows_per_warp * 128u + (uint32_t)lane * 4u;
    const uint32_t cn_w = cn_base + (uint32_t)warp * p.cn_stride * 4u;
    const uint32_t vn_w = vn_base + (uint32_t)warp * p.vn_stride * 4u + (uint32_t)h * 16u;

    float tt[kGrpMaxVS];   // t = exp(8 y) of this thread's variables (static slots)
    uint32_t bits = 0u;    // their current hard decisions (bit s)
#pragma unroll
    for (int s = 0; s < kGrpMaxVS; ++s) tt[s] = 1.0f;
    long long cw = -1;
    bool live = false, done = false, loading = false;
    int it = 0, my_iters = 0;

    auto cn_pass = [&]() -> uint32_t {
        uint32_t unsat = 0u;
        uint32_t tab = cn_w, rrow = e_own;
        for (int cs = 0; cs < CS; ++cs) {
            const int dc = p.cdeg[cs];
#define SP_CASE(D) case D: unsat |= sp_check<D, SUB>(tab, dummy_hb, rrow, c4, h); break;
            switch (dc) {
                SP_CASE(1) SP_CASE(2) SP_CASE(3) SP_CASE(4) SP_CASE(5) SP_CASE(6) SP_CASE(7) SP_CASE(8)
                default:
                    if constexpr (DMAX > 8) {
                        switch (dc) {
                            SP_CASE(9) SP_CASE(10) SP_CASE(11) SP_CASE(12) SP_CASE(13) SP_CASE(14) SP_CASE(15) SP_CASE(16)
                            default:
                                if constexpr (DMAX > 16) {
                                    switch (dc) {
                                        SP_CASE(17) SP_CASE(18) SP_CASE(19) SP_CASE(20)
                                        default: break;
                                    }
                                }
                                break;
                        }
                    }
                    break;
            }
#undef SP_CASE
            rrow += (uint32_t)dc * 128u;
            tab += 16u * (uint32_t)(((dc + 3) >> 2) * SUB);
        }
        return unsat;
    };
    auto vn_pass = [&]() {
        const bool frozen = !(live && !done);
        uint32_t q = vn_w;
#pragma unroll
        for (int s = 0; s < kGrpMaxVS; ++s) {
            if (s < VS) {
                const int d = p.vdeg[s];
                const uint32_t prev = (bits >> s) & 1u;
                uint32_t nb = prev;
#define SP_VCASE(D) case D: nb = sp_variable<D, SUB>(q, dummy_e, tt[s], prev, c4, frozen); break;
                switch (d) {
                    SP_VCASE(1) SP_VCASE(2) SP_VCASE(3) SP_VCASE(4) SP_VCASE(5) SP_VCASE(6) SP_VCASE(7) SP_VCASE(8)
                    default: break;
                }
#undef SP_VCASE
                if (!frozen) {
                    bits = (bits & ~(1u << s)) | (nb << s);
                    sts_u32(hb_own + (uint32_t)s * t_stride, nb);
                }
                q += 16u * (uint32_t)(((d + 3) >> 2) * SUB);
            }
        }
    };
    auto emit = [&](bool sel) {  // toChar, decodeCL.c:188-199
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = warp * SUB + h; b < KB; b += NL) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.K) v |= (lds_u32(hb_base + __ldg(p.pos_of_var + n) * (G * 4) + c4) & 1u) << t;
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
                    if (n < p.N) v |= (lds_u32(hb_base + __ldg(p.pos_of_var + n) * (G * 4) + c4) & 1u) << t;
                }
                if (sel) p.hard[(size_t)cw * NB + b] = (uint8_t)v;
            }
        }
        if (p.iters && warp == 0 && h == 0 && sel) p.iters[cw] = my_iters;
    };
    auto fetch = [&](bool want) {
        if (warp == 0 && h == 0 && want) s_cw[c] = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(p.counter64), 1ull);
        __syncthreads();
        if (want) {
            live = false; done = false;
            cw = s_cw[c];
            if (cw < p.ncw) {
                loading = true;
                const float* src = p.llr + (size_t)cw * p.N;
#pragma unroll
                for (int s = 0; s < kGrpMaxVS; ++s) {
                    if (s < VS) {
                        const uint32_t v = __ldg(p.var_of_pos + s * NL + warp * SUB + h);
                        const uint32_t dst = hb_own + (uint32_t)s * t_stride;  // staged in this lane's (dead) HB slots
                        if (v != 0xffffffffu) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + v) : "memory");
                        else sts_f32(dst, 1.0f);
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto start_loaded = [&]() {  // decodeInit, decodeCL.c:3-22
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (loading) {
            uint32_t q = vn_w;
#pragma unroll
            for (int s = 0; s < kGrpMaxVS; ++s) {
                if (s < VS) {
                    const float y = lds_f32(hb_own + (uint32_t)s * t_stride);
                    tt[s] = sp_expf(__fmul_rn(8.0f, y));
                    sts_u32(hb_own + (uint32_t)s * t_stride, 0u);
                    const int d = p.vdeg[s];
                    sp_init_variable<SUB>(q, d, dummy_e, tt[s], c4);
                    q += 16u * (uint32_t)(((d + 3) >> 2) * SUB);
                }
            }
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

        const uint32_t unsat = cn_pass();  // refreshR + checkResult of the previous hard decision
        const bool check = p.early_term && it >= 1 && live && !done;
        if (check && unsat) s_flag[ph & 1][c] = 1u;
        __syncthreads();
        if (check && s_flag[ph & 1][c] == 0u) { done = true; my_iters = it; }
        if (warp == 0) s_flag[(ph + 1) & 1][lane] = 0u;
        ++ph;

        vn_pass();  // hardDecision + refreshQ
        if (live && !done) {
            ++it;
            if (it == p.max_iter) { done = true; my_iters = it; }
        }
        __syncthreads();
    }
}

}  // namespace ldpc_b200
