// ldpc_qcg.cuh -- the quasi-cyclic kernel of ldpc_qc.cuh with a RUN-TIME profile: any code made of z x z
// circulants (every size and rate Coder::initCheckMatrix builds, MyLdpc.cpp:52-109: N = 24 z, z = 24 .. 96)
// whose blocks split into groups of SUB = 32/G rows, G in {8, 4, 2} codewords per CTA chosen so that two CTAs
// fit one SM.  Same layout and arithmetic; slot degrees come from the kernel parameters (a switch per slot into
// straight-line code), a slot whose groups have different degrees is padded (check side: a T row of -inf and a
// private R block, neutral for min, parity and syndrome; variable side: the zero row), and the wrap copy is
// decided per (warp, slot).  The warp-uniform tables are copied to SHARED memory at kernel start and read with
// broadcast LDS.128: behind a run-time switch the compiler cannot hoist uniform constant loads ahead of their use,
// and their latency (constant cache misses) was exposed once per slot -- measured 2.4x slower than the compiled
// profile with LDCU, see DESIGN.md.
#pragma once
#include <cstddef>
#include "ldpc_qc.cuh"

namespace ldpc_b200 {

constexpr int kQcgMaxW = 12;    // warps per CTA (two CTAs of <= 384 threads per SM: 80 registers)
constexpr int kQcgMaxCS = 12;   // check slots per thread
constexpr int kQcgMaxVS = 16;   // variable slots per thread (channel values in registers)
constexpr int kQcgMaxCE = 96;   // table entries per warp (sum of the slot degrees, each padded to a multiple of 4)
constexpr int kQcgMaxVE = 96;
constexpr int kQcgMaxCD = 16;   // check degree instantiated
constexpr int kQcgMaxVD = 8;    // variable degree instantiated

struct QcgWarpTab {
    alignas(16) uint32_t cn_t[kQcgMaxCE];  // [slot][j] T rows (a slot starts on a multiple of 4 entries; pads -> the -inf row)
    alignas(16) uint32_t vn_r[kQcgMaxVE];  // [slot][k] R rows in ascending-row order (pads -> the zero row)
    uint32_t cn_r[kQcgMaxCS];             // own R rows of the slot's checks, edge j at + j * RS
    uint32_t vn_t[kQcgMaxVS];             // own T rows of the slot's variables
    uint32_t var0[kQcgMaxVS];             // variable index of node lane 0
    uint32_t cdup, vdup;                  // bit s: the slot's group owns wrapped rows
    uint32_t cact, vact;                  // bit s: the slot holds a real group (the last slot of some warps is empty)
};

struct QcgParams {
    const QcgWarpTab* __restrict__ tabs;  // [W] in device memory
    int N, K, Z, W, CS, VS;
    uint32_t RS, WRAP;            // bytes of one padded block; bytes of z rows
    uint32_t t_bytes, r_bytes;    // T at 0, R at t_bytes, then the zero row, the -inf row (128 B each) and the tables
    uint8_t cdeg[kQcgMaxCS];
    uint8_t vdeg[kQcgMaxVS];
    int max_iter, early_term, refill_wait;
    const float* __restrict__ llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;
    unsigned long long* counter64;
    const unsigned long long* avail;
    int* status;
    unsigned long long wait_ns;
};

// One check of exact degree D; run-time block stride (see qc_check).  Split into its loads and the rest so that
// two equal-degree slots can have their loads in flight together.
template <int D>
__device__ __forceinline__ void qcg_check_load(uint32_t tt, uint32_t rrow, uint32_t la, uint32_t RS, float* tv, float* S) {
#pragma unroll
    for (int j = 0; j < D; j += 4) {  // four warp-uniform bases per broadcast LDS.128
        const uint4 e = lds_u128(tt + (uint32_t)j * 4u);
        tv[j] = lds_f32(la + e.x);
        if (j + 1 < D) tv[j + 1] = lds_f32(la + e.y);
        if (j + 2 < D) tv[j + 2] = lds_f32(la + e.z);
        if (j + 3 < D) tv[j + 3] = lds_f32(la + e.w);
    }
#pragma unroll
    for (int j = 0; j < D; ++j) S[j] = lds_f32(la + rrow + (uint32_t)j * RS);
}
template <int D>
__device__ __forceinline__ uint32_t qcg_check_finish(uint32_t rrow, uint32_t la, uint32_t RS, uint32_t WRAP, bool dup, const float* tv, float* S) {
    uint32_t px = 0u, sx = 0u;
#pragma unroll
    for (int j = 0; j < D; ++j) S[j] = __fadd_rn(tv[j], S[j]);  // = -Q_j
#pragma unroll
    for (int j = 0; j + 1 < D; j += 2) {
        px = px ^ __float_as_uint(S[j]) ^ __float_as_uint(S[j + 1]);
        sx = sx ^ __float_as_uint(tv[j]) ^ __float_as_uint(tv[j + 1]);
    }
    if (D & 1) {
        px ^= __float_as_uint(S[D - 1]);
        sx ^= __float_as_uint(tv[D - 1]);
    }
    float rn[D];
    ms_new_messages<D>(S, px, rn);
#pragma unroll
    for (int j = 0; j < D; ++j) sts_f32(la + rrow + (uint32_t)j * RS, rn[j]);
    if (dup) {  // warp-uniform
#pragma unroll
        for (int j = 0; j < D; ++j) sts_f32(la + rrow + (uint32_t)j * RS - WRAP, rn[j]);
    }
    return ((sx >> 31) ^ (uint32_t)D) & 1u;
}
template <int D>
__device__ __forceinline__ uint32_t qcg_check(uint32_t tt, uint32_t rrow, uint32_t la, uint32_t RS, uint32_t WRAP, bool dup) {
    float tv[D + 3], S[D];
    qcg_check_load<D>(tt, rrow, la, RS, tv, S);
    return qcg_check_finish<D>(rrow, la, RS, WRAP, dup, tv, S);
}
// Two checks of the same degree (consecutive slots): both sets of loads first.
template <int D>
__device__ __forceinline__ uint32_t qcg_check2(uint32_t tt, uint32_t rrow0, uint32_t rrow1, uint32_t la, uint32_t RS, uint32_t WRAP,
                                                bool dup0, bool dup1) {
    float tv0[D + 3], S0[D], tv1[D + 3], S1[D];
    qcg_check_load<D>(tt, rrow0, la, RS, tv0, S0);
    qcg_check_load<D>(tt + (uint32_t)((D + 3) & ~3) * 4u, rrow1, la, RS, tv1, S1);
    const uint32_t u0 = qcg_check_finish<D>(rrow0, la, RS, WRAP, dup0, tv0, S0);
    return u0 | qcg_check_finish<D>(rrow1, la, RS, WRAP, dup1, tv1, S1);
}

// One variable slot of exact degree D: returns T = (-y) - R_1 - R_2 ... (ascending-row order).
template <int D>
__device__ __forceinline__ float qcg_var(uint32_t rr, uint32_t la, float acc) {
    float r[D + 3];
#pragma unroll
    for (int k = 0; k < D; k += 4) {
        const uint4 u = lds_u128(rr + (uint32_t)k * 4u);
        r[k] = lds_f32(la + u.x);
        if (k + 1 < D) r[k + 1] = lds_f32(la + u.y);
        if (k + 2 < D) r[k + 2] = lds_f32(la + u.z);
        if (k + 3 < D) r[k + 3] = lds_f32(la + u.w);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) acc = __fsub_rn(acc, r[k]);
    return acc;
}

// Two variable slots of the same degree: loads of both first, then the two subtraction chains interleaved.
template <int D>
__device__ __forceinline__ void qcg_var2(uint32_t rr, uint32_t la, float& acc0, float& acc1) {
    float r0[D + 3], r1[D + 3];
    const uint32_t rr1 = rr + (uint32_t)((D + 3) & ~3) * 4u;
#pragma unroll
    for (int k = 0; k < D; k += 4) {
        const uint4 u = lds_u128(rr + (uint32_t)k * 4u), v = lds_u128(rr1 + (uint32_t)k * 4u);
        r0[k] = lds_f32(la + u.x); r1[k] = lds_f32(la + v.x);
        if (k + 1 < D) { r0[k + 1] = lds_f32(la + u.y); r1[k + 1] = lds_f32(la + v.y); }
        if (k + 2 < D) { r0[k + 2] = lds_f32(la + u.z); r1[k + 2] = lds_f32(la + v.z); }
        if (k + 3 < D) { r0[k + 3] = lds_f32(la + u.w); r1[k + 3] = lds_f32(la + v.w); }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) { acc0 = __fsub_rn(acc0, r0[k]); acc1 = __fsub_rn(acc1, r1[k]); }
}

template <int G>
__global__ void __launch_bounds__(kQcgMaxW * 32, 2) ldpc_ms_qcg_kernel(const __grid_constant__ QcgParams p) {
    constexpr int SUB = 32 / G;
    constexpr uint32_t ROWB = (uint32_t)G * 4u;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_flag[2][32];
    __shared__ long long s_cw[32], s_nxt[32];

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform: the table reads become LDCU
    const int c = lane & (G - 1), h = lane / G;
    const uint32_t sb = smem_u32(smem_raw);
    // this warp's tables, copied to shared memory (read back with warp-uniform addresses: broadcast loads)
    const uint32_t tsm = sb + p.t_bytes + p.r_bytes + 256u + (uint32_t)warp * (uint32_t)sizeof(QcgWarpTab);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.tabs + warp);
        for (int i = lane; i < (int)(sizeof(QcgWarpTab) / 4); i += 32) sts_f32(tsm + (uint32_t)i * 4u, __uint_as_float(__ldg(src + i)));
        __syncwarp();
    }
    auto tab_u32 = [&](uint32_t byte_off) -> uint32_t { return __float_as_uint(lds_f32(tsm + byte_off)); };
    constexpr uint32_t kOffCnT = (uint32_t)offsetof(QcgWarpTab, cn_t), kOffVnR = (uint32_t)offsetof(QcgWarpTab, vn_r);
    constexpr uint32_t kOffCnR = (uint32_t)offsetof(QcgWarpTab, cn_r), kOffVnT = (uint32_t)offsetof(QcgWarpTab, vn_t);
    constexpr uint32_t kOffVar0 = (uint32_t)offsetof(QcgWarpTab, var0);
    const uint32_t cdup_bits = tab_u32((uint32_t)offsetof(QcgWarpTab, cdup)), vdup_bits = tab_u32((uint32_t)offsetof(QcgWarpTab, vdup));
    const uint32_t cact_bits = tab_u32((uint32_t)offsetof(QcgWarpTab, cact)), vact_bits = tab_u32((uint32_t)offsetof(QcgWarpTab, vact));
    const uint32_t la = sb + (uint32_t)lane * 4u;
    const uint32_t c4 = (uint32_t)c * 4u;
    const int NL = p.W * SUB, Z = p.Z, CS = p.CS, VS = p.VS;
    const uint32_t RS = p.RS, WRAP = p.WRAP;

    if (threadIdx.x < 32) {
        sts_f32(sb + p.t_bytes + p.r_bytes + (uint32_t)lane * 4u, 0.0f);             // zero row
        sts_f32(sb + p.t_bytes + p.r_bytes + 128u + (uint32_t)lane * 4u, -INFINITY); // -inf row (check-side padding)
    }

    float yn[kQcgMaxVS];
#pragma unroll
    for (int s = 0; s < kQcgMaxVS; ++s) yn[s] = -1.0f;
    long long cw = -1;
    bool live = false, done = false, loading = false;
    int it = 0, my_iters = 0;

    auto cn_pass = [&]() -> uint32_t {
        uint32_t unsat = 0u;
        uint32_t tt = tsm + kOffCnT;
        for (int cs = 0; cs < CS; ++cs) {
            const int d = p.cdeg[cs];
            const bool dup = (cdup_bits >> cs) & 1u;
            const uint32_t rrow = tab_u32(kOffCnR + (uint32_t)cs * 4u);
            const uint32_t step = (uint32_t)((d + 3) & ~3) * 4u;
            if (!((cact_bits >> cs) & 1u)) { tt += step; continue; }  // empty slot of this warp (warp-uniform)
            if (d <= 8 && cs + 1 < CS && p.cdeg[cs + 1] == d && ((cact_bits >> (cs + 1)) & 1u)) {  // two slots of a low degree: loads of both in flight
                const bool dup1 = (cdup_bits >> (cs + 1)) & 1u;
                const uint32_t rrow1 = tab_u32(kOffCnR + (uint32_t)(cs + 1) * 4u);
#define QCG_CASE2(D) case D: unsat |= qcg_check2<D>(tt, rrow, rrow1, la, RS, WRAP, dup, dup1); break;
                switch (d) {
                    QCG_CASE2(1) QCG_CASE2(2) QCG_CASE2(3) QCG_CASE2(4) QCG_CASE2(5) QCG_CASE2(6) QCG_CASE2(7) QCG_CASE2(8)
                    default: break;
                }
#undef QCG_CASE2
                tt += 2u * step;
                ++cs;
                continue;
            }
#define QCG_CASE(D) case D: unsat |= qcg_check<D>(tt, rrow, la, RS, WRAP, dup); break;
            switch (d) {
                QCG_CASE(1) QCG_CASE(2) QCG_CASE(3) QCG_CASE(4) QCG_CASE(5) QCG_CASE(6) QCG_CASE(7) QCG_CASE(8)
                QCG_CASE(9) QCG_CASE(10) QCG_CASE(11) QCG_CASE(12) QCG_CASE(13) QCG_CASE(14) QCG_CASE(15) QCG_CASE(16)
                default: break;
            }
#undef QCG_CASE
            tt += step;
        }
        return unsat;
    };
    auto vn_pass = [&](bool frozen) {
        uint32_t rr = tsm + kOffVnR;
        auto store_t = [&](int s, float acc) {
            if (!frozen) {
                const uint32_t ta = la + tab_u32(kOffVnT + (uint32_t)s * 4u);
                sts_f32(ta, acc);
                if ((vdup_bits >> s) & 1u) sts_f32(ta + WRAP, acc);
            }
        };
        auto one = [&](int s, int d, float acc) {
            if (!((vact_bits >> s) & 1u)) { rr += (uint32_t)((d + 3) & ~3) * 4u; return; }  // empty slot of this warp
            switch (d) {
                case 1: acc = qcg_var<1>(rr, la, acc); break;
                case 2: acc = qcg_var<2>(rr, la, acc); break;
                case 3: acc = qcg_var<3>(rr, la, acc); break;
                case 4: acc = qcg_var<4>(rr, la, acc); break;
                case 5: acc = qcg_var<5>(rr, la, acc); break;
                case 6: acc = qcg_var<6>(rr, la, acc); break;
                case 7: acc = qcg_var<7>(rr, la, acc); break;
                case 8: acc = qcg_var<8>(rr, la, acc); break;
                default: break;
            }
            store_t(s, acc);
            rr += (uint32_t)((d + 3) & ~3) * 4u;
        };
#pragma unroll
        for (int s = 0; s < kQcgMaxVS; s += 2) {  // slots in pairs: equal degrees (the common case) share one straight-line block
            if (s < VS) {
                const int d0 = p.vdeg[s];
                if (s + 1 < VS && p.vdeg[s + 1] == d0 && ((vact_bits >> (s + 1)) & 1u)) {  // (slot s + 1 real => slot s real)
                    float a0 = yn[s], a1 = yn[s + 1];
                    switch (d0) {
                        case 1: qcg_var2<1>(rr, la, a0, a1); break;
                        case 2: qcg_var2<2>(rr, la, a0, a1); break;
                        case 3: qcg_var2<3>(rr, la, a0, a1); break;
                        case 4: qcg_var2<4>(rr, la, a0, a1); break;
                        case 5: qcg_var2<5>(rr, la, a0, a1); break;
                        case 6: qcg_var2<6>(rr, la, a0, a1); break;
                        case 7: qcg_var2<7>(rr, la, a0, a1); break;
                        case 8: qcg_var2<8>(rr, la, a0, a1); break;
                        default: break;
                    }
                    store_t(s, a0);
                    store_t(s + 1, a1);
                    rr += 2u * (uint32_t)((d0 + 3) & ~3) * 4u;
                } else {
                    one(s, d0, yn[s]);
                    if (s + 1 < VS) one(s + 1, p.vdeg[s + 1], yn[s + 1]);
                }
            }
        }
    };
    // hard bits 8b .. 8b+7 of codeword lane c (one division per byte; a byte may straddle block columns)
    auto pack8 = [&](int b) -> uint32_t {
        int bc = (b * 8) / Z, r = (b * 8) % Z;
        uint32_t v = 0u;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (b * 8 + t < p.N) v |= ((~__float_as_uint(lds_f32(sb + (uint32_t)(bc * (Z + SUB) + r) * ROWB + c4))) >> 31) << t;
            if (++r == Z) { r = 0; ++bc; }
        }
        return v;
    };
    auto emit = [&](bool sel) {
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = warp * SUB + h; b < KB; b += NL) {
                uint32_t v = pack8(b);
                if (b * 8 + 8 > p.K) v &= (1u << (p.K - b * 8)) - 1u;
                if (sel) p.info[(size_t)cw * KB + b] = (uint8_t)v;
            }
        }
        if (p.hard) {
            const int NB8 = (p.N + 7) >> 3;
            for (int b = warp * SUB + h; b < NB8; b += NL) {
                const uint32_t v = pack8(b);
                if (sel) p.hard[(size_t)cw * NB8 + b] = (uint8_t)v;
            }
        }
        if (p.post && sel) {
            for (int n = warp * SUB + h; n < p.N; n += NL)
                p.post[(size_t)cw * p.N + n] = -lds_f32(sb + (uint32_t)((n / Z) * (Z + SUB) + (n % Z)) * ROWB + c4);
        }
        if (p.iters && warp == 0 && h == 0 && sel) p.iters[cw] = my_iters;
    };
    // warp 0: lanes selected by `want` claim their next word and pull its channel values into L2 (see ldpc_qc.cuh)
    auto claim = [&](bool want) {
        long long nn = (h == 0 && want) ? (long long)atomicAdd(p.counter64, 1ull) : p.ncw;
        if (h == 0 && want) s_nxt[c] = nn;
        nn = __shfl_sync(0xffffffffu, nn, c);
        const bool w2 = __shfl_sync(0xffffffffu, (int)want, c) != 0;
        const char* base = reinterpret_cast<const char*>(p.llr + (size_t)nn * p.N);
        const bool pf = w2 && nn < p.ncw;
        for (int i = 0; i < (p.N * 4 + SUB * 128 - 1) / (SUB * 128); ++i) {
            const int off = (i * SUB + h) * 128;
            if (pf && off < p.N * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
    };
    auto fetch = [&](bool want) {
        if (warp == 0) {
            const bool take = h == 0 && want;
            long long w = take ? s_nxt[c] : -1;
            if (p.avail && !qc_wait_input(p.avail, w, take && w < p.ncw, p.status, p.wait_ns)) w = p.ncw;
            if (take) s_cw[c] = w;
        }
        __syncthreads();  // also: every read of the retiring lanes' T (emit) is complete
        if (want) {
            live = false; done = false;
            cw = s_cw[c];
            if (cw < p.ncw) {
                loading = true;
                const float* src = p.llr + (size_t)cw * p.N + h;
#pragma unroll
                for (int s = 0; s < kQcgMaxVS; ++s) {
                    if (s < VS && ((vact_bits >> s) & 1u)) {
                        const uint32_t dst = la + tab_u32(kOffVnT + (uint32_t)s * 4u);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + tab_u32(kOffVar0 + (uint32_t)s * 4u)) : "memory");
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (warp == 0) claim(want);
    };
    auto start_loaded = [&]() {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (loading) {
#pragma unroll
            for (int s = 0; s < kQcgMaxVS; ++s) {
                if (s < VS && ((vact_bits >> s) & 1u)) {
                    const uint32_t ta = la + tab_u32(kOffVnT + (uint32_t)s * 4u);
                    const float y = lds_f32(ta);
                    yn[s] = __fadd_rn(-y, 0.0f);
                    sts_f32(ta, yn[s]);
                    if ((vdup_bits >> s) & 1u) sts_f32(ta + WRAP, yn[s]);
                }
            }
            for (int cs = 0; cs < CS; ++cs) {
                if (!((cact_bits >> cs) & 1u)) continue;
                const bool dup = (cdup_bits >> cs) & 1u;
                const uint32_t rrow = la + tab_u32(kOffCnR + (uint32_t)cs * 4u);
                for (int j = 0; j < p.cdeg[cs]; ++j) {
                    sts_f32(rrow + (uint32_t)j * RS, 0.0f);
                    if (dup) sts_f32(rrow + (uint32_t)j * RS - WRAP, 0.0f);
                }
            }
            loading = false; live = true; done = false; it = 0;
        }
        __syncthreads();
    };

    if (warp == 0) { s_flag[0][lane] = 0u; s_flag[1][lane] = 0u; claim(true); }
    __syncthreads();
    fetch(true);
    uint32_t ph = 0;
    for (;;) {
        if (__any_sync(0xffffffffu, loading)) start_loaded();
        const bool retire = live && done;
        if (__any_sync(0xffffffffu, retire)) {
            emit(retire);
            fetch(retire);
            if ((p.refill_wait || !__any_sync(0xffffffffu, live)) && __any_sync(0xffffffffu, loading)) start_loaded();
        }
        if (!__any_sync(0xffffffffu, live || loading)) break;

        const uint32_t unsat = cn_pass();
        const bool check = p.early_term && it >= 1 && live && !done;
        if (check && unsat) s_flag[ph & 1][c] = 1u;  // same-value race, benign
        __syncthreads();
        if (check && s_flag[ph & 1][c] == 0u) { done = true; my_iters = it; }
        if (warp == 0) s_flag[(ph + 1) & 1][lane] = 0u;
        ++ph;

        vn_pass(!(live && !done));
        if (live && !done) {
            ++it;
            if (it == p.max_iter) { done = true; my_iters = it; }
        }
        __syncthreads();
    }
}

}  // namespace ldpc_b200
