// ldpc_kernels.cuh -- sm_100a device code of the flooding min-sum decoder.
//
// Arithmetic contract (bit-exact with Coder::decodeCPU, reference MyLdpc.cpp:684-784):
//   * check node (refreshRMS, decodeCL.c:126-147 / MyLdpc.cpp:705-721): for edge e of a row,
//     R_e = (-1)^{xor of (Q_p<0), p != e} * min(1000, min_{p != e} |Q_p|).  Held here as the
//     row's compressed state {min1, min2, argmin position, sign bits}; the exclude-self minimum
//     is (pos == argmin ? min2 : min1), ties included (then min2 == min1).
//   * variable node (refreshPostPMS, decodeCL.c:149-171 / MyLdpc.cpp:723-735):
//     P_n = ((y_n + R_e1) + R_e2) + ... in ascending-row order, one fp32 rounding per add
//     (__fadd_rn: never contracted, never reordered); hard bit = !(P_n > 0).
//   * refreshQMS (decodeCL.c:175-186 / MyLdpc.cpp:757-762): Q_e = P_col(e) - R_e (__fsub_rn),
//     evaluated lazily inside the next check-node pass.
//   * checkResult (decodeCL.c:88-108 / MyLdpc.cpp:737-750): the syndrome of iteration t's hard
//     decision is the xor of !(P>0) over each row -- computed for free while the next
//     check-node pass gathers P, so a converged word costs one extra check pass instead of a
//     separate syndrome pass every iteration.
// y is canonicalised with y + 0.0f (-0.0 -> +0.0): then neither P nor Q can ever be -0.0 and the
// sign bit of Q equals the reference's (Q < 0) test (see DESIGN.md "zero signs").
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldpc_b200 {

constexpr int kLanes = 32;
constexpr float kClamp = 1000.0f;  // reference MyLdpc.cpp:708, decodeCL.c:134

struct DecodeParams {
    const int32_t* __restrict__ row_ptr;   // [M+1]
    const uint32_t* __restrict__ cn_col;   // [nnz] column of edge e, check-major
    const int32_t* __restrict__ col_ptr;   // [N+1]
    const uint32_t* __restrict__ vn_edge;  // [nnz] (check << 5) | pos, variable-major, ascending row
    int M, N, K, max_iter, early_term;
    const float* __restrict__ llr;         // [ncw][N]
    long long ncw;
    uint8_t* info;                         // [ncw][ceil(K/8)] or null
    uint8_t* hard;                         // [ncw][ceil(N/8)] or null
    int32_t* iters;                        // [ncw] or null
    float* post;                           // [ncw][N] or null
    float* ws;                             // CTA-private global workspace (LANE_GLOBAL path)
    size_t ws_stride;                      // floats per CTA
    unsigned int* counter;                 // work-queue head, zeroed before launch
    int ngroups;                           // ceil(ncw / 32)
};

// R_e rebuilt from a row's compressed state: magnitude by argmin position, sign from bit j.
__device__ __forceinline__ float msg_from_state(float m1, float m2, uint32_t w, int j) {
    const float mag = (j == (int)(w >> 27)) ? m2 : m1;
    const uint32_t s = (w << (31 - j)) & 0x80000000u;
    return __uint_as_float(__float_as_uint(mag) ^ s);
}

// ---------------------------------------------------------------------------------------
// LANE kernels: one CTA decodes 32 codewords at a time, lane = codeword.  Every array is laid
// out [index][lane], so each shared/global access of a warp is one conflict-free 128-byte row
// and all index arithmetic is warp-uniform.  Warps split the checks (CN pass) and the variables
// (VN pass) of those 32 words; two __syncthreads per iteration give the flooding schedule.
//   P  [N][32]  posterior of the previous iteration (channel value before iteration 1)
//   Y  [N][32]  channel values
//   M1 [M][32], M2 [M][32], SW [M][32]  compressed check state (clamped min1/min2, signs|argmin)
// SMEM=true keeps all five in shared memory ((2N+3M)*128 B); SMEM=false keeps them in a
// CTA-private slice of a global workspace (any code size; L2/HBM-bound).
// ---------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(1024, 1) ldpc_ms_lane_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_group;
    __shared__ uint32_t s_flag[2][kLanes];

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int W = blockDim.x >> 5;
    const int M = p.M, N = p.N;

    float* base = SMEM ? reinterpret_cast<float*>(smem_raw) : (p.ws + (size_t)blockIdx.x * p.ws_stride);
    float* P = base;
    float* Y = P + (size_t)N * kLanes;
    float* M1 = Y + (size_t)N * kLanes;
    float* M2 = M1 + (size_t)M * kLanes;
    uint32_t* SW = reinterpret_cast<uint32_t*>(M2 + (size_t)M * kLanes);

    for (;;) {
        if (threadIdx.x == 0) s_group = (int)atomicAdd(p.counter, 1u);
        __syncthreads();
        const int g = s_group;
        if (g >= p.ngroups) break;

        const long long cw = (long long)g * kLanes + lane;
        const bool active = cw < p.ncw;
        const float* src = p.llr + (size_t)(active ? cw : 0) * N;

        // ---- load: y -> Y and P (decodeInitMS, decodeCL.c:113-124: Q_e = y[col(e)], i.e. R = 0)
        if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(p.llr) & 15u) == 0) {  // (a caller's pointer may be 4-byte aligned only)
            for (int n4 = warp; n4 < (N >> 2); n4 += W) {
                float4 v = active ? __ldg(reinterpret_cast<const float4*>(src) + n4) : make_float4(1.f, 1.f, 1.f, 1.f);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float y = __fadd_rn(vv[k], 0.0f);
                    const int a = (n4 * 4 + k) * kLanes + lane;
                    P[a] = y;
                    Y[a] = y;
                }
            }
        } else {
            for (int n = warp; n < N; n += W) {
                const float y = __fadd_rn(active ? __ldg(src + n) : 1.0f, 0.0f);
                P[n * kLanes + lane] = y;
                Y[n * kLanes + lane] = y;
            }
        }
        for (int c = warp; c < M; c += W) {
            M1[c * kLanes + lane] = 0.0f;
            M2[c * kLanes + lane] = 0.0f;
            SW[c * kLanes + lane] = 0u;
        }
        if (warp == 0) { s_flag[0][lane] = 0u; s_flag[1][lane] = 0u; }
        __syncthreads();

        bool done = !active;
        int my_iters = 0;
        int iter = 0;
        for (;;) {
            // ---- check-node pass (refreshQMS folded in) + syndrome of the previous posterior
            uint32_t unsat = 0u;
            for (int c = warp; c < M; c += W) {
                const int e0 = __ldg(p.row_ptr + c);
                const int dc = __ldg(p.row_ptr + c + 1) - e0;
                const int si = c * kLanes + lane;
                const float m1o = M1[si], m2o = M2[si];
                const uint32_t wo = SW[si];
                const int idxo = (int)(wo >> 27);
                float m1 = INFINITY, m2 = INFINITY;
                uint32_t sg = 0u, idx = 0u, syn = 0u;
                // batches of 8 edges: all column indices, then all posteriors, then the arithmetic --
                // up to 8 independent 128-byte rows in flight per warp (the global-workspace path is
                // bound by memory latency x parallelism, see profiles/r01_m_lane_global_cfg5_ncu.txt)
                for (int j0 = 0; j0 < dc; j0 += 8) {
                    uint32_t col[8];
                    float pvv[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) col[t] = (j0 + t < dc) ? __ldg(p.cn_col + e0 + j0 + t) : 0u;
#pragma unroll
                    for (int t = 0; t < 8; ++t) pvv[t] = (j0 + t < dc) ? P[col[t] * kLanes + lane] : 1.0f;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const int j = j0 + t;
                        if (j < dc) {
                            const float pv = pvv[t];
                            const float mag = (j == idxo) ? m2o : m1o;
                            const float r = __uint_as_float(__float_as_uint(mag) ^ ((wo << (31 - j)) & 0x80000000u));
                            const float q = __fsub_rn(pv, r);
                            syn ^= (pv > 0.0f) ? 0u : 1u;
                            sg |= (__float_as_uint(q) >> 31) << j;
                            const float a = fabsf(q);
                            idx = (a < m1) ? (uint32_t)j : idx;
                            m2 = fminf(m2, fmaxf(m1, a));
                            m1 = fminf(m1, a);
                        }
                    }
                }
                const uint32_t mask = (dc >= 32) ? 0xffffffffu : ((1u << dc) - 1u);
                const uint32_t sr = ((__popc(sg) & 1) ? ~sg : sg) & mask;
                M1[si] = fminf(m1, kClamp);
                M2[si] = fminf(m2, kClamp);
                SW[si] = sr | (idx << 27);
                unsat |= syn;
            }
            const bool check = p.early_term && iter >= 1;
            if (check && unsat) s_flag[iter & 1][lane] = 1u;  // same-value race, benign
            __syncthreads();
            if (check && !done && s_flag[iter & 1][lane] == 0u) { done = true; my_iters = iter; }
            if (__all_sync(0xffffffffu, done)) break;  // identical in every warp
            if (warp == 0) s_flag[(iter + 1) & 1][lane] = 0u;

            // ---- variable-node pass: posterior in ascending-row order; frozen once a word is done
            for (int n = warp; n < N; n += W) {
                const int v0 = __ldg(p.col_ptr + n);
                const int dv = __ldg(p.col_ptr + n + 1) - v0;
                float acc = Y[n * kLanes + lane];
                for (int k0 = 0; k0 < dv; k0 += 8) {  // batched like the check pass: 24 state loads in flight
                    uint32_t pk[8], sw[8];
                    float a1[8], a2[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) pk[t] = (k0 + t < dv) ? __ldg(p.vn_edge + v0 + k0 + t) : 0u;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (k0 + t < dv) {
                            const int si = (int)(pk[t] >> 5) * kLanes + lane;
                            a1[t] = M1[si]; a2[t] = M2[si]; sw[t] = SW[si];
                        }
                    }
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if (k0 + t < dv) acc = __fadd_rn(acc, msg_from_state(a1[t], a2[t], sw[t], (int)(pk[t] & 31u)));
                }
                if (!done) P[n * kLanes + lane] = acc;
            }
            ++iter;
            if (iter == p.max_iter) {
                if (!done) my_iters = iter;
                break;
            }
            __syncthreads();
        }
        __syncthreads();

        // ---- outputs (toChar, decodeCL.c:188-199: LSB-first packing of the first K hard bits)
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = warp; b < KB; b += W) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.K) v |= ((P[n * kLanes + lane] > 0.0f) ? 0u : 1u) << t;
                }
                if (active) p.info[(size_t)cw * KB + b] = (uint8_t)v;
            }
        }
        if (p.hard) {
            const int NB = (N + 7) >> 3;
            for (int b = warp; b < NB; b += W) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < N) v |= ((P[n * kLanes + lane] > 0.0f) ? 0u : 1u) << t;
                }
                if (active) p.hard[(size_t)cw * NB + b] = (uint8_t)v;
            }
        }
        if (p.post && active) {
            for (int n = warp; n < N; n += W) p.post[(size_t)cw * N + n] = P[n * kLanes + lane];
        }
        if (p.iters && warp == 0 && active) p.iters[cw] = my_iters;
    }
}

// ---------------------------------------------------------------------------------------
// LANE16 kernel: the tuned shared-memory path for short codes (Test.cpp's N=576 code).
// Same schedule as above (32 codewords per CTA, lane = codeword, two barriers per iteration),
// arranged so that the binding resource -- the 16-lane ALU pipe (LOP3/SHF/FMNMX/ISETP/SEL) --
// sees as few instructions per edge as the arithmetic allows:
//   T  [pos][32] f32    NEGATED posterior T = -P, zero canonicalised to +0.0, so that
//                       hard bit = !signbit(T) and a row's syndrome is one XOR per edge
//   ST [cpos][32] uint4 {min1, min2, sign word, argmin key}: one LDS.128 per edge in the VN pass;
//                       argmin key = byte offset of the argmin's row in T
//   channel values      stay in REGISTERS: warp w owns variable positions w, w+W, w+2W ... for
//                       the whole decode (static slots); thread (w, lane) keeps -y of its words
//   ownership order     variables and checks are sorted by degree (host side) and dealt to
//                       (slot, warp) in that order, so a slot has ONE degree for all warps:
//                       no per-variable / per-check degree loads, padding only at class edges
//   index tables        per-warp flat lists in shared memory, pre-multiplied:
//                       cn: byte offset of T row;  vn: {byte offset of ST row, 1 << shift}
//                       (the sign bit is fetched with an IMAD on the FMA pipe, not a shift)
//   padding             dummy T row = -inf (acts as P = +inf: never the minimum, positive sign,
//                       hard bit 0); dummy ST row = {0,0,0,nokey} (R = +0.0, adds nothing)
// Arithmetic is the same fp32 sequence as the reference: T = (-y) - R1 - R2 ... is the exact
// negation of y + R1 + R2 ... (round-to-nearest is sign-symmetric); Q = P - R = -(T + R).
// The sign of an exact zero never influences a non-zero value or a decision.
// ---------------------------------------------------------------------------------------
constexpr int kL16MaxVS = 24;  // variable slots per warp (channel values held in registers)
constexpr int kL16MaxCS = 32;  // check slots per warp

struct Lane16Params {
    const uint32_t* __restrict__ cn_tab;      // [W][CS][DCP] byte offsets pos*128 (dummy = PD*128)
    const uint2* __restrict__ vn_tab;         // [W][vn_stride] {cpos*512, 1 << shift} (dummy = {CPD*512, 0})
    const uint32_t* __restrict__ var_of_pos;  // [VS*W] variable index, 0xffffffff = phantom
    const uint32_t* __restrict__ pos_of_var;  // [N]
    int M, N, K, W, CS, VS, DCP, vn_stride;
    int max_iter, early_term;
    const float* __restrict__ llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;
    unsigned int* counter;
    int ngroups;
    uint8_t vdeg[kL16MaxVS];  // slot degree of variable slot s (0 beyond VS)
    uint8_t cdeg[kL16MaxCS];  // slot degree of check slot cs
};

// Explicit 32-bit shared-window addressing: keeps every hot-loop access an `LDS [R + imm]` with a
// single 2-input add in front of it (no generic-address reconstruction in the loop).
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_u128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Channel value i of a buffer in one of the host formats of ldpc_b200_decode_host_packed (0 fp32, 1 IEEE binary16,
// 2 int8): (float)x * scale, one fp32 rounding -- what the widening kernel of the chunked pipeline computes.
__device__ __forceinline__ float llr_at(const void* __restrict__ base, int fmt, float scale, size_t i) {
    if (fmt == 1) return __fmul_rn(__half2float(__ldg(static_cast<const __half*>(base) + i)), scale);
    if (fmt == 2) return __fmul_rn((float)__ldg(static_cast<const signed char*>(base) + i), scale);
    return __ldg(static_cast<const float*>(base) + i);
}

// The new messages of one check from S_j = -Q_j (refreshRMS, decodeCL.c:126-147 / MyLdpc.cpp:705-721):
//   R_j = sign_j * min(1000, min over the OTHER edges of |S|),  sign_j negative iff an odd number of the other Q are
//   negative = parity ^ 1 ^ signbit(S_j)  (Q_j < 0 <=> !signbit(S_j); px = xor of all S_j, D = degree).
// The exclude-self minimum is taken as min3(prefix, neighbour, suffix) over PAIRS of edges: two three-input min
// instructions per edge.  Tracking min1/min2 and selecting by comparison gives the same value bit for bit (ties
// included: the minimum of the others is the same number either way) but costs 4.5 instructions per edge on the
// half-rate ALU pipe, which bound the check pass.  The sign is applied by one multiplication with +-1.0 (FMA pipe; exact,
// and a zero magnitude takes the sign bit as the xor did).
__device__ __forceinline__ float ms_min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
// emit(j, R_j) is called for j = 0 .. D-1 as soon as each message is known (kernels whose stores may follow their
// messages one by one: measured 4 % faster in ldpc_qcw.cuh than storing all at the end).
template <int D, class Emit>
__device__ __forceinline__ void ms_new_messages_each(const float* S, uint32_t px, Emit emit) {
    constexpr int H = (D + 1) / 2;
    const uint32_t one = (((px >> 31) ^ (uint32_t)D ^ 1u) << 31) ^ 0x3f800000u;  // +-1.0f, flipped per edge by signbit(S_j)
    // pe[t] = min(1000, |S_0| .. |S_{2t-1}|), se[t] = min(1000, |S_{2t}| .. |S_{D-1}|)
    float pe[H + 1], se[H + 1];
    pe[0] = kClamp;
#pragma unroll
    for (int t = 0; t + 1 < H; ++t) pe[t + 1] = ms_min3(pe[t], fabsf(S[2 * t]), fabsf(S[2 * t + 1]));
    se[H] = kClamp;
#pragma unroll
    for (int t = H - 1; t >= 1; --t)
        se[t] = (2 * t + 1 < D) ? ms_min3(se[t + 1], fabsf(S[2 * t]), fabsf(S[2 * t + 1])) : fminf(se[t + 1], fabsf(S[2 * t]));
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int t = j >> 1, o = j ^ 1;   // the other edge of the pair (none for the last edge of an odd row)
        const float m = o < D ? ms_min3(pe[t], fabsf(S[o]), se[t + 1]) : fminf(pe[t], se[t + 1]);
        uint32_t sg;  // (S_j & 0x80000000) ^ one in ONE LOP3
        asm("lop3.b32 %0, %1, 0x80000000, %2, 0x6a;" : "=r"(sg) : "r"(__float_as_uint(S[j])), "r"(one));
        emit(j, __fmul_rn(m, __uint_as_float(sg)));
    }
}
template <int D>
__device__ __forceinline__ void ms_new_messages(const float* S, uint32_t px, float* rn) {
    ms_new_messages_each<D>(S, px, [&](int j, float v) { rn[j] = v; });
}

// CNT consecutive variable-node edges: entries at shared address `q` ({ST row address, 1 << shift}).
template <int CNT>
__device__ __forceinline__ void l16_vn_edges(uint32_t q, uint32_t lane16, uint32_t key, float& acc) {
    uint2 e[CNT];
#pragma unroll
    for (int k = 0; k < CNT; ++k) e[k] = lds_u64(q + 8 * k);
    uint4 st[CNT];
#pragma unroll
    for (int k = 0; k < CNT; ++k) st[k] = lds_u128(e[k].x + lane16);
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        const uint32_t mag = (st[k].w == key) ? st[k].y : st[k].x;
        const uint32_t sg = st[k].z * e[k].y;  // bit `shift` of the sign word -> bit 31 (IMAD, FMA pipe)
        acc = __fsub_rn(acc, __uint_as_float((sg & 0x80000000u) ^ mag));
    }
}

template <int MAX_THREADS>  // 768 (<= 24 warps: 80 registers per thread) or 1024 (64 registers)
__global__ void __launch_bounds__(MAX_THREADS, 1) ldpc_ms_lane16_kernel(const __grid_constant__ Lane16Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_group;
    __shared__ uint32_t s_flag[2][kLanes];

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int W = p.W, CS = p.CS, VS = p.VS, DCP = p.DCP;
    const int PD = VS * W, CPD = CS * W;  // dummy rows

    // shared-window byte addresses of the four regions
    const uint32_t st_base = smem_u32(smem_raw);
    const uint32_t t_base = st_base + (uint32_t)(CPD + 1) * kLanes * 16;
    const uint32_t cn_base = t_base + (uint32_t)(PD + 1) * kLanes * 4;
    const uint32_t vn_base = cn_base + (uint32_t)W * CS * DCP * 4;
    {
        // tables are stored with the region base already added: an entry IS a shared address
        uint32_t* cn_tab = reinterpret_cast<uint32_t*>(smem_raw + (cn_base - st_base));
        uint2* vn_tab = reinterpret_cast<uint2*>(smem_raw + (vn_base - st_base));
        for (int i = threadIdx.x; i < W * CS * DCP; i += blockDim.x) cn_tab[i] = __ldg(p.cn_tab + i) + t_base;
        for (int i = threadIdx.x; i < W * p.vn_stride; i += blockDim.x) {
            uint2 e = __ldg(p.vn_tab + i);
            e.x += st_base;
            vn_tab[i] = e;
        }
    }
    const uint32_t lane4 = (uint32_t)lane * 4u, lane16 = (uint32_t)lane * 16u;
    if (warp == 0) {
        sts_f32(t_base + (uint32_t)PD * 128u + lane4, -INFINITY);
        sts_u128(st_base + (uint32_t)CPD * 512u + lane16, make_uint4(0u, 0u, 0u, 0xffffffffu));
    }
    const uint32_t cn_w = cn_base + (uint32_t)warp * CS * DCP * 4;
    const uint32_t vn_w = vn_base + (uint32_t)warp * p.vn_stride * 8;
    const uint32_t key_w = t_base + (uint32_t)warp * 128u;  // T row (lane 0) of this warp's slot 0
    const uint32_t sta_w = st_base + (uint32_t)warp * 512u + lane16;
    const uint32_t t_stride = (uint32_t)W * 128u, st_stride = (uint32_t)W * 512u;

    for (;;) {
        if (threadIdx.x == 0) s_group = (int)atomicAdd(p.counter, 1u);
        __syncthreads();
        const int g = s_group;
        if (g >= p.ngroups) break;

        const long long cw = (long long)g * kLanes + lane;
        const bool active = cw < p.ncw;
        const float* src = p.llr + (size_t)(active ? cw : 0) * p.N;

        // ---- load: -y into registers (static slots) and T; R = 0 (decodeInitMS, decodeCL.c:113-124)
        float yn[kL16MaxVS];
#pragma unroll
        for (int s = 0; s < kL16MaxVS; ++s) {
            yn[s] = -1.0f;
            if (s < VS) {
                const uint32_t v = __ldg(p.var_of_pos + s * W + warp);
                float y = 1.0f;
                if (v != 0xffffffffu && active) y = __ldg(src + v);
                yn[s] = __fadd_rn(-y, 0.0f);  // canonical: -0.0 is never stored
                sts_f32(key_w + (uint32_t)s * t_stride + lane4, yn[s]);
            }
        }
        for (int cs = 0; cs < CS; ++cs) sts_u128(sta_w + (uint32_t)cs * st_stride, make_uint4(0u, 0u, 0u, 0xffffffffu));
        if (warp == 0) { s_flag[0][lane] = 0u; s_flag[1][lane] = 0u; }
        __syncthreads();

        bool done = !active;
        int my_iters = 0;
        int iter = 0;
        for (;;) {
            // ---- check-node pass: S = T + R_old = -Q; new {min1, min2, signs, argmin}; syndrome of T
            uint32_t unsat = 0u;
            for (int cs = 0; cs < CS; ++cs) {
                const int dc = p.cdeg[cs];
                const uint32_t sta = sta_w + (uint32_t)cs * st_stride;
                const uint4 so = lds_u128(sta);
                const float m1o = __uint_as_float(so.x), m2o = __uint_as_float(so.y);
                uint32_t wsh = so.z << ((32 - dc) & 31);
                const uint32_t argo = so.w;
                uint32_t tab = cn_w + (uint32_t)(cs * DCP) * 4u;
                float m1 = INFINITY, m2 = INFINITY;
                uint32_t sS = 0u, arg = 0xffffffffu, sx = 0u;
                auto edge = [&](uint32_t ent) {
                    const float t = lds_f32(ent + lane4);
                    const float mag = (argo == ent) ? m2o : m1o;
                    const float r = __uint_as_float(__float_as_uint(mag) ^ (wsh & 0x80000000u));
                    wsh <<= 1;
                    const float sv = __fadd_rn(t, r);
                    sS = __funnelshift_l(__float_as_uint(sv), sS, 1);
                    sx ^= __float_as_uint(t);
                    const float a = fabsf(sv);
                    arg = (a < m1) ? ent : arg;
                    m2 = fminf(m2, fmaxf(m1, a));
                    m1 = fminf(m1, a);
                };
                int j = dc;
#pragma unroll 1
                for (; j >= 4; j -= 4) {
                    const uint4 o = lds_u128(tab);
                    tab += 16;
                    edge(o.x); edge(o.y); edge(o.z); edge(o.w);
                }
                if (j > 0) {
                    const uint4 o = lds_u128(tab);
                    edge(o.x);
                    if (j > 1) edge(o.y);
                    if (j > 2) edge(o.z);
                }
                // bit (dc-1-j) of sS = signbit(S_j) = !(Q_j < 0); sign of R_j = parity ^ (Q_j < 0)
                const uint32_t mask = (dc >= 32) ? 0xffffffffu : ((1u << dc) - 1u);
                const uint32_t par = (uint32_t)(dc - __popc(sS & mask)) & 1u;
                const uint32_t sr = (par ? sS : ~sS) & mask;
                sts_u128(sta, make_uint4(__float_as_uint(fminf(m1, kClamp)), __float_as_uint(fminf(m2, kClamp)), sr, arg));
                // hard bit = !signbit(T): row syndrome = xor of signbits ^ (dc & 1)
                unsat |= ((sx >> 31) ^ (uint32_t)dc) & 1u;
            }
            const bool check = p.early_term && iter >= 1;
            if (check && unsat) s_flag[iter & 1][lane] = 1u;  // same-value race, benign
            __syncthreads();
            if (check && !done && s_flag[iter & 1][lane] == 0u) { done = true; my_iters = iter; }
            if (__all_sync(0xffffffffu, done)) break;
            if (warp == 0) s_flag[(iter + 1) & 1][lane] = 0u;

            // ---- variable-node pass: T = (-y) - R_e1 - R_e2 ... in ascending-row order
            {
                uint32_t q = vn_w;
                uint32_t key = key_w;
#pragma unroll
                for (int s = 0; s < kL16MaxVS; ++s) {
                    if (s < VS) {
                        int d = p.vdeg[s];
                        float acc = yn[s];
#pragma unroll 1
                        for (; d >= 4; d -= 4) { l16_vn_edges<4>(q, lane16, key, acc); q += 32; }
                        if (d & 2) { l16_vn_edges<2>(q, lane16, key, acc); q += 16; }
                        if (d & 1) { l16_vn_edges<1>(q, lane16, key, acc); q += 8; }
                        if (!done) sts_f32(key + lane4, acc);
                        key += t_stride;
                    }
                }
            }
            ++iter;
            if (iter == p.max_iter) {
                if (!done) my_iters = iter;
                break;
            }
            __syncthreads();
        }
        __syncthreads();

        // ---- outputs (toChar, decodeCL.c:188-199): bit n = !(P > 0) = !signbit(T)
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = warp; b < KB; b += W) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.K) {
                        const uint32_t pos = __ldg(p.pos_of_var + n);
                        v |= ((~__float_as_uint(lds_f32(t_base + pos * 128u + lane4))) >> 31) << t;
                    }
                }
                if (active) p.info[(size_t)cw * KB + b] = (uint8_t)v;
            }
        }
        if (p.hard) {
            const int NB = (p.N + 7) >> 3;
            for (int b = warp; b < NB; b += W) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.N) {
                        const uint32_t pos = __ldg(p.pos_of_var + n);
                        v |= ((~__float_as_uint(lds_f32(t_base + pos * 128u + lane4))) >> 31) << t;
                    }
                }
                if (active) p.hard[(size_t)cw * NB + b] = (uint8_t)v;
            }
        }
        if (p.post && active) {
            for (int n = warp; n < p.N; n += W)
                p.post[(size_t)cw * p.N + n] = -lds_f32(t_base + __ldg(p.pos_of_var + n) * 128u + lane4);
        }
        if (p.iters && warp == 0 && active) p.iters[cw] = my_iters;
    }
}

// ---------------------------------------------------------------------------------------
// GROUP kernel: explicit per-edge messages on chip.  A CTA decodes G codewords at a time; a warp
// instruction covers SUB = 32/G graph nodes x G codewords: lane = (h, c), c = lane % G the
// codeword, h = lane / G the node lane.  Shared memory holds
//   T [pos][G]      f32  negated posterior (as in LANE16: hard bit = !signbit, zero canonical)
//   R [row][32]     f32  the check-to-variable message of every edge: row (warp, slot, j) holds
//                        edge j of the SUB checks that warp processes together, so the check pass
//                        reads R_old and writes R_new with conflict-free `[base + j*128]` accesses
//   tables               check pass: byte offset of the T row of each edge, [slot][j/4][h][4]
//                        variable pass: byte address of the R element of each edge, [entry][h]
// and the channel values stay in registers (static variable slots), as in LANE16.
// Per edge the check pass does 2 loads + 1 store + ~8 ALU-pipe ops (no message reconstruction,
// no sign-word shifting: R_new = ((j==argmin) ? min2 : min1) ^ parity ^ sign(S_j)), the variable
// pass one table load, one message load and one FADD.  With G = 16 Test.cpp's code fits one SM
// (T 37 KB + R 135 KB); with G = 1 (tables in global memory) an N = 8192 code does.
// Same arithmetic contract as above: Q = P - R = -(T + R_old); T = (-y) - R_1 - R_2 ... in
// ascending-row order.
// ---------------------------------------------------------------------------------------
constexpr int kGrpMaxVS = 16;  // variable slots per thread (channel values in registers)
constexpr int kGrpMaxCS = 16;  // check slots per thread

struct GroupParams {
    const uint32_t* __restrict__ cn_tab;      // [W][cn_stride] quads: [slot][jq][h][4] T-row byte offsets
    const uint32_t* __restrict__ vn_tab;      // [W][vn_stride] quads: [slot][kq][h][4] R element byte offsets (row*128 + h_e*G*4)
    const uint32_t* __restrict__ var_of_pos;  // [VS*NL] variable index, 0xffffffff = phantom
    const uint32_t* __restrict__ pos_of_var;  // [N]
    int M, N, K, W, CS, VS;
    int cn_stride, vn_stride;                 // uint32 words per warp
    int r_rows_per_warp;                      // sum of check-slot degrees
    int max_iter, early_term;
    const float* __restrict__ llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;
    unsigned int* counter;
    unsigned long long* counter64;            // work-queue head: next codeword index
    int refill_wait;                          // lane refill policy (see the kernel's loop top)
    int ngroups;
    uint8_t vdeg[kGrpMaxVS];
    uint8_t cdeg[kGrpMaxCS];
    int n_vclass;                             // Y_SMEM: runs of equal variable-slot degree
    uint8_t vclass_deg[kGrpMaxVS];
    uint8_t vclass_cnt[kGrpMaxVS];
};

// Message / posterior gathers go through the CTA's own shared window, or -- in the cluster kernel --
// through the cluster window (distributed shared memory: the row may live in a peer CTA).
__device__ __forceinline__ float ld_cluster_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
template <bool DSM>
__device__ __forceinline__ float ld_node_f32(uint32_t a) {
    if constexpr (DSM) return ld_cluster_f32(a);
    else return lds_f32(a);
}

// One check of exact degree D, straight-line: D x {T gather, R_old load, S = T + R_old}, running
// min1/min2 and sign parities, then D x {R_new = ((|S_j| == min1) ? min2 : min1) ^ parity ^ sign(S_j)}.
// (|S_j| == min1 picks the argmin; with a tie min2 == min1, so either choice is the same value.)
// Returns the row's syndrome bit.
template <int D, int SUB, bool TAB_SMEM, bool DSM = false, bool T16 = false>
__device__ __forceinline__ uint32_t grp_check(uint32_t tab, const uint32_t* __restrict__ gtab, uint32_t t_base,
                                               uint32_t rrow, uint32_t c4, int h) {
    constexpr int NQ = (D + 3) / 4;
    uint32_t ent[(T16 ? (D + 7) / 8 * 8 : NQ * 4)];
    if constexpr (T16) {
        // 16-bit entries (T row index): eight per LDS.128, address = index * row bytes + base in one IMAD
        static_assert(!DSM, "16-bit tables address the CTA's own shared memory");
        constexpr uint32_t ROWB = 128u / SUB;
        const uint32_t base = t_base + c4;
#pragma unroll
        for (int jo = 0; jo < (D + 7) / 8; ++jo) {
            uint4 o;
            if (TAB_SMEM) o = lds_u128(tab + (uint32_t)(jo * SUB + h) * 16u);
            else o = __ldg(reinterpret_cast<const uint4*>(gtab + tab) + (jo * SUB + h));
            const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ent[jo * 8 + 2 * i] = (w[i] & 0xffffu) * ROWB + base;
                ent[jo * 8 + 2 * i + 1] = (w[i] >> 16) * ROWB + base;
            }
        }
    } else {
#pragma unroll
        for (int jq = 0; jq < NQ; ++jq) {
            uint4 o;
            if (TAB_SMEM) o = lds_u128(tab + (uint32_t)(jq * SUB + h) * 16u);
            else {
                o = __ldg(reinterpret_cast<const uint4*>(gtab + tab) + (jq * SUB + h));
                o.x += t_base; o.y += t_base; o.z += t_base; o.w += t_base;
            }
            ent[jq * 4 + 0] = o.x; ent[jq * 4 + 1] = o.y; ent[jq * 4 + 2] = o.z; ent[jq * 4 + 3] = o.w;
        }
    }
    float tv[D], S[D];
#pragma unroll
    for (int j = 0; j < D; ++j) tv[j] = ld_node_f32<DSM>(ent[j] + (T16 ? 0u : c4));
#pragma unroll
    for (int j = 0; j < D; ++j) S[j] = lds_f32(rrow + (uint32_t)j * 128u);
    uint32_t px = 0u, sx = 0u;
#pragma unroll
    for (int j = 0; j < D; ++j) S[j] = __fadd_rn(tv[j], S[j]);  // = -Q_j
    // sign parities, two edges per 3-input LOP3
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
    for (int j = 0; j < D; ++j) sts_f32(rrow + (uint32_t)j * 128u, rn[j]);
    return ((sx >> 31) ^ (uint32_t)D) & 1u;  // hard bit = !signbit(T)
}

template <int CNT, bool DSM = false>
__device__ __forceinline__ void grp_vn_part(const uint4 o, uint32_t c4, float& acc) {
    const uint32_t e[4] = {o.x, o.y, o.z, o.w};
    float r[CNT];
#pragma unroll
    for (int k = 0; k < CNT; ++k) r[k] = ld_node_f32<DSM>(e[k] + c4);
#pragma unroll
    for (int k = 0; k < CNT; ++k) acc = __fsub_rn(acc, r[k]);
}

// NS variable slots of exact degree D processed together (independent FADD chains interleaved):
// channel value from the Y array, D messages each through the quad table, new T stored.
template <int D, int NS, int SUB, bool TAB_SMEM>
__device__ __forceinline__ void grp_vn_slots(uint32_t& q, const uint32_t* __restrict__ gtab, uint32_t r_base, uint32_t& ta,
                                             uint32_t& ya, uint32_t t_stride, uint32_t c4, bool done) {
    constexpr int NQ = (D + 3) / 4;
    float acc[NS];
    uint32_t e[NS][NQ * 4];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        acc[i] = lds_f32(ya + (uint32_t)i * t_stride);
#pragma unroll
        for (int jq = 0; jq < NQ; ++jq) {
            uint4 o;
            if (TAB_SMEM) o = lds_u128(q + (uint32_t)((i * NQ + jq) * SUB) * 16u);
            else {
                o = __ldg(reinterpret_cast<const uint4*>(gtab + q) + (i * NQ + jq) * SUB);
                o.x += r_base; o.y += r_base; o.z += r_base; o.w += r_base;
            }
            e[i][jq * 4 + 0] = o.x; e[i][jq * 4 + 1] = o.y; e[i][jq * 4 + 2] = o.z; e[i][jq * 4 + 3] = o.w;
        }
    }
    float r[NS][D];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int k = 0; k < D; ++k) r[i][k] = lds_f32(e[i][k] + c4);
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
        for (int i = 0; i < NS; ++i) acc[i] = __fsub_rn(acc[i], r[i][k]);
#pragma unroll
    for (int i = 0; i < NS; ++i)
        if (!done) sts_f32(ta + (uint32_t)i * t_stride, acc[i]);
    q += (TAB_SMEM ? 16u : 4u) * (uint32_t)(NS * NQ * SUB);
    ta += (uint32_t)NS * t_stride;
    ya += (uint32_t)NS * t_stride;
}

// ---- degree profiles ---------------------------------------------------------------------------
// The slot degrees of a (code, G, W) combination are a handful of small integers.  When they are
// known at compile time the check and variable passes become straight-line code: no per-slot degree
// loads, switches or loop counters, and the scheduler can overlap the loads of several slots.
// GenericProfile reads the degrees from the kernel parameters; a static profile is picked by the
// host when the plan's degrees match it exactly (ldpc_b200.cu: launch_group).
struct GenericProfile {
    static constexpr bool kStatic = false;
};
// Test.cpp's code (802.16e rate 3/4B, z = 24: N = 576, M = 144) dealt over 48 node lanes
// (G = 8: 12 warps x 4, or G = 16: 24 warps x 2): 3 check slots, 12 variable slots.
struct ProfileWimax34B576 {
    static constexpr bool kStatic = true;
    static constexpr int CS = 3, VS = 12;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[3] = {15, 15, 14}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int i) { constexpr int d[12] = {6, 6, 6, 6, 3, 3, 3, 3, 3, 3, 2, 2}; return d[i]; }
};

// The same code dealt over 72 node lanes (G = 4: 9 warps x 8, three CTAs per SM): 2 check slots, 8 variable slots.
struct ProfileWimax34B576L72 {
    static constexpr bool kStatic = true;
    static constexpr int CS = 2, VS = 8;
    __host__ __device__ static constexpr int cdeg(int) { return 15; }
    __host__ __device__ static constexpr int vdeg(int i) { constexpr int d[8] = {6, 6, 6, 3, 3, 3, 3, 2}; return d[i]; }
};

// Regular (3,6) code with N = 8192 (BASELINE config 3) dealt over 1024 node lanes (G = 1, 32 warps).
struct ProfileRegular36N8192 {
    static constexpr bool kStatic = true;
    static constexpr int CS = 4, VS = 8;
    __host__ __device__ static constexpr int cdeg(int) { return 6; }
    __host__ __device__ static constexpr int vdeg(int) { return 3; }
};

// NS variable slots of exact degree D with the channel values in registers.
template <int D, int NS, int SUB, bool TAB_SMEM, bool T16 = false>
__device__ __forceinline__ void grp_vn_slots_reg(uint32_t& q, const uint32_t* __restrict__ gtab, uint32_t r_base, uint32_t& ta,
                                                 const float* yv, uint32_t t_stride, uint32_t c4, bool done, int h = 0) {
    constexpr int NQ = (D + 3) / 4;
    float acc[NS];
    uint32_t e[NS][T16 ? 8 : NQ * 4];
    if constexpr (T16) {
        // 16-bit entries (R element index in row units): one LDS.32 / .64 / .128 per slot and lane
        static_assert(D <= 8, "16-bit variable tables: degree <= 8");
        constexpr uint32_t ROWB = 128u / SUB;
        constexpr uint32_t SD = D <= 2 ? 4u : (D <= 4 ? 8u : 16u);
        const uint32_t base = r_base + c4;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            acc[i] = yv[i];
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            if constexpr (TAB_SMEM) {
                const uint32_t a = q + (uint32_t)(i * SUB + h) * SD;  // T16 layout: [slot][lane][SD bytes], q = slot base
                if constexpr (SD == 4u) w[0] = __float_as_uint(lds_f32(a));
                else if constexpr (SD == 8u) { const uint2 o = lds_u64(a); w[0] = o.x; w[1] = o.y; }
                else { const uint4 o = lds_u128(a); w[0] = o.x; w[1] = o.y; w[2] = o.z; w[3] = o.w; }
            } else {  // same layout in global memory, q = slot base in words
                const uint32_t* a = gtab + q;
                if constexpr (SD == 4u) w[0] = __ldg(a + (i * SUB + h));
                else if constexpr (SD == 8u) { const uint2 o = __ldg(reinterpret_cast<const uint2*>(a) + (i * SUB + h)); w[0] = o.x; w[1] = o.y; }
                else { const uint4 o = __ldg(reinterpret_cast<const uint4*>(a) + (i * SUB + h)); w[0] = o.x; w[1] = o.y; w[2] = o.z; w[3] = o.w; }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                e[i][2 * k] = (w[k] & 0xffffu) * ROWB + base;
                e[i][2 * k + 1] = (w[k] >> 16) * ROWB + base;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            acc[i] = yv[i];
#pragma unroll
            for (int jq = 0; jq < NQ; ++jq) {
                uint4 o;
                if (TAB_SMEM) o = lds_u128(q + (uint32_t)((i * NQ + jq) * SUB) * 16u);
                else {
                    o = __ldg(reinterpret_cast<const uint4*>(gtab + q) + (i * NQ + jq) * SUB);
                    o.x += r_base; o.y += r_base; o.z += r_base; o.w += r_base;
                }
                e[i][jq * 4 + 0] = o.x; e[i][jq * 4 + 1] = o.y; e[i][jq * 4 + 2] = o.z; e[i][jq * 4 + 3] = o.w;
            }
        }
    }
    float r[NS][D];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int k = 0; k < D; ++k) r[i][k] = lds_f32(e[i][k] + (T16 ? 0u : c4));
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
        for (int i = 0; i < NS; ++i) acc[i] = __fsub_rn(acc[i], r[i][k]);
#pragma unroll
    for (int i = 0; i < NS; ++i)
        if (!done) sts_f32(ta + (uint32_t)i * t_stride, acc[i]);
    if constexpr (T16) q += (uint32_t)NS * SUB * (D <= 2 ? 4u : (D <= 4 ? 8u : 16u)) / (TAB_SMEM ? 1u : 4u);
    else q += (TAB_SMEM ? 16u : 4u) * (uint32_t)(NS * NQ * SUB);
    ta += (uint32_t)NS * t_stride;
}

// Static variable pass: runs of equal-degree slots, up to 4 (degree <= 3) or 2 slots in flight.
template <class P, int S0, int SUB, bool TAB_SMEM, bool T16 = false>
__device__ __forceinline__ void grp_vn_static(uint32_t& q, const uint32_t* __restrict__ gtab, uint32_t r_base, uint32_t& ta,
                                              const float* yn, uint32_t t_stride, uint32_t c4, bool done, int h = 0) {
    if constexpr (S0 < P::VS) {
        constexpr int D = P::vdeg(S0);
        constexpr int same2 = (S0 + 1 < P::VS) && P::vdeg(S0 + 1 < P::VS ? S0 + 1 : S0) == D;
        constexpr int same4 = same2 && (S0 + 3 < P::VS) && P::vdeg(S0 + 2 < P::VS ? S0 + 2 : S0) == D &&
                              P::vdeg(S0 + 3 < P::VS ? S0 + 3 : S0) == D;
        constexpr int NS = (same4 && D <= 3) ? 4 : (same2 ? 2 : 1);
        if constexpr (D > 0) grp_vn_slots_reg<D, NS, SUB, TAB_SMEM, T16>(q, gtab, r_base, ta, yn + S0, t_stride, c4, done, h);
        else ta += (uint32_t)NS * t_stride;
        grp_vn_static<P, S0 + NS, SUB, TAB_SMEM, T16>(q, gtab, r_base, ta, yn, t_stride, c4, done, h);
    }
}

// Static check pass: one straight-line check per slot.
template <class P, int CS0, int SUB, bool TAB_SMEM, bool T16 = false>
__device__ __forceinline__ uint32_t grp_cn_static(uint32_t tab, const uint32_t* __restrict__ gtab, uint32_t t_base, uint32_t rrow,
                                                   uint32_t c4, int h) {
    if constexpr (CS0 < P::CS) {
        constexpr int D = P::cdeg(CS0);
        const uint32_t u = grp_check<D, SUB, TAB_SMEM, false, T16>(tab, gtab, t_base, rrow, c4, h);
        constexpr uint32_t adv = T16 ? (TAB_SMEM ? 16u : 4u) * (uint32_t)(((D + 7) >> 3) * SUB)
                                     : (TAB_SMEM ? 16u : 4u) * (uint32_t)(((D + 3) >> 2) * SUB);
        return u | grp_cn_static<P, CS0 + 1, SUB, TAB_SMEM, T16>(tab + adv, gtab, t_base, rrow + (uint32_t)D * 128u, c4, h);
    } else {
        return 0u;
    }
}

template <int G, int DMAX, bool TAB_SMEM, int MAX_THREADS, bool Y_SMEM, class PROF = GenericProfile, bool T16 = false>
__global__ void __launch_bounds__(MAX_THREADS, (MAX_THREADS <= 288 ? 3 : (MAX_THREADS <= 384 ? 2 : 1))) ldpc_ms_group_kernel(const __grid_constant__ GroupParams p) {
    constexpr int SUB = 32 / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_flag[2][32];

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int c = lane & (G - 1), h = lane / G;
    const int W = p.W, CS = p.CS, VS = p.VS;
    const int NL = W * SUB;
    const int PD = VS * NL;                         // dummy T row (= -inf)
    const int RD = W * p.r_rows_per_warp;           // dummy R row (= 0)

    const uint32_t t_base = smem_u32(smem_raw);
    const uint32_t t_bytes = ((uint32_t)(PD + 1) * G * 4 + 127u) & ~127u;
    const uint32_t y_base = t_base + t_bytes;                                        // channel values (Y_SMEM)
    const uint32_t r_base = y_base + (Y_SMEM ? t_bytes : 0u);                        // 128-byte aligned rows
    const uint32_t cn_base = r_base + (uint32_t)(RD + 1) * 128;
    const uint32_t vn_base = cn_base + (TAB_SMEM ? (uint32_t)W * p.cn_stride * 4 : 0u);
    if (TAB_SMEM) {
        uint32_t* cn_s = reinterpret_cast<uint32_t*>(smem_raw + (cn_base - t_base));
        uint32_t* vn_s = reinterpret_cast<uint32_t*>(smem_raw + (vn_base - t_base));
        // 32-bit entries are stored with the region base added (an entry IS a shared address);
        // 16-bit entries (T16) are row indices and stay as they are
        for (int i = threadIdx.x; i < W * p.cn_stride; i += blockDim.x) cn_s[i] = __ldg(p.cn_tab + i) + (T16 ? 0u : t_base);
        for (int i = threadIdx.x; i < W * p.vn_stride; i += blockDim.x) vn_s[i] = __ldg(p.vn_tab + i) + (T16 ? 0u : r_base);
    }
    const uint32_t c4 = (uint32_t)c * 4u;
    if (warp == 0) {
        if (lane < G) sts_f32(t_base + (uint32_t)PD * G * 4 + lane * 4, -INFINITY);
        sts_f32(r_base + (uint32_t)RD * 128 + lane * 4, 0.0f);
    }
    // this thread's own T elements: position (s*NL + warp*SUB + h), codeword c -> contiguous in tid
    const uint32_t t_own = t_base + (uint32_t)threadIdx.x * 4u;
    const uint32_t t_stride = (uint32_t)NL * G * 4u;
    const uint32_t r_own = r_base + (uint32_t)warp * p.r_rows_per_warp * 128u + (uint32_t)lane * 4u;
    // table cursors (shared addresses or global word offsets)
    const uint32_t cn_w = TAB_SMEM ? cn_base + (uint32_t)warp * p.cn_stride * 4u : (uint32_t)warp * p.cn_stride;
    const uint32_t vn_w = TAB_SMEM ? vn_base + (uint32_t)warp * p.vn_stride * 4u : (uint32_t)warp * p.vn_stride;

    // ---- per-lane decode state.  Every thread of a codeword lane c keeps the same copy. ---------------
    float yn[kGrpMaxVS];
#pragma unroll
    for (int s = 0; s < kGrpMaxVS; ++s) yn[s] = -1.0f;
    long long cw = -1;
    bool live = false;     // decoding a codeword
    bool done = false;     // live and finished (converged or at the cap): retires at the next loop top
    bool loading = false;  // channel values in flight (cp.async into this lane's own T elements)
    int it = 0, my_iters = 0;
    __shared__ long long s_cw[32];

    // passes -----------------------------------------------------------------------------------------
    auto cn_pass = [&]() -> uint32_t {
        // ---- check-node pass
        uint32_t unsat = 0u;
        if constexpr (PROF::kStatic) {
            unsat = grp_cn_static<PROF, 0, SUB, TAB_SMEM, T16>(cn_w, p.cn_tab, t_base, r_own, c4, h);
        } else {
            uint32_t tab = cn_w;     // quads for this warp, slot by slot
            uint32_t rrow = r_own;   // this lane's R column, row by row
            for (int cs = 0; cs < CS; ++cs) {
                const int dc = p.cdeg[cs];
#define GRP_CASE(D) case D: unsat |= grp_check<D, SUB, TAB_SMEM>(tab, p.cn_tab, t_base, rrow, c4, h); break;
                switch (dc) {
                    GRP_CASE(1) GRP_CASE(2) GRP_CASE(3) GRP_CASE(4) GRP_CASE(5) GRP_CASE(6) GRP_CASE(7) GRP_CASE(8)
                    default:
                        if constexpr (DMAX > 8) {
                            switch (dc) {
                                GRP_CASE(9) GRP_CASE(10) GRP_CASE(11) GRP_CASE(12)
                                GRP_CASE(13) GRP_CASE(14) GRP_CASE(15) GRP_CASE(16)
                                default: break;
                            }
                        }
                        break;
                }
#undef GRP_CASE
                rrow += (uint32_t)dc * 128u;
                tab += (TAB_SMEM ? 16u : 4u) * (uint32_t)(((dc + 3) >> 2) * SUB);
            }
        }
        return unsat;
    };
    auto vn_pass = [&]() {
        const bool done_mask = !(live && !done);  // only a live, unfinished word moves its posterior
        // ---- variable-node pass: T = (-y) - R_e1 - R_e2 ... in ascending-row order
        if constexpr (Y_SMEM) {
            // channel values in shared memory: the slot loop is dynamic, two equal-degree slots at a
            // time (slots are degree-sorted into at most kGrpMaxCls runs)
            uint32_t q = vn_w + (TAB_SMEM ? (uint32_t)h * 16u : (uint32_t)h * 4u);
            uint32_t ta = t_own, ya = t_own + t_bytes;
            for (int ci = 0; ci < p.n_vclass; ++ci) {
                const int d = p.vclass_deg[ci];
                int n = p.vclass_cnt[ci];
#define GRP_VCASE(D)                                                                                              \
case D:                                                                                                       \
    if (D <= 3)                                                                                               \
        for (; n >= 4; n -= 4) grp_vn_slots<D, (D <= 3 ? 4 : 1), SUB, TAB_SMEM>(q, p.vn_tab, r_base, ta, ya, t_stride, c4, done_mask); \
    for (; n >= 2; n -= 2) grp_vn_slots<D, 2, SUB, TAB_SMEM>(q, p.vn_tab, r_base, ta, ya, t_stride, c4, done_mask); \
    if (n) grp_vn_slots<D, 1, SUB, TAB_SMEM>(q, p.vn_tab, r_base, ta, ya, t_stride, c4, done_mask);                 \
    break;
                switch (d) {
                    GRP_VCASE(1) GRP_VCASE(2) GRP_VCASE(3) GRP_VCASE(4) GRP_VCASE(5) GRP_VCASE(6) GRP_VCASE(7) GRP_VCASE(8)
                    GRP_VCASE(9) GRP_VCASE(10) GRP_VCASE(11) GRP_VCASE(12)
                    default:
                        if (d == 0) { ta += (uint32_t)n * t_stride; ya += (uint32_t)n * t_stride; }
                        break;
                }
#undef GRP_VCASE
            }
        } else if constexpr (PROF::kStatic) {
            uint32_t q = vn_w + (T16 ? 0u : (TAB_SMEM ? (uint32_t)h * 16u : (uint32_t)h * 4u));  // T16: lane offset per slot
            uint32_t ta = t_own;
            grp_vn_static<PROF, 0, SUB, TAB_SMEM, T16>(q, p.vn_tab, r_base, ta, yn, t_stride, c4, done_mask, h);
        } else {
            uint32_t q = vn_w + (TAB_SMEM ? (uint32_t)h * 16u : (uint32_t)h * 4u);  // quads [slot][kq][h][4]
            auto next_quad = [&]() {
                uint4 o;
                if (TAB_SMEM) { o = lds_u128(q); q += SUB * 16; }
                else {
                    o = __ldg(reinterpret_cast<const uint4*>(p.vn_tab + q));
                    q += SUB * 4;
                    o.x += r_base; o.y += r_base; o.z += r_base; o.w += r_base;
                }
                return o;
            };
#pragma unroll
            for (int s = 0; s < kGrpMaxVS; ++s) {
                if (s < VS) {
                    int d = p.vdeg[s];
                    float acc = yn[s];
                    // straight-line code per degree (slots are degree-sorted, so consecutive slots
                    // take the same case); degrees above 8 fall back to a quad loop
                    switch (d) {
                        case 1: grp_vn_part<1>(next_quad(), c4, acc); break;
                        case 2: grp_vn_part<2>(next_quad(), c4, acc); break;
                        case 3: grp_vn_part<3>(next_quad(), c4, acc); break;
                        case 4: grp_vn_part<4>(next_quad(), c4, acc); break;
                        case 5: grp_vn_part<4>(next_quad(), c4, acc); grp_vn_part<1>(next_quad(), c4, acc); break;
                        case 6: grp_vn_part<4>(next_quad(), c4, acc); grp_vn_part<2>(next_quad(), c4, acc); break;
                        case 7: grp_vn_part<4>(next_quad(), c4, acc); grp_vn_part<3>(next_quad(), c4, acc); break;
                        case 8: grp_vn_part<4>(next_quad(), c4, acc); grp_vn_part<4>(next_quad(), c4, acc); break;
                        default:
#pragma unroll 1
                            for (; d >= 4; d -= 4) grp_vn_part<4>(next_quad(), c4, acc);
                            if (d > 0) {
                                const uint4 o = next_quad();
                                if (d == 1) grp_vn_part<1>(o, c4, acc);
                                else if (d == 2) grp_vn_part<2>(o, c4, acc);
                                else grp_vn_part<3>(o, c4, acc);
                            }
                            break;
                    }
                    if (!done_mask) sts_f32(t_own + (uint32_t)s * t_stride, acc);
                }
            }
        }
    };
    auto emit = [&](bool sel) {
        // ---- outputs (toChar, decodeCL.c:188-199): bit n = !(P > 0) = !signbit(T); node lanes share the bytes
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = warp * SUB + h; b < KB; b += NL) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.K) {
                        const uint32_t pos = __ldg(p.pos_of_var + n);
                        v |= ((~__float_as_uint(lds_f32(t_base + pos * (G * 4) + c4))) >> 31) << t;
                    }
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
                    if (n < p.N) {
                        const uint32_t pos = __ldg(p.pos_of_var + n);
                        v |= ((~__float_as_uint(lds_f32(t_base + pos * (G * 4) + c4))) >> 31) << t;
                    }
                }
                if (sel) p.hard[(size_t)cw * NB + b] = (uint8_t)v;
            }
        }
        if (p.post && sel) {
            for (int n = warp * SUB + h; n < p.N; n += NL)
                p.post[(size_t)cw * p.N + n] = -lds_f32(t_base + __ldg(p.pos_of_var + n) * (G * 4) + c4);
        }
        if (p.iters && warp == 0 && h == 0 && sel) p.iters[cw] = my_iters;
    };
    // fetch the next codeword of lanes selected by `want` and start its channel values on their way
    auto fetch = [&](bool want) {
        if (warp == 0 && h == 0 && want) s_cw[c] = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(p.counter64), 1ull);
        __syncthreads();  // also: every read of the retiring lanes' T (emit) is complete
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
                        const uint32_t dst = t_own + (uint32_t)s * t_stride;
                        if (v != 0xffffffffu) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + v) : "memory");
                        else sts_f32(dst, 1.0f);
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // lanes whose values have landed start decoding: T = -y (canonical zero), R = 0 (decodeInitMS)
    auto start_loaded = [&]() {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (loading) {
#pragma unroll
            for (int s = 0; s < kGrpMaxVS; ++s) {
                if (s < VS) {
                    const float y = lds_f32(t_own + (uint32_t)s * t_stride);
                    yn[s] = __fadd_rn(-y, 0.0f);
                    sts_f32(t_own + (uint32_t)s * t_stride, yn[s]);
                    if (Y_SMEM) sts_f32(t_own + t_bytes + (uint32_t)s * t_stride, yn[s]);
                }
            }
            for (int r = 0; r < p.r_rows_per_warp; ++r) sts_f32(r_own + (uint32_t)r * 128u, 0.0f);
            loading = false; live = true; done = false; it = 0;
        }
        __syncthreads();
    };

    if (warp == 0) { s_flag[0][lane] = 0u; s_flag[1][lane] = 0u; }
    __syncthreads();
    if constexpr (G == 1) {
        // One codeword per CTA: nothing to refill around -- plain word-at-a-time loop with direct loads.
        for (;;) {
            if (threadIdx.x == 0) s_cw[0] = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(p.counter64), 1ull);
            __syncthreads();
            cw = s_cw[0];
            if (cw >= p.ncw) break;
            const float* src = p.llr + (size_t)cw * p.N;
#pragma unroll
            for (int s = 0; s < kGrpMaxVS; ++s) {
                if (s < VS) {
                    const uint32_t v = __ldg(p.var_of_pos + s * NL + warp * SUB + h);
                    const float y = (v != 0xffffffffu) ? __ldg(src + v) : 1.0f;
                    yn[s] = __fadd_rn(-y, 0.0f);
                    sts_f32(t_own + (uint32_t)s * t_stride, yn[s]);
                    if (Y_SMEM) sts_f32(t_own + t_bytes + (uint32_t)s * t_stride, yn[s]);
                }
            }
            for (int r = 0; r < p.r_rows_per_warp; ++r) sts_f32(r_own + (uint32_t)r * 128u, 0.0f);
            if (warp == 0) { s_flag[0][lane] = 0u; s_flag[1][lane] = 0u; }
            live = true; done = false; it = 0;
            __syncthreads();
            for (;;) {
                const uint32_t unsat = cn_pass();
                const bool check = p.early_term && it >= 1;
                if (check && unsat) s_flag[it & 1][0] = 1u;
                __syncthreads();
                if (check && s_flag[it & 1][0] == 0u) { done = true; my_iters = it; break; }  // CTA-uniform
                if (warp == 0) s_flag[(it + 1) & 1][lane] = 0u;
                vn_pass();
                ++it;
                if (it == p.max_iter) { done = true; my_iters = it; break; }
                __syncthreads();
            }
            __syncthreads();
            emit(true);
            __syncthreads();
        }
        return;
    }
    fetch(true);
    uint32_t ph = 0;
    for (;;) {
        // ---- loop top: start words whose values arrived, retire finished words and refill their lanes.
        // (A warp holds every codeword lane and all threads of a lane agree, so these votes are CTA-uniform.)
        if (__any_sync(0xffffffffu, loading)) start_loaded();  // refilled at an earlier loop top
        const bool retire = live && done;
        if (__any_sync(0xffffffffu, retire)) {
            emit(retire);
            fetch(retire);
            // refill_wait = 0: a lane refilled now starts at the next loop top, after one more iteration of
            // the others (its load overlaps them); when nothing else is running there is nothing to
            // overlap with and it starts at once.  refill_wait = 1: always wait for the values now.
            if ((p.refill_wait || !__any_sync(0xffffffffu, live)) && __any_sync(0xffffffffu, loading)) start_loaded();
        }
        if (!__any_sync(0xffffffffu, live || loading)) break;

        // ---- check-node pass + syndrome of the previous posterior
        const uint32_t unsat = cn_pass();
        const bool check = p.early_term && it >= 1 && live && !done;
        if (check && unsat) s_flag[ph & 1][c] = 1u;  // same-value race, benign
        __syncthreads();
        if (check && s_flag[ph & 1][c] == 0u) { done = true; my_iters = it; }
        if (warp == 0) s_flag[(ph + 1) & 1][lane] = 0u;
        ++ph;

        // ---- variable-node pass (posterior of finished / idle lanes is frozen)
        vn_pass();
        if (live && !done) {
            ++it;
            if (it == p.max_iter) { done = true; my_iters = it; }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Synthetic BPSK + AWGN channel (Coder::test, reference MyLdpc.cpp:1061-1078): bit 0 -> +1,
// bit 1 -> -1, plus sigma * N(0,1).  Counter-based: element i of the stream depends only on
// (seed, i), so any shard of any GPU generates the same floats for the same codeword index.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static __global__ void __launch_bounds__(256) ldpc_synth_llr_kernel(float* __restrict__ out, long long total, int N,
                                                             float sigma, uint64_t seed,
                                                             const uint8_t* __restrict__ bits, long long first_index) {
    const int NB = (N + 7) >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const uint64_t h = splitmix64(seed ^ splitmix64((uint64_t)(i + first_index)));
        const float u1 = ((float)((uint32_t)(h >> 40)) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
        const float u2 = (float)((uint32_t)(h & 0xFFFFFFu)) * (1.0f / 16777216.0f);     // [0,1)
        const float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
        float s = 1.0f;
        if (bits) {
            const long long cw = i / N;
            const int n = (int)(i - cw * N);
            if ((bits[(size_t)cw * NB + (n >> 3)] >> (n & 7)) & 1) s = -1.0f;
        }
        out[i] = __fmaf_rn(sigma, g, s);
    }
}

// ---------------------------------------------------------------------------------------
// Shared-memory bandwidth probe (measurement aid for bench.py's on-chip roofline): every
// thread streams conflict-free 16-byte loads from a 128 KB tile.  Bytes moved per launch =
// gridDim.x * blockDim.x * 16 * 8 * loops.
// ---------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(1024, 1) ldpc_smem_probe_kernel(uint32_t* __restrict__ sink, int loops) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* buf = reinterpret_cast<uint4*>(smem_raw);
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int l = 0; l < loops; ++l) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 v = buf[(threadIdx.x + k * 1024 + l * 32) & 8191];
            acc.x += v.x; acc.y ^= v.y; acc.z += v.z; acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345u) sink[blockIdx.x] = acc.x;  // keep the loads alive
}

}  // namespace ldpc_b200
