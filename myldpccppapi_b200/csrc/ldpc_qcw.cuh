// ldpc_qcw.cuh -- flooding min-sum for short quasi-cyclic codes in the EARLY-TERMINATION regime: ONE WARP = ONE
// CODEWORD, lane = row (check pass) / column (variable pass) of a z x z block, z <= 32.
//
// ldpc_ms_qc_kernel (ldpc_qc.cuh) decodes G words per CTA in lockstep.  When words converge after 3-5 iterations
// (BASELINE config 4, the region a decoder operates in) it spends most of its time outside the two passes
// (profiles/r02_ring_kernel.md, r02_et_kernel.md): CTA-wide barriers around every retire/start event, the leaving word's
// bits packed from shared memory, the starting word's messages cleared, and one extra loop trip per word because the
// syndrome of iteration t is only seen by the check pass of t + 1.  A CTA-level variant with an incremental syndrome
// (built, measured, removed: profiles/r02_et_kernel.md) took the extra trip away but not the barriers and paid for it
// in divergent shared-memory atomics.
//
// Here nothing is shared between codewords, so nothing is synchronised between them:
//   * a warp owns one codeword: T[24][2z] (negated posterior, each block column stored twice so that the cyclic
//     wrap is a plain offset) and R[E][z] (one message row per non-zero circulant) in its own slice of shared memory.
//     Check pass: lane r handles row r of every block row; variable pass: lane c handles column c of every block
//     column, the messages of its edges sit at row (c - s) mod z of their circulant -- the lanes below s read z rows
//     further (one select per edge).  The circulant structure of the 802.16e codes is COMPILED IN (qcw_tables.h,
//     generated from the standard's base matrices): every address is lane + immediate, there are no index tables.
//     __syncwarp between the passes, no CTA barrier anywhere in the loop.
//   * EXCLUDE-SELF MINIMA BY PREFIX / SUFFIX.  R_e = +-min(1000, min over the other edges |Q|) is taken as
//     min3(prefix, neighbour, suffix) over pairs of edges: 2 three-input min/max instructions per edge instead of
//     min1/min2 tracking plus compare-and-select (4.5), the same value bit for bit; the sign is applied by one
//     multiplication with +-1.0 on the FMA pipe (with min1/min2 tracking the kernel was bound by the half-rate ALU
//     pipe; now by shared-memory wavefronts: 93 % of peak at the cap, profiles/r02_ncu_qcw_fixed40.txt).
//   * SYNDROME FROM PACKED BITS.  The variable pass ballots the sign of every posterior it writes: lane b keeps the
//     z hard bits of block column b in a register (a select per block column; keeping the words in shared memory
//     instead cost 24 wavefronts per iteration of a kernel that is bound by them: +5 %).  The syndrome of ALL z rows
//     of a block row is the XOR over its circulants of that word rotated by the shift -- one shuffle and one rotate
//     per CIRCULANT (88 for Test.cpp's code), not per edge, reduced by a butterfly.  A word is finished the moment its syndrome is clean after a
//     variable pass: it costs exactly `iters` trips (reference stop rule, MyLdpc.cpp:751-755).
//   * the leaving word's info bytes ARE those registers (toChar, decodeCL.c:188-199: three bytes per block column);
//     the starting word's channel values were loaded into registers while the previous word was decoded; its
//     messages are not cleared (decodeInitMS: R = 0): the first check pass is a specialised copy that takes R = 0 as
//     known and loads no messages (2 E wavefronts per word less: 4 dB 0.398 -> 0.374 ms per 65,536 words).
//   * a warp that finishes takes the next word from the global queue on its own; the other warps never notice.
// Price: z of 32 lanes work (z = 24: 75 %); the host picks the kernel per launch (ldpc_b200.cu).
// Arithmetic and outputs are those of ldpc_ms_qc_kernel: bit-exact with Coder::decodeCPU (MyLdpc.cpp:684-784).
#pragma once
#include "ldpc_qc.cuh"
#include "qcw_tables.h"

namespace ldpc_b200 {

constexpr int kQcwMaxWarps = 16;  // codewords in flight per SM for z = 24 (13,184 B each; 17 would fit, but a fifth warp on one
                                  // scheduler caps the kernel at 96 registers)

template <class C>
struct QcwProfile : C {
    static constexpr int Z = C::Z, NB = C::NB, MB = C::MB, E = C::E;
    static constexpr uint32_t ZB = (uint32_t)Z * 4u;
    static constexpr uint32_t T_BYTES = (uint32_t)NB * 2u * ZB;   // T[NB][2z]: every block column twice
    static constexpr uint32_t WARP_BYTES = T_BYTES + (uint32_t)E * ZB + 128u;   // R[E][z] behind T (+ slack: idle lanes read up to 32 B past the last circulant)
    static constexpr int SEG = C::DMAX <= 8 ? 8 : (C::DMAX <= 16 ? 16 : 32);   // lanes per block row in the syndrome rounds
    static constexpr int ROUNDS = (MB * SEG + 31) / 32;
};

struct QcwParams {
    int N, K;
    int max_iter, early_term;
    const float* __restrict__ llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;
    unsigned long long* counter64;
    const unsigned long long* avail;   // streamed input (see QcParams)
    int* status;
    unsigned long long wait_ns;
    int fmt;                           // format of llr: 0 fp32, 1 binary16, 2 int8 (ldpc_b200_decode_host_packed), widened at the load
    float scale;
    const uint32_t* syn_tab;           // [ROUNDS][32]: (block column << 8) | shift of the circulant a lane rotates, 0xffffffff = none
};

#ifdef LDPC_QCW_DEVICE   // the kernel: only the unit that instantiates it (k_qcw.cu)

// One block row: lane = row.  refreshRMS (decodeCL.c:126-147): S_j = T + R_old = -Q_j, R_new_j = sign * min(1000, min of
// the other |S|), sign(R_new_j) = parity of the other negative Q's = parity ^ 1 ^ signbit(S_j).
// FIRST: the word's first check pass.  decodeInitMS left R = 0, so S_j = T + 0 = T exactly (T is never -0): the
// messages are neither cleared when a word starts nor loaded here -- 2 E wavefronts per word less, of ~4.5 E per trip,
// which is what counts when a word leaves after three to five trips.
template <class P, int I, bool FIRST>
__device__ __forceinline__ void qcw_check(uint32_t la, bool act) {
    constexpr int D = P::cdeg(I), E0 = P::e0(I);
    float S[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float t = lds_f32(la + (uint32_t)(P::cbc(E0 + j) * 2 * P::Z + P::csh(E0 + j)) * 4u);
        if constexpr (FIRST) {
            S[j] = t;
        } else {
            const float r = lds_f32(la + P::T_BYTES + (uint32_t)(E0 + j) * P::ZB);
            S[j] = __fadd_rn(t, r);
        }
    }
    uint32_t px = 0u;
#pragma unroll
    for (int j = 0; j + 1 < D; j += 2) px = px ^ __float_as_uint(S[j]) ^ __float_as_uint(S[j + 1]);
    if (D & 1) px ^= __float_as_uint(S[D - 1]);
    // exclude-self minima by prefix / suffix, sign by multiplication (ldpc_kernels.cuh); every message is stored right
    // behind its computation (measured 4 % faster here than storing them all at the end)
    ms_new_messages_each<D>(S, px, [&](int j, float rn) {
        if (act) sts_f32(la + P::T_BYTES + (uint32_t)(E0 + j) * P::ZB, rn);
    });
}

template <class P, int I, bool FIRST>
__device__ __forceinline__ void qcw_cn(uint32_t la, bool act) {
    if constexpr (I < P::MB) {
        qcw_check<P, I, FIRST>(la, act);
        qcw_cn<P, I + 1, FIRST>(la, act);
    }
}

// One block column: lane = column.  refreshPostPMS (decodeCL.c:149-171): T = (-y) - R_1 - R_2 ... in ascending-row
// order; edge k's message sits at row (lane - s) mod z of its circulant: lane*4 + (circulant - s*4), plus z*4 for the
// lanes below s (laz = la + z*4).  The z hard bits of the column (bit = !signbit(T)) go to lane B's register.
template <class P, int B>
__device__ __forceinline__ void qcw_vn(uint32_t la, uint32_t laz, uint32_t lane, const float* yn, bool act, uint32_t& hb) {
    if constexpr (B < P::NB) {
        constexpr int D = P::vdeg(B), V0 = P::v0(B);
        float r[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const uint32_t s = (uint32_t)P::csh(P::ve(V0 + k));
            r[k] = lds_f32((lane < s ? laz : la) + (P::T_BYTES + (uint32_t)P::ve(V0 + k) * P::ZB - s * 4u));
        }
        float acc = yn[B];
#pragma unroll
        for (int k = 0; k < D; ++k) acc = __fsub_rn(acc, r[k]);
        if (act) {
            sts_f32(la + (uint32_t)B * 2u * P::ZB, acc);
            sts_f32(la + (uint32_t)B * 2u * P::ZB + P::ZB, acc);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, act && (__float_as_uint(acc) >> 31) == 0u);
        if (lane == (uint32_t)B) hb = bal;   // (a register select: the kernel is bound by shared-memory wavefronts, not by the ALU pipe)
        qcw_vn<P, B + 1>(la, laz, lane, yn, act, hb);
    }
}

// PACKED: p.llr holds float16 / int8 values (ldpc_b200_decode_host_packed), widened at the load -- a separate instantiation,
// so that the fp32 kernel stays exactly the code that was tuned.
template <class P, bool PACKED = false>
__global__ void __launch_bounds__(kQcwMaxWarps * 32, 1) ldpc_ms_qcw_kernel(const __grid_constant__ QcwParams p) {
    constexpr int Z = P::Z, NB = P::NB;
    constexpr uint32_t ZB = P::ZB;
    constexpr uint32_t ZMASK = Z == 32 ? 0xffffffffu : ((1u << (Z & 31)) - 1u);
    static_assert(Z % 8 == 0 && Z <= 32, "a byte of hard bits sits in one block column; a lane per row");
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const uint32_t lane = threadIdx.x & 31u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const uint32_t wb = smem_u32(smem_raw) + (uint32_t)warp * P::WARP_BYTES;
    const uint32_t la = wb + lane * 4u;
    const bool act = lane < (uint32_t)Z;
    const int KB = (p.K + 7) >> 3, NB8 = (p.N + 7) >> 3;

    uint32_t syn[P::ROUNDS];
#pragma unroll
    for (int r = 0; r < P::ROUNDS; ++r) syn[r] = p.syn_tab[r * 32 + (int)lane];

    // lane 0 takes a ticket from the work queue; it is broadcast only when it is consumed, one word later, so the
    // atomic's latency is never waited for
    auto claim = [&]() -> long long { return lane == 0 ? (long long)atomicAdd(p.counter64, 1ull) : 0ll; };
    auto landed = [&](long long w) -> bool {   // streamed input: wait (bounded) until word w is in device memory
        return w < p.ncw && (!p.avail || qc_wait_input(p.avail, w, true, p.status, p.wait_ns));
    };
    uint32_t hb = 0u;   // lane b: the z hard bits of block column b (bit = !signbit(T)) after the last variable pass
    // hard-bit word of the block column that holds byte b of the codeword, shifted to that byte
    auto byte_of = [&](int b) -> uint32_t {
        const int bc = (8 * b) / Z;
        return (__shfl_sync(0xffffffffu, hb, bc < NB ? bc : 0) >> (8 * b - bc * Z)) & 0xffu;
    };

    // fp32 input: the channel values of the word decoded next are loaded into registers (yq) while the current word is
    // decoded -- a starting word finds them there, no global-memory latency between two words of a warp.  Packed input
    // (its widening needs the registers): pulled into L2 one word ahead (a lane per 128-byte line), loaded at the start.
    float yn[NB], yq[PACKED ? 1 : NB];
    auto load_y = [&](long long w) {
        if constexpr (!PACKED) {
            const float* src = p.llr + (size_t)w * p.N + lane;
#pragma unroll
            for (int b = 0; b < NB; ++b) yq[b] = act ? __ldg(src + b * Z) : 0.0f;
        } else {
            const int esz = p.fmt == 1 ? 2 : 1;   // bytes per channel value in p.llr
            const char* src = reinterpret_cast<const char*>(p.llr) + (size_t)w * p.N * esz;
            for (int o0 = 0; o0 < p.N * esz; o0 += 32 * 128) {   // (same trip count in every lane)
                const int o = o0 + (int)lane * 128;
                if (o < p.N * esz) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
            }
        }
    };
    long long wn = __shfl_sync(0xffffffffu, claim(), 0);   // the word decoded next
    long long tick = claim();                                // lane 0: the word after it, claimed, not touched yet
    if (!landed(wn)) wn = p.ncw;
    if (wn < p.ncw) load_y(wn);
    for (;;) {
        const long long w = wn;
        if (w >= p.ncw) break;
        // ---- start word w (decodeInitMS, decodeCL.c:113-124): T = -y (canonical zero), R = 0
        if constexpr (!PACKED) {
#pragma unroll
            for (int b = 0; b < NB; ++b) yn[b] = yq[b];
        } else {   // packed host formats, widened here
            const size_t i0 = (size_t)w * p.N + lane;
#pragma unroll
            for (int b = 0; b < NB; ++b) yn[b] = act ? llr_at(p.llr, p.fmt, p.scale, i0 + (size_t)(b * Z)) : 0.0f;
        }
        // (R = 0 is not stored: the first check pass below takes it as known)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            yn[b] = __fadd_rn(-yn[b], 0.0f);
            if (act) {
                sts_f32(la + (uint32_t)b * 2u * ZB, yn[b]);
                sts_f32(la + (uint32_t)b * 2u * ZB + ZB, yn[b]);
            }
        }
        // the next word's values travel while this one is decoded; the queue ticket after it is consumed at the next start
        wn = __shfl_sync(0xffffffffu, tick, 0);
        if (!landed(wn)) wn = p.ncw;
        if (wn < p.ncw) load_y(wn);
        tick = claim();
        __syncwarp();

        int it = 0;
        for (;;) {
            if (it == 0) qcw_cn<P, 0, true>(la, act);
            else qcw_cn<P, 0, false>(la, act);
            __syncwarp();
            qcw_vn<P, 0>(la, la + ZB, lane, yn, act, hb);
            __syncwarp();
            ++it;
            if (it >= p.max_iter) break;
            if (p.early_term) {
                // checkResult (decodeCL.c:88-108): row r of block row i is the XOR over its circulants (bc, s) of bit
                // (r + s) mod z of column block bc -- the column's word rotated right by s.  A lane rotates one circulant
                // per round; a butterfly over the SEG lanes of a block row gives its z syndrome bits.
                uint32_t unsat = 0u;
#pragma unroll
                for (int r = 0; r < P::ROUNDS; ++r) {
                    const uint32_t e = syn[r];
                    const bool valid = e != 0xffffffffu;
                    const uint32_t wv = __shfl_sync(0xffffffffu, hb, (int)((e >> 8) & 31u));
                    const uint32_t s = valid ? (e & 31u) : 0u;
                    uint32_t x;
                    if constexpr (Z == 32) x = __funnelshift_r(wv, wv, s);
                    else x = ((wv >> s) | (wv << (Z - s))) & ZMASK;
                    if (!valid) x = 0u;
#pragma unroll
                    for (int o = P::SEG / 2; o; o >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, o);
                    unsat |= x;
                }
                if (!__any_sync(0xffffffffu, unsat != 0u)) break;
            }
        }

        // ---- word w leaves (toChar, decodeCL.c:188-199): byte b holds variables 8b .. 8b+7, all in block column 8b / z
        if (p.info) {
            for (int b0 = 0; b0 < KB; b0 += 32) {
                const int b = b0 + (int)lane;
                uint32_t v = byte_of(b);
                if (b == KB - 1 && (p.K & 7)) v &= (1u << (p.K & 7)) - 1u;   // (the last info byte may hold parity bits)
                if (b < KB) p.info[(size_t)w * KB + b] = (uint8_t)v;
            }
        }
        if (p.hard) {
            for (int b0 = 0; b0 < NB8; b0 += 32) {
                const int b = b0 + (int)lane;
                const uint32_t v = byte_of(b);
                if (b < NB8) p.hard[(size_t)w * NB8 + b] = (uint8_t)v;
            }
        }
        if (p.post && act) {
#pragma unroll
            for (int b = 0; b < NB; ++b) p.post[(size_t)w * p.N + b * Z + (int)lane] = -lds_f32(la + (uint32_t)b * 2u * ZB);
        }
        if (p.iters && lane == 0) p.iters[w] = it;
        __syncwarp();   // every read of this word's T and hard bits is done before the next word overwrites them
    }
}

#endif  // LDPC_QCW_DEVICE

}  // namespace ldpc_b200
