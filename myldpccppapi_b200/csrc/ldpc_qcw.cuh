// ldpc_qcw.cuh -- flooding min-sum for short quasi-cyclic codes in the EARLY-TERMINATION regime: ONE WARP = ONE
// CODEWORD, lane = row (check pass) / column (variable pass) of a z x z block, z <= 32.
//
// ldpc_ms_qc_kernel (ldpc_qc.cuh) decodes G words per CTA in lockstep.  When words converge after 3-5 iterations
// (BASELINE config 4, the region a decoder operates in) it spends most of its time outside the two passes
// (profiles/r02_ring_kernel.md, r02_et_kernel.md): CTA-wide barriers around every retire/start event, the leaving word's
// bits packed from shared memory, the starting word's messages cleared, and one extra loop trip per word because the
// syndrome of iteration t is only seen by the check pass of t + 1.  A CTA-level variant with an incremental syndrome
// (ldpc_qc_et.cuh) removed the extra trip but not the barriers and paid for it in divergent shared-memory atomics.
//
// Here nothing is shared between codewords, so nothing is synchronised between them:
//   * a warp owns one codeword: T[24][2z] (negated posterior, each block column stored twice so that the cyclic
//     wrap is a plain offset) and R[E][z] (one message row per non-zero circulant) in its own slice of shared memory.
//     Check pass: lane r handles row r of every block row -- T gathered at a warp-uniform base (__constant__ ->
//     uniform register, as in ldpc_qc.cuh) + lane, own message rows at compile-time offsets.  Variable pass: lane c
//     handles column c of every block column; the messages of its edges sit at (c - s) mod z of their block: two
//     predicated loads with one uniform base.  __syncwarp between the passes, no CTA barrier anywhere in the loop.
//   * SYNDROME FROM PACKED BITS.  The variable pass ballots the sign of every posterior it writes: lane b keeps the
//     z hard bits of block column b in a register.  The syndrome of ALL z rows of a block row is the XOR over its
//     circulants of that word rotated by the shift -- one shuffle and one rotate per CIRCULANT (88 for Test.cpp's
//     code), not per edge, reduced by a butterfly.  A word is finished the moment its syndrome is clean after a
//     variable pass: it costs exactly `iters` trips (reference stop rule, MyLdpc.cpp:751-755).
//   * the leaving word's info bytes ARE those registers (toChar, decodeCL.c:188-199: three bytes per block column);
//     the starting word's channel values were pulled into L2 one word ahead; its messages are not cleared --
//     the first check pass reads them under a predicate (decodeInitMS: R = 0).
//   * a warp that finishes takes the next word from the global queue on its own; the other warps never notice.
// Price: z of 32 lanes work (z = 24: 75 %) and the variable pass issues two loads per edge, so at a fixed 40
// iterations this kernel is slower than the lockstep one; the host picks per launch (ldpc_b200.cu).
// Arithmetic and outputs are those of ldpc_ms_qc_kernel: bit-exact with Coder::decodeCPU (MyLdpc.cpp:684-784).
#pragma once
#include "ldpc_qc.cuh"

namespace ldpc_b200 {

// Degree sequences of the 802.16e rates in NATURAL block order (seed tables MyLdpc.h:40-102: the same for every z).
struct QcwRate12 {
    static constexpr int MB = 12;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[12] = {6, 7, 7, 6, 6, 7, 6, 6, 7, 6, 6, 6}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int b) { constexpr int d[24] = {3, 3, 6, 3, 3, 6, 3, 6, 3, 6, 3, 6, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2}; return d[b]; }
};
struct QcwRate23A {
    static constexpr int MB = 8;
    __host__ __device__ static constexpr int cdeg(int) { return 10; }
    __host__ __device__ static constexpr int vdeg(int b) { constexpr int d[24] = {3, 3, 6, 3, 3, 6, 3, 3, 6, 3, 3, 6, 3, 3, 6, 3, 3, 2, 2, 2, 2, 2, 2, 2}; return d[b]; }
};
struct QcwRate23B {
    static constexpr int MB = 8;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[8] = {10, 10, 10, 10, 10, 10, 11, 10}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int b) { constexpr int d[24] = {4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 2, 2, 2, 2, 2, 2, 2}; return d[b]; }
};
struct QcwRate34A {
    static constexpr int MB = 6;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[6] = {14, 14, 14, 15, 14, 14}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int b) { constexpr int d[24] = {4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 2, 2, 2, 2, 2}; return d[b]; }
};
struct QcwRate34B {
    static constexpr int MB = 6;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[6] = {14, 15, 15, 15, 14, 15}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int b) { constexpr int d[24] = {3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 6, 6, 6, 6, 6, 6, 6, 3, 2, 2, 2, 2, 2}; return d[b]; }
};
struct QcwRate56 {
    static constexpr int MB = 4;
    __host__ __device__ static constexpr int cdeg(int) { return 20; }
    __host__ __device__ static constexpr int vdeg(int b) { constexpr int d[24] = {3, 3, 3, 3, 3, 3, 3, 3, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 2, 2, 2}; return d[b]; }
};

constexpr int kQcwNB = 24;        // block columns (802.16e)
constexpr int kQcwMaxWarps = 16;  // codewords in flight per SM for z = 24 (13,056 B each; 17 would fit, but a fifth warp on one scheduler caps the kernel at 96 registers)

template <class R, int Z_>
struct QcwProfile : R {
    static constexpr int Z = Z_, NB = kQcwNB;
    __host__ __device__ static constexpr int e0(int i) { int o = 0; for (int k = 0; k < i; ++k) o += R::cdeg(k); return o; }   // first circulant of block row i
    __host__ __device__ static constexpr int coff(int i) { int o = 0; for (int k = 0; k < i; ++k) o += (R::cdeg(k) + 1) & ~1; return o; }
    __host__ __device__ static constexpr int voff(int b) { int o = 0; for (int k = 0; k < b; ++k) o += R::vdeg(k); return o; }
    __host__ __device__ static constexpr int dmax() { int m = 0; for (int k = 0; k < R::MB; ++k) m = R::cdeg(k) > m ? R::cdeg(k) : m; return m; }
    static constexpr int E = e0(R::MB), CE = coff(R::MB), VE = voff(kQcwNB);
    static constexpr int SEG = dmax() <= 8 ? 8 : (dmax() <= 16 ? 16 : 32);   // lanes per block row in the syndrome rounds
    static constexpr int ROUNDS = (R::MB * SEG + 31) / 32;
    static constexpr uint32_t T_BYTES = (uint32_t)kQcwNB * 2u * Z * 4u;
    static constexpr uint32_t WARP_BYTES = T_BYTES + (uint32_t)E * Z * 4u + 128u;   // (+ slack: idle lanes read past the last block)
};

template <class P>
struct QcwTab {
    alignas(8) uint32_t cn_t[P::CE];   // [block row][j]: T bytes of (block column, shift) = (bc * 2z + s) * 4
    alignas(8) uint2 vn[P::VE];        // [block column][k]: {R bytes of the circulant minus s * 4 (from the warp's base), s}
};

constexpr int kQcwBankBytes = 1280;

struct QcwParams {
    int tab_slot;
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
    const uint32_t* syn_tab;           // [ROUNDS][32]: (block column << 8) | shift of the circulant a lane rotates, 0xffffffff = none
};

#ifdef LDPC_QCW_DEVICE   // the kernel and its table bank: only the unit that instantiates them (k_qcw.cu)
static __constant__ uint4 g_qcw_bank[kQcTabSlots][kQcwBankBytes / 16];

// One block row: lane = row.  qc_check's arithmetic with the message rows at compile-time offsets, no wrap copies,
// stores only by the z working lanes, old messages of a starting word taken as 0.
template <int D, uint32_t ZB, uint32_t RBASE>
__device__ __forceinline__ void qcw_check(const uint32_t* __restrict__ tt, uint32_t la, uint32_t fresh, bool act) {
    float tv[D + 1], S[D];
#pragma unroll
    for (int j = 0; j < D; j += 2) {
        const uint2 e = *reinterpret_cast<const uint2*>(tt + j);
        tv[j] = lds_f32(la + e.x);
        if (j + 1 < D) tv[j + 1] = lds_f32(la + e.y);
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
        asm volatile(
            "{\n\t.reg .pred q;\n\t"
            "setp.ne.u32 q, %2, 0;\n\t"
            "mov.f32 %0, 0f00000000;\n\t"
            "@!q ld.shared.f32 %0, [%1];\n\t}"
            : "=f"(S[j]) : "r"(la + RBASE + (uint32_t)j * ZB), "r"(fresh) : "memory");
    }
    float m1 = INFINITY, m2 = INFINITY;
    uint32_t px = 0u;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        S[j] = __fadd_rn(tv[j], S[j]);  // = -Q_j
        const float a = fabsf(S[j]);
        m2 = fminf(m2, fmaxf(m1, a));
        m1 = fminf(m1, a);
    }
#pragma unroll
    for (int j = 0; j + 1 < D; j += 2) px = px ^ __float_as_uint(S[j]) ^ __float_as_uint(S[j + 1]);
    if (D & 1) px ^= __float_as_uint(S[D - 1]);
    const uint32_t flip = (((px >> 31) ^ (uint32_t)D ^ 1u) & 1u) << 31;
    uint32_t m1x = __float_as_uint(fminf(m1, kClamp)) ^ flip;
    uint32_t m2x = __float_as_uint(fminf(m2, kClamp)) ^ flip;
    asm("" : "+r"(m1x), "+r"(m2x));
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const uint32_t mag = (fabsf(S[j]) == m1) ? m2x : m1x;
        uint32_t rn;
        asm("lop3.b32 %0, %1, 0x80000000, %2, 0x6a;" : "=r"(rn) : "r"(__float_as_uint(S[j])), "r"(mag));
        if (act) sts_f32(la + RBASE + (uint32_t)j * ZB, __uint_as_float(rn));
    }
}

template <class P, int I>
__device__ __forceinline__ void qcw_cn(const QcwTab<P>& tb, uint32_t la, uint32_t fresh, bool act) {
    if constexpr (I < P::MB) {
        qcw_check<P::cdeg(I), (uint32_t)P::Z * 4u, P::T_BYTES + (uint32_t)P::e0(I) * P::Z * 4u>(tb.cn_t + P::coff(I), la, fresh, act);
        qcw_cn<P, I + 1>(tb, la, fresh, act);
    }
}

// One block column: lane = column.  T = (-y) - R_1 - R_2 ... in ascending-row order; edge k's message sits at row
// (lane - s) mod z of its circulant: base - s*4 + lane*4, plus z*4 for the lanes below s (laz = la + z*4).  The z hard bits of the
// column (bit = !signbit(T)) go to lane B's register.
template <class P, int B>
__device__ __forceinline__ void qcw_vn(const QcwTab<P>& tb, uint32_t la, uint32_t laz, uint32_t lane, const float* yn, bool act, uint32_t& hb) {
    if constexpr (B < P::NB) {
        constexpr int D = P::vdeg(B);
        constexpr uint32_t ZB = (uint32_t)P::Z * 4u;
        float r[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const uint2 u = tb.vn[P::voff(B) + k];
            r[k] = lds_f32((lane < u.y ? laz : la) + u.x);   // one load, the wrap folded into the lane's base
        }
        float acc = yn[B];
#pragma unroll
        for (int k = 0; k < D; ++k) acc = __fsub_rn(acc, r[k]);
        if (act) {
            sts_f32(la + (uint32_t)B * 2u * ZB, acc);
            sts_f32(la + (uint32_t)B * 2u * ZB + ZB, acc);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, act && (__float_as_uint(acc) >> 31) == 0u);
        if (lane == (uint32_t)B) hb = bal;
        qcw_vn<P, B + 1>(tb, la, laz, lane, yn, act, hb);
    }
}

template <class P>
__global__ void __launch_bounds__(kQcwMaxWarps * 32, 1) ldpc_ms_qcw_kernel(const __grid_constant__ QcwParams p) {
    constexpr int Z = P::Z, NB = P::NB;
    constexpr uint32_t ZB = (uint32_t)Z * 4u;
    constexpr uint32_t ZMASK = Z == 32 ? 0xffffffffu : ((1u << (Z & 31)) - 1u);
    static_assert(Z % 8 == 0 && Z <= 32, "a byte of hard bits sits in one block column; a lane per row");
    static_assert(sizeof(QcwTab<P>) <= kQcwBankBytes, "profile tables exceed a bank slot");
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const uint32_t lane = threadIdx.x & 31u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: the table reads stay LDCU
    const QcwTab<P>& tb = *reinterpret_cast<const QcwTab<P>*>(&g_qcw_bank[p.tab_slot][0]);
    const uint32_t la = smem_u32(smem_raw) + (uint32_t)warp * P::WARP_BYTES + lane * 4u;
    const bool act = lane < (uint32_t)Z;
    const int KB = (p.K + 7) >> 3, NB8 = (p.N + 7) >> 3;

    uint32_t syn[P::ROUNDS];
#pragma unroll
    for (int r = 0; r < P::ROUNDS; ++r) syn[r] = p.syn_tab[r * 32 + (int)lane];

    // word w's channel values are pulled into L2 one word ahead (a lane per 128-byte line) ...
    auto prefetch_y = [&](long long w) {
        const char* src = reinterpret_cast<const char*>(p.llr + (size_t)w * p.N);
        for (int o0 = 0; o0 < p.N * 4; o0 += 32 * 128) {   // (same trip count in every lane)
            const int o = o0 + (int)lane * 128;
            if (o < p.N * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
        }
    };
    // lane 0 takes a ticket from the work queue; it is broadcast only when it is consumed, one word later, so the
    // atomic's latency is never waited for
    auto claim = [&]() -> long long { return lane == 0 ? (long long)atomicAdd(p.counter64, 1ull) : 0ll; };
    auto landed = [&](long long w) -> bool {   // streamed input: wait (bounded) until word w is in device memory
        return w < p.ncw && (!p.avail || qc_wait_input(p.avail, w, true, p.status, p.wait_ns));
    };

    float yn[NB];
    long long wn = __shfl_sync(0xffffffffu, claim(), 0);   // the word decoded next (its channel values are on their way to L2)
    long long tick = claim();                                // lane 0: the word after it, claimed, not touched yet
    if (!landed(wn)) wn = p.ncw;
    if (wn < p.ncw) prefetch_y(wn);
    for (;;) {
        const long long w = wn;
        if (w >= p.ncw) break;
        // ---- start word w (decodeInitMS, decodeCL.c:113-124): T = -y (canonical zero); R = 0 is the predicate of the first check pass
        // ... and read here: this lane's column of every block column (lanes >= z: nothing)
        uint32_t hb = 0u;
        {
            const float* src = p.llr + (size_t)w * p.N + lane;
#pragma unroll
            for (int b = 0; b < NB; ++b) yn[b] = act ? __ldg(src + b * Z) : 0.0f;
        }
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
        if (wn < p.ncw) prefetch_y(wn);
        tick = claim();
        __syncwarp();

        int it = 0;
        for (;;) {
            qcw_cn<P, 0>(tb, la, it == 0 ? 1u : 0u, act);
            __syncwarp();
            qcw_vn<P, 0>(tb, la, la + ZB, lane, yn, act, hb);
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
                const int bc = (8 * b) / Z;
                const uint32_t wv = __shfl_sync(0xffffffffu, hb, bc < NB ? bc : 0);
                uint32_t v = (wv >> (8 * b - bc * Z)) & 0xffu;
                if (b == KB - 1 && (p.K & 7)) v &= (1u << (p.K & 7)) - 1u;   // (the last info byte may hold parity bits)
                if (b < KB) p.info[(size_t)w * KB + b] = (uint8_t)v;
            }
        }
        if (p.hard) {
            for (int b0 = 0; b0 < NB8; b0 += 32) {
                const int b = b0 + (int)lane;
                const int bc = (8 * b) / Z;
                const uint32_t wv = __shfl_sync(0xffffffffu, hb, bc < NB ? bc : 0);
                if (b < NB8) p.hard[(size_t)w * NB8 + b] = (uint8_t)((wv >> (8 * b - bc * Z)) & 0xffu);
            }
        }
        if (p.post && act) {
#pragma unroll
            for (int b = 0; b < NB; ++b) p.post[(size_t)w * p.N + b * Z + (int)lane] = -lds_f32(la + (uint32_t)b * 2u * ZB);
        }
        if (p.iters && lane == 0) p.iters[w] = it;
        __syncwarp();   // every read of this word's T is done before the next word's values overwrite it
    }
}

#endif  // LDPC_QCW_DEVICE

}  // namespace ldpc_b200
