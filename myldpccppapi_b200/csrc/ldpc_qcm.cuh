// ldpc_qcm.cuh -- flooding min-sum for quasi-cyclic codes of ANY block size z <= 96: a codeword is decoded by its own
// group of NW = ceil(z / 32) warps, lane = row (check pass) / column (variable pass) of a z x z block.
//
// The lockstep kernel (ldpc_qc.cuh) deals the 32 / G node lanes of a warp to consecutive rows of one block and needs
// z to split evenly over its warps: compiled profiles exist for z = 24, 32, 40, 48, 64, 80, 96, and the other twelve
// block sizes Coder::initCheckMatrix accepts (MyLdpc.cpp:55: z = N / 24 from 28 to 92) fell back to one codeword per
// CTA at 1.8-2.7 Gbit/s.  ldpc_qcw.cuh showed the alternative for z <= 32: no lockstep between codewords at all.  This
// is the same layout for every z, with run-time tables instead of a compiled-in code:
//   * a codeword owns a slice of shared memory: T[24][2z] (negated posterior, every block column stored twice so that
//     the cyclic wrap of the check pass is a plain offset), R[E][z] (one message row per non-zero circulant), a bit
//     buffer for its hard decisions.  Warp sw of its group handles rows / columns [sw * RW, (sw + 1) * RW), RW =
//     ceil(z / NW) <= 32: lane addresses are base + 4 * (sw * RW + lane), so both passes are the straight-line code of
//     ldpc_qcw.cuh with warp-uniform offsets from __constant__ memory (LDCU.64, one pair per edge).
//   * the NW warps of a codeword meet at their own named barrier (bar.sync id, 32 * NW) twice per iteration; NW = 1:
//     __syncwarp.  Nothing is synchronised between codewords: a group that finishes takes the next word from the
//     global queue on its own.
//   * the syndrome of iteration t is seen by the check pass of t + 1 (parity of the gathered signs, a vote per warp, a
//     flag per group): a converged word costs iters + 1 trips.  (z = 24 / 32 of the 802.16e codes have the
//     packed-bit syndrome of ldpc_qcw.cuh and stop after exactly iters trips.)
// Templated on the degree sequence of the rate only (block rows and block columns in natural order): any quasi-cyclic
// code with 24 block columns and those degrees runs, whatever its shifts and z.
// Arithmetic and outputs are those of ldpc_ms_qc_kernel: bit-exact with Coder::decodeCPU (MyLdpc.cpp:684-784).
#pragma once
#include "ldpc_qc.cuh"
#include "qcw_tables.h"

namespace ldpc_b200 {

constexpr int kQcmMaxWarps = 16;   // 128 registers per thread
constexpr int kQcmMaxGroups = 15;  // named barriers 1 .. 15

// table offsets: check-pass entries are read in pairs (LDCU.64), every block row starts on an even entry
template <class R>
struct QcmLayout {
    __host__ __device__ static constexpr int coff(int i) { int o = 0; for (int k = 0; k < i; ++k) o += (R::cdeg(k) + 1) & ~1; return o; }
    static constexpr int CE = coff(R::MB), VE = R::E;
};

template <class R>
struct QcmTab {
    alignas(8) uint32_t cn_t[QcmLayout<R>::CE];   // [block row][j]: T bytes of (block column, shift) = (bc * 2z + s) * 4
    alignas(8) uint2 vn[QcmLayout<R>::VE];        // [block column][k]: {R bytes of the circulant minus s * 4 (from the word's base), s}
};

constexpr int kQcmBankBytes = 1152;

struct QcmParams {
    int tab_slot;
    int z, NW, RW;                // block size, warps per group, rows per warp
    int PK;                       // codewords per group of warps (ldpc_ms_qcm_multi_kernel; 1 = ldpc_ms_qcm_kernel)
    uint32_t zb, t_bytes;         // z * 4; bytes of T (= 24 * 2z * 4): R starts there
    uint32_t bits_off;            // bit buffer of the word's hard decisions (N / 8 bytes + 4), byte offset in its slice
    uint32_t word_bytes;          // shared memory per codeword in flight
    uint32_t hb_bytes;            // sum-product kernel (ldpc_spq.cuh): bytes of its hard-decision array HB[24][2z] (one BYTE per
                                  // variable); E starts there.  0 in the min-sum launches.
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
    int fmt;                      // format of llr: 0 fp32, 1 binary16, 2 int8 (ldpc_b200_decode_host_packed), widened at the load
    float scale;
};

#ifdef LDPC_QCM_DEVICE   // the kernel and its table bank: only the unit that instantiates them (k_qcm.cu)
static __constant__ uint4 g_qcm_bank[kQcTabSlots][kQcmBankBytes / 16];

__device__ __forceinline__ void qcm_atoms_or(uint32_t a, uint32_t v) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// One block row: lane = row.  Returns the row's syndrome bit (hard bit = !signbit(T)).
template <class R, int I>
__device__ __forceinline__ uint32_t qcm_check(const QcmTab<R>& tb, uint32_t la, uint32_t rb, uint32_t zb, bool act) {
    constexpr int D = R::cdeg(I);
    float tv[D + 1], S[D];
#pragma unroll
    for (int j = 0; j < D; j += 2) {  // two warp-uniform bases per LDCU.64
        const uint2 e = *reinterpret_cast<const uint2*>(tb.cn_t + QcmLayout<R>::coff(I) + j);
        tv[j] = lds_f32(la + e.x);
        if (j + 1 < D) tv[j + 1] = lds_f32(la + e.y);
    }
    const uint32_t r0 = la + rb + (uint32_t)R::e0(I) * zb;   // this lane's row of the block row's first circulant
#pragma unroll
    for (int j = 0; j < D; ++j) S[j] = lds_f32(r0 + (uint32_t)j * zb);
#pragma unroll
    for (int j = 0; j < D; ++j) S[j] = __fadd_rn(tv[j], S[j]);  // = -Q_j
    uint32_t px = 0u, sx = 0u;
#pragma unroll
    for (int j = 0; j + 1 < D; j += 2) {
        px = px ^ __float_as_uint(S[j]) ^ __float_as_uint(S[j + 1]);
        sx = sx ^ __float_as_uint(tv[j]) ^ __float_as_uint(tv[j + 1]);
    }
    if (D & 1) {
        px ^= __float_as_uint(S[D - 1]);
        sx ^= __float_as_uint(tv[D - 1]);
    }
    ms_new_messages_each<D>(S, px, [&](int j, float rn) {
        if (act) sts_f32(r0 + (uint32_t)j * zb, rn);
    });
    return ((sx >> 31) ^ (uint32_t)D) & 1u;
}

template <class R, int I>
__device__ __forceinline__ uint32_t qcm_cn(const QcmTab<R>& tb, uint32_t la, uint32_t rb, uint32_t zb, bool act) {
    if constexpr (I < R::MB) {
        const uint32_t u = qcm_check<R, I>(tb, la, rb, zb, act);
        return u | qcm_cn<R, I + 1>(tb, la, rb, zb, act);
    } else {
        return 0u;
    }
}

// One block column: lane = column c.  T = (-y) - R_1 - R_2 ... in ascending-row order; edge k's message sits at row
// (c - s) mod z of its circulant: the columns below s read z rows further (laz = la + z * 4).
template <class R, int B>
__device__ __forceinline__ void qcm_vn(const QcmTab<R>& tb, uint32_t la, uint32_t laz, uint32_t c, uint32_t zb, const float* yn, bool act) {
    if constexpr (B < R::NB) {
        constexpr int D = R::vdeg(B), V0 = R::v0(B);
        float r[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const uint2 u = tb.vn[V0 + k];
            r[k] = lds_f32((c < u.y ? laz : la) + u.x);
        }
        float acc = yn[B];
#pragma unroll
        for (int k = 0; k < D; ++k) acc = __fsub_rn(acc, r[k]);
        if (act) {
            sts_f32(la + (uint32_t)(2 * B) * zb, acc);
            sts_f32(la + (uint32_t)(2 * B + 1) * zb, acc);
        }
        qcm_vn<R, B + 1>(tb, la, laz, c, zb, yn, act);
    }
}

// PACKED: p.llr holds float16 / int8 values, widened at the load (a separate instantiation, see ldpc_qcw.cuh)
template <class R, bool PACKED = false>
__global__ void __launch_bounds__(kQcmMaxWarps * 32, 1) ldpc_ms_qcm_kernel(const __grid_constant__ QcmParams p) {
    constexpr int NB = R::NB;
    static_assert(sizeof(QcmTab<R>) <= kQcmBankBytes, "profile tables exceed a bank slot");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_word[kQcmMaxWarps];       // the word a group decodes (from its leader warp)
    __shared__ uint32_t s_flag[kQcmMaxWarps][2];     // "some check of the word is unsatisfied", double buffered

    const uint32_t lane = threadIdx.x & 31u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: the table reads stay LDCU
    const int ws = warp / p.NW, sw = warp - ws * p.NW;                       // group (= codeword slot) and warp inside it
    const QcmTab<R>& tb = *reinterpret_cast<const QcmTab<R>*>(&g_qcm_bank[p.tab_slot][0]);
    const uint32_t wb = smem_u32(smem_raw) + (uint32_t)ws * p.word_bytes;
    const uint32_t c = (uint32_t)(sw * p.RW) + lane;                          // this lane's row / column inside a block
    const bool act = lane < (uint32_t)p.RW && c < (uint32_t)p.z;
    const uint32_t zb = p.zb;
    const uint32_t la = wb + c * 4u, laz = la + zb;
    const uint32_t bits = wb + p.bits_off;
    const int KB = (p.K + 7) >> 3, NB8 = (p.N + 7) >> 3;
    const int gl = sw * 32 + (int)lane, gn = p.NW * 32;                       // lane index and size of the group
    const bool single = p.NW == 1;
    auto gsync = [&]() {
        if (single) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(ws + 1), "r"(gn) : "memory");
    };

    // the leader warp's lane 0 holds a ticket from the work queue one word ahead (see ldpc_qcw.cuh)
    const int esz = !PACKED ? 4 : (p.fmt == 1 ? 2 : 1);   // bytes per channel value in p.llr
    auto claim = [&]() -> long long { return (sw == 0 && lane == 0) ? (long long)atomicAdd(p.counter64, 1ull) : 0ll; };
    auto prefetch_y = [&](long long w) {
        const char* src = reinterpret_cast<const char*>(p.llr) + (size_t)w * p.N * esz;
        for (int o0 = 0; o0 < p.N * esz; o0 += 32 * 128) {   // (same trip count in every lane)
            const int o = o0 + (int)lane * 128;
            if (o < p.N * esz) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
        }
    };
    auto landed = [&](long long w) -> bool {
        return w < p.ncw && (!p.avail || qc_wait_input(p.avail, w, true, p.status, p.wait_ns));
    };

    float yn[NB];
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
            // hard-bit buffer of the word: cleared here, filled when it leaves
            for (int o0 = 0; o0 < NB8 + 4; o0 += 128) {
                const int o = o0 + (int)lane * 4;
                if (o < NB8 + 4) qc_sts_u32(bits + (uint32_t)o, 0u);
            }
        }
        gsync();
        const long long w = s_word[ws];
        if (w >= p.ncw) break;
        // ---- start word w (decodeInitMS, decodeCL.c:113-124): T = -y (canonical zero), R = 0
        if constexpr (!PACKED) {
            const float* src = p.llr + (size_t)w * p.N + c;
#pragma unroll
            for (int b = 0; b < NB; ++b) yn[b] = act ? __ldg(src + b * p.z) : 0.0f;
        } else {   // packed host formats, widened here
            const size_t i0 = (size_t)w * p.N + c;
#pragma unroll
            for (int b = 0; b < NB; ++b) yn[b] = act ? llr_at(p.llr, p.fmt, p.scale, i0 + (size_t)(b * p.z)) : 0.0f;
        }
        if (act) {
#pragma unroll 8
            for (int e = 0; e < R::E; ++e) sts_f32(la + p.t_bytes + (uint32_t)e * zb, 0.0f);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            yn[b] = __fadd_rn(-yn[b], 0.0f);
            if (act) {
                sts_f32(la + (uint32_t)(2 * b) * zb, yn[b]);
                sts_f32(la + (uint32_t)(2 * b + 1) * zb, yn[b]);
            }
        }
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
            // check-node pass + syndrome of the previous posterior (checkResult, decodeCL.c:88-108)
            const uint32_t unsat = qcm_cn<R, 0>(tb, la, p.t_bytes, zb, act);
            const bool check = p.early_term && it >= 1;
            if (check && __any_sync(0xffffffffu, act && unsat != 0u) && lane == 0) s_flag[ws][ph] = 1u;  // same-value race, benign
            gsync();
            if (check && s_flag[ws][ph] == 0u) break;            // stop rule MyLdpc.cpp:751-755: T of iteration `it` is final
            if (sw == 0 && lane == 0) s_flag[ws][ph ^ 1u] = 0u;
            ph ^= 1u;
            qcm_vn<R, 0>(tb, la, laz, c, zb, yn, act);
            gsync();
            ++it;
            if (it >= p.max_iter) break;
        }

        // ---- word w leaves (toChar, decodeCL.c:188-199): bit n = !(P > 0) = !signbit(T).  A warp ballots its columns
        // of every block column into the word's bit buffer (bit bc * z + column), then the group copies the bytes out.
        for (int b = 0; b < NB; ++b) {
            const float t = lds_f32(la + (uint32_t)(2 * b) * zb);
            const uint32_t bal = __ballot_sync(0xffffffffu, act && (__float_as_uint(t) >> 31) == 0u);
            if (p.post && act) p.post[(size_t)w * p.N + b * p.z + (int)c] = -t;
            if (lane == 0) {
                const uint32_t g = (uint32_t)(b * p.z + sw * p.RW), sh = g & 31u;
                qcm_atoms_or(bits + (g >> 5) * 4u, bal << sh);
                if (sh) qcm_atoms_or(bits + (g >> 5) * 4u + 4u, bal >> (32u - sh));
            }
        }
        gsync();
        {
            const int nby = p.hard ? NB8 : KB;
            for (int b0 = 0; b0 < nby; b0 += gn) {
                const int b = b0 + gl;
                if (b < nby) {
                    const uint32_t v = (qc_lds_u32(bits + (uint32_t)(b & ~3)) >> (8 * (b & 3))) & 0xffu;
                    if (p.hard) p.hard[(size_t)w * NB8 + b] = (uint8_t)v;
                    if (p.info && b < KB) {
                        const uint32_t keep = (b == KB - 1 && (p.K & 7)) ? ((1u << (p.K & 7)) - 1u) : 0xffu;   // (the last info byte may hold parity bits)
                        p.info[(size_t)w * KB + b] = (uint8_t)(v & keep);
                    }
                }
            }
        }
        if (p.iters && gl == 0) p.iters[w] = it;
        gsync();   // every read of this word's T and bits is done before the next word overwrites them
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Several codewords per group of warps.  With one codeword per group a block size just above a multiple of 32 leaves
// most lanes of the last warp idle (z = 36: 2 warps, 18 of 32 lanes each; z = 68: 3 warps, 23 lanes each), and an idle
// lane costs what a working one does -- a warp's shared-memory access is a wavefront whether 18 or 32 lanes take part.
// Here the rows of PK = 2 or 3 codewords are laid side by side over the group's lanes (z = 36: 3 codewords = 108 rows on
// 4 warps; z = 68: 2 codewords = 136 rows on 5 warps): lane (sw, l) serves row (sw * RW + l) mod z of codeword
// (sw * RW + l) / z, in that codeword's own slice of shared memory, so both passes stay the straight-line code above with
// per-lane bases.  The slices are z words apart modulo the 32 banks (qcm_multi_geometry), so the lanes of a warp that
// straddles two codewords still touch 32 different banks.  The group's codewords start together; each stops by its own
// syndrome (its lanes freeze, reference stop rule MyLdpc.cpp:751-755) and the group takes its next PK words when all
// have stopped -- the price in the early-termination regime is the spread of the iteration counts inside a group, so the
// host uses this kernel only while the words run long (ldpc_b200.cu: launch_decode).  Bit-exact with ldpc_ms_qcm_kernel.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kQcmMaxPack = 3;

template <class R>
__global__ void __launch_bounds__(kQcmMaxWarps * 32, 1) ldpc_ms_qcm_multi_kernel(const __grid_constant__ QcmParams p) {
    constexpr int NB = R::NB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_word[kQcmMaxWarps];       // first word of the group's pack
    __shared__ uint32_t s_flag[kQcmMaxWarps][2];     // bit k: "some check of the group's k-th word is unsatisfied", double buffered

    const uint32_t lane = threadIdx.x & 31u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int ws = warp / p.NW, sw = warp - ws * p.NW;
    const QcmTab<R>& tb = *reinterpret_cast<const QcmTab<R>*>(&g_qcm_bank[p.tab_slot][0]);
    const uint32_t cc = (uint32_t)(sw * p.RW) + lane;                         // row of the side-by-side layout
    const bool inr = lane < (uint32_t)p.RW && cc < (uint32_t)(p.PK * p.z);
    const uint32_t k = min(cc / (uint32_t)p.z, (uint32_t)(p.PK - 1));         // which of the group's codewords
    const uint32_t c = cc - k * (uint32_t)p.z;                                // this lane's row / column inside a block
    const uint32_t gb = smem_u32(smem_raw) + (uint32_t)(ws * p.PK) * p.word_bytes;
    const uint32_t wb = gb + k * p.word_bytes;
    const uint32_t zb = p.zb;
    const uint32_t la = wb + c * 4u, laz = la + zb;
    const int KB = (p.K + 7) >> 3, NB8 = (p.N + 7) >> 3;
    const int gl = sw * 32 + (int)lane, gn = p.NW * 32;
    // this warp's lanes of each codeword (warp-uniform; a codeword's lanes are contiguous)
    uint32_t km[kQcmMaxPack];
#pragma unroll
    for (int kk = 0; kk < kQcmMaxPack; ++kk) km[kk] = __ballot_sync(0xffffffffu, inr && k == (uint32_t)kk);
    auto gsync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(ws + 1), "r"(gn) : "memory"); };

    auto claim = [&]() -> long long { return (sw == 0 && lane == 0) ? (long long)atomicAdd(p.counter64, (unsigned long long)p.PK) : 0ll; };
    auto prefetch_y = [&](long long w0) {   // the pack's words are consecutive: one contiguous range
        const long long n = min((long long)p.PK, p.ncw - w0);
        const char* src = reinterpret_cast<const char*>(p.llr + (size_t)w0 * p.N);
        const int bytes = (int)n * p.N * 4;
        for (int o0 = 0; o0 < bytes; o0 += 32 * 128) {
            const int o = o0 + (int)lane * 128;
            if (o < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
        }
    };
    auto landed = [&](long long w0) -> bool {   // streamed input: the pack's last word has arrived
        if (w0 >= p.ncw) return false;
        const long long wl = min(w0 + p.PK, p.ncw) - 1;
        return !p.avail || qc_wait_input(p.avail, wl, true, p.status, p.wait_ns);
    };

    float yn[NB];
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
            for (int kk = 0; kk < p.PK; ++kk)   // hard-bit buffers of the pack: cleared here, filled when the words leave
                for (int o0 = 0; o0 < NB8 + 4; o0 += 128) {
                    const int o = o0 + (int)lane * 4;
                    if (o < NB8 + 4) qc_sts_u32(gb + (uint32_t)kk * p.word_bytes + p.bits_off + (uint32_t)o, 0u);
                }
        }
        gsync();
        const long long w0 = s_word[ws];
        if (w0 >= p.ncw) break;
        const long long w = w0 + (long long)k;
        const bool has = inr && w < p.ncw;                         // (the last pack of a batch may be short)
        const uint32_t livemask = (1u << (int)min((long long)p.PK, p.ncw - w0)) - 1u;
        bool act = has;
        // ---- start the pack (decodeInitMS, decodeCL.c:113-124): T = -y (canonical zero), R = 0
        {
            const float* src = p.llr + (size_t)w * p.N + c;
#pragma unroll
            for (int b = 0; b < NB; ++b) yn[b] = has ? __ldg(src + b * p.z) : 0.0f;
        }
        if (has) {
#pragma unroll 8
            for (int e = 0; e < R::E; ++e) sts_f32(la + p.t_bytes + (uint32_t)e * zb, 0.0f);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            yn[b] = __fadd_rn(-yn[b], 0.0f);
            if (has) {
                sts_f32(la + (uint32_t)(2 * b) * zb, yn[b]);
                sts_f32(la + (uint32_t)(2 * b + 1) * zb, yn[b]);
            }
        }
        if (sw == 0) {   // the next pack's values travel while this one is decoded
            wn = __shfl_sync(0xffffffffu, tick, 0);
            if (!landed(wn)) wn = p.ncw;
            if (wn < p.ncw) prefetch_y(wn);
            tick = claim();
        }
        gsync();

        int it = 0, my_it = 0;
        uint32_t ph = 0u, donemask = 0u;
        for (;;) {
            const uint32_t unsat = qcm_cn<R, 0>(tb, la, p.t_bytes, zb, act);
            const bool check = p.early_term && it >= 1;
            if (check) {
                const uint32_t bal = __ballot_sync(0xffffffffu, act && unsat != 0u);
                uint32_t f = 0u;
#pragma unroll
                for (int kk = 0; kk < kQcmMaxPack; ++kk) f |= (bal & km[kk]) ? (1u << kk) : 0u;
                if (f && lane == 0) qcm_atoms_or(smem_u32(&s_flag[ws][ph]), f);
            }
            gsync();
            if (check) {
                const uint32_t stopped = livemask & ~donemask & ~s_flag[ws][ph];   // syndrome zero: T of iteration `it` is final
                if (act && ((stopped >> k) & 1u)) { act = false; my_it = it; }
                donemask |= stopped;
            }
            if (donemask == livemask) break;
            if (sw == 0 && lane == 0) s_flag[ws][ph ^ 1u] = 0u;
            ph ^= 1u;
            qcm_vn<R, 0>(tb, la, laz, c, zb, yn, act);
            gsync();
            ++it;
            if (it >= p.max_iter) break;
        }
        if (act) my_it = it;   // stopped by the cap
        // (a flag left set by the last trip is cleared by trips 0 / 1 of the next pack before that buffer is used again)

        // ---- the pack leaves (toChar, decodeCL.c:188-199): bit n = !(P > 0) = !signbit(T)
        for (int b = 0; b < NB; ++b) {
            const float t = lds_f32(la + (uint32_t)(2 * b) * zb);
            const uint32_t bal = __ballot_sync(0xffffffffu, has && (__float_as_uint(t) >> 31) == 0u);
            if (p.post && has) p.post[(size_t)w * p.N + b * p.z + (int)c] = -t;
            if (lane == 0) {
#pragma unroll
                for (int kk = 0; kk < kQcmMaxPack; ++kk) {
                    if (km[kk]) {
                        const int f = __ffs((int)km[kk]) - 1;                       // first lane of codeword kk in this warp
                        const uint32_t c0 = (uint32_t)(sw * p.RW + f) - (uint32_t)(kk * p.z);
                        const uint32_t g = (uint32_t)(b * p.z) + c0, sh = g & 31u;
                        const uint32_t v = (bal & km[kk]) >> f;
                        const uint32_t bits = gb + (uint32_t)kk * p.word_bytes + p.bits_off;
                        qcm_atoms_or(bits + (g >> 5) * 4u, v << sh);
                        if (sh) qcm_atoms_or(bits + (g >> 5) * 4u + 4u, v >> (32u - sh));
                    }
                }
            }
        }
        gsync();
        {
            const int nby = p.hard ? NB8 : KB;
            for (int kk = 0; kk < p.PK; ++kk) {
                if (!((livemask >> kk) & 1u)) break;
                const uint32_t bits = gb + (uint32_t)kk * p.word_bytes + p.bits_off;
                const long long wk = w0 + kk;
                for (int b0 = 0; b0 < nby; b0 += gn) {
                    const int b = b0 + gl;
                    if (b < nby) {
                        const uint32_t v = (qc_lds_u32(bits + (uint32_t)(b & ~3)) >> (8 * (b & 3))) & 0xffu;
                        if (p.hard) p.hard[(size_t)wk * NB8 + b] = (uint8_t)v;
                        if (p.info && b < KB) {
                            const uint32_t keep = (b == KB - 1 && (p.K & 7)) ? ((1u << (p.K & 7)) - 1u) : 0xffu;
                            p.info[(size_t)wk * KB + b] = (uint8_t)(v & keep);
                        }
                    }
                }
            }
        }
        if (p.iters && has && c == 0u) p.iters[w] = my_it;
        gsync();   // every read of the pack's T and bits is done before the next pack overwrites them
    }
}

#endif  // LDPC_QCM_DEVICE

}  // namespace ldpc_b200
