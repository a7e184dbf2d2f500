// ldpc_qc.cuh -- flooding min-sum for QUASI-CYCLIC codes: the group kernel's arithmetic with
// warp-uniform index tables.
//
// Every code the reference can build is quasi-cyclic (Coder::initCheckMatrix, MyLdpc.cpp:52-109:
// H is a grid of z x z blocks, each zero or the identity shifted by s: row r of block row br meets
// column (r + s) mod z of block column bc).  The generic group kernel (ldpc_kernels.cuh) ignores
// that: it reads one table entry per (node lane, edge), four node lanes per warp instruction, and
// those lane-varying LDS.128 table reads are 22 % of its shared-memory wavefronts
// (profiles/r01_j_group_g8_profile_ncu.txt: 5.58 wavefronts per edge-instruction against 4.35 for
// the data itself).  Here the SUB = 32/G node lanes of a warp instruction take SUB CONSECUTIVE rows
// (or columns) of one block, so the address of lane (h, c) is
//       base(warp, slot, edge) + h * G*4 + c * 4  =  base + lane * 4
// -- one warp-uniform base per edge, read from __constant__ memory into a uniform register
// (LDCU.64 c[0x3][UR + imm]; no shared-memory wavefront, no vector ALU op) and used directly as
// `LDS R, [R_lane + UR]`.  A warp access is 128 contiguous bytes: conflict-free by construction.
// The cyclic wrap is absorbed by padding: every block column of T carries SUB extra rows that
// repeat its first rows, every R block SUB leading rows that repeat its last rows; the group that
// owns them stores twice (one group in z/SUB does; the branch is warp-uniform).
//
// Layout in shared memory (G codewords per CTA, rows of G floats = 32 B for G = 8):
//   T [NB][z + SUB][G]     negated posterior, natural column order inside a block column
//   R [E ][SUB + z][G]     E = number of non-zero blocks; block e = (block row, j) holds the message
//                          of edge j of every row of that block row, indexed by the ROW
//   zero row, -inf row      targets of padded variable-pass / check-pass entries (slots whose two blocks differ in degree)
// Arithmetic contract, lane refill and outputs are those of ldpc_ms_group_kernel (bit-exact with
// Coder::decodeCPU, reference MyLdpc.cpp:684-784).
#pragma once
#include "ldpc_kernels.cuh"

namespace ldpc_b200 {

// Compiled profiles.  A block has z / SUB groups of SUB = 32/G rows; with W = 2 z / SUB warps a slot = two blocks, and
// the slot degrees then depend on the 802.16e rate only (seed tables MyLdpc.h:40-102; blocks sorted by degree, a
// slot's degree = its larger block's, the smaller one is padded).  Instantiated for (z, G, W) = (24, 8, 12),
// (48, 4, 12), (96, 2, 12), (40, 4, 10), (80, 2, 10), (32, 4, 8), (64, 2, 8).
// QcProfileWimax34B576 = Test.cpp's code (rate 3/4B, z = 24: N = 576, M = 144).
template <int Z_, int G_, int W_ = 12>
struct QcProfile34B {
    static constexpr int Z = Z_, G = G_, W = W_, CS = 3, VS = 12;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[3] = {15, 15, 14}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int i) { constexpr int d[12] = {6, 6, 6, 6, 3, 3, 3, 3, 3, 3, 2, 2}; return d[i]; }
};
template <int Z_, int G_, int W_ = 12>
struct QcProfile34A {
    static constexpr int Z = Z_, G = G_, W = W_, CS = 3, VS = 12;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[3] = {15, 14, 14}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int i) { constexpr int d[12] = {4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 2, 2}; return d[i]; }
};
template <int Z_, int G_, int W_ = 12>
struct QcProfile23B {
    static constexpr int Z = Z_, G = G_, W = W_, CS = 4, VS = 12;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[4] = {11, 10, 10, 10}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int i) { constexpr int d[12] = {4, 4, 4, 4, 4, 4, 4, 4, 3, 2, 2, 2}; return d[i]; }
};
template <int Z_, int G_, int W_ = 12>
struct QcProfile23A {
    static constexpr int Z = Z_, G = G_, W = W_, CS = 4, VS = 12;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[4] = {10, 10, 10, 10}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int i) { constexpr int d[12] = {6, 6, 6, 3, 3, 3, 3, 3, 3, 2, 2, 2}; return d[i]; }
};
template <int Z_, int G_, int W_ = 12>
struct QcProfile12 {
    static constexpr int Z = Z_, G = G_, W = W_, CS = 6, VS = 12;
    __host__ __device__ static constexpr int cdeg(int i) { constexpr int d[6] = {7, 7, 6, 6, 6, 6}; return d[i]; }
    __host__ __device__ static constexpr int vdeg(int i) { constexpr int d[12] = {6, 6, 6, 3, 3, 3, 3, 2, 2, 2, 2, 2}; return d[i]; }
};
template <int Z_, int G_, int W_ = 12>
struct QcProfile56 {
    static constexpr int Z = Z_, G = G_, W = W_, CS = 2, VS = 12;
    __host__ __device__ static constexpr int cdeg(int) { return 20; }
    __host__ __device__ static constexpr int vdeg(int i) { constexpr int d[12] = {4, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 2}; return d[i]; }
};
using QcProfileWimax34B576 = QcProfile34B<24, 8>;

// Table offsets shared by the host builder and the kernel.  Entries are read two at a time (LDCU.64), so
// every check slot and every run of variable slots that is processed together starts on an even entry.
template <class P>
struct QcLayout {
    __host__ __device__ static constexpr int coff(int i) { int o = 0; for (int k = 0; k < i; ++k) o += (P::cdeg(k) + 1) & ~1; return o; }
    // variable slots are processed in runs of NS equal-degree slots (4 for degree <= 3, else 2, else 1)
    __host__ __device__ static constexpr int vrun(int s0) {
        const int d = P::vdeg(s0);
        int n = 1;
        while (s0 + n < P::VS && P::vdeg(s0 + n) == d) ++n;
        return (n >= 4 && d <= 6) ? 4 : (n >= 2 ? 2 : 1);
    }
    __host__ __device__ static constexpr int voff(int s) {  // s may be P::VS (total)
        int o = 0, s0 = 0;
        while (s0 < P::VS) {
            const int ns = vrun(s0), d = P::vdeg(s0);
            if (s < s0 + ns) return o + (s - s0) * d;
            o += (ns * d + 1) & ~1;
            s0 += ns;
        }
        return o;
    }
    static constexpr int CE = coff(P::CS), VE = voff(P::VS);
};

// One warp's tables.  All shared-memory values are byte offsets from the dynamic shared base.
template <class P>
struct QcWarpTab {
    alignas(8) uint32_t cn_t[QcLayout<P>::CE];  // [slot][j]  T rows gathered by edge j of the slot's 4 checks
    alignas(8) uint32_t vn_r[QcLayout<P>::VE];  // [slot][k]  R rows gathered by the k-th edge (ascending row) of the slot's 4 variables
    uint32_t cn_r[P::CS];             // own R rows of the slot's checks, edge j at + j * RS
    uint32_t vn_t[P::VS];             // own T rows of the slot's variables
    uint32_t var0[P::VS];             // variable index of node lane 0 (node lane h holds var0 + h)
    uint32_t cdup, vdup;              // non-zero: this warp's groups own wrapped rows -- store the padded copy too
                                      // (a warp's groups sit at the same place of their blocks: all or none)
};

// The tables live in __constant__ memory (bank 3), not in the kernel parameters: launches with several KB of
// parameters neither overlap with each other nor launch quickly (measured: the 3-stream host pipeline lost
// 47 us per launch).  kQcTabSlots decoders per device can hold tables at once; a handle owns one slot.
constexpr int kQcTabSlots = 8;
constexpr int kQcBankBytes = 6656;  // per slot: the largest compiled profile's tables (12 warps)
// (static: the library is built from several translation units, one per 802.16e rate for these kernels -- each
// has its own bank; a handle's tables are uploaded through the unit that holds its profile.)
static __constant__ uint4 g_qc_bank[kQcTabSlots][kQcBankBytes / 16];

template <class P>
__device__ __forceinline__ const QcWarpTab<P>& qc_tab(int slot, int warp) {
    static_assert(sizeof(QcWarpTab<P>) * P::W <= kQcBankBytes, "profile tables exceed a bank slot");
    return reinterpret_cast<const QcWarpTab<P>*>(&g_qc_bank[slot][0])[warp];
}

struct QcParams {
    int tab_slot;                 // which entry of the __constant__ table bank
    int N, K, NB;                 // NB = N / z block columns
    uint32_t t_bytes, r_bytes;    // region sizes (T at 0, R at t_bytes, then the zero row and the -inf row, 128 B each)
    int max_iter, early_term, refill_wait;
    const float* __restrict__ llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;
    unsigned long long* counter64;
    // Streamed input (host-buffer pipeline): words [0, *avail) have landed in `llr`; the copy stream advances
    // *avail after each chunk while this kernel is already running.  null = everything is there.
    const unsigned long long* avail;
    int* status;                  // set to 1 if the wait for input timed out
    unsigned long long wait_ns;   // bound of that wait (ldpc_b200_set_option "wait_timeout_ms")
    uint32_t ring_off;            // ring kernel: staging ring, byte offset in dynamic shared memory
};

// Warp 0 waits until the words its lanes just took from the queue have landed (streamed batches).  Every branch
// is decided by a warp vote, so the warp never diverges here: ptxas keeps treating the table reads of the hot loop
// as warp-uniform (a divergent spin loop turned every LDCU into a vector LDC + address add and halved the speed).
// The poll reads through L2 (ld.acquire.gpu); the channel values were never cached by this SM before, so the
// cp.async that follows sees the DMA's data.  Bounded: a stalled copy stream sets *status instead of hanging.
__device__ __forceinline__ bool qc_wait_input(const unsigned long long* avail, long long w, bool need, int* status,
                                              unsigned long long wait_ns) {
    unsigned long long t0 = 0ull;
    for (uint32_t spins = 0;; ++spins) {
        unsigned long long a;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(avail) : "memory");
        if (__all_sync(0xffffffffu, !need || a > (unsigned long long)w)) return true;
        __nanosleep(spins < 64 ? 100 : 1000);
        if ((spins & 1023u) == 1023u) {
            // once one lane group anywhere has timed out the whole launch is given up: do not wait a second time
            int st = 0;
            if (status) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(st) : "l"(status) : "memory");
            if (__any_sync(0xffffffffu, st != 0)) return false;
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0ull) t0 = now;
            if (__any_sync(0xffffffffu, now - t0 > wait_ns)) {
                if (status && (threadIdx.x & 31) == 0) atomicExch(status, 1);
                return false;
            }
        }
    }
}

// One check of exact degree D (see grp_check / ms_new_messages): the T rows come from warp-uniform bases, the R rows
// are this lane's own column of the block (edge j at + j * RS).  Returns the row's syndrome bit.
template <int D, uint32_t RS, uint32_t WRAP, bool DUP>
__device__ __forceinline__ uint32_t qc_check(const uint32_t* __restrict__ tt, uint32_t rrow, uint32_t la) {
    float tv[D + 1], S[D];
#pragma unroll
    for (int j = 0; j < D; j += 2) {  // two warp-uniform bases per LDCU.64
        const uint2 e = *reinterpret_cast<const uint2*>(tt + j);
        tv[j] = lds_f32(la + e.x);
        if (j + 1 < D) tv[j + 1] = lds_f32(la + e.y);
    }
#pragma unroll
    for (int j = 0; j < D; ++j) S[j] = lds_f32(la + rrow + (uint32_t)j * RS);
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
    float rn[D];
    ms_new_messages<D>(S, px, rn);
#pragma unroll
    for (int j = 0; j < D; ++j) sts_f32(la + rrow + (uint32_t)j * RS, rn[j]);
    if constexpr (DUP) {  // this group owns the block's last rows, which are also read through the leading pad
#pragma unroll
        for (int j = 0; j < D; ++j) sts_f32(la + rrow + (uint32_t)j * RS - WRAP, rn[j]);
    }
    return ((sx >> 31) ^ (uint32_t)D) & 1u;  // hard bit = !signbit(T)
}

template <class P, int CS0, bool DUP>
__device__ __forceinline__ uint32_t qc_cn_static(const QcWarpTab<P>& tb, uint32_t la) {
    if constexpr (CS0 < P::CS) {
        constexpr int D = P::cdeg(CS0);
        constexpr uint32_t RS = (uint32_t)(P::Z + 32 / P::G) * P::G * 4u;
        constexpr uint32_t WRAP = (uint32_t)P::Z * P::G * 4u;
        const uint32_t u = qc_check<D, RS, WRAP, DUP>(tb.cn_t + QcLayout<P>::coff(CS0), tb.cn_r[CS0], la);
        return u | qc_cn_static<P, CS0 + 1, DUP>(tb, la);
    } else {
        return 0u;
    }
}

// NS variable slots of exact degree D together (independent FADD chains interleaved): T = (-y) - R_1 - R_2 ...
// in ascending-row order; the new T goes to this lane's own row (and to the trailing pad for the block's first group).
template <class P, int S0, int D, int NS, bool DUP>
__device__ __forceinline__ void qc_vn_slots(const QcWarpTab<P>& tb, uint32_t la, const float* yn, bool done) {
    constexpr uint32_t WRAP = (uint32_t)P::Z * P::G * 4u;
    constexpr int NE = NS * D;
    float acc[NS], r[NE + 1];
#pragma unroll
    for (int i = 0; i < NS; ++i) acc[i] = yn[S0 + i];
#pragma unroll
    for (int e = 0; e < NE; e += 2) {  // entries of the run are contiguous: [slot][k]
        const uint2 u = *reinterpret_cast<const uint2*>(tb.vn_r + QcLayout<P>::voff(S0) + e);
        r[e] = lds_f32(la + u.x);
        if (e + 1 < NE) r[e + 1] = lds_f32(la + u.y);
    }
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
        for (int i = 0; i < NS; ++i) acc[i] = __fsub_rn(acc[i], r[i * D + k]);
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        if (!done) {
            sts_f32(la + tb.vn_t[S0 + i], acc[i]);
            if constexpr (DUP) sts_f32(la + tb.vn_t[S0 + i] + WRAP, acc[i]);
        }
    }
}

template <class P, int S0, bool DUP>
__device__ __forceinline__ void qc_vn_static(const QcWarpTab<P>& tb, uint32_t la, const float* yn, bool done) {
    if constexpr (S0 < P::VS) {
        constexpr int NS = QcLayout<P>::vrun(S0);
        qc_vn_slots<P, S0, P::vdeg(S0), NS, DUP>(tb, la, yn, done);
        qc_vn_static<P, S0 + NS, DUP>(tb, la, yn, done);
    }
}

template <class P>
__global__ void __launch_bounds__(P::W * 32, (P::W * 32 <= 288 ? 3 : (P::W * 32 <= 384 ? 2 : 1)))
ldpc_ms_qc_kernel(const __grid_constant__ QcParams p) {
    constexpr int G = P::G, SUB = 32 / G, NL = P::W * SUB, Z = P::Z;
    constexpr uint32_t ROWB = (uint32_t)G * 4u;              // bytes of one row (G codewords)
    constexpr uint32_t RS = (uint32_t)(Z + SUB) * ROWB;      // bytes of one padded block
    constexpr uint32_t WRAP = (uint32_t)Z * ROWB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_flag[2][32];
    __shared__ long long s_cw[32], s_nxt[32];

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform: the table reads become LDCU
    const int c = lane & (G - 1), h = lane / G;
    const QcWarpTab<P>& tb = qc_tab<P>(p.tab_slot, warp);
    const uint32_t sb = smem_u32(smem_raw);
    const uint32_t la = sb + (uint32_t)lane * 4u;            // every hot-loop access is [la + uniform (+ imm)]
    const uint32_t c4 = (uint32_t)c * 4u;

    if (threadIdx.x < 32) {
        sts_f32(sb + p.t_bytes + p.r_bytes + (uint32_t)lane * 4u, 0.0f);              // zero row (padded variable-pass entries)
        sts_f32(sb + p.t_bytes + p.r_bytes + 128u + (uint32_t)lane * 4u, -INFINITY);  // -inf row (padded check-pass entries)
    }

    float yn[P::VS];
#pragma unroll
    for (int s = 0; s < P::VS; ++s) yn[s] = -1.0f;
    long long cw = -1;
    bool live = false, done = false, loading = false;
    int it = 0, my_iters = 0;

    auto t_addr = [&](int n) -> uint32_t {  // T element of variable n, codeword lane c
        return sb + (uint32_t)((n / Z) * (Z + SUB) + (n % Z)) * ROWB + c4;
    };
    auto emit = [&](bool sel) {
        // toChar (decodeCL.c:188-199): bit n = !(P > 0) = !signbit(T); node lanes share the bytes
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = warp * SUB + h; b < KB; b += NL) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.K) v |= ((~__float_as_uint(lds_f32(t_addr(n)))) >> 31) << t;
                }
                if (sel) p.info[(size_t)cw * KB + b] = (uint8_t)v;
            }
        }
        if (p.hard) {
            const int NB8 = (p.N + 7) >> 3;
            for (int b = warp * SUB + h; b < NB8; b += NL) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.N) v |= ((~__float_as_uint(lds_f32(t_addr(n)))) >> 31) << t;
                }
                if (sel) p.hard[(size_t)cw * NB8 + b] = (uint8_t)v;
            }
        }
        if (p.post && sel) {
            for (int n = warp * SUB + h; n < p.N; n += NL) p.post[(size_t)cw * p.N + n] = -lds_f32(t_addr(n));
        }
        if (p.iters && warp == 0 && h == 0 && sel) p.iters[cw] = my_iters;
    };
    // warp 0: lanes selected by `want` claim their next word and pull its channel values into L2
    auto claim = [&](bool want) {
        long long nn = (h == 0 && want) ? (long long)atomicAdd(p.counter64, 1ull) : p.ncw;
        if (h == 0 && want) s_nxt[c] = nn;
        nn = __shfl_sync(0xffffffffu, nn, c);          // node lanes 1..SUB-1 of warp 0 help with the prefetch
        const bool w2 = __shfl_sync(0xffffffffu, (int)want, c) != 0;
        const char* base = reinterpret_cast<const char*>(p.llr + (size_t)nn * p.N);
        const bool pf = w2 && nn < p.ncw;
        for (int i = 0; i < (p.N * 4 + SUB * 128 - 1) / (SUB * 128); ++i) {  // same trip count in every lane
            const int off = (i * SUB + h) * 128;
            if (pf && off < p.N * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
    };
    // Refill the lanes selected by `want`: each takes the word its lane claimed at the PREVIOUS refill (s_nxt),
    // whose channel values were prefetched into L2 then, and sends them by cp.async into this thread's own T
    // elements; after that warp 0 claims the word after it from the queue and starts its L2 prefetch.  The queue's
    // atomic and the HBM latency are thus off the critical path of a refill (they used to cost ~2 us per event,
    // which at 3 iterations per word doubled the time of a batch).
    auto fetch = [&](bool want) {
        if (warp == 0) {
            const bool take = h == 0 && want;
            long long w = take ? s_nxt[c] : -1;
            if (p.avail && !qc_wait_input(p.avail, w, take && w < p.ncw, p.status, p.wait_ns)) w = p.ncw;  // timed out: give the words up
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
                for (int s = 0; s < P::VS; ++s) {
                    const uint32_t dst = la + tb.vn_t[s];
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + tb.var0[s]) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (warp == 0) claim(want);
    };
    // lanes whose values have landed start decoding: T = -y (canonical zero), R = 0 (decodeInitMS)
    auto start_loaded = [&]() {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (loading) {
#pragma unroll
            for (int s = 0; s < P::VS; ++s) {
                const float y = lds_f32(la + tb.vn_t[s]);
                yn[s] = __fadd_rn(-y, 0.0f);
                sts_f32(la + tb.vn_t[s], yn[s]);
                if (tb.vdup) sts_f32(la + tb.vn_t[s] + WRAP, yn[s]);
            }
#pragma unroll
            for (int cs = 0; cs < P::CS; ++cs) {
                const bool dup = tb.cdup != 0u;
                for (int j = 0; j < P::cdeg(cs); ++j) {
                    sts_f32(la + tb.cn_r[cs] + (uint32_t)j * RS, 0.0f);
                    if (dup) sts_f32(la + tb.cn_r[cs] + (uint32_t)j * RS - WRAP, 0.0f);
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
        // loop top: start words whose values arrived, retire finished words and refill their lanes
        // (a warp holds every codeword lane and all threads of a lane agree: the votes are CTA-uniform)
        if (__any_sync(0xffffffffu, loading)) start_loaded();
        const bool retire = live && done;
        if (__any_sync(0xffffffffu, retire)) {
            emit(retire);
            fetch(retire);
            if ((p.refill_wait || !__any_sync(0xffffffffu, live)) && __any_sync(0xffffffffu, loading)) start_loaded();
        }
        if (!__any_sync(0xffffffffu, live || loading)) break;

        // check-node pass + syndrome of the previous posterior
        const uint32_t unsat = tb.cdup ? qc_cn_static<P, 0, true>(tb, la) : qc_cn_static<P, 0, false>(tb, la);  // warp-uniform
        const bool check = p.early_term && it >= 1 && live && !done;
        if (check && unsat) s_flag[ph & 1][c] = 1u;  // same-value race, benign
        __syncthreads();
        if (check && s_flag[ph & 1][c] == 0u) { done = true; my_iters = it; }
        if (warp == 0) s_flag[(ph + 1) & 1][lane] = 0u;
        ++ph;

        // variable-node pass (the posterior of finished / idle lanes is frozen)
        if (tb.vdup) qc_vn_static<P, 0, true>(tb, la, yn, !(live && !done));  // warp-uniform
        else qc_vn_static<P, 0, false>(tb, la, yn, !(live && !done));
        if (live && !done) {
            ++it;
            if (it == p.max_iter) { done = true; my_iters = it; }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// The same decoder with the refill taken off the compute warps' critical path (the early-termination regime: at
// Eb/N0 = 3.5 dB a word leaves after ~5 iterations, so some lane of a CTA retires in almost every iteration).
//
// ldpc_ms_qc_kernel above retires and refills inside the loop top: atomic on the work queue, cp.async of 4-byte
// elements from L2, three CTA-wide barriers, all exposed.  Here:
//   * a STAGING RING of NS = G/2 whole codewords sits in shared memory.  Thread 0 claims word indices from the global
//     queue at the loop top -- exactly as many as the ring can take, the atomic's latency passing under the retire /
//     start work that follows -- and warp 0 pulls each codeword with ONE bulk asynchronous copy
//     (cp.async.bulk.shared.global, the 1-D TMA path: N*4 contiguous bytes) that signals an mbarrier when the bytes
//     have landed.  Nobody waits for a copy: once per iteration warp 0 probes the barriers (test_wait, non-blocking) and
//     ASSIGNS landed codewords to the lanes that will be free at the next loop top.
//   * warp 0 publishes ONE control word per iteration (who retires, who starts from which slot, who is live); a loop
//     top at which nothing happens costs one shared-memory read.  Retiring and starting are out of line
//     (qc_ring_event): no barrier when lanes only retire, one when lanes only start, two when both; a lane retires
//     and restarts in the same loop top, and no global-memory latency is on any warp's path.
// Arithmetic, table layout and outputs are those of ldpc_ms_qc_kernel (bit-exact with Coder::decodeCPU).
// Requires p.llr 16-byte aligned (bulk copy); the host falls back to ldpc_ms_qc_kernel otherwise.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// NON-blocking probe (test_wait; try_wait may suspend the thread for a system-dependent time when the phase is not complete,
// which stalled warp 0 -- a compute warp -- on every probe of a slot still in flight)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Static shared state of the ring kernel.  The hot loop touches only `flag`.  The out-of-line helpers address it
// through its 32-bit shared-window address (explicit LDS / STS: a reference would decay to generic loads).
struct QcRingShared {
    uint32_t flag[2][32];               // syndrome flags of the check pass, double buffered
    alignas(8) uint2 ctl;               // published by warp 0 after the mid barrier, read at the loop top when needed:
                                        //   x = lanes retiring | lanes starting << 8 | lanes live after the loop top << 16 | exit << 31
                                        //   y = staging slot of lane c in bits 4c .. 4c+3
    long long cw[8];                    // word held by lane c
    long long stage_cw[8];              // word held (or being loaded) by a ring slot
    alignas(8) unsigned long long full[8];  // mbarriers of the ring slots
    long long q[2];                     // words claimed from the work queue but not yet in the ring: [q0, q1)
    long long taken;                    // words this CTA has claimed so far
    uint32_t filled, phase;             // slots holding / loading a word; barrier parity per slot
    uint32_t exhausted;                 // a claim came back past the end of the batch
    unsigned long long t_idle;          // since when the CTA has been waiting for streamed input
};
#define QC_SH(field) ((uint32_t)offsetof(QcRingShared, field))
__device__ __forceinline__ uint32_t qc_lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void qc_sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ long long lds_s64(uint32_t a) { long long v; asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_s64(uint32_t a, long long v) { asm volatile("st.shared.s64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts_u32x2(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }

// Thread 0: how many words to claim from the global queue now.  `freed` = ring slots the lanes starting at this loop top
// are about to release, `unserved` = lanes that are free at this loop top and got no codeword.  Words are claimed when a
// ring slot can take them -- no chunks held back, so the split of a batch over the CTAs stays as even as claiming word
// by word -- and ahead of need only while the CTA is below its fair share of the batch (the tail is claimed on demand).
template <int NS>
__device__ __noinline__ int qc_ring_want(const QcParams& p, uint32_t sha, uint32_t freed, int unserved) {
    if (qc_lds_u32(sha + QC_SH(exhausted))) return 0;
    const uint32_t filled = qc_lds_u32(sha + QC_SH(filled)) & ~freed;
    const int pending = (int)(lds_s64(sha + QC_SH(q[1])) - lds_s64(sha + QC_SH(q[0])));
    if (pending > 0) return 0;  // (streamed input that has not landed yet: stage what was claimed before claiming more)
    int room = NS - __popc(filled);
    if (room <= 0) return 0;
    const long long fair = p.ncw / gridDim.x;
    if (lds_s64(sha + QC_SH(taken)) >= fair) {  // past the fair share: only what free lanes need beyond what is on its way
        const int shortfall = unserved - __popc(filled);
        room = room < shortfall ? room : shortfall;
    }
    return room > 0 ? room : 0;
}

// Warp 0, all lanes: put claimed words into the empty slots of the ring -- lane i serves slot i, one bulk copy per
// codeword.  (first, k) = a fresh claim of k words starting at `first` (k may be 0).
template <int NS>
__device__ __noinline__ void qc_ring_stage(const QcParams& p, uint32_t sha, uint32_t ring, int lane, long long first, int k) {
    const uint32_t wbytes = (uint32_t)p.N * 4u;
    long long q0 = lds_s64(sha + QC_SH(q[0])), q1 = lds_s64(sha + QC_SH(q[1]));
    __syncwarp();
    if (k > 0) {
        if (q0 >= q1) { q0 = first; q1 = first + k; }   // (claims are made only when nothing is pending)
        if (q1 > p.ncw) { q1 = p.ncw; if (lane == 0) qc_sts_u32(sha + QC_SH(exhausted), 1u); }
        if (q0 > q1) q0 = q1;
        if (lane == 0) sts_s64(sha + QC_SH(taken), lds_s64(sha + QC_SH(taken)) + (q1 - q0));
    }
    uint32_t filled = qc_lds_u32(sha + QC_SH(filled));
    long long have = q1 - q0;
    if (have > 0 && p.avail) {  // streamed input: only words that have landed in device memory
        unsigned long long a;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p.avail) : "memory");
        const long long lim = (long long)a - q0;
        have = have < lim ? have : lim;
    }
    const uint32_t empty = ~filled & ((1u << NS) - 1u);
    if (have > 0 && empty) {
        const int rank = __popc(empty & ((1u << lane) - 1u));
        const bool mine = lane < NS && ((empty >> lane) & 1u) && rank < have;
        if (mine) {
            const long long w = q0 + rank;
            sts_s64(sha + QC_SH(stage_cw[0]) + 8u * (uint32_t)lane, w);
            const uint32_t bar = sha + QC_SH(full[0]) + 8u * (uint32_t)lane;
            mbar_expect_tx(bar, wbytes);
            bulk_load(ring + (uint32_t)lane * wbytes, p.llr + (size_t)w * p.N, wbytes, bar);
        }
        const uint32_t took = __ballot_sync(0xffffffffu, mine);
        filled |= took;
        q0 += __popc(took);
    }
    __syncwarp();
    if (lane == 0) {
        sts_s64(sha + QC_SH(q[0]), q0); sts_s64(sha + QC_SH(q[1]), q1);
        qc_sts_u32(sha + QC_SH(filled), filled);
    }
    __syncwarp();
}

// Loop-top event (every thread of the CTA calls it, `ctl` is CTA-uniform): lanes in ctl.x[0..7] retire, lanes in
// ctl.x[8..15] start the codeword in the ring slot ctl.y names.  Kept out of line so that its registers and address
// arithmetic stay out of the hot loop.  Barriers: none when lanes only retire, one when lanes only start, two when
// both (the bytes of a retiring word are gathered from every thread's T rows before its lane is overwritten).
template <class P>
__device__ __noinline__ void qc_ring_event(const QcParams& p, uint32_t sha, uint32_t sb, int warp_in, int lane, uint2 ctl, int it) {
    constexpr int G = P::G, SUB = 32 / G, Z = P::Z, NT = P::W * 32;
    constexpr uint32_t ROWB = (uint32_t)G * 4u;
    constexpr uint32_t RS = (uint32_t)(Z + SUB) * ROWB;
    constexpr uint32_t WRAP = (uint32_t)Z * ROWB;
    const int warp = __shfl_sync(0xffffffffu, warp_in, 0);  // warp-uniform again: the table reads below stay LDCU
    const int c = lane & (G - 1), h = lane / G;
    const uint32_t rmask = ctl.x & 0xffu, smask = (ctl.x >> 8) & 0xffu;
    if (rmask) {
        // toChar (decodeCL.c:188-199): bit n = !(P > 0) = !signbit(T).  A thread takes one variable of the retiring
        // word, a ballot packs 32 of them, lanes 0..3 of the warp store the four bytes.
        const int KB = (p.K + 7) >> 3, NB8 = (p.N + 7) >> 3;
        const int nbits = p.hard ? p.N : p.K;
        for (uint32_t m = rmask; m; m &= m - 1u) {
            const int cc = __ffs((int)m) - 1;
            const long long w = lds_s64(sha + QC_SH(cw[0]) + 8u * (uint32_t)cc);
            for (int n0 = warp * 32; n0 < nbits; n0 += NT) {   // (warp-uniform trip count: every lane reaches the ballot)
                const int n = n0 + lane;
                float t = -1.0f;
                if (n < p.N) t = lds_f32(sb + (uint32_t)((n / Z) * (Z + SUB) + (n % Z)) * ROWB + (uint32_t)cc * 4u);
                const uint32_t bal = __ballot_sync(0xffffffffu, n < p.N && (__float_as_uint(t) >> 31) == 0u);
                if (p.post && n < p.N) p.post[(size_t)w * p.N + n] = -t;
                const int b = (n0 >> 3) + lane;
                if (lane < 4) {
                    const uint32_t v = (bal >> (8 * lane)) & 0xffu;
                    if (p.info && b < KB) {
                        // (the last info byte may hold parity bits when K is not a multiple of 8: mask them)
                        const uint32_t keep = (b == KB - 1 && (p.K & 7)) ? ((1u << (p.K & 7)) - 1u) : 0xffu;
                        p.info[(size_t)w * KB + b] = (uint8_t)(v & keep);
                    }
                    if (p.hard && b < NB8) p.hard[(size_t)w * NB8 + b] = (uint8_t)v;
                }
            }
            if (p.post && !p.hard)  // posteriors of the parity part too
                for (int n = p.K + (int)threadIdx.x; n < p.N; n += NT)
                    p.post[(size_t)w * p.N + n] = -lds_f32(sb + (uint32_t)((n / Z) * (Z + SUB) + (n % Z)) * ROWB + (uint32_t)cc * 4u);
        }
        if (p.iters && warp == 0 && h == 0 && ((rmask >> c) & 1u)) p.iters[lds_s64(sha + QC_SH(cw[0]) + 8u * (uint32_t)c)] = it;
        if (smask) __syncthreads();  // every read of the retiring lanes' T is done before a starting lane overwrites it
    }
    if (smask) {
        if ((smask >> c) & 1u) {
            // decodeInitMS (decodeCL.c:113-124): T = -y (canonical zero), R = 0
            const QcWarpTab<P>& tb = qc_tab<P>(p.tab_slot, warp);
            const uint32_t la = sb + (uint32_t)lane * 4u;
            const int slot = (int)((ctl.y >> (4 * c)) & 15u);
            const uint32_t src = sb + p.ring_off + (uint32_t)slot * ((uint32_t)p.N * 4u) + (uint32_t)h * 4u;
            float y[P::VS];
#pragma unroll
            for (int s = 0; s < P::VS; ++s) y[s] = lds_f32(src + tb.var0[s] * 4u);
#pragma unroll
            for (int s = 0; s < P::VS; ++s) {
                const float v = __fadd_rn(-y[s], 0.0f);
                sts_f32(la + tb.vn_t[s], v);
                if (tb.vdup) sts_f32(la + tb.vn_t[s] + WRAP, v);
            }
            const bool dup = tb.cdup != 0u;
#pragma unroll
            for (int cs = 0; cs < P::CS; ++cs) {
#pragma unroll
                for (int j = 0; j < P::cdeg(cs); ++j) {
                    sts_f32(la + tb.cn_r[cs] + (uint32_t)j * RS, 0.0f);
                    if (dup) sts_f32(la + tb.cn_r[cs] + (uint32_t)j * RS - WRAP, 0.0f);
                }
            }
            if (warp == 0 && h == 0) sts_s64(sha + QC_SH(cw[0]) + 8u * (uint32_t)c, lds_s64(sha + QC_SH(stage_cw[0]) + 8u * (uint32_t)slot));
        }
        __syncthreads();  // T / R of the starting lanes are in place before the check pass gathers them
        if (threadIdx.x == 0) {  // every assigned slot was taken: free it, flip its barrier parity
            uint32_t freed = 0u;
            for (uint32_t m = smask; m; m &= m - 1u) freed |= 1u << ((ctl.y >> (4 * (__ffs((int)m) - 1))) & 15u);
            qc_sts_u32(sha + QC_SH(filled), qc_lds_u32(sha + QC_SH(filled)) & ~freed);
            qc_sts_u32(sha + QC_SH(phase), qc_lds_u32(sha + QC_SH(phase)) ^ freed);
        }
    }
}

// After the mid barrier, warp 0 (all lanes): hand landed codewords to the lanes that will be free at the next loop top
// and publish the control word.  need = lanes idle or about to retire, lv = lanes live (a finished lane still has to
// emit), rnext = lanes that retire at the next loop top.  Lane i probes slot i; the k-th needy lane gets the k-th
// landed slot.
template <int NS, int G>
__device__ __noinline__ void qc_ring_assign(const QcParams& p, uint32_t sha, int lane, uint32_t need, uint32_t lv, uint32_t rnext) {
    constexpr uint32_t GMASK = (1u << G) - 1u;
    const uint32_t filled = qc_lds_u32(sha + QC_SH(filled));
    const uint32_t phase = qc_lds_u32(sha + QC_SH(phase));
    bool ready = false;
    if (lane < NS && ((filled >> lane) & 1u)) ready = mbar_test_wait(sha + QC_SH(full[0]) + 8u * (uint32_t)lane, (phase >> lane) & 1u);
    const uint32_t rdy = __ballot_sync(0xffffffffu, ready);
    int slot = -1;
    if (lane < G && ((need >> lane) & 1u)) {
        const int k = __popc(need & ((1u << lane) - 1u));
        if (k < __popc(rdy)) slot = (int)__fns(rdy, 0, k + 1);
    }
    const uint32_t smask = __ballot_sync(0xffffffffu, slot >= 0) & GMASK;
    const uint32_t slots = __reduce_or_sync(0xffffffffu, slot >= 0 ? (uint32_t)slot << (4 * lane) : 0u);
    uint32_t ex = 0u;
    if (lv == 0u && smask == 0u && filled == 0u) {  // every word has left, nothing was handed out, nothing is on its way
        const bool pending = lds_s64(sha + QC_SH(q[0])) < lds_s64(sha + QC_SH(q[1]));
        if (!pending && qc_lds_u32(sha + QC_SH(exhausted))) {
            ex = 1u;                 // queue exhausted, ring empty: done
        } else if (pending && p.avail) {
            unsigned long long now;  // waiting for streamed input: bounded
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            int st = 0;
            if (p.status) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(st) : "l"(p.status) : "memory");
            unsigned long long t0 = (unsigned long long)lds_s64(sha + QC_SH(t_idle));
            __syncwarp();
            if (t0 == 0ull) { t0 = now; if (lane == 0) sts_s64(sha + QC_SH(t_idle), (long long)now); }
            if (st != 0 || now - t0 > p.wait_ns) {
                if (p.status && lane == 0) atomicExch(p.status, 1);
                ex = 1u;
            }
            ex = __shfl_sync(0xffffffffu, ex, 0);
        }
    } else if (lane == 0) {
        sts_s64(sha + QC_SH(t_idle), 0ll);
    }
    const uint32_t live_after = ((lv & ~rnext) | smask) & GMASK;
    if (lane == 0) sts_u32x2(sha + QC_SH(ctl), make_uint2(rnext | (smask << 8) | (live_after << 16) | (ex << 31), slots));
}

template <class P>
__global__ void __launch_bounds__(P::W * 32, (P::W * 32 <= 288 ? 3 : (P::W * 32 <= 384 ? 2 : 1)))
ldpc_ms_qc_ring_kernel(const __grid_constant__ QcParams p) {
    constexpr int G = P::G, SUB = 32 / G, Z = P::Z, NS = (G >= 2 ? G / 2 : 1);
    constexpr uint32_t GMASK = (1u << G) - 1u;
    static_assert(G <= 8 && NS <= 8, "control word holds 8 lanes");
    static_assert(Z % 8 == 0, "the eight variables of a byte sit in one block column");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ QcRingShared sh;

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int c = lane & (G - 1), h = lane / G;
    const QcWarpTab<P>& tb = qc_tab<P>(p.tab_slot, warp);
    const uint32_t sb = smem_u32(smem_raw);
    const uint32_t la = sb + (uint32_t)lane * 4u;
    const uint32_t sha = smem_u32(&sh);
    unsigned int* const ctr32 = reinterpret_cast<unsigned int*>(p.counter64);  // (a launch holds fewer than 2^31 words)

    if (threadIdx.x < 32) {
        sts_f32(sb + p.t_bytes + p.r_bytes + (uint32_t)lane * 4u, 0.0f);
        sts_f32(sb + p.t_bytes + p.r_bytes + 128u + (uint32_t)lane * 4u, -INFINITY);
        sh.flag[0][lane] = 0u; sh.flag[1][lane] = 0u;
        if (lane < 8) { sh.cw[lane] = -1; sh.stage_cw[lane] = -1; }
        if (lane < NS) mbar_init(sha + QC_SH(full[0]) + 8u * (uint32_t)lane, 1u);
        long long first = 0;
        if (lane == 0) {
            sh.ctl = make_uint2(0u, 0u);
            sh.filled = 0u; sh.phase = 0u; sh.t_idle = 0ull; sh.exhausted = 0u; sh.taken = 0; sh.q[0] = 0; sh.q[1] = 0;
            first = (long long)atomicAdd(ctr32, (unsigned int)NS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        first = __shfl_sync(0xffffffffu, first, 0);
        qc_ring_stage<NS>(p, sha, sb + p.ring_off, lane, first, NS);
    }
    __syncthreads();

    float yn[P::VS];
#pragma unroll
    for (int s = 0; s < P::VS; ++s) yn[s] = -1.0f;
    bool live = false, done = false;
    int it = 0;
    uint32_t ph = 0;
    for (;;) {
        // ---- loop top: nothing to do unless some lane is idle or has finished (a warp holds every lane: the vote is CTA-uniform)
        bool any_live = true;
        // (refill_wait = n > 1: lanes retire and start only at every n-th loop top -- fewer, fuller events at the price of
        // finished lanes idling; measured in profiles/r02_ring_gate.txt)
        const bool gate = p.refill_wait <= 1 || (ph % (uint32_t)p.refill_wait) == 0u || !__any_sync(0xffffffffu, live);
        if (gate && __any_sync(0xffffffffu, !live || done)) {
            const uint2 ctl = sh.ctl;
            // thread 0 claims the words the ring can take after this loop top; the atomic's latency passes under the event
            unsigned int ticket = 0u;
            int k = 0;
            if (threadIdx.x == 0) {
                uint32_t freed = 0u;
                for (uint32_t m = (ctl.x >> 8) & 0xffu; m; m &= m - 1u) freed |= 1u << ((ctl.y >> (4 * (__ffs((int)m) - 1))) & 15u);
                const uint32_t free_lanes = (~(ctl.x >> 16)) & GMASK;  // lanes with no word after this loop top
                k = qc_ring_want<NS>(p, sha, freed, __popc(free_lanes));
                if (k > 0) ticket = atomicAdd(ctr32, (unsigned int)k);
            }
            if (ctl.x & 0xffffu) {
                qc_ring_event<P>(p, sha, sb, warp, lane, ctl, it);
                if ((ctl.x >> c) & 1u) live = false;                       // retired
                if ((ctl.x >> (8 + c)) & 1u) {                              // started: channel values back into registers
#pragma unroll
                    for (int s = 0; s < P::VS; ++s) yn[s] = lds_f32(la + tb.vn_t[s]);
                    live = true; done = false; it = 0;
                }
            }
            if (ctl.x >> 31) break;
            if (warp == 0) {
                k = __shfl_sync(0xffffffffu, k, 0);
                const long long first = (long long)__shfl_sync(0xffffffffu, ticket, 0);
                qc_ring_stage<NS>(p, sha, sb + p.ring_off, lane, first, k);
            }
            any_live = ((ctl.x >> 16) & GMASK) != 0u;
        }

        if (any_live) {
            // check-node pass + syndrome of the previous posterior
            const uint32_t unsat = tb.cdup ? qc_cn_static<P, 0, true>(tb, la) : qc_cn_static<P, 0, false>(tb, la);
            const bool check = p.early_term && it >= 1 && live && !done;
            if (check && unsat) sh.flag[ph & 1][c] = 1u;  // same-value race, benign
            __syncthreads();
            if (check && sh.flag[ph & 1][c] == 0u) done = true;
            const bool last = live && !done && it + 1 == p.max_iter;  // this variable pass is the word's last
            if (warp == 0) {
                // control word of the next loop top (only if a lane will be idle or retiring then), written now so that
                // it is long in place at the end barrier
                sh.flag[(ph + 1) & 1][lane] = 0u;
                const uint32_t lv = __ballot_sync(0xffffffffu, h == 0 && live) & GMASK;
                const uint32_t rnext = __ballot_sync(0xffffffffu, h == 0 && live && (done || last)) & GMASK;
                if (lv != GMASK || rnext != 0u) qc_ring_assign<NS, G>(p, sha, lane, (~lv | rnext) & GMASK, lv, rnext);
            }
            ++ph;
            // variable-node pass (the posterior of finished / idle lanes is frozen)
            if (tb.vdup) qc_vn_static<P, 0, true>(tb, la, yn, !(live && !done));
            else qc_vn_static<P, 0, false>(tb, la, yn, !(live && !done));
            if (live && !done) {
                ++it;
                if (last) done = true;
            }
        } else {
            __nanosleep(100);  // nothing to decode yet (start of the launch, or streamed input not there)
            __syncthreads();   // every warp has read ctl before warp 0 rewrites it
            if (warp == 0) qc_ring_assign<NS, G>(p, sha, lane, GMASK, 0u, 0u);
        }
        __syncthreads();
    }
}

}  // namespace ldpc_b200
