// ldpc_warp.cuh -- the textbook mapping, kept as a measured alternative (opt-in, LDPC_B200_PATH_WARP):
// one codeword per CTA, a SUB-WARP PER CHECK whose row reductions run on shuffles and ballots.
//
//   check pass (refreshRMS, decodeCL.c:126-147): lane j of a sub-warp of SW = 8/16/32 lanes holds edge j of its
//     check: S_j = T[col_j] + R_old_j (= -Q_j); {min1, min2} of |S| by a log2(SW)-stage __shfl_xor butterfly;
//     the sign parity of the row and the syndrome bit by __ballot_sync + popc; R_new_j written back in place.
//   variable pass (refreshPostPMS / refreshQMS, decodeCL.c:149-186): a thread per variable,
//     T = (-y) - R_1 - R_2 ... in ascending-row order (the reference's column-list order).
//
// Same arithmetic contract as every other path (bit-exact with Coder::decodeCPU, MyLdpc.cpp:684-784).  Why it is
// not the default: a warp instruction here covers 32 edges of one codeword, exactly as many edge-units as the
// group/QC kernels' (4 nodes x 8 codewords), but every butterfly stage costs two SHFLs that go down the same
// MIO/LSU pipe as the shared-memory accesses -- ~11 pipe slots per edge-instruction against 4.75 -- and the
// gathers are in natural order (bank conflicts).  Measured numbers: profiles/r01_r_warp_per_check.txt.
#pragma once
#include "ldpc_kernels.cuh"

namespace ldpc_b200 {

struct WarpParams {
    const uint32_t* __restrict__ cn_col;   // [M][SW] column of edge j of each check, 0xffffffff = padding
    const int32_t* __restrict__ col_ptr;   // [N+1]
    const uint32_t* __restrict__ vn_pos;   // [nnz] position check*SW + j of each variable's edges, ascending row
    int M, N, K, SW;                       // SW = sub-warp width (power of two >= max check degree)
    int max_iter, early_term;
    const float* __restrict__ llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;
    unsigned long long* counter64;
};

template <int SW>
__global__ void __launch_bounds__(1024, 1) ldpc_ms_warp_kernel(const __grid_constant__ WarpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Y = reinterpret_cast<float*>(smem_raw);  // -y, canonical zero
    float* T = Y + p.N;                             // negated posterior
    float* R = T + p.N;                             // [M][SW] check-to-variable messages, check-major
    __shared__ long long s_cw;
    __shared__ uint32_t s_flag[2];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const int sub = lane & (SW - 1);                // edge slot inside the check
    const uint32_t submask = (SW == 32 ? 0xffffffffu : ((1u << SW) - 1u)) << (lane & ~(SW - 1));
    const int nslots = p.M * SW;

    for (;;) {
        if (tid == 0) s_cw = (long long)atomicAdd(p.counter64, 1ull);
        __syncthreads();
        const long long cw = s_cw;
        if (cw >= p.ncw) break;
        const float* src = p.llr + (size_t)cw * p.N;
        for (int n = tid; n < p.N; n += nthr) {
            const float y = __fadd_rn(-__ldg(src + n), 0.0f);
            Y[n] = y;
            T[n] = y;
        }
        for (int e = tid; e < nslots; e += nthr) R[e] = 0.0f;  // decodeInitMS: Q = y
        if (tid < 2) s_flag[tid] = 0u;
        __syncthreads();
        int it = 0;
        for (;;) {
            // ---- check pass: one sub-warp per check, reductions on shuffles / ballots
            uint32_t unsat = 0u;
            for (int e = tid; e < ((nslots + 31) & ~31); e += nthr) {  // whole warps stay together for the shuffles
                const bool inr = e < nslots;
                const uint32_t col = inr ? __ldg(p.cn_col + e) : 0xffffffffu;
                const bool valid = col != 0xffffffffu;
                const float t = valid ? T[col] : 0.0f;
                const float S = valid ? __fadd_rn(t, R[e]) : 0.0f;  // = -Q
                float m1 = valid ? fabsf(S) : INFINITY, m2 = INFINITY;
                const float a = m1;
#pragma unroll
                for (int off = SW / 2; off >= 1; off >>= 1) {
                    const float o1 = __shfl_xor_sync(0xffffffffu, m1, off);
                    const float o2 = __shfl_xor_sync(0xffffffffu, m2, off);
                    m2 = fminf(fmaxf(m1, o1), fminf(m2, o2));
                    m1 = fminf(m1, o1);
                }
                const uint32_t bS = __ballot_sync(0xffffffffu, valid && (__float_as_uint(S) >> 31)) & submask;
                const uint32_t bT = __ballot_sync(0xffffffffu, valid && (__float_as_uint(t) >> 31)) & submask;
                const uint32_t D = (uint32_t)__popc(__ballot_sync(0xffffffffu, valid) & submask);
                // (Q_j < 0) = !signbit(S_j): parity of the negative Q's = (D & 1) ^ parity(signbit S);
                // sign(R_j) = that parity ^ (Q_j < 0)
                const uint32_t flip = ((((uint32_t)__popc(bS) ^ D ^ 1u) & 1u) << 31);
                const uint32_t mag = __float_as_uint(fminf((a == m1) ? m2 : m1, kClamp));
                if (valid) R[e] = __uint_as_float((__float_as_uint(S) & 0x80000000u) ^ mag ^ flip);
                unsat |= ((uint32_t)__popc(bT) ^ D) & 1u;  // hard bit = !signbit(T)
            }
            const bool check = p.early_term && it >= 1;
            if (check && unsat) s_flag[it & 1] = 1u;
            __syncthreads();
            if (check && s_flag[it & 1] == 0u) break;  // converged: T holds the posterior that passed
            if (tid == 0) s_flag[(it + 1) & 1] = 0u;
            // ---- variable pass: a thread per variable, ascending-row sum
            for (int n = tid; n < p.N; n += nthr) {
                float acc = Y[n];
                const int k1 = __ldg(p.col_ptr + n + 1);
                for (int k = __ldg(p.col_ptr + n); k < k1; ++k) acc = __fsub_rn(acc, R[__ldg(p.vn_pos + k)]);
                T[n] = acc;
            }
            ++it;
            __syncthreads();
            if (it == p.max_iter) break;
        }
        // ---- outputs (toChar, decodeCL.c:188-199)
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = tid; b < KB; b += nthr) {
                uint32_t v = 0u;
                for (int t8 = 0; t8 < 8; ++t8) {
                    const int n = b * 8 + t8;
                    if (n < p.K) v |= ((~__float_as_uint(T[n])) >> 31) << t8;
                }
                p.info[(size_t)cw * KB + b] = (uint8_t)v;
            }
        }
        if (p.hard) {
            const int NB = (p.N + 7) >> 3;
            for (int b = tid; b < NB; b += nthr) {
                uint32_t v = 0u;
                for (int t8 = 0; t8 < 8; ++t8) {
                    const int n = b * 8 + t8;
                    if (n < p.N) v |= ((~__float_as_uint(T[n])) >> 31) << t8;
                }
                p.hard[(size_t)cw * NB + b] = (uint8_t)v;
            }
        }
        if (p.post)
            for (int n = tid; n < p.N; n += nthr) p.post[(size_t)cw * p.N + n] = -T[n];
        if (p.iters && tid == 0) p.iters[cw] = it;
        __syncthreads();
    }
}

}  // namespace ldpc_b200
