// ldpc_stream.cuh -- flooding min-sum for codes whose messages do not fit on chip (N = 64800 class), sm_100a.
//
// Same schedule and arithmetic as ldpc_ms_lane_kernel<false> (32 codewords per CTA, lane = codeword, every
// array [index][32 lanes] in a CTA-private global workspace, so each warp access is one coalesced 128-byte
// row), rebuilt around what bounds that path: HBM latency x bytes in flight.
//   * explicit per-edge messages R[check*8 + j][32] instead of the compressed check state: 16 bytes per edge
//     and iteration (P gather, R read, R write in the check pass; R read in the variable pass) -- exactly the
//     algorithmic traffic, against 16 + 24/check for the compressed form -- and a check's messages are one
//     contiguous 1 KB block;
//   * fixed-stride index tables (8 entries per check, 8 entries per variable bundle, 0xffffffff = none), read
//     as two warp-uniform 16-byte loads: no row_ptr -> column -> data dependency chain;
//   * the next node's indices are fetched while the current node's data is in flight, and low-degree variables
//     are processed in bundles (1 x degree<=8, 2 x degree<=4, 4 x degree<=2), so every warp keeps 9..16
//     independent 128-byte rows in flight instead of 3..11.
// Codes with check or variable degree above 8 use ldpc_ms_lane_kernel<false>.
#pragma once

#include "ldpc_kernels.cuh"

namespace ldpc_b200 {

constexpr uint32_t kStreamNone = 0xffffffffu;

struct StreamParams {
    const uint4* __restrict__ cn_tab;        // [M][2]      column positions of the check's edges
    const uint4* __restrict__ vn_tab;        // [bundles][2] R row (check*8 + j) of each edge, ascending row per variable
    const uint32_t* __restrict__ var_of_pos; // [NP] variable at a position, kStreamNone = phantom
    const uint32_t* __restrict__ pos_of_var; // [N]
    int M, N, K, NP;
    int nb8, nb4, nb2;  // bundles of 1 / 2 / 4 variables; positions: [0, nb8) | 2 per bundle | 4 per bundle
    int max_iter, early_term;
    const float* __restrict__ llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;
    float* ws;
    size_t ws_stride;  // floats per CTA: (2 NP + 8 M) * 32
    unsigned int* counter;
    int ngroups;
};

__device__ __forceinline__ void stream_unpack(const uint4 a, const uint4 b, uint32_t (&e)[8]) {
    e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
}

// Variable bundles of VPB variables x (8 / VPB) edges, bundle index b0, b0 + W, ... < nb; the first variable
// of bundle b sits at position pos0 + b * VPB.
template <int VPB>
__device__ __forceinline__ void stream_vn_class(const StreamParams& p, const uint4* __restrict__ tab, int nb, int pos0, int warp, int W,
                                                int lane, const float* __restrict__ R, const float* __restrict__ Y, float* __restrict__ P, bool done) {
    constexpr int S = 8 / VPB;
    int b = warp;
    if (b >= nb) return;
    uint32_t idx[8];
    stream_unpack(__ldg(tab + 2 * (size_t)b), __ldg(tab + 2 * (size_t)b + 1), idx);
    while (b < nb) {
        const int bn = b + W;
        uint4 na = make_uint4(0, 0, 0, 0), nb4 = make_uint4(0, 0, 0, 0);
        if (bn < nb) { na = __ldg(tab + 2 * (size_t)bn); nb4 = __ldg(tab + 2 * (size_t)bn + 1); }  // next indices: in flight with this bundle's data
        float r[8], acc[VPB];
#pragma unroll
        for (int t = 0; t < 8; ++t) r[t] = (idx[t] != kStreamNone) ? R[(size_t)idx[t] * kLanes + lane] : 0.0f;
        const size_t pa = (size_t)(pos0 + b * VPB) * kLanes + lane;
#pragma unroll
        for (int i = 0; i < VPB; ++i) acc[i] = Y[pa + (size_t)i * kLanes];
#pragma unroll
        for (int i = 0; i < VPB; ++i)
#pragma unroll
            for (int k = 0; k < S; ++k) acc[i] = __fadd_rn(acc[i], r[i * S + k]);  // ascending-row order; x + 0.0f = x (P is never -0)
        if (!done) {
#pragma unroll
            for (int i = 0; i < VPB; ++i) P[pa + (size_t)i * kLanes] = acc[i];
        }
        stream_unpack(na, nb4, idx);
        b = bn;
    }
}

template <int MAX_THREADS>
__global__ void __launch_bounds__(MAX_THREADS, 1) ldpc_ms_stream_kernel(const __grid_constant__ StreamParams p) {
    __shared__ int s_group;
    __shared__ uint32_t s_flag[2][kLanes];

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int W = blockDim.x >> 5;
    const int M = p.M, N = p.N, NP = p.NP;

    float* P = p.ws + (size_t)blockIdx.x * p.ws_stride;
    float* Y = P + (size_t)NP * kLanes;
    float* R = Y + (size_t)NP * kLanes;

    for (;;) {
        if (threadIdx.x == 0) s_group = (int)atomicAdd(p.counter, 1u);
        __syncthreads();
        const int g = s_group;
        if (g >= p.ngroups) break;
        const long long cw = (long long)g * kLanes + lane;
        const bool active = cw < p.ncw;
        const float* src = p.llr + (size_t)(active ? cw : 0) * N;

        // ---- load: y -> Y and P (decodeInitMS, decodeCL.c:113-124); messages start at zero (iteration 0 reads none)
        for (int pos = warp; pos < NP; pos += W) {
            const uint32_t v = __ldg(p.var_of_pos + pos);
            const float y = __fadd_rn((active && v != kStreamNone) ? __ldg(src + v) : 1.0f, 0.0f);
            P[(size_t)pos * kLanes + lane] = y;
            Y[(size_t)pos * kLanes + lane] = y;
        }
        if (warp == 0) { s_flag[0][lane] = 0u; s_flag[1][lane] = 0u; }
        __syncthreads();

        bool done = !active;
        int my_iters = 0;
        int iter = 0;
        for (;;) {
            // ---- check-node pass (refreshQMS folded in) + syndrome of the previous posterior
            uint32_t unsat = 0u;
            {
                int c = warp;
                uint32_t col[8];
                if (c < M) stream_unpack(__ldg(p.cn_tab + 2 * (size_t)c), __ldg(p.cn_tab + 2 * (size_t)c + 1), col);
                while (c < M) {
                    const int cn = c + W;
                    uint4 na = make_uint4(0, 0, 0, 0), nb = make_uint4(0, 0, 0, 0);
                    if (cn < M) { na = __ldg(p.cn_tab + 2 * (size_t)cn); nb = __ldg(p.cn_tab + 2 * (size_t)cn + 1); }
                    float q[8], ro[8];
                    float* rrow = R + (size_t)c * 8 * kLanes + lane;
#pragma unroll
                    for (int t = 0; t < 8; ++t) q[t] = (col[t] != kStreamNone) ? P[(size_t)col[t] * kLanes + lane] : INFINITY;
#pragma unroll
                    for (int t = 0; t < 8; ++t) ro[t] = (col[t] != kStreamNone && iter > 0) ? rrow[t * kLanes] : 0.0f;
                    uint32_t syn = 0u, par = 0u;
                    float m1 = INFINITY, m2 = INFINITY;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        syn ^= (q[t] > 0.0f) ? 0u : 1u;        // hard bit of the previous posterior
                        q[t] = __fsub_rn(q[t], ro[t]);         // Q = P - R (refreshQMS)
                        par ^= __float_as_uint(q[t]);
                        const float a = fabsf(q[t]);
                        m2 = fminf(m2, fmaxf(m1, a));
                        m1 = fminf(m1, a);
                    }
                    const uint32_t m1c = __float_as_uint(fminf(m1, kClamp)), m2c = __float_as_uint(fminf(m2, kClamp));
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const uint32_t mag = (fabsf(q[t]) == m1) ? m2c : m1c;  // exclude-self minimum (ties: m2 == m1)
                        const uint32_t rn = mag ^ ((par ^ __float_as_uint(q[t])) & 0x80000000u);
                        if (col[t] != kStreamNone) rrow[t * kLanes] = __uint_as_float(rn);
                    }
                    unsat |= syn;
                    stream_unpack(na, nb, col);
                    c = cn;
                }
            }
            const bool check = p.early_term && iter >= 1;
            if (check && unsat) s_flag[iter & 1][lane] = 1u;
            __syncthreads();
            if (check && !done && s_flag[iter & 1][lane] == 0u) { done = true; my_iters = iter; }
            if (__all_sync(0xffffffffu, done)) break;
            if (warp == 0) s_flag[(iter + 1) & 1][lane] = 0u;

            // ---- variable-node pass: posterior in ascending-row order; frozen once a word is done
            stream_vn_class<1>(p, p.vn_tab, p.nb8, 0, warp, W, lane, R, Y, P, done);
            stream_vn_class<2>(p, p.vn_tab + 2 * (size_t)p.nb8, p.nb4, p.nb8, warp, W, lane, R, Y, P, done);
            stream_vn_class<4>(p, p.vn_tab + 2 * ((size_t)p.nb8 + p.nb4), p.nb2, p.nb8 + 2 * p.nb4, warp, W, lane, R, Y, P, done);
            ++iter;
            if (iter == p.max_iter) {
                if (!done) my_iters = iter;
                break;
            }
            __syncthreads();
        }
        __syncthreads();

        // ---- outputs (toChar, decodeCL.c:188-199)
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = warp; b < KB; b += W) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.K) v |= ((P[(size_t)__ldg(p.pos_of_var + n) * kLanes + lane] > 0.0f) ? 0u : 1u) << t;
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
                    if (n < N) v |= ((P[(size_t)__ldg(p.pos_of_var + n) * kLanes + lane] > 0.0f) ? 0u : 1u) << t;
                }
                if (active) p.hard[(size_t)cw * NB + b] = (uint8_t)v;
            }
        }
        if (p.post && active) {
            for (int n = warp; n < N; n += W) p.post[(size_t)cw * N + n] = P[(size_t)__ldg(p.pos_of_var + n) * kLanes + lane];
        }
        if (p.iters && warp == 0 && active) p.iters[cw] = my_iters;
        __syncthreads();
    }
}

}  // namespace ldpc_b200
