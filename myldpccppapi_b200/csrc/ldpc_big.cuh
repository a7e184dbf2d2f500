// ldpc_big.cuh -- sum-product and layered min-sum for codes of ANY size (sm_100a).
//
// The reference's OpenCL kernels have no size limit (every message lives in global memory: decodeCL.c:25-62, 203-292;
// host loops MyLdpc.cpp:889-1059).  The tuned kernels of ldpc_sp.cuh / ldpc_tdmp.cuh keep all messages of G
// codewords in shared memory and therefore accept only short codes; these two kernels take over where those do not
// fit (most of the reference's own family beyond z = 28, and the long synthetic codes), so that DecodeSP / DecodeTDMP
// never have to fall back to another algorithm.  Same arithmetic contract as the on-chip kernels, i.e. bit-exact
// with the oracle restatements that are pinned against the reference's own kernels (tests/test_oracle_vs_refcl.py).
//
// Layout: lane = codeword.  A CTA decodes 32 codewords at a time; every array sits in a CTA-private slice of a global
// workspace as [index][32 lanes], so a warp access is one 128-byte row and all index arithmetic is warp-uniform.
// Warps share out the checks and the variables; __syncthreads separates the phases of the reference's host loop.
// Hard bits are one 32-bit word per variable (bit l = codeword lane l): the syndrome of a row is the XOR of its
// columns' words, the tie rule "keeps the previous bit" by masking.  HBM/L2-bound by design -- correctness first.
#pragma once
#include "ldpc_sp.cuh"
#include "ldpc_tables.h"

namespace ldpc_b200 {

constexpr int kBigMaxDeg = 32;  // check / variable degree bound (the handle rejects check degree > 32 at create)

struct BigParams {
    const int32_t* __restrict__ row_ptr;   // [M+1]
    const uint32_t* __restrict__ cn_col;   // [nnz] column of edge e (check-major)
    const int32_t* __restrict__ col_ptr;   // [N+1]
    const uint32_t* __restrict__ vn_edge;  // [nnz] (check << 5) | position, per variable in ascending row order
    int M, N, K, nnz, max_iter, early_term;
    int z;                                 // layered: rows per layer
    int fused_layered;                     // fused-kernel arithmetic: 1 = decodeOnceTDMP (layered), 0 = decodeOnceMS (flooding)
    const float* __restrict__ llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;                           // layered only
    float* ws;                             // workspace, ws_stride floats per CTA
    size_t ws_stride;
    unsigned int* counter;
    int ngroups;
};

#ifdef LDPC_B200_BIG_KERNELS  // the kernels are instantiated in one translation unit (k_misc.cu); the host logic only needs BigParams
__device__ __forceinline__ int big_edge(const BigParams& p, uint32_t ve) { return p.row_ptr[ve >> kPosBits] + (int)(ve & ((1u << kPosBits) - 1u)); }

// shared tail of both kernels: per-lane stop rule (MyLdpc.cpp:1035-1039 / 946-958) and outputs (toChar, decodeCL.c:188-199)
__device__ __forceinline__ void big_emit(const BigParams& p, const uint32_t* HB, long long cw, bool valid, int my_iters, int lane, int warp, int nwarps) {
    const int KB = (p.K + 7) >> 3, NB8 = (p.N + 7) >> 3;
    if (p.info)
        for (int b = warp; b < KB; b += nwarps) {
            uint32_t v = 0u;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int n = b * 8 + t;
                if (n < p.K) v |= ((HB[n] >> lane) & 1u) << t;
            }
            if (valid) p.info[(size_t)cw * KB + b] = (uint8_t)v;
        }
    if (p.hard)
        for (int b = warp; b < NB8; b += nwarps) {
            uint32_t v = 0u;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int n = b * 8 + t;
                if (n < p.N) v |= ((HB[n] >> lane) & 1u) << t;
            }
            if (valid) p.hard[(size_t)cw * NB8 + b] = (uint8_t)v;
        }
    if (p.iters && warp == 0 && valid) p.iters[cw] = my_iters;
}

// ---------------------------------------------------------------------------------------------------------------------
// Sum-product (DecodeSP): decodeInit / refreshR / hardDecision / checkResult / refreshQ of decodeCL.c:3-108 under the
// loop of decodeOnceSP (MyLdpc.cpp:977-1059).  Workspace: QD[nnz] = q0 - q1 per edge, RD[nnz] = d per edge
// (r0 = (1+d)/2, r1 = (1-d)/2 are recomputed exactly), P0[N], P1[N], HB[N].
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) ldpc_sp_big_kernel(const __grid_constant__ BigParams p) {
    __shared__ uint32_t s_unsat;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* QD = p.ws + (size_t)blockIdx.x * p.ws_stride;
    float* RD = QD + (size_t)p.nnz * 32;
    float* P0 = RD + (size_t)p.nnz * 32;
    float* P1 = P0 + (size_t)p.N * 32;
    uint32_t* HB = reinterpret_cast<uint32_t*>(P1 + (size_t)p.N * 32);
    for (;;) {
        __shared__ int s_group;
        __syncthreads();
        if (threadIdx.x == 0) s_group = (int)atomicAdd(p.counter, 1u);
        __syncthreads();
        const int g = s_group;
        if (g >= p.ngroups) break;
        const long long cw = (long long)g * 32 + lane;
        const bool valid = cw < p.ncw;
        const float* y = p.llr + (size_t)(valid ? cw : 0) * p.N;
        // decodeInit (decodeCL.c:3-22): q0 = t/(1+t), q1 = 1/(1+t), t = exp(8 y); the prior is the same pair
        for (int n = warp; n < p.N; n += nwarps) {
            const float t = sp_expf(__fmul_rn(8.0f, y[n]));
            const float a = __fadd_rn(1.0f, t);
            const float p0 = __fdiv_rn(t, a), p1 = __fdiv_rn(1.0f, a);
            P0[(size_t)n * 32 + lane] = p0;
            P1[(size_t)n * 32 + lane] = p1;
            const float qd = __fsub_rn(p0, p1);
            for (int k = p.col_ptr[n]; k < p.col_ptr[n + 1]; ++k) QD[(size_t)big_edge(p, p.vn_edge[k]) * 32 + lane] = qd;
            if (lane == 0) HB[n] = 0u;
        }
        bool live = valid, done = !valid;
        int it = 0, my_iters = 0;
        if (threadIdx.x == 0) s_unsat = 0u;
        __syncthreads();
        for (;;) {
            // refreshR (decodeCL.c:25-41): d_e = product of (q0 - q1) over the row's OTHER edges, in row order
            for (int r = warp; r < p.M; r += nwarps) {
                const int e0 = p.row_ptr[r], d = p.row_ptr[r + 1] - e0;
                float x[kBigMaxDeg];
#pragma unroll
                for (int j = 0; j < kBigMaxDeg; ++j)
                    if (j < d) x[j] = QD[(size_t)(e0 + j) * 32 + lane];
                float pre = 1.0f;  // (((1 x_0) x_1) ... x_{j-1}): the literal loop's running product up to its skip
#pragma unroll
                for (int j = 0; j < kBigMaxDeg; ++j) {
                    if (j < d) {
                        float dt = pre;
#pragma unroll
                        for (int k = 0; k < kBigMaxDeg; ++k)
                            if (k > j && k < d) dt = __fmul_rn(dt, x[k]);
                        RD[(size_t)(e0 + j) * 32 + lane] = dt;
                        pre = __fmul_rn(pre, x[j]);
                    }
                }
            }
            __syncthreads();
            // hardDecision (decodeCL.c:64-86): all column edges; > -> 0, < -> 1, tie keeps the bit
            for (int n = warp; n < p.N; n += nwarps) {
                float t0 = P0[(size_t)n * 32 + lane], t1 = P1[(size_t)n * 32 + lane];
                for (int k = p.col_ptr[n]; k < p.col_ptr[n + 1]; ++k) {
                    const float dd = RD[(size_t)big_edge(p, p.vn_edge[k]) * 32 + lane];
                    t0 = __fmul_rn(t0, __fmul_rn(__fadd_rn(1.0f, dd), 0.5f));
                    t1 = __fmul_rn(t1, __fmul_rn(__fsub_rn(1.0f, dd), 0.5f));
                }
                const uint32_t one = __ballot_sync(0xffffffffu, t0 < t1), keep = __ballot_sync(0xffffffffu, !(t0 > t1) && !(t0 < t1));
                const uint32_t frozen = __ballot_sync(0xffffffffu, !live || done);
                if (lane == 0) {
                    const uint32_t old = HB[n];
                    HB[n] = (old & (keep | frozen)) | (one & ~frozen);
                }
            }
            __syncthreads();
            // checkResult (decodeCL.c:88-108)
            uint32_t un = 0u;
            for (int r = warp; r < p.M; r += nwarps) {
                uint32_t xw = 0u;
                for (int e = p.row_ptr[r] + lane; e < p.row_ptr[r + 1]; e += 32) xw ^= HB[p.cn_col[e]];
#pragma unroll
                for (int o = 16; o; o >>= 1) xw ^= __shfl_xor_sync(0xffffffffu, xw, o);
                un |= xw;
            }
            if (lane == 0 && un) atomicOr(&s_unsat, un);
            __syncthreads();
            const uint32_t unsat = s_unsat;
            if (live && !done) {
                ++it;
                if ((p.early_term && !((unsat >> lane) & 1u)) || it == p.max_iter) { done = true; my_iters = it; }
            }
            const bool all_done = __all_sync(0xffffffffu, done);
            __syncthreads();
            if (threadIdx.x == 0) s_unsat = 0u;
            if (all_done) break;
            // refreshQ (decodeCL.c:43-62): per edge the column products WITHOUT that edge, then normalise
            for (int n = warp; n < p.N; n += nwarps) {
                const int k0 = p.col_ptr[n], dv = p.col_ptr[n + 1] - k0;
                const float p0 = P0[(size_t)n * 32 + lane], p1 = P1[(size_t)n * 32 + lane];
                for (int k = 0; k < dv; ++k) {
                    float t0 = p0, t1 = p1;
                    for (int k2 = 0; k2 < dv; ++k2) {
                        if (k2 == k) continue;
                        const float dd = RD[(size_t)big_edge(p, p.vn_edge[k0 + k2]) * 32 + lane];
                        t0 = __fmul_rn(t0, __fmul_rn(__fadd_rn(1.0f, dd), 0.5f));
                        t1 = __fmul_rn(t1, __fmul_rn(__fsub_rn(1.0f, dd), 0.5f));
                    }
                    const float s = __fadd_rn(t0, t1);
                    QD[(size_t)big_edge(p, p.vn_edge[k0 + k]) * 32 + lane] = __fsub_rn(__fdiv_rn(t0, s), __fdiv_rn(t1, s));
                }
            }
            __syncthreads();
        }
        big_emit(p, HB, cw, valid, my_iters, lane, warp, nwarps);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Layered min-sum (DecodeTDMP): decodeInitTDMP / refreshQTDMP / refreshRTDMP / refreshPostPTDMP / hardDecisionTDMP /
// checkResult of decodeCL.c:203-292 under the loop of decodeOnceTDMP (MyLdpc.cpp:889-976), layers of z rows.
// Workspace: P[N], R[nnz], HB[N].
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) ldpc_tdmp_big_kernel(const __grid_constant__ BigParams p) {
    __shared__ uint32_t s_unsat;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* P = p.ws + (size_t)blockIdx.x * p.ws_stride;
    float* R = P + (size_t)p.N * 32;
    uint32_t* HB = reinterpret_cast<uint32_t*>(R + (size_t)p.nnz * 32);
    for (;;) {
        __shared__ int s_group;
        __syncthreads();
        if (threadIdx.x == 0) s_group = (int)atomicAdd(p.counter, 1u);
        __syncthreads();
        const int g = s_group;
        if (g >= p.ngroups) break;
        const long long cw = (long long)g * 32 + lane;
        const bool valid = cw < p.ncw;
        const float* y = p.llr + (size_t)(valid ? cw : 0) * p.N;
        for (int n = warp; n < p.N; n += nwarps) {
            P[(size_t)n * 32 + lane] = y[n];
            if (lane == 0) HB[n] = 0u;
        }
        for (int e = warp; e < p.nnz; e += nwarps) R[(size_t)e * 32 + lane] = 0.0f;
        bool live = valid, done = !valid;
        int it = 0, my_iters = 0;
        if (threadIdx.x == 0) s_unsat = 0u;
        __syncthreads();
        for (;;) {
            for (int r0 = 0; r0 < p.M; r0 += p.z) {
                for (int r = r0 + warp; r < r0 + p.z; r += nwarps) {
                    const int e0 = p.row_ptr[r], d = p.row_ptr[r + 1] - e0;
                    float q[kBigMaxDeg];
                    uint32_t col[kBigMaxDeg];
                    float m1 = INFINITY, m2 = INFINITY;
                    uint32_t par = 0u;
                    int arg = -1;
#pragma unroll
                    for (int j = 0; j < kBigMaxDeg; ++j) {
                        if (j < d) {
                            col[j] = p.cn_col[e0 + j];
                            q[j] = __fsub_rn(P[(size_t)col[j] * 32 + lane], R[(size_t)(e0 + j) * 32 + lane]);  // refreshQTDMP
                            const float a = fabsf(q[j]);
                            par ^= (uint32_t)(q[j] < 0.0f);
                            if (a < m1) { m2 = m1; m1 = a; arg = j; } else if (a < m2) { m2 = a; }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < kBigMaxDeg; ++j) {
                        if (j < d) {
                            // refreshRTDMP: sign = xor of the others' (q < 0), magnitude = fmin(1000, min of the others' |q|)
                            const float mag = fminf(kClamp, j == arg ? m2 : m1);
                            const float rn = (par ^ (uint32_t)(q[j] < 0.0f)) ? -mag : mag;
                            if (live && !done) {
                                R[(size_t)(e0 + j) * 32 + lane] = rn;
                                P[(size_t)col[j] * 32 + lane] = __fadd_rn(q[j], rn);  // refreshPostPTDMP
                            }
                        }
                    }
                }
                __syncthreads();  // layers are column-disjoint inside, not between
            }
            // hardDecisionTDMP (decodeCL.c:263-281): > 0 -> 0, < 0 -> 1, == 0 keeps
            for (int n = warp; n < p.N; n += nwarps) {
                const float t = P[(size_t)n * 32 + lane];
                const uint32_t one = __ballot_sync(0xffffffffu, t < 0.0f), keep = __ballot_sync(0xffffffffu, !(t > 0.0f) && !(t < 0.0f));
                const uint32_t frozen = __ballot_sync(0xffffffffu, !live || done);
                if (lane == 0) {
                    const uint32_t old = HB[n];
                    HB[n] = (old & (keep | frozen)) | (one & ~frozen);
                }
            }
            __syncthreads();
            uint32_t un = 0u;
            for (int r = warp; r < p.M; r += nwarps) {
                uint32_t xw = 0u;
                for (int e = p.row_ptr[r] + lane; e < p.row_ptr[r + 1]; e += 32) xw ^= HB[p.cn_col[e]];
#pragma unroll
                for (int o = 16; o; o >>= 1) xw ^= __shfl_xor_sync(0xffffffffu, xw, o);
                un |= xw;
            }
            if (lane == 0 && un) atomicOr(&s_unsat, un);
            __syncthreads();
            const uint32_t unsat = s_unsat;
            if (live && !done) {
                ++it;
                if ((p.early_term && !((unsat >> lane) & 1u)) || it == p.max_iter) { done = true; my_iters = it; }
            }
            const bool all_done = __all_sync(0xffffffffu, done);
            __syncthreads();
            if (threadIdx.x == 0) s_unsat = 0u;
            if (all_done) break;
        }
        if (p.post && valid)
            for (int n = warp; n < p.N; n += nwarps) p.post[(size_t)cw * p.N + n] = P[(size_t)n * 32 + lane];
        big_emit(p, HB, cw, valid, my_iters, lane, warp, nwarps);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// The reference's two FUSED OpenCL kernels with their own arithmetic, reproduced exactly (DecodeMSCL -> decodeOnceMS,
// decodeCL.c:432-567; DecodeTDMPCL -> decodeOnceTDMP, decodeCL.c:307-426).  They differ from decodeCPU / the layered
// schedule in corner cases only: the message sign comes from the float PRODUCT of the row's Q (an exact zero, an
// underflow or inf * 0 zeroes every message of the row), the minimum search starts from (1000, 1001) with `<=`, the hard
// decision is bit = (P < 0), and the caps are the literals 120 / 40 (passed in max_iter).  Operation for operation the
// test checker.s literal restatement of those kernels, which is pinned against the executed kernels themselves.
// Workspace: P[N], R[nnz], HB[N].
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cl_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? x : 0.0f)); }

__device__ __forceinline__ void fused_row(const BigParams& p, float* P, float* R, int r, int lane, bool layered, bool active) {
    const int e0 = p.row_ptr[r], d = p.row_ptr[r + 1] - e0;
    float a = 1.0f, b = 1000.0f, c = 1001.0f;
    int bInd = -1;
    float sg[kBigMaxDeg];
    uint32_t col[kBigMaxDeg];
#pragma unroll
    for (int j = 0; j < kBigMaxDeg; ++j) {
        if (j < d) {
            col[j] = p.cn_col[e0 + j];
            float tmp = __fsub_rn(P[(size_t)col[j] * 32 + lane], R[(size_t)(e0 + j) * 32 + lane]);
            sg[j] = cl_sign(tmp);
            a = __fmul_rn(a, tmp);
            if (layered && active) P[(size_t)col[j] * 32 + lane] = tmp;
            tmp = fabsf(tmp);
            if (tmp <= b) { c = b; b = tmp; bInd = j; }
            else if (tmp > b && tmp <= c) { c = tmp; }
        }
    }
    a = cl_sign(a);
    const float ac = __fmul_rn(a, c), ab = __fmul_rn(a, b);
#pragma unroll
    for (int j = 0; j < kBigMaxDeg; ++j) {
        if (j < d) {
            const float rn = __fmul_rn(sg[j], j == bInd ? ac : ab);
            if (active) {
                R[(size_t)(e0 + j) * 32 + lane] = rn;
                if (layered) P[(size_t)col[j] * 32 + lane] = __fadd_rn(P[(size_t)col[j] * 32 + lane], rn);
            }
        }
    }
}

__global__ void __launch_bounds__(512, 1) ldpc_fused_big_kernel(const __grid_constant__ BigParams p) {
    __shared__ uint32_t s_unsat;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool layered = p.fused_layered != 0;
    float* P = p.ws + (size_t)blockIdx.x * p.ws_stride;
    float* R = P + (size_t)p.N * 32;
    uint32_t* HB = reinterpret_cast<uint32_t*>(R + (size_t)p.nnz * 32);
    for (;;) {
        __shared__ int s_group;
        __syncthreads();
        if (threadIdx.x == 0) s_group = (int)atomicAdd(p.counter, 1u);
        __syncthreads();
        const int g = s_group;
        if (g >= p.ngroups) break;
        const long long cw = (long long)g * 32 + lane;
        const bool valid = cw < p.ncw;
        const float* y = p.llr + (size_t)(valid ? cw : 0) * p.N;
        for (int n = warp; n < p.N; n += nwarps) {
            P[(size_t)n * 32 + lane] = y[n];
            if (lane == 0) HB[n] = 0u;
        }
        for (int e = warp; e < p.nnz; e += nwarps) R[(size_t)e * 32 + lane] = 0.0f;
        bool live = valid, done = !valid;
        int it = 0, my_iters = 0;
        if (threadIdx.x == 0) s_unsat = 0u;
        __syncthreads();
        for (;;) {
            const bool active = live && !done;
            if (layered) {
                for (int r0 = 0; r0 < p.M; r0 += p.z) {
                    for (int r = r0 + warp; r < r0 + p.z; r += nwarps) fused_row(p, P, R, r, lane, true, active);
                    __syncthreads();
                }
            } else {
                for (int r = warp; r < p.M; r += nwarps) fused_row(p, P, R, r, lane, false, active);
                __syncthreads();
                for (int n = warp; n < p.N; n += nwarps) {  // decodeCL.c:517-537: lP = y + the column's lR in ascending row order
                    float t = y[n];
                    for (int k = p.col_ptr[n]; k < p.col_ptr[n + 1]; ++k) t = __fadd_rn(t, R[(size_t)big_edge(p, p.vn_edge[k]) * 32 + lane]);
                    if (active) P[(size_t)n * 32 + lane] = t;
                }
                __syncthreads();
            }
            for (int n = warp; n < p.N; n += nwarps) {  // srcBool = lP < 0 (decodeCL.c:386, 541)
                const uint32_t one = __ballot_sync(0xffffffffu, P[(size_t)n * 32 + lane] < 0.0f);
                const uint32_t frozen = __ballot_sync(0xffffffffu, !active);
                if (lane == 0) HB[n] = (HB[n] & frozen) | (one & ~frozen);
            }
            __syncthreads();
            uint32_t un = 0u;
            for (int r = warp; r < p.M; r += nwarps) {
                uint32_t xw = 0u;
                for (int e = p.row_ptr[r] + lane; e < p.row_ptr[r + 1]; e += 32) xw ^= HB[p.cn_col[e]];
#pragma unroll
                for (int o = 16; o; o >>= 1) xw ^= __shfl_xor_sync(0xffffffffu, xw, o);
                un |= xw;
            }
            if (lane == 0 && un) atomicOr(&s_unsat, un);
            __syncthreads();
            const uint32_t unsat = s_unsat;
            if (active) {
                ++it;
                if ((p.early_term && !((unsat >> lane) & 1u)) || it == p.max_iter) { done = true; my_iters = it; }
            }
            const bool all_done = __all_sync(0xffffffffu, done);
            __syncthreads();
            if (threadIdx.x == 0) s_unsat = 0u;
            if (all_done) break;
        }
        if (p.post && valid)
            for (int n = warp; n < p.N; n += nwarps) p.post[(size_t)cw * p.N + n] = P[(size_t)n * 32 + lane];
        big_emit(p, HB, cw, valid, my_iters, lane, warp, nwarps);
    }
}
#endif  // LDPC_B200_BIG_KERNELS

}  // namespace ldpc_b200
