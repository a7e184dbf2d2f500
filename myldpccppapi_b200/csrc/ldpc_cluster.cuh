// ldpc_cluster.cuh -- long codes (N = 64800): one codeword per THREAD-BLOCK CLUSTER.
//
// A codeword of the DVB-S2-sized code needs 259 KB of posteriors and 907 KB of edge messages: more
// than one SM holds, well within the 8 x 227 KB of a cluster.  CTA r of the cluster owns a slice of
// the checks (with the R rows of their edges) and a slice of the variables (with their T entries),
// both in its own shared memory.  The passes are those of the group kernel with G = 1 (lane = graph
// node); the only difference is that a gathered T entry / R message may live in a peer CTA, so the
// gathers are `ld.shared::cluster` (distributed shared memory) and the two barriers per iteration are
// cluster barriers.  All remote accesses are reads; every store is local.  Nothing but the channel
// values and the final bits touches HBM.
// Table entries are cluster-window offsets (rank << 24) + byte offset: `mapa` places CTA `rank`'s
// shared window at rank * 2^24 inside the cluster window (checked at kernel start; trap otherwise).
#pragma once
#include "ldpc_kernels.cuh"

namespace ldpc_b200 {

constexpr int kClusterSize = 8;

struct ClusterParams {
    const uint32_t* __restrict__ cn_tab;      // [CL][W][cn_stride] quads [slot][jq][lane][4]: (rank<<24) + T byte offset
    const uint32_t* __restrict__ vn_tab;      // [CL][W][vn_stride] quads [slot][kq][lane][4]: (rank<<24) + R byte offset
    const uint32_t* __restrict__ var_of_pos;  // [CL][VS*NL] variable index, 0xffffffff = phantom
    const uint32_t* __restrict__ out_addr;    // [N] (rank<<24) + T byte offset of variable n
    int M, N, K, W, CS, VS;
    int cn_stride, vn_stride, r_rows_per_warp;
    int max_iter, early_term;
    const float* __restrict__ llr;
    long long ncw;
    uint8_t* info;
    uint8_t* hard;
    int32_t* iters;
    float* post;
    unsigned long long* counter64;
    uint8_t vdeg[kGrpMaxVS];
    uint8_t cdeg[kGrpMaxCS];
};

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(local_addr), "r"(rank));
    return a;
}
__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

template <int DMAX>
__global__ void __launch_bounds__(1024, 1) ldpc_ms_cluster_kernel(const __grid_constant__ ClusterParams p) {
    constexpr int SUB = 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_flag[2];
    __shared__ uint32_t s_any;
    __shared__ unsigned long long s_cw;

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int h = lane;
    const uint32_t rank = cluster_ctarank();
    const int W = p.W, CS = p.CS, VS = p.VS;
    const int NL = W * SUB;
    const int PD = VS * NL;                // dummy T entry (= -inf), local
    const int RD = W * p.r_rows_per_warp;  // dummy R row (= 0), local

    const uint32_t t_base = smem_u32(smem_raw);
    const uint32_t t_bytes = ((uint32_t)(PD + 1) * 4 + 127u) & ~127u;
    const uint32_t r_base = t_base + t_bytes;
    // cluster-window addresses of rank 0's T and R regions; peers sit at + rank * 2^24
    const uint32_t t_cl0 = mapa_u32(t_base, 0), r_cl0 = mapa_u32(r_base, 0);
    if (mapa_u32(t_base, 1) - t_cl0 != (1u << 24)) __trap();  // layout assumption of the tables

    if (threadIdx.x == 0) {
        sts_f32(t_base + (uint32_t)PD * 4, -INFINITY);
        s_flag[0] = 0u; s_flag[1] = 0u;
    }
    if (warp == 0) sts_f32(r_base + (uint32_t)RD * 128 + lane * 4, 0.0f);
    const uint32_t t_own = t_base + (uint32_t)threadIdx.x * 4u;
    const uint32_t t_stride = (uint32_t)NL * 4u;
    const uint32_t r_own = r_base + (uint32_t)warp * p.r_rows_per_warp * 128u + (uint32_t)lane * 4u;
    const uint32_t cn_w = ((uint32_t)rank * W + warp) * p.cn_stride;   // uint32 word offsets into the global tables
    const uint32_t vn_w = ((uint32_t)rank * W + warp) * p.vn_stride;
    const uint32_t* vop = p.var_of_pos + (size_t)rank * PD;
    const uint32_t flag_cl = mapa_u32(smem_u32(&s_flag[0]), (uint32_t)(lane < kClusterSize ? lane : 0));
    const uint32_t cw_cl0 = mapa_u32(smem_u32(&s_cw), 0);
    const uint32_t gtid = rank * blockDim.x + threadIdx.x, gthreads = kClusterSize * blockDim.x;

    float yn[kGrpMaxVS];
    for (;;) {
        // ---- next codeword of this cluster (rank 0 draws it, everyone reads it through DSMEM)
        if (rank == 0 && threadIdx.x == 0) s_cw = atomicAdd(p.counter64, 1ull);
        cluster_sync_all();
        const long long cw = (long long)((unsigned long long)ld_cluster_u32(cw_cl0) | ((unsigned long long)ld_cluster_u32(cw_cl0 + 4) << 32));
        if (cw >= p.ncw) break;
        const float* src = p.llr + (size_t)cw * p.N;
#pragma unroll
        for (int s = 0; s < kGrpMaxVS; ++s) {
            yn[s] = -1.0f;
            if (s < VS) {
                const uint32_t v = __ldg(vop + s * NL + threadIdx.x);
                const float y = (v != 0xffffffffu) ? __ldg(src + v) : 1.0f;
                yn[s] = __fadd_rn(-y, 0.0f);
                sts_f32(t_own + (uint32_t)s * t_stride, yn[s]);
            }
        }
        for (int r = 0; r < p.r_rows_per_warp; ++r) sts_f32(r_own + (uint32_t)r * 128u, 0.0f);
        if (threadIdx.x == 0) { s_flag[0] = 0u; s_flag[1] = 0u; }
        cluster_sync_all();

        int it = 0, my_iters = 0;
        for (;;) {
            // ---- check-node pass over this CTA's checks (T gathered through the cluster window)
            uint32_t unsat = 0u;
            {
                uint32_t tab = cn_w;
                uint32_t rrow = r_own;
                for (int cs = 0; cs < CS; ++cs) {
                    const int dc = p.cdeg[cs];
#define CL_CASE(D) case D: unsat |= grp_check<D, SUB, false, true>(tab, p.cn_tab, t_cl0, rrow, 0u, h); break;
                    switch (dc) {
                        CL_CASE(1) CL_CASE(2) CL_CASE(3) CL_CASE(4) CL_CASE(5) CL_CASE(6) CL_CASE(7) CL_CASE(8)
                        default:
                            if constexpr (DMAX > 8) {
                                switch (dc) {
                                    CL_CASE(9) CL_CASE(10) CL_CASE(11) CL_CASE(12)
                                    CL_CASE(13) CL_CASE(14) CL_CASE(15) CL_CASE(16)
                                    default: break;
                                }
                            }
                            break;
                    }
#undef CL_CASE
                    rrow += (uint32_t)dc * 128u;
                    tab += 4u * (uint32_t)(((dc + 3) >> 2) * SUB);
                }
            }
            const bool check = p.early_term && it >= 1;
            if (check && unsat) s_flag[it & 1] = 1u;  // same-value race, benign
            cluster_sync_all();
            if (check) {
                if (warp == 0) {
                    const uint32_t f = lane < kClusterSize ? ld_cluster_u32(flag_cl + (uint32_t)(it & 1) * 4u) : 0u;
                    const bool any = __any_sync(0xffffffffu, f != 0u);
                    if (lane == 0) s_any = any ? 1u : 0u;
                }
                __syncthreads();
                if (s_any == 0u) { my_iters = it; break; }  // identical in every CTA of the cluster
            }
            if (threadIdx.x == 0) s_flag[(it + 1) & 1] = 0u;

            // ---- variable-node pass over this CTA's variables (R gathered through the cluster window)
            {
                uint32_t q = vn_w + (uint32_t)h * 4u;
                auto next_quad = [&]() {
                    uint4 o = __ldg(reinterpret_cast<const uint4*>(p.vn_tab + q));
                    q += SUB * 4;
                    o.x += r_cl0; o.y += r_cl0; o.z += r_cl0; o.w += r_cl0;
                    return o;
                };
#pragma unroll
                for (int s = 0; s < kGrpMaxVS; ++s) {
                    if (s < VS) {
                        int d = p.vdeg[s];
                        float acc = yn[s];
                        switch (d) {
                            case 1: grp_vn_part<1, true>(next_quad(), 0u, acc); break;
                            case 2: grp_vn_part<2, true>(next_quad(), 0u, acc); break;
                            case 3: grp_vn_part<3, true>(next_quad(), 0u, acc); break;
                            case 4: grp_vn_part<4, true>(next_quad(), 0u, acc); break;
                            case 8: grp_vn_part<4, true>(next_quad(), 0u, acc); grp_vn_part<4, true>(next_quad(), 0u, acc); break;
                            default:
#pragma unroll 1
                                for (; d >= 4; d -= 4) grp_vn_part<4, true>(next_quad(), 0u, acc);
                                if (d > 0) {
                                    const uint4 o = next_quad();
                                    if (d == 1) grp_vn_part<1, true>(o, 0u, acc);
                                    else if (d == 2) grp_vn_part<2, true>(o, 0u, acc);
                                    else grp_vn_part<3, true>(o, 0u, acc);
                                }
                                break;
                        }
                        sts_f32(t_own + (uint32_t)s * t_stride, acc);
                    }
                }
            }
            ++it;
            if (it == p.max_iter) { my_iters = it; break; }
            cluster_sync_all();
        }
        cluster_sync_all();  // final posteriors of every CTA are in place

        // ---- outputs: the cluster's threads share the bytes; bits are gathered through DSMEM
        if (p.info) {
            const int KB = (p.K + 7) >> 3;
            for (int b = (int)gtid; b < KB; b += (int)gthreads) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.K) v |= ((~__float_as_uint(ld_cluster_f32(t_cl0 + __ldg(p.out_addr + n)))) >> 31) << t;
                }
                p.info[(size_t)cw * KB + b] = (uint8_t)v;
            }
        }
        if (p.hard) {
            const int NB = (p.N + 7) >> 3;
            for (int b = (int)gtid; b < NB; b += (int)gthreads) {
                uint32_t v = 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = b * 8 + t;
                    if (n < p.N) v |= ((~__float_as_uint(ld_cluster_f32(t_cl0 + __ldg(p.out_addr + n)))) >> 31) << t;
                }
                p.hard[(size_t)cw * NB + b] = (uint8_t)v;
            }
        }
        if (p.post) {
            for (int n = (int)gtid; n < p.N; n += (int)gthreads)
                p.post[(size_t)cw * p.N + n] = -ld_cluster_f32(t_cl0 + __ldg(p.out_addr + n));
        }
        if (p.iters && gtid == 0) p.iters[cw] = my_iters;
        cluster_sync_all();  // nobody overwrites T while a peer still reads it
    }
    cluster_sync_all();  // no CTA exits while a peer may still read its shared memory (s_cw of rank 0)
}

}  // namespace ldpc_b200
