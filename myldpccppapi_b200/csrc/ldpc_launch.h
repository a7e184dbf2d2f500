// ldpc_launch.h -- internal seam between the host logic (ldpc_b200.cu) and the translation units that
// instantiate the sm_100a kernels (k_*.cu).  One unit per kernel family (and one per 802.16e rate for the
// compiled quasi-cyclic profiles) so that nvcc builds them in parallel; nothing here is part of the C-ABI.
//
// Every launcher returns 0 on success, a cudaError_t (> 0) when the runtime refused, or kNoKernel when no
// instantiation matches the requested shape (the caller turns that into LDPC_B200_ERR_UNSUPPORTED).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

#include "ldpc_tables.h"

namespace ldpc_b200 {

constexpr int kNoKernel = -1;
constexpr int kNoCluster = -2;  // cluster kernel: no 8-CTA cluster of this size can be resident

struct GroupParams;
struct Lane16Params;
struct StreamParams;
struct WarpParams;
struct ClusterParams;
struct QcParams;
struct QcgParams;
struct QcwParams;
struct QcmParams;
struct TdmpParams;
struct BigParams;

// Which ldpc_ms_group_kernel instantiation a plan asks for (ldpc_b200.cu: Plan).
struct GroupSel {
    int G, dmax, threads, CS, VS;
    bool tab_smem, y_smem, t16, allow_profile;
};
int k_launch_group(const GroupSel& sel, const GroupParams& q, int grid, size_t smem, cudaStream_t stream);
int k_launch_sp(int G, int threads, const GroupParams& q, int grid, size_t smem, cudaStream_t stream);
int k_launch_tdmp(int G, const TdmpParams& q, int grid, int threads, size_t smem, cudaStream_t stream);
int k_launch_qcg(int G, const QcgParams& q, int grid, size_t smem, cudaStream_t stream);
int k_launch_warp(int sw, const WarpParams& q, int grid, int threads, size_t smem, cudaStream_t stream);
int k_launch_cluster(int dmax, const ClusterParams& q, int nclusters_wanted, int threads, size_t smem, cudaStream_t stream);
int k_launch_lane16(const Lane16Params& q, int grid, int threads, size_t smem, cudaStream_t stream);
int k_launch_stream(const StreamParams& q, int grid, int threads, cudaStream_t stream);
int k_launch_sp_big(const BigParams& q, int grid, cudaStream_t stream);    // any size, messages in a global workspace
int k_launch_tdmp_big(const BigParams& q, int grid, cudaStream_t stream);
int k_launch_fused_big(const BigParams& q, int grid, cudaStream_t stream);  // the reference's fused kernels, their arithmetic
// stats[0] = sum of up to 4096 evenly spaced entries of iters[0, ncw), stats[1] = how many (the regime of the last launch)
// format 1: IEEE binary16, 2: int8; out[i] = (float)in[i] * scale
int k_launch_widen(int format, const void* in, float* out, long long n, float scale, cudaStream_t stream);
int k_launch_iter_stats(const int32_t* iters, long long ncw, unsigned long long* stats, cudaStream_t stream);

// Quasi-cyclic block structure of H for block size z: rows[br] = the circulants (block column, shift) of block
// row br in ascending column order.
struct QcBlk { int bc, s; };

// A compiled quasi-cyclic profile (ldpc_qc.cuh): rate x (z, G, W).  `upload` writes a handle's tables into slot
// `slot` of the __constant__ bank of the translation unit that holds the kernel.
struct QcProfileEntry {
    int z, G, W;
    bool (*build)(const HostTables&, const std::vector<std::vector<QcBlk>>&, QcParams*, std::vector<unsigned char>*, size_t*);
    int (*launch)(const QcParams&, int grid, size_t smem, cudaStream_t stream);
    int (*upload)(int slot, const void* tab, size_t bytes);
    // the same profile with the refill off the critical path (ldpc_ms_qc_ring_kernel): needs extra shared memory for
    // the staging ring and the hard-bit table; ring_ctas_per_sm = resident CTAs with that much (current device)
    int (*launch_ring)(const QcParams&, int grid, size_t smem, cudaStream_t stream);
    int (*ring_ctas_per_sm)(size_t smem);
};
// The warp-per-codeword kernel (ldpc_qcw.cuh): 802.16e rate x z, z = 24 or 32.
struct QcwProfileEntry {
    int z, warp_bytes;   // shared memory per codeword in flight (= per warp)
    // true if H is exactly the code compiled into the instantiation (its circulants are immediates); fills the per-lane
    // table of the syndrome rounds
    bool (*build)(const HostTables&, const std::vector<std::vector<QcBlk>>&, std::vector<uint32_t>* syn_tab);
    int (*launch)(const QcwParams&, int grid, int warps, cudaStream_t stream);
};
const QcwProfileEntry* qcw_profiles(int* n);

// The group-of-warps-per-codeword kernel (ldpc_qcm.cuh): one instantiation per 802.16e degree profile, any z <= 96.
struct QcmProfileEntry {
    bool (*build)(const HostTables&, int z, const std::vector<std::vector<QcBlk>>&, size_t smem_limit, QcmParams*,
                  std::vector<unsigned char>* tab, int* groups);
    int (*launch)(const QcmParams&, int grid, int groups, cudaStream_t stream);
    int (*upload)(int slot, const void* tab, size_t bytes);
    // several codewords per group of warps (ldpc_ms_qcm_multi_kernel): geometry for `pack` codewords, and its launch
    bool (*multi_geometry)(const QcmParams& single, int pack, size_t smem_limit, QcmParams* out, int* groups);
    int (*launch_multi)(const QcmParams&, int grid, int groups, cudaStream_t stream);
};
const QcmProfileEntry* qcm_profiles(int* n);

// quasi-cyclic sum-product kernel (k_spq.cu, ldpc_spq.cuh): same tables, geometry and profile order as qcm_profiles()
struct SpqProfileEntry {
    int (*launch)(const QcmParams&, int grid, int groups, cudaStream_t stream);
    int (*upload)(int slot, const void* tab, size_t bytes);
};
const SpqProfileEntry* spq_profiles(int* n);

// one table per 802.16e rate (k_qc.cu compiled with -DLDPC_QC_RATE=...)
const QcProfileEntry* qc_profiles_34B(int* n);
const QcProfileEntry* qc_profiles_34A(int* n);
const QcProfileEntry* qc_profiles_23B(int* n);
const QcProfileEntry* qc_profiles_23A(int* n);
const QcProfileEntry* qc_profiles_12(int* n);
const QcProfileEntry* qc_profiles_56(int* n);

}  // namespace ldpc_b200
