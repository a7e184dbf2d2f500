// k_misc.cu -- instantiations of the alternative / fallback kernels: sub-warp per check (ldpc_warp.cuh), one codeword
// per 8-CTA cluster (ldpc_cluster.cuh), lane = codeword with compressed check state (ldpc_kernels.cuh: lane16) and the
// long-code kernel with its messages in a global workspace (ldpc_stream.cuh); sum-product and layered min-sum for codes
// of any size (ldpc_big.cuh).
#include <algorithm>

#include <cuda_fp16.h>

#include "ldpc_launch.h"
#include "ldpc_kernels.cuh"
#include "ldpc_cluster.cuh"
#include "ldpc_stream.cuh"
#include "ldpc_warp.cuh"
#define LDPC_B200_BIG_KERNELS
#include "ldpc_big.cuh"

namespace ldpc_b200 {
namespace {
// Mean iteration count of a launch, from a sample of its per-word counts: one CTA, one pass.
__global__ void __launch_bounds__(256) ldpc_iter_stats_kernel(const int32_t* __restrict__ iters, long long ncw, unsigned long long* stats) {
    const long long nsamp = ncw < 4096 ? ncw : 4096;
    unsigned long long s = 0;
    for (long long i = threadIdx.x; i < nsamp; i += 256) s += (unsigned long long)iters[i * ncw / nsamp];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) s += part[w];
        stats[0] = s;
        stats[1] = (unsigned long long)nsamp;
    }
}

// Packed channel values -> fp32 (ldpc_b200_decode_host_packed): y = (float)x * scale, one rounding.
template <class T>
__global__ void __launch_bounds__(256) ldpc_widen_kernel(const T* __restrict__ in, float* __restrict__ out, long long n, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v;
        if constexpr (sizeof(T) == 2) v = __half2float(in[i]);
        else v = (float)in[i];
        out[i] = __fmul_rn(v, scale);
    }
}

template <int SW>
int launch_warp_t(const WarpParams& q, int grid, int threads, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ldpc_ms_warp_kernel<SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ldpc_ms_warp_kernel<SW><<<grid, threads, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

template <int DMAX>
int launch_cluster_t(const ClusterParams& q, int nclusters_wanted, int threads, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ldpc_ms_cluster_kernel<DMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kClusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(kClusterSize);
    int maxc = 0;
    e = cudaOccupancyMaxActiveClusters(&maxc, ldpc_ms_cluster_kernel<DMAX>, &cfg);
    if (e != cudaSuccess) return (int)e;
    if (maxc < 1) return kNoCluster;
    cfg.gridDim = dim3(kClusterSize * std::min(maxc, nclusters_wanted));
    return (int)cudaLaunchKernelEx(&cfg, ldpc_ms_cluster_kernel<DMAX>, q);
}

template <int MAXT>
int launch_lane16_t(const Lane16Params& q, int grid, int threads, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ldpc_ms_lane16_kernel<MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ldpc_ms_lane16_kernel<MAXT><<<grid, threads, smem, stream>>>(q);
    return (int)cudaGetLastError();
}
}  // namespace

int k_launch_warp(int sw, const WarpParams& q, int grid, int threads, size_t smem, cudaStream_t stream) {
    switch (sw) {
        case 8: return launch_warp_t<8>(q, grid, threads, smem, stream);
        case 16: return launch_warp_t<16>(q, grid, threads, smem, stream);
        case 32: return launch_warp_t<32>(q, grid, threads, smem, stream);
    }
    return kNoKernel;
}

int k_launch_cluster(int dmax, const ClusterParams& q, int nclusters_wanted, int threads, size_t smem, cudaStream_t stream) {
    return dmax == 8 ? launch_cluster_t<8>(q, nclusters_wanted, threads, smem, stream)
                     : launch_cluster_t<16>(q, nclusters_wanted, threads, smem, stream);
}

int k_launch_lane16(const Lane16Params& q, int grid, int threads, size_t smem, cudaStream_t stream) {
    return threads <= 768 ? launch_lane16_t<768>(q, grid, threads, smem, stream) : launch_lane16_t<1024>(q, grid, threads, smem, stream);
}

int k_launch_sp_big(const BigParams& q, int grid, cudaStream_t stream) {
    ldpc_sp_big_kernel<<<grid, 512, 0, stream>>>(q);
    return (int)cudaGetLastError();
}

int k_launch_tdmp_big(const BigParams& q, int grid, cudaStream_t stream) {
    ldpc_tdmp_big_kernel<<<grid, 512, 0, stream>>>(q);
    return (int)cudaGetLastError();
}

int k_launch_fused_big(const BigParams& q, int grid, cudaStream_t stream) {
    ldpc_fused_big_kernel<<<grid, 512, 0, stream>>>(q);
    return (int)cudaGetLastError();
}

int k_launch_iter_stats(const int32_t* iters, long long ncw, unsigned long long* stats, cudaStream_t stream) {
    ldpc_iter_stats_kernel<<<1, 256, 0, stream>>>(iters, ncw, stats);
    return (int)cudaGetLastError();
}

int k_launch_widen(int format, const void* in, float* out, long long n, float scale, cudaStream_t stream) {
    const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 8);
    if (format == 1) ldpc_widen_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(in), out, n, scale);
    else if (format == 2) ldpc_widen_kernel<signed char><<<grid, 256, 0, stream>>>(static_cast<const signed char*>(in), out, n, scale);
    else return kNoKernel;
    return (int)cudaGetLastError();
}

int k_launch_stream(const StreamParams& q, int grid, int threads, cudaStream_t stream) {
    if (threads <= 512) ldpc_ms_stream_kernel<512><<<grid, threads, 0, stream>>>(q);
    else if (threads <= 768) ldpc_ms_stream_kernel<768><<<grid, threads, 0, stream>>>(q);
    else ldpc_ms_stream_kernel<1024><<<grid, threads, 0, stream>>>(q);
    return (int)cudaGetLastError();
}
}  // namespace ldpc_b200
