// k_tdmp.cu -- instantiations of the layered (TDMP) min-sum kernel (ldpc_tdmp.cuh).
#include "ldpc_launch.h"
#include "ldpc_tdmp.cuh"

namespace ldpc_b200 {
namespace {
template <int G, int MAXT>
int launch_tdmp_t(const TdmpParams& q, int grid, int threads, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ldpc_tdmp_group_kernel<G, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ldpc_tdmp_group_kernel<G, MAXT><<<grid, threads, smem, stream>>>(q);
    return (int)cudaGetLastError();
}
}  // namespace

int k_launch_tdmp(int G, const TdmpParams& q, int grid, int threads, size_t smem, cudaStream_t stream) {
    const bool big = threads > 384;
    switch (G) {
        case 4: return big ? launch_tdmp_t<4, 1024>(q, grid, threads, smem, stream) : launch_tdmp_t<4, 384>(q, grid, threads, smem, stream);
        case 8: return big ? launch_tdmp_t<8, 1024>(q, grid, threads, smem, stream) : launch_tdmp_t<8, 384>(q, grid, threads, smem, stream);
        case 16: return big ? launch_tdmp_t<16, 1024>(q, grid, threads, smem, stream) : launch_tdmp_t<16, 384>(q, grid, threads, smem, stream);
    }
    return kNoKernel;
}
}  // namespace ldpc_b200
