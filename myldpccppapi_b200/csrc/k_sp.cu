// k_sp.cu -- instantiations of the sum-product kernel (ldpc_sp.cuh).
#include "ldpc_launch.h"
#include "ldpc_sp.cuh"

namespace ldpc_b200 {
namespace {
template <int G, int DMAX, int MAXT>
int launch_sp_t(const GroupParams& q, int grid, int threads, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ldpc_sp_group_kernel<G, DMAX, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ldpc_sp_group_kernel<G, DMAX, MAXT><<<grid, threads, smem, stream>>>(q);
    return (int)cudaGetLastError();
}
}  // namespace

int k_launch_sp(int G, int threads, const GroupParams& q, int grid, size_t smem, cudaStream_t stream) {
    if (G == 8 && threads <= 384) return launch_sp_t<8, 20, 384>(q, grid, threads, smem, stream);
    if (G == 16 && threads <= 768) return launch_sp_t<16, 20, 768>(q, grid, threads, smem, stream);
    if (G == 16) return launch_sp_t<16, 20, 1024>(q, grid, threads, smem, stream);
    return kNoKernel;
}
}  // namespace ldpc_b200
