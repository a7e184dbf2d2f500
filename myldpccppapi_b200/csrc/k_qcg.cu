// k_qcg.cu -- instantiations of the run-time-profile quasi-cyclic kernel (ldpc_qcg.cuh).
#include "ldpc_launch.h"
#include "ldpc_qcg.cuh"

namespace ldpc_b200 {
namespace {
template <int G>
int launch_qcg_t(const QcgParams& q, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ldpc_ms_qcg_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ldpc_ms_qcg_kernel<G><<<grid, q.W * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}
}  // namespace

int k_launch_qcg(int G, const QcgParams& q, int grid, size_t smem, cudaStream_t stream) {
    switch (G) {
        case 8: return launch_qcg_t<8>(q, grid, smem, stream);
        case 4: return launch_qcg_t<4>(q, grid, smem, stream);
        case 2: return launch_qcg_t<2>(q, grid, smem, stream);
    }
    return kNoKernel;
}
}  // namespace ldpc_b200
