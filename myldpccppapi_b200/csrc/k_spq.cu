// k_spq.cu -- the quasi-cyclic sum-product kernel (ldpc_spq.cuh): one instantiation per 802.16e degree profile (any z,
// any shifts) and this unit's __constant__ table bank.  Tables and geometry are those of the group-of-warps min-sum
// kernel (k_qcm.cu: qcm_build); the entries of spq_profiles() are in the order of qcm_profiles().
#include "ldpc_launch.h"
#define LDPC_SPQ_DEVICE
#include "ldpc_spq.cuh"

namespace ldpc_b200 {
namespace {

template <class R>
int launch_spq_t(const QcmParams& q, int grid, int groups, cudaStream_t stream) {
    const size_t smem = (size_t)groups * q.word_bytes;
    auto kernel = ldpc_sp_qcm_kernel<R>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kernel<<<grid, groups * q.NW * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

int upload_spq_bank(int slot, const void* tab, size_t bytes) {
    if (slot < 0 || slot >= kQcTabSlots || bytes > (size_t)kQcmBankBytes) return (int)cudaErrorInvalidValue;
    return (int)cudaMemcpyToSymbol(g_spq_bank, tab, bytes, (size_t)slot * kQcmBankBytes, cudaMemcpyHostToDevice);
}

#define SPQ_PROFILE(C) {&launch_spq_t<C>, &upload_spq_bank}
const SpqProfileEntry kTable[] = {
    SPQ_PROFILE(QcwCode34B_24), SPQ_PROFILE(QcwCode34A_24), SPQ_PROFILE(QcwCode23B_24),
    SPQ_PROFILE(QcwCode23A_24), SPQ_PROFILE(QcwCode12_24),  SPQ_PROFILE(QcwCode56_24),
};
#undef SPQ_PROFILE

}  // namespace

const SpqProfileEntry* spq_profiles(int* n) {
    *n = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace ldpc_b200
