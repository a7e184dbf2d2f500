// k_qc.cu -- the compiled quasi-cyclic profiles of ONE 802.16e rate (ldpc_qc.cuh), built once per rate:
//   nvcc -DLDPC_QC_RATE=QcProfile34B -DLDPC_QC_RATE_FN=qc_profiles_34B ...
// Each unit owns its copy of the __constant__ table bank; `upload` is how the host logic reaches it.
#include "ldpc_qc_host.h"

#ifndef LDPC_QC_RATE
#error "compile with -DLDPC_QC_RATE=<profile template> -DLDPC_QC_RATE_FN=<table function>"
#endif

namespace ldpc_b200 {
namespace {

template <class P>
int launch_qc_t(const QcParams& q, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ldpc_ms_qc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ldpc_ms_qc_kernel<P><<<grid, P::W * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

template <class P>
int launch_qc_ring_t(const QcParams& q, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(ldpc_ms_qc_ring_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ldpc_ms_qc_ring_kernel<P><<<grid, P::W * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

// resident CTAs per SM of the ring kernel with `smem` bytes of dynamic shared memory (0 if it cannot launch)
template <class P>
int ring_ctas_per_sm(size_t smem) {
    if (cudaFuncSetAttribute(ldpc_ms_qc_ring_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ldpc_ms_qc_ring_kernel<P>, P::W * 32, smem) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int upload_bank(int slot, const void* tab, size_t bytes) {
    if (slot < 0 || slot >= kQcTabSlots || bytes > (size_t)kQcBankBytes) return (int)cudaErrorInvalidValue;
    return (int)cudaMemcpyToSymbol(g_qc_bank, tab, bytes, (size_t)slot * kQcBankBytes, cudaMemcpyHostToDevice);
}

#define QC_PROFILE(T, Z, G, W) {Z, G, W, &qc_build<T<Z, G, W>>, &launch_qc_t<T<Z, G, W>>, &upload_bank, nullptr, nullptr}
// with the ring-staged variant of the kernel (opt-in experiment, measured slower: profiles/r02_ring_kernel.md) -- only
// where its test and measurements live, z = 24
#define QC_PROFILE_RING(T, Z, G, W) {Z, G, W, &qc_build<T<Z, G, W>>, &launch_qc_t<T<Z, G, W>>, &upload_bank, &launch_qc_ring_t<T<Z, G, W>>, &ring_ctas_per_sm<T<Z, G, W>>}
const QcProfileEntry kTable[] = {
    QC_PROFILE_RING(LDPC_QC_RATE, 24, 8, 12), QC_PROFILE(LDPC_QC_RATE, 48, 4, 12), QC_PROFILE(LDPC_QC_RATE, 96, 2, 12),
    QC_PROFILE(LDPC_QC_RATE, 40, 4, 10), QC_PROFILE(LDPC_QC_RATE, 80, 2, 10),
    QC_PROFILE(LDPC_QC_RATE, 32, 4, 8),  QC_PROFILE(LDPC_QC_RATE, 64, 2, 8),
    QC_PROFILE_RING(LDPC_QC_RATE, 24, 4, 6),  // experiment (option qc_prefer_g = 4): half-size CTAs, three per SM
};
#undef QC_PROFILE_RING
#undef QC_PROFILE

}  // namespace

const QcProfileEntry* LDPC_QC_RATE_FN(int* n) {
    *n = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace ldpc_b200
