// ldpc_encode.cuh -- systematic encoder on the device (sm_100a).
//
// Coder::encode (reference MyLdpc.cpp:554-569, 633-682) computes the parity bits of H c = 0 with a
// Richardson-Urbanke split and dense integer "inverses" through Eigen.  The parity part B of H = [A | B] is
// invertible, so the parity bits are fixed by the info bits: p = X u over GF(2), X = B^-1 A (M x K bits,
// solved once on the host by Gauss-Jordan, ldpc_tables.cpp: build_encoder).  Here a warp encodes one codeword:
// lanes hold the info bits as 32-bit words, and for every 32 parity bits lane l accumulates
// xor_w (XT[w][32 pw + l] & u_w) -- XT is X transposed into [info word][parity bit] so the loads coalesce --
// takes the parity of the popcount, and a ballot packs the 32 bits.  Codeword layout as the reference's
// priorCode: N/8 bytes per word, bit i of the word at byte i/8, bit i%8 (LSB first), info bits first.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldpc_b200 {

constexpr int kEncMaxKW = 4;  // info words per lane: K <= 32 * 32 * 4 = 4096 bits

struct EncodeParams {
    const uint32_t* __restrict__ xt;  // [KW][MW * 32]
    int K, N, M, KW, MW;
    const uint8_t* __restrict__ info;  // [ncw][K/8]
    uint8_t* __restrict__ out;         // [ncw][N/8]
    long long ncw;
};

static __global__ void __launch_bounds__(256) ldpc_encode_kernel(const EncodeParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int KB = p.K >> 3, NB = p.N >> 3, MP = p.MW * 32;
    for (long long cw = warp0; cw < p.ncw; cw += nwarps) {
        const uint8_t* src = p.info + (size_t)cw * KB;
        uint8_t* dst = p.out + (size_t)cw * NB;
        // info bytes -> systematic part, and into 32-bit words spread over the lanes (word w on lane w % 32)
        uint32_t u[kEncMaxKW];
#pragma unroll
        for (int i = 0; i < kEncMaxKW; ++i) {
            u[i] = 0u;
            const int w = i * 32 + lane;
            if (w < p.KW) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int byte = w * 4 + b;
                    if (byte < KB) {
                        const uint32_t v = src[byte];
                        dst[byte] = (uint8_t)v;
                        u[i] |= v << (8 * b);
                    }
                }
            }
        }
        for (int pw = 0; pw < p.MW; ++pw) {
            uint32_t acc = 0u;
            const uint32_t* col = p.xt + pw * 32 + lane;
#pragma unroll
            for (int i = 0; i < kEncMaxKW; ++i) {
                const int wn = min(32, p.KW - i * 32);
                for (int w = 0; w < wn; ++w) acc ^= __ldg(col + (size_t)(i * 32 + w) * MP) & __shfl_sync(0xffffffffu, u[i], w);
            }
            const uint32_t word = __ballot_sync(0xffffffffu, __popc(acc) & 1);
            // parity bits 32 pw .. 32 pw + 31 start at codeword bit K + 32 pw (K is a multiple of 8)
            const int byte = KB + pw * 4 + lane;
            if (lane < 4 && byte < NB) dst[byte] = (uint8_t)(word >> (8 * lane));
        }
    }
}

}  // namespace ldpc_b200
