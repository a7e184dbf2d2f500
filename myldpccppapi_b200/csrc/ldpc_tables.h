// ldpc_tables.h -- host-side construction of the sparse-H edge layout.
//
// Replaces the reference's seven int tables (hRows/hCols + per-row and per-column singly
// linked lists + hRowRange, MyLdpc.cpp:167-222) with two flat, coalesced index tables:
//   check-major    : row_ptr[M+1], cn_col[nnz]            (edge e = CSR position)
//   variable-major : col_ptr[N+1], vn_edge[nnz] = (check << 5) | position-in-check,
//                    listed per variable in ascending edge id (= ascending row), which is
//                    the reference's column-list order and therefore the fp32 summation
//                    order of the posterior (MyLdpc.cpp:723-728).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace ldpc_b200 {

constexpr int kPosBits = 5;          // position-in-check field of vn_edge
constexpr int kMaxCheckDegree = 27;  // 27 sign bits + 5-bit argmin index in one 32-bit word

struct HostTables {
    int M = 0, N = 0, nnz = 0;
    int max_row_weight = 0, max_col_weight = 0;
    std::vector<int32_t> row_ptr;   // [M+1]
    std::vector<int32_t> col_idx;   // [nnz] plain CSR (kept for get_csr / Coder::checkMatrix)
    std::vector<int32_t> col_ptr;   // [N+1]
    std::vector<uint32_t> vn_edge;  // [nnz]
};

// Validates the CSR and fills every table.  Returns empty string on success, else a message.
std::string build_tables(int M, int N, const int32_t* row_ptr, const int32_t* col_idx, HostTables* out);

// Coder::initCheckMatrix semantics (reference MyLdpc.cpp:52-109) without the O(z^2) scan:
// block (sr, sc) with seed p >= 0 contributes H[sr*z + r, sc*z + (r + shift) % z] = 1 where
// shift = p*z/96 (integer floor), or p % z for rate_2_3_a.  Emits CSR with ascending columns.
std::string wimax_csr(int K, int N, int rate, std::vector<int32_t>* row_ptr, std::vector<int32_t>* col_idx, int* M);

// Systematic encoder matrix: with H = [A | B] (B = the last M columns, invertible), the parity bits are
// p = X u over GF(2), X = B^-1 A (the unique solution of H c = 0; what Coder::forEncoder / Coder::encode compute
// through the Richardson-Urbanke split, reference MyLdpc.cpp:137-165, 633-682).  Returns X transposed and packed
// for the device: xt[w * (MW*32) + r] holds info bits 32w .. 32w+31 of row r; KW = ceil(K/32), MW = ceil(M/32).
std::string build_encoder(const HostTables& t, int K, std::vector<uint32_t>* xt, int* KW, int* MW);

// Copy of a pageable caller buffer into a page-locked staging slot with NON-TEMPORAL stores (the host-buffer pipeline of
// ldpc_b200_decode_host): a plain memcpy reads the destination lines for ownership before overwriting them and leaves
// them in the cache, from where the DMA engine's read forces a write-back -- four memory transfers per byte where
// three are needed (source read, streaming write, DMA read).  Ends with a store fence: the caller may queue the DMA.
void stage_copy_nt(void* dst, const void* src, size_t bytes);

}  // namespace ldpc_b200
