// ldpc_tables.cpp -- see ldpc_tables.h
#include "ldpc_tables.h"

#include <algorithm>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "wimax_tables.h"

namespace ldpc_b200 {

std::string build_tables(int M, int N, const int32_t* row_ptr, const int32_t* col_idx, HostTables* out) {
    if (M <= 0 || N <= 0) return "M and N must be positive";
    if (!row_ptr || !col_idx) return "null CSR pointer";
    if (row_ptr[0] != 0) return "row_ptr[0] must be 0";
    for (int r = 0; r < M; ++r)
        if (row_ptr[r + 1] < row_ptr[r]) return "row_ptr must be non-decreasing";
    const int nnz = row_ptr[M];
    if ((uint64_t)M >= (1ull << (32 - kPosBits))) return "too many checks for the packed edge table";

    HostTables& t = *out;
    t.M = M; t.N = N; t.nnz = nnz;
    t.row_ptr.assign(row_ptr, row_ptr + M + 1);
    t.col_idx.assign(col_idx, col_idx + nnz);
    t.col_ptr.assign(N + 1, 0);
    t.max_row_weight = 0;
    std::vector<int32_t> seen(N, -1);
    for (int r = 0; r < M; ++r) {
        t.max_row_weight = std::max(t.max_row_weight, row_ptr[r + 1] - row_ptr[r]);
        for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
            int c = col_idx[e];
            if (c < 0 || c >= N) return "column index out of range";
            if (seen[c] == r) return "duplicate (row, column) entry";
            seen[c] = r;
            t.col_ptr[c + 1]++;
        }
    }
    t.max_col_weight = 0;
    for (int c = 0; c < N; ++c) {
        t.max_col_weight = std::max(t.max_col_weight, t.col_ptr[c + 1]);
        t.col_ptr[c + 1] += t.col_ptr[c];
    }
    // Counting sort by column, stable in edge id => each variable's list is in ascending
    // edge id, the order of the reference's hColFirstPtr/hColNextPtr walk.
    t.vn_edge.assign(nnz, 0);
    std::vector<int32_t> fill(t.col_ptr.begin(), t.col_ptr.end() - 1);
    for (int r = 0; r < M; ++r) {
        for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
            int pos = e - row_ptr[r];
            uint32_t packed = ((uint32_t)r << kPosBits) | (uint32_t)(pos & ((1 << kPosBits) - 1));
            t.vn_edge[fill[col_idx[e]]++] = packed;
        }
    }
    return std::string();
}

std::string wimax_csr(int K, int N, int rate, std::vector<int32_t>* row_ptr, std::vector<int32_t>* col_idx, int* M_out) {
    const int8_t* base = nullptr;
    int brows = 0;
    switch (rate) {
        case 0: base = &kBase_1_2[0][0];   brows = 12; break;
        case 1: base = &kBase_2_3_A[0][0]; brows = 8;  break;
        case 2: base = &kBase_2_3_B[0][0]; brows = 8;  break;
        case 3: base = &kBase_3_4_A[0][0]; brows = 6;  break;
        case 4: base = &kBase_3_4_B[0][0]; brows = 6;  break;
        case 5: base = &kBase_5_6[0][0];   brows = 4;  break;
        default: return "rate must be 0..5 (rate_1_2 .. rate_5_6)";
    }
    if (N <= 0 || N % kBaseCols != 0) return "N must be a positive multiple of 24";
    const int z = N / kBaseCols;
    const int M = brows * z;
    if (K != N - M) return "K does not match N and the rate (K must equal N - rows*z)";
    row_ptr->assign(1, 0);
    col_idx->clear();
    std::vector<int32_t> cols;
    for (int br = 0; br < brows; ++br) {
        for (int r = 0; r < z; ++r) {
            cols.clear();
            for (int bc = 0; bc < kBaseCols; ++bc) {
                int p = base[br * kBaseCols + bc];
                if (p < 0) continue;
                int shift = (rate != 1) ? (p * z / 96) : (p % z);
                cols.push_back(bc * z + (r + shift) % z);
            }
            // one entry per block column and blocks are disjoint column ranges => already ascending
            col_idx->insert(col_idx->end(), cols.begin(), cols.end());
            row_ptr->push_back((int32_t)col_idx->size());
        }
    }
    *M_out = M;
    return std::string();
}

std::string build_encoder(const HostTables& t, int K, std::vector<uint32_t>* xt, int* KW, int* MW) {
    const int M = t.M, N = t.N;
    if (N - K != M || K < 1) return "encoder needs a square parity part (N - K == M)";
    const int W = (N + 63) / 64;
    std::vector<uint64_t> a((size_t)M * W, 0);  // row r: columns permuted to [parity (M) | info (K)]
    auto getbit = [&](int r, int c) { return (a[(size_t)r * W + (c >> 6)] >> (c & 63)) & 1ull; };
    for (int r = 0; r < M; ++r)
        for (int e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) {
            const int c = t.col_idx[e], cc = c >= K ? c - K : M + c;
            a[(size_t)r * W + (cc >> 6)] ^= 1ull << (cc & 63);
        }
    for (int c = 0; c < M; ++c) {  // Gauss-Jordan over GF(2) on the parity block
        int piv = -1;
        for (int r = c; r < M; ++r)
            if (getbit(r, c)) { piv = r; break; }
        if (piv < 0) return "parity part of H is singular";
        if (piv != c)
            for (int w = 0; w < W; ++w) std::swap(a[(size_t)piv * W + w], a[(size_t)c * W + w]);
        const uint64_t* src = &a[(size_t)c * W];
        for (int r = 0; r < M; ++r)
            if (r != c && getbit(r, c)) {
                uint64_t* dst = &a[(size_t)r * W];
                for (int w = c >> 6; w < W; ++w) dst[w] ^= src[w];
            }
    }
    *KW = (K + 31) / 32;
    *MW = (M + 31) / 32;
    const size_t MP = (size_t)*MW * 32;
    xt->assign((size_t)*KW * MP, 0u);
    for (int r = 0; r < M; ++r)
        for (int k = 0; k < K; ++k)
            if (getbit(r, M + k)) (*xt)[(size_t)(k >> 5) * MP + r] |= 1u << (k & 31);
    return "";
}

// ---- staging copy with streaming stores (see ldpc_tables.h) ------------------------------------------------------------
#if defined(__x86_64__)
namespace {
__attribute__((target("avx2"))) void stage_copy_avx2(char* d, const char* s, size_t n) {  // d 32-byte aligned, n % 128 == 0
    for (size_t i = 0; i < n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), e);
    }
}
void stage_copy_sse2(char* d, const char* s, size_t n) {  // d 16-byte aligned, n % 64 == 0
    for (size_t i = 0; i < n; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
    }
}
}  // namespace

void stage_copy_nt(void* dst, const void* src, size_t bytes) {
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    const size_t head = std::min(bytes, (size_t)((64u - (reinterpret_cast<uintptr_t>(d) & 63u)) & 63u));
    if (head) { std::memcpy(d, s, head); d += head; s += head; bytes -= head; }
    static const bool avx2 = __builtin_cpu_supports("avx2");
    const size_t body = bytes & ~(size_t)127;
    if (body) {
        if (avx2) stage_copy_avx2(d, s, body);
        else stage_copy_sse2(d, s, body);
    }
    if (bytes > body) std::memcpy(d + body, s + body, bytes - body);
    _mm_sfence();
}
#else
void stage_copy_nt(void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); }
#endif

}  // namespace ldpc_b200
