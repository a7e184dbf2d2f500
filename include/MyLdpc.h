// MyLdpc.h -- drop-in replacement for the reference's public header (wing02/MyLdpcCppApi MyLdpc.h).
//
// Same `class Coder` public interface, enums, macros and seed-table names, so code written against
// the reference (its own Test.cpp included) compiles unchanged; the OpenCL / cl.hpp machinery and the
// Eigen dependency are gone.  Decoding runs on the sm_100a CUDA decoder in libldpc_b200.so
// (include/ldpc_b200.h): DecodeSP selects the probability-domain sum-product kernel (the semantics of the
// reference's decodeOnceSP, MyLdpc.cpp:977-1059 / decodeCL.c:3-108; codes too large for its on-chip layout
// fall back to min-sum, reported through lastError()); DecodeTDMP and DecodeTDMPCL select the layered
// min-sum kernel (the schedule decodeOnceTDMP intends, MyLdpc.cpp:889-976 / decodeCL.c:203-292, with the
// same fallback); DecodeCPU, DecodeMS and DecodeMSCL run flooding min-sum with the semantics of the
// reference's Coder::decodeCPU (MyLdpc.cpp:684-784) -- DecodeMSCL with the iteration cap 120 that the reference's
// fused kernel hard-codes (decodeCL.c:479) unless setMaxIter() was called.  There is no CPU decode path in
// this library, DecodeCPU included.
//
//   reference                         here
//   --------------------------------  ------------------------------------------------------------
//   Coder(K, N, rate)                 same; builds H exactly like initCheckMatrix (MyLdpc.cpp:52-109)
//   forDecoder(batchSize)             creates the CUDA decoder(s) (ldpc_b200_create)      (:167-305)
//   addDecodeType(t)                  reserves device staging for batchSize words         (:307-552)
//   decode(post, src, len, t)         ldpc_b200_decode_host, sharded over the devices     (:571-618)
//   forEncoder / encode               systematic GF(2) encoder (same codewords: H fixes the parity)
//   test(prior, post, len, sd)        BPSK + Box-Muller AWGN with rand(), as the reference (:1061-1105)
//   checkMatrix                       CSR view with the Eigen accessors the reference's users touch
//
// Return values: 0 (LDPC_SUCCESS) on success like the reference; a negative ldpc_b200 error code on
// failure (the reference ignores its OpenCL errors; see lastError()).
// Additive members are marked [B200].
#ifndef MYLDPC_H_
#define MYLDPC_H_

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <vector>

#define LDPC_SUCCESS 0
#define LDPC_FAIL 0
#define MAX 200000
#define DataType int
#define LOG std::cout << __FILE__ << " " << __LINE__ << std::endl
#define LOGA(a) std::cout << __FILE__ << " " << __LINE__ << " " << a << std::endl
#define ASSERT(a) if (!(a)) { LOG; std::cout << "Assert! error"; exit(0); }
#define GETTIME LOGA((double)clock() / CLOCKS_PER_SEC)

enum rate_type { rate_1_2, rate_2_3_a, rate_2_3_b, rate_3_4_a, rate_3_4_b, rate_5_6 };

enum decodeType { DecodeCPU, DecodeMS, DecodeSP, DecodeTDMP, DecodeTDMPCL, DecodeMSCL };

#include "MyLdpcSeeds.h"

// Row-major sparse 0/1 matrix with the slice of Eigen::SparseMatrix<int, RowMajor>'s interface that
// the reference's `checkMatrix` member offers to callers.
class CheckMatrixCSR {
public:
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    int nonZeros() const { return (int)col_.size(); }
    int outerSize() const { return rows_; }
    int coeff(int r, int c) const {
        for (int e = ptr_[r]; e < ptr_[r + 1]; ++e)
            if (col_[e] == c) return 1;
        return 0;
    }
    const int *outerIndexPtr() const { return ptr_.data(); }
    const int *innerIndexPtr() const { return col_.data(); }
    class InnerIterator {
    public:
        InnerIterator(const CheckMatrixCSR &m, int outer) : m_(&m), r_(outer), e_(m.ptr_[outer]) {}
        operator bool() const { return e_ < m_->ptr_[r_ + 1]; }
        InnerIterator &operator++() { ++e_; return *this; }
        int row() const { return r_; }
        int col() const { return m_->col_[e_]; }
        int index() const { return m_->col_[e_]; }
        int value() const { return 1; }
    private:
        const CheckMatrixCSR *m_;
        int r_, e_;
    };
private:
    friend class Coder;
    int rows_ = 0, cols_ = 0;
    std::vector<int> ptr_, col_;
};

class Coder {
public:
    Coder(int ldpcK, int ldpcN, enum rate_type rate);
    ~Coder();
    int forEncoder();
    int forDecoder(int batchSize);
    int addDecodeType(enum decodeType deType);
    int forTest();

    int encode(char *srcCode, char *priorCode, int srcLength);
    // code length in decode is 8 times of code length in encode (one float per code bit)
    int decode(float *postCode, char *srcCode, int srcLength, enum decodeType deType);

    int test(char *priorCode, float *postCode, int priorCodeLength, float rate = 0.2f);

    int getPriorCodeLength(int srcLength);
    int getPostCodeLength(int srcLength);
    int getCodeSize(int srcLength);

    CheckMatrixCSR checkMatrix;
    const char *kernelSourceCode;

    // ---- [B200] additive interface ------------------------------------------------------------
    // Arbitrary parity-check matrix in CSR (the reference can only build 802.16e codes, N = 24 z).
    Coder(int ldpcM, int ldpcN, int ldpcK, const int *rowPtr, const int *colIdx);
    int setMaxIter(int times);                       // the reference fixes times = 40 (MyLdpc.cpp:24)
    int setDevices(const int *deviceIds, int count); // shard codewords over these GPUs (before forDecoder)
    int setEarlyTermination(bool on);
    // DecodeSP / DecodeTDMP / DecodeTDMPCL need their kernel's layout to hold the code.  Default (false): a code that
    // does not fit is decoded with flooding min-sum instead, decode() returns 0 and lastAlgorithm() tells; strict
    // (true): decode() returns LDPC_B200_ERR_UNSUPPORTED (-3) and writes nothing.
    int setStrictDecodeType(bool strict);
    // DecodeMSCL / DecodeTDMPCL: by default the fast flooding / layered kernels (the reference kernels' caps, 120 / 40,
    // and schedules; results differ from the reference's fused OpenCL kernels only where a message or posterior is
    // EXACTLY zero).  true: reproduce those kernels' arithmetic exactly (sign through the product of the row's Q,
    // bit = (P < 0); decodeCL.c:307-567) with the any-size kernel -- byte-identical to the reference, slower.
    int setFusedKernelArithmetic(bool exact);
    // true: page-lock (cudaHostRegister) a malloc'd postCode buffer the first time decode() sees it and keep it so until
    // the Coder is destroyed -- the caller promises not to free it before.  Repeated decodes out of one buffer (Test.cpp's
    // loop) then cost what pinned memory costs; default false (pageable input is staged through pinned buffers).
    int setRegisterHostBuffers(bool on);
    int lastAlgorithm() const;                       // LDPC_B200_ALG_* the last decode() actually ran (-1: none yet)
    const int *lastIterations() const;               // per-codeword iteration counts of the last decode()
    int lastCodeSize() const;
    const char *lastError() const;
    // Time of the last decode() by phase, in seconds (the reference keeps clock() sums per kernel in stepTime[],
    // MyLdpc.cpp:26-28, 987-1056; the phases of a fused decoder are different): [0] whole call, wall clock;
    // [1] host->device copies, [2] kernels, [3] device->host copies (CUDA events on the busiest device; the three
    // overlap, so they do not add up to [0]).  Returns the number of entries written (<= n).
    int lastStepTimes(double *seconds, int n) const;
    Coder(const Coder &) = delete;
    Coder &operator=(const Coder &) = delete;

private:
    struct Impl;
    Impl *impl;
};

char *load_program_source(const char *filename);
float gaussian(float ave, float sd);

#endif /* MYLDPC_H_ */
