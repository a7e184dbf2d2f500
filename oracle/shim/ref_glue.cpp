// oracle/shim/ref_glue.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" doorway into the reference's own `Coder`, compiled UNMODIFIED from
// /root/reference/MyLdpc.cpp (see oracle/Makefile).  Used by tests/test_oracle_vs_ref.py to pin the
// oracle restatement against the reference's decodeCPU, and by bench.py --impl reference.
// `private` is opened only in this translation unit so the iteration cap (`times`, fixed at 40 in
// the reference's constructor, MyLdpc.cpp:24) can be varied: running the reference's decodeCPU
// with caps 1..40 exposes its per-iteration hard decisions and therefore its stopping iteration.
#define private public
#include "MyLdpc.h"
#undef private

#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

// The container's g++ wrapper links libstdc++ statically; a dlopen'ed library then carries its own std::cout, which
// nobody initialises unless an ios_base::Init object is constructed in it (Python has no C++ runtime of its own).
static std::ios_base::Init s_ios_init;

extern "C" {

void *ref_coder_new(int K, int N, int rate) { return new Coder(K, N, (enum rate_type)rate); }
void ref_coder_free(void *c) { delete static_cast<Coder *>(c); }
int ref_forDecoder(void *c, int batch) { return static_cast<Coder *>(c)->forDecoder(batch); }
int ref_forEncoder(void *c) { return static_cast<Coder *>(c)->forEncoder(); }
void ref_set_times(void *c, int times) { static_cast<Coder *>(c)->times = times; }
int ref_decode_cpu(void *c, float *post, char *src, int srcLength) {
    return static_cast<Coder *>(c)->decode(post, src, srcLength, DecodeCPU);
}
// The OpenCL decode variants (meaningful only in libmyldpc_refcl.so, where oracle/shim/cl_exec.h executes the
// reference's kernels; in the inert-shim builds they return without decoding).  The reference prints the number
// of iterations of every chunk it decodes ("Time=", MyLdpc.cpp:838,966,1048): std::cout is captured during the
// call and the numbers are returned in times_out (one per chunk of batchSize words; with forDecoder(1) that is
// the reference's own per-codeword iteration count).  Returns the number of "Time=" lines seen.
int ref_addDecodeType(void *c, int type) { return static_cast<Coder *>(c)->addDecodeType((enum decodeType)type); }
int ref_decode(void *c, float *post, char *src, int srcLength, int type, int *times_out, int times_cap) {
    std::ostringstream cap;
    std::streambuf *old = std::cout.rdbuf(cap.rdbuf());
    static_cast<Coder *>(c)->decode(post, src, srcLength, (enum decodeType)type);
    std::cout.rdbuf(old);
    const std::string text = cap.str();
    int n = 0;
    for (size_t pos = 0; (pos = text.find("Time=", pos)) != std::string::npos; pos += 5) {
        if (times_out && n < times_cap) times_out[n] = std::atoi(text.c_str() + pos + 5);
        ++n;
    }
    return n;
}
double ref_stepTime(void *c, int i) { return (i >= 0 && i < 10) ? static_cast<Coder *>(c)->stepTime[i] : -1.0; }
int ref_z(void *c) { return static_cast<Coder *>(c)->z; }
int ref_encode(void *c, char *src, char *prior, int srcLength) { return static_cast<Coder *>(c)->encode(src, prior, srcLength); }
int ref_getCodeSize(void *c, int srcLength) { return static_cast<Coder *>(c)->getCodeSize(srcLength); }
int ref_getPostCodeLength(void *c, int srcLength) { return static_cast<Coder *>(c)->getPostCodeLength(srcLength); }
int ref_getPriorCodeLength(void *c, int srcLength) { return static_cast<Coder *>(c)->getPriorCodeLength(srcLength); }
int ref_nnz(void *c) { return static_cast<Coder *>(c)->checkMatrix.nonZeros(); }
int ref_M(void *c) { return static_cast<Coder *>(c)->ldpcM; }
// H as the reference built it (public member checkMatrix), walked exactly like forDecoder does
void ref_csr(void *c, int *row_ptr, int *col_idx) {
    Coder *k = static_cast<Coder *>(c);
    int off = 0;
    for (int r = 0; r < k->checkMatrix.outerSize(); ++r) {
        row_ptr[r] = off;
        for (Eigen::SparseMatrix<DataType, Eigen::RowMajor>::InnerIterator it(k->checkMatrix, r); it; ++it) col_idx[off++] = it.col();
    }
    row_ptr[k->checkMatrix.outerSize()] = off;
}
// the edge tables forDecoder built (private members)
void ref_edge_tables(void *c, int *hRows, int *hCols, int *rowFirst, int *rowNext, int *colFirst, int *colNext, int *rowRange) {
    Coder *k = static_cast<Coder *>(c);
    std::memcpy(hRows, k->hRows, sizeof(int) * k->nonZeros);
    std::memcpy(hCols, k->hCols, sizeof(int) * k->nonZeros);
    std::memcpy(rowFirst, k->hRowFirstPtr, sizeof(int) * k->ldpcM);
    std::memcpy(rowNext, k->hRowNextPtr, sizeof(int) * k->nonZeros);
    std::memcpy(colFirst, k->hColFirstPtr, sizeof(int) * k->ldpcN);
    std::memcpy(colNext, k->hColNextPtr, sizeof(int) * k->nonZeros);
    std::memcpy(rowRange, k->hRowRange, sizeof(int) * (k->ldpcM + 1));
}
// Coder::test's bit -> BPSK map with sd = 0 would still call rand(); expose only the map through
// the reference's own code by passing rate (= sd) 0: gaussian(0, 0) returns exactly 0.
int ref_test(void *c, char *prior, float *post, int priorLen, float sd) { return static_cast<Coder *>(c)->test(prior, post, priorLen, sd); }
}
