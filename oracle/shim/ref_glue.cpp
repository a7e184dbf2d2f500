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

extern "C" {

void *ref_coder_new(int K, int N, int rate) { return new Coder(K, N, (enum rate_type)rate); }
void ref_coder_free(void *c) { delete static_cast<Coder *>(c); }
int ref_forDecoder(void *c, int batch) { return static_cast<Coder *>(c)->forDecoder(batch); }
int ref_forEncoder(void *c) { return static_cast<Coder *>(c)->forEncoder(); }
void ref_set_times(void *c, int times) { static_cast<Coder *>(c)->times = times; }
int ref_decode_cpu(void *c, float *post, char *src, int srcLength) {
    return static_cast<Coder *>(c)->decode(post, src, srcLength, DecodeCPU);
}
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
