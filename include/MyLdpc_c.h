/* MyLdpc_c.h -- C doorway to the drop-in `class Coder` of include/MyLdpc.h (libmyldpc_b200.so).
 *
 * One function per public member of the reference's Coder (wing02/MyLdpcCppApi MyLdpc.h:104-129), same names,
 * argument meaning and return values (0 = LDPC_SUCCESS; a negative ldpc_b200 code on failure), plus the additive
 * [B200] members.  For callers that bind through a C FFI (ctypes, cgo, JNI ...): the Python `Coder` in
 * myldpccppapi_b200/decoder.py is a thin binding of exactly these symbols, so the C++ class is the only
 * implementation of the Coder logic.  rate / deType are the reference's enum values (MyLdpc.h:33-39).
 */
#ifndef MYLDPC_C_H_
#define MYLDPC_C_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct myldpc_coder myldpc_coder;

myldpc_coder *myldpc_coder_new(int ldpcK, int ldpcN, int rate);                /* Coder::Coder, MyLdpc.cpp:20-29 */
myldpc_coder *myldpc_coder_new_csr(int ldpcM, int ldpcN, int ldpcK, const int *rowPtr, const int *colIdx); /* [B200] */
void myldpc_coder_free(myldpc_coder *c);                                       /* Coder::~Coder, :31-51          */
int myldpc_forEncoder(myldpc_coder *c);                                        /* :137-165                        */
int myldpc_forDecoder(myldpc_coder *c, int batchSize);                         /* :167-305                        */
int myldpc_addDecodeType(myldpc_coder *c, int deType);                         /* :307-552                        */
int myldpc_encode(myldpc_coder *c, char *srcCode, char *priorCode, int srcLength);            /* :554-569 */
int myldpc_decode(myldpc_coder *c, float *postCode, char *srcCode, int srcLength, int deType); /* :571-618 */
int myldpc_test(myldpc_coder *c, char *priorCode, float *postCode, int priorCodeLength, float sd); /* :1061-1078 */
int myldpc_getPriorCodeLength(myldpc_coder *c, int srcLength);                 /* :620-622                        */
int myldpc_getPostCodeLength(myldpc_coder *c, int srcLength);                  /* :624-626                        */
int myldpc_getCodeSize(myldpc_coder *c, int srcLength);                        /* :628-631                        */
/* public member checkMatrix (MyLdpc.h:128) as CSR; any pointer may be NULL (query rows/cols/nnz first) */
int myldpc_checkMatrix(myldpc_coder *c, int *rows, int *cols, int *nnz, int *rowPtr, int *colIdx);

/* [B200] additive members of include/MyLdpc.h */
int myldpc_setMaxIter(myldpc_coder *c, int times);
int myldpc_setDevices(myldpc_coder *c, const int *deviceIds, int count);
int myldpc_setEarlyTermination(myldpc_coder *c, int on);
int myldpc_setStrictDecodeType(myldpc_coder *c, int strict);
int myldpc_setFusedKernelArithmetic(myldpc_coder *c, int exact);
int myldpc_setRegisterHostBuffers(myldpc_coder *c, int on);
int myldpc_lastAlgorithm(myldpc_coder *c);
const int *myldpc_lastIterations(myldpc_coder *c);
int myldpc_lastCodeSize(myldpc_coder *c);
const char *myldpc_lastError(myldpc_coder *c);
int myldpc_lastStepTimes(myldpc_coder *c, double *seconds, int n);

#ifdef __cplusplus
}
#endif
#endif /* MYLDPC_C_H_ */
