/*
 * oracle/ldpc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's min-sum decode path (wing02/MyLdpcCppApi).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the stated CPU baseline.
 * The product (libldpc_b200.so, libmyldpc_b200.so) never links or calls it.
 *
 * Parity status: the reference ships no golden vectors (its only test, Test.cpp, is
 * unseeded).  The pin is oracle/_ref: the reference's own MyLdpc.cpp compiled unmodified
 * against container-only shims (see oracle/Makefile, oracle/shim/), whose decodeCPU output
 * is compared with this restatement in tests/test_oracle_vs_ref.py, plus committed golden
 * fixtures generated from it (tests/golden/).  The sum-product, layered and fused-kernel
 * restatements are pinned the same way against oracle/_ref/libmyldpc_refcl.so, which also compiles
 * the reference's decodeCL.c unmodified and executes its kernels on the CPU
 * (tests/test_oracle_vs_refcl.py).
 */
#ifndef LDPC_ORACLE_H_
#define LDPC_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rate_type values, reference MyLdpc.h:33-35 */
enum { ORACLE_RATE_1_2 = 0, ORACLE_RATE_2_3_A, ORACLE_RATE_2_3_B, ORACLE_RATE_3_4_A,
       ORACLE_RATE_3_4_B, ORACLE_RATE_5_6 };

/* Edge tables exactly as Coder::forDecoder builds them (reference MyLdpc.cpp:167-222). */
typedef struct oracle_tables {
    int M, N, nnz;
    int *hRows, *hCols;            /* [nnz]  row / column of edge e (e = CSR position)   */
    int *hRowFirstPtr, *hRowNextPtr; /* [M], [nnz]  singly linked list per row, -1 ends  */
    int *hColFirstPtr, *hColNextPtr; /* [N], [nnz]  singly linked list per column        */
    int *hRowRange;                /* [M+1] CSR offsets                                   */
} oracle_tables;

/* Base-matrix tables live in the oracle's own copy (oracle/ldpc_oracle.c) so that the
 * oracle does not depend on product code.  `seed` is rows x 24, row-major.              */
int oracle_wimax_seed(int rate, const signed char **seed, int *seed_rows);

/* Restates Coder::initCheckMatrix (reference MyLdpc.cpp:52-109): expands the 802.16e seed
 * for code length N (z = N/24) into CSR (row-major, ascending column inside a row).
 * Caller frees *row_ptr and *col_idx with oracle_free().  Returns nnz or <0.             */
int oracle_wimax_H(int N, int rate, int32_t **row_ptr, int32_t **col_idx, int *M_out);

oracle_tables *oracle_tables_build(int M, int N, const int32_t *row_ptr, const int32_t *col_idx);
void oracle_tables_free(oracle_tables *t);
void oracle_free(void *p);

/* Restates Coder::decodeCPU (reference MyLdpc.cpp:684-784) over the stream API:
 *   postCode : codeSize*N floats, codeSize = ceil(srcLength/(K/8))  (MyLdpc.cpp:628-631)
 *   srcCode  : >= srcLength+1 bytes (the reference's guard is `charOffset <= srcLength`)
 *   times    : iteration cap (reference: 40, MyLdpc.cpp:24)
 * The reference frees `time`, `src` and `lPostP`; we expose them (any may be NULL):
 *   iters[codeSize], hard[codeSize*N] (0/1 bytes), post[codeSize*N].                     */
int oracle_decodeCPU(const oracle_tables *t, int K, int times, const float *postCode,
                     char *srcCode, int srcLength, int32_t *iters, uint8_t *hard, float *post);

/* Same results, per-codeword records (K/8 bytes each, K%8==0), codewords split over
 * nthreads pthreads.  `literal`!=0 uses the reference's O(d_c^2) check-node loop (the
 * faithful CPU baseline); 0 uses the min1/min2 form (checked equal in tests).           */
int oracle_decode_batch(const oracle_tables *t, int K, int times, const float *llr,
                        int64_t ncw, uint8_t *info, int32_t *iters, uint8_t *hard,
                        float *post, int nthreads, int literal);

/* ---- sum-product (DecodeSP) -------------------------------------------------------------------
 * Restates the reference's probability-domain sum-product decoder, which exists only as OpenCL
 * kernels + host loop (decodeInit / refreshR / hardDecision / checkResult / refreshQ, decodeCL.c:3-108;
 * Coder::decodeOnceSP, MyLdpc.cpp:977-1059), per codeword (the reference freezes a word with isDones
 * once its syndrome is clean, decodeCL.c:45-49, so per-word results equal a per-word loop):
 *   init      q0 = t/(1+t), q1 = 1/(1+t), t = exp(8*y) per edge; prior0/1 the same per column
 *   refreshR  d_e = product over the row's OTHER edges (row-list order) of (q0-q1); r0=(1+d)/2, r1=(1-d)/2
 *   hardDecision  t0 = prior0 * prod r0, t1 = prior1 * prod r1 over the column list; bit 0 if t0>t1,
 *             1 if t0<t1, unchanged on a tie (bits start at 0 here; the reference's buffer is uninitialised)
 *   checkResult / ++time / stop if clean or time == times
 *   refreshQ  t0 = prior0 * prod_{others} r0, t1 likewise (column-list order); q0 = t0/(t0+t1), q1 = t1/(t0+t1)
 * `exp` is an OpenCL built-in in the reference (implementation-defined rounding, no canonical value):
 * oracle and CUDA kernel both use oracle_sp_expf / the identical device routine, a fixed sequence of
 * IEEE fp32 operations, so the two sides can be compared bit for bit.
 * PINNED: the reference's own decodeOnceSP host loop driving its own kernels (oracle/_ref/libmyldpc_refcl.so,
 * exp supplied by oracle_sp_expf, every other operation the reference's) gives the bytes and the per-word
 * iteration counts of this restatement for all six rates, saturating inputs included
 * (tests/test_oracle_vs_refcl.py::test_reference_opencl_sum_product_*).                                  */
float oracle_sp_expf(float x);
int oracle_decode_sp_batch(const oracle_tables *t, int K, int times, const float *llr, int64_t ncw,
                           uint8_t *info, int32_t *iters, uint8_t *hard, float *post0, float *post1,
                           int nthreads);

/* ---- layered min-sum (DecodeTDMP) ----------------------------------------------------------------
 * Restates the schedule the reference's TDMP path intends (host loop Coder::decodeOnceTDMP,
 * MyLdpc.cpp:889-976; kernels decodeInitTDMP / refreshRTDMP / refreshPostPTDMP / refreshQTDMP /
 * hardDecisionTDMP, decodeCL.c:203-292): the rows are processed in layers of z consecutive rows (one block
 * row of the QC matrix); inside a layer  lQ = lPostP - lR,  lR = min-sum over the row's other edges
 * (clamp 1000, as decodeCPU),  lPostP = lQ + lR;  after the last layer a hard decision
 * (>0 -> 0, <0 -> 1, ==0 keeps the previous bit; bits start at 0 here, the reference's buffer is
 * uninitialised), the syndrome check, ++time, stop if clean or time == times.
 * PINNED where the reference's own loop is sound: decodeOnceTDMP sizes layer b as
 * hRowRange[b + z] - hRowRange[b] (MyLdpc.cpp:907,958; b is a BLOCK row index), which is the layer's edge count
 * exactly when all rows have one weight -- rates 2/3A and 5/6.  There the reference's own host loop and kernels,
 * executed on the CPU (oracle/_ref/libmyldpc_refcl.so), give the bytes and per-word iteration counts of this
 * restatement.  For mixed row weights (rates 1/2, 2/3B, 3/4A, 3/4B) the shipped loop launches the wrong number
 * of work-items from the second layer on and its output is no longer a layered schedule
 * (tests/test_oracle_vs_refcl.py::test_reference_host_looped_tdmp_*); this restatement uses the true layer
 * boundaries hRowRange[(b+1) z] there.  (lQ is layer-relative in all four kernels, decodeCL.c:232-258,284-292, and
 * decodeInitTDMP's first N entries cover the first layer: those two are consistent, not defects.)
 * Returns -1 if some column appears twice inside a layer (the reference's per-edge threads would race there).  */
int oracle_tdmp_layering_ok(const oracle_tables *t, int z);
int oracle_decode_tdmp_batch(const oracle_tables *t, int K, int times, int z, const float *llr, int64_t ncw,
                             uint8_t *info, int32_t *iters, uint8_t *hard, float *post, int nthreads);

/* ---- the fused one-kernel decoders (DecodeMSCL / DecodeTDMPCL) ------------------------------------------------
 * Literal restatement of decodeOnceMS (mode 0, decodeCL.c:432-567) and decodeOnceTDMP (mode 1, decodeCL.c:307-426):
 * the min-sum / layered min-sum schedules with those kernels' own arithmetic -- message sign through the float
 * PRODUCT of the row's Q (a zero or an underflow zeroes the whole row), min search from (1000, 1001), hard decision
 * bit = (P < 0), caps passed in `times` (the kernels hard-code 120 and 40).  Per-word iteration counts and the
 * final posterior are exposed (the kernels keep them private).  Pinned bit for bit against the kernels themselves,
 * executed by oracle/shim/cl_exec.h (tests/test_oracle_vs_refcl.py).  On inputs that never produce an exact-zero
 * message or posterior the results equal oracle_decode_batch (mode 0) / oracle_decode_tdmp_batch (mode 1).     */
int oracle_decode_fused_batch(const oracle_tables *t, int K, int times, int z, int mode, const float *llr, int64_t ncw,
                              uint8_t *info, int32_t *iters, uint8_t *hard, float *post, int nthreads);

/* Restates Coder::test's bit->BPSK map (reference MyLdpc.cpp:1061-1072), noise supplied by
 * the caller (the reference's rand()-based Box-Muller is unseeded).                      */
void oracle_bpsk(const uint8_t *bytes, int nbytes, float *out);

/* Size helpers, reference MyLdpc.cpp:620-631. */
int oracle_getCodeSize(int K, int srcLength);
int oracle_getPostCodeLength(int K, int N, int srcLength);
int oracle_getPriorCodeLength(int K, int N, int srcLength);

#ifdef __cplusplus
}
#endif
#endif
