/*
 * ldpc_b200.h -- C-ABI of the B200-native min-sum LDPC decoder (libldpc_b200.so).
 *
 * This is the drop-in boundary: the entry points below are what Coder::{forDecoder,
 * addDecodeType, decode} bind instead of the reference's OpenCL host calls.  Plain pointers
 * and sizes only -- no C++ types, no torch types, no exceptions cross this boundary.
 * Every function returns LDPC_B200_OK (0) or a negative error code; ldpc_b200_last_error()
 * gives the message for the calling thread.  There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with LDPC_B200_ERR_CUDA.
 *
 * Reference interfaces replaced (wing02/MyLdpcCppApi):
 *   ldpc_b200_create          <- Coder::forDecoder edge tables + OpenCL context/queue/program
 *                                + table upload              (MyLdpc.cpp:167-305)
 *   ldpc_b200_set_max_iter    <- `times = 40` in the ctor    (MyLdpc.cpp:24)
 *   ldpc_b200_reserve         <- Coder::addDecodeType(DecodeMS) buffer allocation
 *                                                            (MyLdpc.cpp:387-437)
 *   ldpc_b200_decode_host     <- Coder::decodeOnceMS: enqueueWriteBuffer, decodeInitMS,
 *                                loop{refreshRMS, refreshPostPMS, checkResult, read flags,
 *                                refreshQMS}, toChar, enqueueReadBuffer
 *                                (MyLdpc.cpp:786-848; kernels decodeCL.c:88-199)
 *   ldpc_b200_decode_device   <- the same with device-resident buffers (no PCIe leg)
 *   ldpc_b200_destroy         <- Coder::~Coder / cl::Buffer RAII (MyLdpc.cpp:31-51)
 *   ldpc_b200_synth_llr       <- Coder::test BPSK + AWGN          (MyLdpc.cpp:1061-1105)
 *   ldpc_b200_encode_*        <- Coder::forEncoder / Coder::encode (MyLdpc.cpp:137-165, 554-569, 633-682)
 *
 * Decode semantics are those of Coder::decodeCPU (MyLdpc.cpp:684-784), per codeword:
 * flooding min-sum in fp32, messages clamped at 1000, posterior accumulated from the channel
 * value in ascending-row order, hard bit = !(posterior > 0), syndrome check after every
 * iteration from 1, stop at `max_iter` inclusive; output = the first K hard bits packed
 * LSB-first.  Input values are the raw channel samples (+1/-1 plus noise), as in the
 * reference -- min-sum is scale-free.  NaN inputs are outside the contract.
 */
#ifndef LDPC_B200_H_
#define LDPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldpc_b200_decoder *ldpc_b200_handle;

enum {
    LDPC_B200_OK = 0,
    LDPC_B200_ERR_ARG = -1,         /* bad argument / malformed parity-check matrix */
    LDPC_B200_ERR_CUDA = -2,        /* CUDA runtime failure or no device             */
    LDPC_B200_ERR_UNSUPPORTED = -3, /* e.g. check degree above the packed-state limit */
    LDPC_B200_ERR_NOMEM = -4
};

/* Which kernel family a handle uses (ldpc_b200_info.path). */
enum {
    LDPC_B200_PATH_LANE_SMEM = 0,  /* 32 codewords per CTA, lane = codeword, state in shared memory */
    LDPC_B200_PATH_LANE_GLOBAL = 1,/* same schedule, state in a CTA-private L2/HBM workspace        */
    /* 2 is unassigned */
    LDPC_B200_PATH_GROUP = 4,      /* explicit per-edge messages in shared memory; G codewords per CTA,
                                      32/G graph nodes per warp instruction (G=16 short codes, G=1 N~8k) */
    LDPC_B200_PATH_CLUSTER = 5,    /* long codes: one codeword per 8-CTA thread-block cluster, posteriors and
                                      messages in distributed shared memory (DSMEM gathers, cluster barriers) */
    LDPC_B200_PATH_STREAM = 6,     /* long codes with check and variable degree <= 8: explicit per-edge messages in a
                                      global workspace, fixed-stride tables, prefetched indices, variable bundles   */
    LDPC_B200_PATH_LANE16 = 3,     /* tuned short-code path: lane = codeword, channel values in
                                      registers, 16-byte check state and index tables in shared memory */
    LDPC_B200_PATH_WARP = 8,       /* opt-in, measured alternative: one codeword per CTA, a sub-warp per check with
                                      shuffle / ballot row reductions, a thread per variable (the textbook mapping) */
    LDPC_B200_PATH_QC = 7          /* quasi-cyclic codes (every code Coder::initCheckMatrix builds): the GROUP
                                      arithmetic with warp-uniform index tables (constant memory -> uniform
                                      registers for the compiled profile of Test.cpp's code; shared memory for the
                                      run-time profile used for z > 24); cyclic wrap absorbed by padded rows     */
};

/* Decoding algorithm of a handle (ldpc_b200_set_algorithm). */
enum {
    LDPC_B200_ALG_MIN_SUM = 0,     /* Coder::decodeCPU / DecodeMS semantics (default)                    */
    LDPC_B200_ALG_SUM_PRODUCT = 1, /* DecodeSP: probability-domain sum-product of decodeCL.c:3-108 under
                                      decodeOnceSP (MyLdpc.cpp:977-1059); on-chip kernel for short codes (group
                                      layout, variable degree <= 8, check degree <= 20), the any-size kernel
                                      (global workspace) otherwise; no posterior output                 */
    LDPC_B200_ALG_LAYERED_MIN_SUM = 2, /* DecodeTDMP: layered min-sum, the schedule decodeOnceTDMP intends
                                      (MyLdpc.cpp:889-976, decodeCL.c:203-292): layers of z consecutive rows,
                                      Q = P - R, R = min-sum, P = Q + R per layer, then hard decision
                                      (P == 0 keeps the bit) and syndrome.  Needs the layer height
                                      (known for ldpc_b200_create_wimax, else ldpc_b200_set_layer_height),
                                      column-disjoint layers; on-chip kernel for <= 16 layers, N <= 32 z that fit
                                      shared memory, the any-size kernel (global workspace) otherwise        */
    LDPC_B200_ALG_FUSED_MIN_SUM = 3,  /* DecodeMSCL with the ARITHMETIC of the reference's fused kernel decodeOnceMS
                                      (decodeCL.c:432-567) reproduced exactly: message sign through the float
                                      product of the row's Q (a zero, an underflow or inf*0 zeroes the row), minimum
                                      search from (1000, 1001), bit = (P < 0).  Differs from MIN_SUM only on such
                                      corner events.  Any-size kernel (global workspace): exact, not fast.
                                      The kernel's cap is 120: set it with ldpc_b200_set_max_iter.          */
    LDPC_B200_ALG_FUSED_LAYERED = 4   /* DecodeTDMPCL likewise (decodeOnceTDMP, decodeCL.c:307-426; cap 40)   */
};

typedef struct ldpc_b200_info {
    int M, N, K, nnz;
    int max_row_weight, max_col_weight;
    int max_iter, early_termination;
    int device, sm_count;
    int path;            /* LDPC_B200_PATH_* */
    int threads_per_cta; /* launch shape                                            */
    int ctas;            /* persistent grid size                                    */
    int codewords_per_cta;
    size_t smem_bytes;   /* dynamic shared memory per CTA                           */
    size_t workspace_bytes;
    size_t table_bytes;  /* device bytes of the check-major + variable-major tables */
    int kernel_variant;  /* quasi-cyclic path with a compiled lockstep profile, most recent launch: 0 lockstep kernel,
                            1 warp-per-codeword kernel, 3 group-of-warps kernel (chosen per launch from the previous
                            launches' mean iteration count; option "qc_et": -1 auto, 0 never, 1 always), 2 ring-staged,
                            4 group-of-warps kernel with several codewords per group (block sizes without a lockstep
                            profile, while the words run long; option "qcm_multi_pct"), 5 quasi-cyclic sum-product
                            kernel (LDPC_B200_ALG_SUM_PRODUCT on codes the on-chip sum-product layout cannot hold) */
    int et_available;    /* which per-codeword kernel this handle can switch to: 0 none, 1 warp per codeword (802.16e
                            codes with z = 24 / 32), 2 group of warps */
} ldpc_b200_info;

/* Build a decoder for the M x N parity-check matrix given in CSR (row_ptr[M+1],
 * col_idx[nnz]; each row's columns distinct).  K = number of leading (systematic) bits
 * reported per codeword.  The edge order that fixes the fp32 summation order is
 * (row, column) ascending, i.e. the order Eigen's RowMajor InnerIterator yields in the
 * reference (MyLdpc.cpp:188-191).  `device` is the CUDA ordinal.                       */
int ldpc_b200_create(ldpc_b200_handle *out, int M, int N, int K, const int32_t *row_ptr,
                     const int32_t *col_idx, int device);

/* Convenience: the reference's 802.16e code for (K, N, rate) -- rate is the reference's
 * enum rate_type value 0..5 (MyLdpc.h:33-35); H as built by Coder::initCheckMatrix.     */
int ldpc_b200_create_wimax(ldpc_b200_handle *out, int K, int N, int rate, int device);

int ldpc_b200_destroy(ldpc_b200_handle h);

int ldpc_b200_set_max_iter(ldpc_b200_handle h, int max_iter);        /* default 40, 1..65535 */
/* 1 (default) = the reference's per-codeword syndrome stop; 0 = always run max_iter
 * iterations (throughput runs; results then differ from the reference for words that
 * would have converged earlier only in the iteration count and later posteriors).       */
int ldpc_b200_set_early_termination(ldpc_b200_handle h, int on);
int ldpc_b200_set_algorithm(ldpc_b200_handle h, int algorithm);     /* LDPC_B200_ALG_* */
/* Rows per layer for LDPC_B200_ALG_LAYERED_MIN_SUM (the reference's z, MyLdpc.cpp:22: one block row of the
 * quasi-cyclic matrix).  ldpc_b200_create_wimax sets it to N/24.                                          */
int ldpc_b200_set_layer_height(ldpc_b200_handle h, int z);
/* Force a kernel path (LDPC_B200_PATH_*) or -1 for automatic choice. */
int ldpc_b200_set_path(ldpc_b200_handle h, int path);
/* Experiment switches of the host-buffer pipeline.  Every switch is seeded ONCE, in
 * ldpc_b200_create, from the environment variable LDPC_B200_<NAME>; nothing on the decode path reads the
 * environment.  The ones below can be changed per handle afterwards (the kernel-choice switches shape the
 * plan made at create time and are refused here with LDPC_B200_ERR_UNSUPPORTED):
 *   "refill_wait", "no_streamed", "streamed_pageable", "no_staged", "staged_min_kb", "stream_chunk",
 *   "stream_batch_kb", "wait_timeout_ms" (bound of the persistent kernel's wait for streamed input; when it
 *   expires the call is rerun through the chunked pipeline), "register_host" (page-lock a pageable input buffer
 *   with cudaHostRegister the first time it is seen and keep it so until ldpc_b200_destroy: the caller promises
 *   not to free it before; repeated decodes out of one malloc'd buffer then run at pinned-memory speed).        */
int ldpc_b200_set_option(ldpc_b200_handle h, const char *name, long long value);
int ldpc_b200_get_info(ldpc_b200_handle h, ldpc_b200_info *info);

/* Copy the CSR of H back (for Coder::checkMatrix); arrays sized M+1 and nnz. */
int ldpc_b200_get_csr(ldpc_b200_handle h, int32_t *row_ptr, int32_t *col_idx);

/* Pre-allocate device staging for host-buffer decodes of up to `batch` codewords per
 * chunk (what forDecoder(batchSize)/addDecodeType do in the reference).                 */
int ldpc_b200_reserve(ldpc_b200_handle h, int64_t batch);

/* Decode `ncw` codewords whose channel values already sit in device memory.
 *   d_llr   : [ncw][N] float32, row-major (the reference's postCode layout)
 *   d_info  : [ncw][ceil(K/8)] bytes, LSB-first (may be NULL)
 *   d_hard  : [ncw][ceil(N/8)] bytes, all N hard bits LSB-first (may be NULL)
 *   d_iters : [ncw] int32 iteration count at exit, 1..max_iter (may be NULL)
 *   d_post  : [ncw][N] float32 posterior values lPostP (may be NULL)
 * Buffers need only the alignment of their element type (4 bytes for floats and counts,
 * 1 for packed bits): a slice of a larger array is fine (tests/test_gpu_bounds.py).
 * Asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream).    */
int ldpc_b200_decode_device(ldpc_b200_handle h, const float *d_llr, int64_t ncw,
                            uint8_t *d_info, uint8_t *d_hard, int32_t *d_iters,
                            float *d_post, void *stream);

/* Same with HOST buffers: chunks of the reserved batch are copied in, decoded and copied
 * out on rotating streams so PCIe transfers overlap the kernels; returns when all outputs
 * are in host memory.  Pinned buffers get full-speed async copies.                      */
int ldpc_b200_decode_host(ldpc_b200_handle h, const float *llr, int64_t ncw, uint8_t *info,
                          uint8_t *hard, int32_t *iters, float *post);

/* Systematic encoder on the device (Coder::forEncoder + Coder::encode, MyLdpc.cpp:137-165, 554-569, 633-682).
 * info: [ncw][K/8] bytes, bits LSB-first (the reference's srcCode split per codeword); codewords: [ncw][N/8]
 * bytes in the reference's priorCode layout (info bits first, then the M parity bits of H c = 0).
 * encoder_init solves the parity part of H over GF(2) once (needs N - K == M, an invertible parity part,
 * K and N multiples of 8, K <= 4096) and is implied by the first encode call.                          */
int ldpc_b200_encoder_init(ldpc_b200_handle h);
int ldpc_b200_encode_device(ldpc_b200_handle h, const uint8_t *d_info, int64_t ncw, uint8_t *d_codewords, void *stream);
int ldpc_b200_encode_host(ldpc_b200_handle h, const uint8_t *info, int64_t ncw, uint8_t *codewords);

/* The same call with the channel values in a PACKED host format, widened to fp32 on the device: the decoder sees
 * exactly (float)x * scale (one fp32 multiplication; scale = 1 widens fp16 exactly), so results equal
 * ldpc_b200_decode_host on those floats bit for bit.  Additive: the reference's API has fp32 only (MyLdpc.h:118).
 * Why: in the early-termination regime and on multi-GPU hosts the host-to-device copy, not the kernel, bounds the
 * call; fp16 halves those bytes, int8 quarters them.  Pinned buffers (ldpc_b200_host_alloc)
 * are copied by DMA under the running kernels; pageable ones work but are staged by the driver. */
enum {
    LDPC_B200_LLR_F32 = 0,  /* const float*   */
    LDPC_B200_LLR_F16 = 1,  /* IEEE binary16  */
    LDPC_B200_LLR_I8 = 2    /* const int8_t*  */
};
int ldpc_b200_decode_host_packed(ldpc_b200_handle h, const void *llr, int format, float scale, int64_t ncw,
                                 uint8_t *info, uint8_t *hard, int32_t *iters, float *post);


/* Synthetic BPSK-AWGN channel on the device: y = (bit ? -1 : +1) + sigma * n, n ~ N(0,1)
 * from a counter-based generator keyed by (seed, codeword, position).  d_bits: packed
 * codeword bits [ncw][ceil(N/8)] LSB-first, or NULL for the all-zero codeword.          */
int ldpc_b200_synth_llr(float *d_llr, int64_t ncw, int N, float sigma, uint64_t seed,
                        const uint8_t *d_bits, int device, void *stream);

/* Host-only helpers (no device needed).
 * ldpc_b200_wimax_csr: H of the reference's 802.16e code as CSR; pass NULL arrays to query
 * M and nnz first (row_ptr needs M+1 ints, col_idx nnz ints).
 * ldpc_b200_edge_tables: the variable-major table the kernels use -- col_ptr[N+1] and
 * vn_edge[nnz] = (check << 5) | position-in-check, per variable in ascending row order.  */
int ldpc_b200_wimax_csr(int K, int N, int rate, int32_t *row_ptr, int32_t *col_idx, int *M, int *nnz);
int ldpc_b200_edge_tables(int M, int N, const int32_t *row_ptr, const int32_t *col_idx,
                          int32_t *col_ptr, uint32_t *vn_edge, int *max_row_weight,
                          int *max_col_weight);

/* Measurement aid: streams conflict-free 16-byte shared-memory loads on every SM and
 * reports the achieved GB/s (the denominator of the on-chip roofline in bench.py).       */
int ldpc_b200_probe_smem_bandwidth(int device, double *gbytes_per_s);

/* Run-time phase timers of ldpc_b200_decode_host, accumulated since creation / the last reset (the reference keeps
 * clock() sums per kernel launch in Coder::stepTime[], MyLdpc.cpp:26-28, 987-1056; here the iterations are fused into
 * one kernel, so the phases a caller can see are the copies and the kernel).  CUDA events on the handle's streams:
 * h2d_s = host->device copies of the channel values, kernel_s = decode kernels (on the streamed path the kernel
 * runs under the copies and its time includes waiting for them), d2h_s = read-back of the results; wall_s = host
 * wall clock of the calls.  The three device phases overlap, so they do not add up to wall_s.                       */
typedef struct ldpc_b200_timing {
    int64_t calls, codewords;
    double wall_s, h2d_s, kernel_s, d2h_s;
} ldpc_b200_timing;
int ldpc_b200_get_timing(ldpc_b200_handle h, ldpc_b200_timing *out);
int ldpc_b200_reset_timing(ldpc_b200_handle h);

/* Page-locked host memory for callers without the CUDA headers (the drop-in Coder keeps its result staging in it, so
 * that the device-to-host copies of a decode are true DMA).  NULL / an error code when the runtime refuses.          */
void *ldpc_b200_host_alloc(size_t bytes);
int ldpc_b200_host_free(void *p);

/* Number of kernel launches this handle has issued (bench.py's gpu_launches). */
int64_t ldpc_b200_launch_count(ldpc_b200_handle h);

const char *ldpc_b200_last_error(void);
const char *ldpc_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H_ */
