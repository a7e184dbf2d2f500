/*
 * oracle/ldpc_oracle.c -- TEST INFRASTRUCTURE ONLY (see ldpc_oracle.h).
 *
 * Plain-C restatement of the reference's CPU min-sum decode path.  Every function cites the
 * reference file:line it follows.  Build: gcc -O2 -ffp-contract=off -fno-fast-math (IEEE
 * binary32, one rounding per add, no FMA contraction -- the reference is built -O0 for
 * x86-64 SSE scalar math, Makefile:1, which has the same arithmetic).
 */
#include "ldpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "wimax_seed_tables.inc"

#define ORACLE_NB 24 /* n_b, reference MyLdpc.h:102 */

void oracle_free(void *p) { free(p); }

/* reference MyLdpc.cpp:58-83: seed table + number of block rows per rate */
int oracle_wimax_seed(int rate, const signed char **seed, int *seed_rows) {
    switch (rate) {
    case ORACLE_RATE_1_2:   *seed = oracle_seed_1_2;   *seed_rows = 12; return 0;
    case ORACLE_RATE_2_3_A: *seed = oracle_seed_2_3_a; *seed_rows = 8;  return 0;
    case ORACLE_RATE_2_3_B: *seed = oracle_seed_2_3_b; *seed_rows = 8;  return 0;
    case ORACLE_RATE_3_4_A: *seed = oracle_seed_3_4_a; *seed_rows = 6;  return 0;
    case ORACLE_RATE_3_4_B: *seed = oracle_seed_3_4_b; *seed_rows = 6;  return 0;
    case ORACLE_RATE_5_6:   *seed = oracle_seed_5_6;   *seed_rows = 4;  return 0;
    }
    return -1;
}

typedef struct { int r, c; } rc_t;
static int rc_cmp(const void *a, const void *b) {
    const rc_t *x = (const rc_t *)a, *y = (const rc_t *)b;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    return 0;
}

/* reference MyLdpc.cpp:52-109.  The triplet loop is restated literally (including the
 * O(z^2) scan of every block and the `(z + permutCol - permutRow) % z == permut` test);
 * Eigen's setFromTriplets into a RowMajor matrix is restated as a (row, col) sort, which is
 * the order its InnerIterator later yields (MyLdpc.cpp:188-191). */
int oracle_wimax_H(int N, int rate, int32_t **row_ptr, int32_t **col_idx, int *M_out) {
    const signed char *hSeed;
    int seedRowLength;
    if (oracle_wimax_seed(rate, &hSeed, &seedRowLength)) return -1;
    const int seedColLength = ORACLE_NB;
    int z = N / ORACLE_NB;                       /* :55 */
    int M = seedRowLength * z;
    size_t cap = (size_t)seedRowLength * seedColLength * z;
    rc_t *trip = (rc_t *)malloc(cap * sizeof(rc_t));
    size_t nt = 0;
    int permut;
    for (int seedRow = 0; seedRow < seedRowLength; ++seedRow) {
        for (int seedCol = 0; seedCol < seedColLength; ++seedCol) {
            if ((permut = hSeed[seedRow * seedColLength + seedCol]) >= 0) {
                if (rate != ORACLE_RATE_2_3_A) {
                    permut = permut * z / 96;    /* :91 integer floor */
                } else {
                    permut = permut % z;         /* :93 */
                }
                for (int permutRow = 0; permutRow < z; ++permutRow) {
                    for (int permutCol = 0; permutCol < z; ++permutCol) {
                        if ((z + permutCol - permutRow) % z == permut) {
                            trip[nt].r = seedRow * z + permutRow;
                            trip[nt].c = seedCol * z + permutCol;
                            ++nt;
                        }
                    }
                }
            }
        }
    }
    qsort(trip, nt, sizeof(rc_t), rc_cmp);
    int32_t *rp = (int32_t *)calloc((size_t)M + 1, sizeof(int32_t));
    int32_t *ci = (int32_t *)malloc(nt * sizeof(int32_t));
    for (size_t i = 0; i < nt; ++i) {
        rp[trip[i].r + 1]++;
        ci[i] = trip[i].c;
    }
    for (int r = 0; r < M; ++r) rp[r + 1] += rp[r];
    free(trip);
    *row_ptr = rp;
    *col_idx = ci;
    *M_out = M;
    return (int)nt;
}

/* reference MyLdpc.cpp:172-222: edge id = CSR position; hRows/hCols; per-row and per-column
 * linked lists built by tail-walk append; hRowRange. */
oracle_tables *oracle_tables_build(int M, int N, const int32_t *row_ptr, const int32_t *col_idx) {
    oracle_tables *t = (oracle_tables *)calloc(1, sizeof(*t));
    int nonZeros = row_ptr[M];
    t->M = M; t->N = N; t->nnz = nonZeros;
    t->hColFirstPtr = (int *)malloc(sizeof(int) * (size_t)N);
    t->hColNextPtr = (int *)malloc(sizeof(int) * (size_t)(nonZeros > 0 ? nonZeros : 1));
    t->hRowFirstPtr = (int *)malloc(sizeof(int) * (size_t)M);
    t->hRowNextPtr = (int *)malloc(sizeof(int) * (size_t)(nonZeros > 0 ? nonZeros : 1));
    memset(t->hColFirstPtr, -1, sizeof(int) * (size_t)N);
    memset(t->hColNextPtr, -1, sizeof(int) * (size_t)nonZeros);
    memset(t->hRowFirstPtr, -1, sizeof(int) * (size_t)M);
    memset(t->hRowNextPtr, -1, sizeof(int) * (size_t)nonZeros);
    t->hCols = (int *)malloc(sizeof(int) * (size_t)(nonZeros > 0 ? nonZeros : 1));
    t->hRows = (int *)malloc(sizeof(int) * (size_t)(nonZeros > 0 ? nonZeros : 1));
    t->hRowRange = (int *)malloc(sizeof(int) * ((size_t)M + 1));
    /* The reference appends by walking to the tail (O(nnz * degree)); a tail cache gives the
     * identical lists.  Kept explicit so long codes (N = 64800) set up quickly. */
    int *rowTail = (int *)malloc(sizeof(int) * (size_t)M);
    int *colTail = (int *)malloc(sizeof(int) * (size_t)N);
    int offset = 0;
    for (int k = 0; k < M; ++k) {
        t->hRowRange[k] = offset;
        for (int it = row_ptr[k]; it < row_ptr[k + 1]; ++it) {
            int row = k, col = col_idx[it];
            t->hRows[offset] = row;
            t->hCols[offset] = col;
            if (t->hRowFirstPtr[row] == -1) t->hRowFirstPtr[row] = offset;
            else t->hRowNextPtr[rowTail[row]] = offset;
            rowTail[row] = offset;
            if (t->hColFirstPtr[col] == -1) t->hColFirstPtr[col] = offset;
            else t->hColNextPtr[colTail[col]] = offset;
            colTail[col] = offset;
            ++offset;
        }
    }
    t->hRowRange[M] = offset;
    free(rowTail);
    free(colTail);
    return t;
}

void oracle_tables_free(oracle_tables *t) {
    if (!t) return;
    free(t->hRows); free(t->hCols);
    free(t->hRowFirstPtr); free(t->hRowNextPtr);
    free(t->hColFirstPtr); free(t->hColNextPtr);
    free(t->hRowRange);
    free(t);
}

/* reference MyLdpc.cpp:620-631 */
int oracle_getCodeSize(int K, int srcLength) { return (srcLength + (K / 8) - 1) / (K / 8); }
int oracle_getPostCodeLength(int K, int N, int srcLength) {
    return (srcLength + (K / 8) - 1) / (K / 8) * N;
}
int oracle_getPriorCodeLength(int K, int N, int srcLength) {
    return (srcLength + (K / 8) - 1) / (K / 8) * (N / 8);
}

typedef struct {
    unsigned char *lQA; /* bool in the reference */
    float *lQB, *lR, *lPostP;
    unsigned char *src;
} scratch_t;

static void scratch_alloc(scratch_t *s, int nnz, int N) {
    s->lQA = (unsigned char *)malloc((size_t)(nnz > 0 ? nnz : 1));
    s->lQB = (float *)malloc(sizeof(float) * (size_t)(nnz > 0 ? nnz : 1));
    s->lR = (float *)malloc(sizeof(float) * (size_t)(nnz > 0 ? nnz : 1));
    s->lPostP = (float *)malloc(sizeof(float) * (size_t)N);
    s->src = (unsigned char *)malloc((size_t)N);
}
static void scratch_free(scratch_t *s) {
    free(s->lQA); free(s->lQB); free(s->lR); free(s->lPostP); free(s->src);
}

/* One codeword of reference MyLdpc.cpp:695-763, literal loops.  Returns `time`. */
static int decode_one_literal(const oracle_tables *t, int times, const float *y, scratch_t *s) {
    const int nonZeros = t->nnz, ldpcN = t->N, ldpcM = t->M;
    int time = 0;
    for (int nodeInd = 0; nodeInd < nonZeros; ++nodeInd) {       /* :697-702 */
        int hCol = t->hCols[nodeInd];
        float code = y[hCol];
        s->lQA[nodeInd] = (code < 0);
        s->lQB[nodeInd] = fabsf(code);
    }
    while (1) {
        for (int nodeInd = 0; nodeInd < nonZeros; ++nodeInd) {   /* :705-721 refreshRMS */
            int hRow = t->hRows[nodeInd];
            unsigned char a = 0;
            float b = 1000;
            for (int ptr = t->hRowFirstPtr[hRow]; ptr != -1; ptr = t->hRowNextPtr[ptr]) {
                if (nodeInd == ptr) continue;
                if (s->lQA[ptr]) a = a ^ 1;
                b = fminf(b, s->lQB[ptr]);
            }
            if (a) s->lR[nodeInd] = -b;
            else s->lR[nodeInd] = b;
        }
        for (int nodeInd = 0; nodeInd < ldpcN; ++nodeInd) {      /* :723-735 refreshPostPMS */
            float tmp = y[nodeInd];
            for (int ptr = t->hColFirstPtr[nodeInd]; ptr != -1; ptr = t->hColNextPtr[ptr]) {
                tmp += s->lR[ptr];
            }
            if (tmp > 0) s->src[nodeInd] = 0;
            else s->src[nodeInd] = 1;
            s->lPostP[nodeInd] = tmp;
        }
        unsigned char flag = 0;                                  /* :737-750 checkResult */
        for (int nodeInd = 0; nodeInd < ldpcM; ++nodeInd) {
            unsigned char result = 0;
            for (int ptr = t->hRowFirstPtr[nodeInd]; ptr != -1; ptr = t->hRowNextPtr[ptr]) {
                if (s->src[t->hCols[ptr]]) result ^= 1;
            }
            if (result) { flag = 1; break; }
        }
        ++time;                                                  /* :751-755 */
        if (flag == 0) break;
        if (time == times) break;
        for (int nodeInd = 0; nodeInd < nonZeros; ++nodeInd) {   /* :757-762 refreshQ */
            int hCol = t->hCols[nodeInd];
            float lQ = s->lPostP[hCol] - s->lR[nodeInd];
            s->lQA[nodeInd] = (lQ < 0);
            s->lQB[nodeInd] = fabsf(lQ);
        }
    }
    return time;
}

/* Same arithmetic with the check-node loop in min1/min2/parity form over CSR rows.
 * Used for large test sizes; tests assert it equals decode_one_literal bit for bit. */
static int decode_one_fast(const oracle_tables *t, int times, const float *y, scratch_t *s) {
    const int nonZeros = t->nnz, ldpcN = t->N, ldpcM = t->M;
    int time = 0;
    for (int e = 0; e < nonZeros; ++e) {
        float code = y[t->hCols[e]];
        s->lQA[e] = (code < 0);
        s->lQB[e] = fabsf(code);
    }
    while (1) {
        for (int r = 0; r < ldpcM; ++r) {
            int e0 = t->hRowRange[r], e1 = t->hRowRange[r + 1];
            float m1 = INFINITY, m2 = INFINITY;
            int i1 = -1;
            unsigned char par = 0;
            for (int e = e0; e < e1; ++e) {
                float a = s->lQB[e];
                par ^= s->lQA[e];
                if (a < m1) { m2 = m1; m1 = a; i1 = e; }
                else if (a < m2) { m2 = a; }
            }
            for (int e = e0; e < e1; ++e) {
                float b = fminf(1000.0f, (e == i1) ? m2 : m1);
                unsigned char a = par ^ s->lQA[e];
                s->lR[e] = a ? -b : b;
            }
        }
        for (int n = 0; n < ldpcN; ++n) {
            float tmp = y[n];
            for (int ptr = t->hColFirstPtr[n]; ptr != -1; ptr = t->hColNextPtr[ptr]) tmp += s->lR[ptr];
            s->src[n] = (tmp > 0) ? 0 : 1;
            s->lPostP[n] = tmp;
        }
        unsigned char flag = 0;
        for (int r = 0; r < ldpcM && !flag; ++r) {
            unsigned char result = 0;
            for (int e = t->hRowRange[r]; e < t->hRowRange[r + 1]; ++e) result ^= s->src[t->hCols[e]];
            if (result) flag = 1;
        }
        ++time;
        if (!flag) break;
        if (time == times) break;
        for (int e = 0; e < nonZeros; ++e) {
            float lQ = s->lPostP[t->hCols[e]] - s->lR[e];
            s->lQA[e] = (lQ < 0);
            s->lQB[e] = fabsf(lQ);
        }
    }
    return time;
}

/* reference MyLdpc.cpp:684-784 */
int oracle_decodeCPU(const oracle_tables *t, int K, int times, const float *postCode,
                     char *srcCode, int srcLength, int32_t *iters, uint8_t *hard, float *post) {
    const int ldpcN = t->N, ldpcK = K;
    memset(srcCode, 0, (size_t)srcLength);                        /* :685 */
    int codeSize = oracle_getCodeSize(K, srcLength);              /* :686 */
    scratch_t s;
    scratch_alloc(&s, t->nnz, ldpcN);
    for (int batch = 0; batch < codeSize; ++batch) {              /* :694 */
        int time = decode_one_literal(t, times, postCode + (size_t)batch * ldpcN, &s);
        for (int tmp = 0; tmp < ldpcK; ++tmp) {                   /* :765-774 */
            if (s.src[tmp]) {
                int offset = batch * ldpcK + tmp;
                int charOffset = offset / 8;
                if (charOffset <= srcLength) {
                    int bitOffset = offset % 8;
                    srcCode[charOffset] |= (char)(1 << bitOffset);
                }
            }
        }
        if (iters) iters[batch] = time;
        if (hard) memcpy(hard + (size_t)batch * ldpcN, s.src, (size_t)ldpcN);
        if (post) memcpy(post + (size_t)batch * ldpcN, s.lPostP, sizeof(float) * (size_t)ldpcN);
    }
    scratch_free(&s);
    return 0;
}

typedef struct {
    const oracle_tables *t;
    int K, times, literal;
    const float *llr;
    int64_t b0, b1;
    uint8_t *info, *hard;
    int32_t *iters;
    float *post;
} job_t;

static void *job_run(void *arg) {
    job_t *j = (job_t *)arg;
    const int N = j->t->N, K = j->K, KB = (K + 7) / 8;
    scratch_t s;
    scratch_alloc(&s, j->t->nnz, N);
    for (int64_t b = j->b0; b < j->b1; ++b) {
        const float *y = j->llr + (size_t)b * N;
        int time = j->literal ? decode_one_literal(j->t, j->times, y, &s)
                              : decode_one_fast(j->t, j->times, y, &s);
        if (j->info) {
            uint8_t *o = j->info + (size_t)b * KB;
            memset(o, 0, (size_t)KB);
            for (int i = 0; i < K; ++i)
                if (s.src[i]) o[i >> 3] |= (uint8_t)(1u << (i & 7)); /* LSB first, :765-774 */
        }
        if (j->iters) j->iters[b] = time;
        if (j->hard) memcpy(j->hard + (size_t)b * N, s.src, (size_t)N);
        if (j->post) memcpy(j->post + (size_t)b * N, s.lPostP, sizeof(float) * (size_t)N);
    }
    scratch_free(&s);
    return NULL;
}

int oracle_decode_batch(const oracle_tables *t, int K, int times, const float *llr,
                        int64_t ncw, uint8_t *info, int32_t *iters, uint8_t *hard,
                        float *post, int nthreads, int literal) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > ncw) nthreads = ncw > 0 ? (int)ncw : 1;
    job_t jobs[256];
    pthread_t th[256];
    for (int i = 0; i < nthreads; ++i) {
        job_t *j = &jobs[i];
        j->t = t; j->K = K; j->times = times; j->literal = literal; j->llr = llr;
        j->b0 = ncw * i / nthreads; j->b1 = ncw * (i + 1) / nthreads;
        j->info = info; j->iters = iters; j->hard = hard; j->post = post;
    }
    if (nthreads == 1) { job_run(&jobs[0]); return 0; }
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, job_run, &jobs[i]);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    return 0;
}

/* ---- sum-product restatement (decodeCL.c:3-108, MyLdpc.cpp:977-1059) ---------------------------- */

/* exp(x) as a fixed sequence of IEEE binary32 operations (no FMA), shared verbatim with the CUDA
 * kernel (csrc/ldpc_sp.cuh: sp_expf).  Cody-Waite reduction + degree-6 polynomial, ~2 ulp; saturates
 * like expf: +inf above 88.72, 0 below -103.  */
float oracle_sp_expf(float x) {
    if (x > 88.72283f) return INFINITY;
    if (x < -103.0f) return 0.0f;
    const float kf = (x * 1.44269504f + 12582912.0f) - 12582912.0f; /* round to nearest integer */
    float r = x - kf * 0.693359375f;                                 /* ln2 high part: 9 bits, exact product */
    r = r - kf * -2.12194440e-4f;                                    /* ln2 low part */
    float p = 1.38888889e-3f;                                        /* 1/720 */
    p = p * r + 8.33333377e-3f;
    p = p * r + 4.16666679e-2f;
    p = p * r + 1.66666672e-1f;
    p = p * r + 0.5f;
    p = p * r + 1.0f;
    p = p * r + 1.0f;
    int k = (int)kf;
    /* scale by 2^k in two steps so that results in the denormal range are produced by multiplication */
    union { float f; uint32_t u; } s1, s2;
    int k1 = k / 2, k2 = k - k1;
    s1.u = (uint32_t)(k1 + 127) << 23;
    s2.u = (uint32_t)(k2 + 127) << 23;
    return (p * s1.f) * s2.f;
}

typedef struct {
    const oracle_tables *t;
    int K, times;
    const float *llr;
    int64_t b0, b1;
    uint8_t *info, *hard;
    int32_t *iters;
    float *post0, *post1;
} spjob_t;

static void *spjob_run(void *arg) {
    spjob_t *j = (spjob_t *)arg;
    const oracle_tables *t = j->t;
    const int nonZeros = t->nnz, ldpcN = t->N, ldpcM = t->M, K = j->K, KB = (K + 7) / 8;
    float *q0 = (float *)malloc(sizeof(float) * (size_t)nonZeros), *q1 = (float *)malloc(sizeof(float) * (size_t)nonZeros);
    float *r0 = (float *)malloc(sizeof(float) * (size_t)nonZeros), *r1 = (float *)malloc(sizeof(float) * (size_t)nonZeros);
    float *priorP0 = (float *)malloc(sizeof(float) * (size_t)ldpcN), *priorP1 = (float *)malloc(sizeof(float) * (size_t)ldpcN);
    float *pt0 = (float *)malloc(sizeof(float) * (size_t)ldpcN), *pt1 = (float *)malloc(sizeof(float) * (size_t)ldpcN);
    unsigned char *src = (unsigned char *)malloc((size_t)ldpcN);
    for (int64_t b = j->b0; b < j->b1; ++b) {
        const float *codes = j->llr + (size_t)b * ldpcN;
        /* decodeInit, decodeCL.c:3-22 */
        for (int nodeInd = 0; nodeInd < nonZeros; ++nodeInd) {
            float tmp = oracle_sp_expf(8 * codes[t->hCols[nodeInd]]);
            q0[nodeInd] = tmp / (1 + tmp);
            q1[nodeInd] = 1 / (1 + tmp);
        }
        for (int n = 0; n < ldpcN; ++n) {
            float tmp = oracle_sp_expf(8 * codes[n]);
            priorP0[n] = tmp / (1 + tmp);
            priorP1[n] = 1 / (1 + tmp);
            src[n] = 0;
            pt0[n] = priorP0[n]; pt1[n] = priorP1[n];
        }
        int time = 0;
        while (1) {
            /* refreshR, decodeCL.c:25-41 */
            for (int nodeInd = 0; nodeInd < nonZeros; ++nodeInd) {
                int hRow = t->hRows[nodeInd];
                float dTmp = 1;
                for (int ptr = t->hRowFirstPtr[hRow]; ptr != -1; ptr = t->hRowNextPtr[ptr]) {
                    if (nodeInd == ptr) continue;
                    dTmp *= q0[ptr] - q1[ptr];
                }
                r0[nodeInd] = (1 + dTmp) / 2;
                r1[nodeInd] = (1 - dTmp) / 2;
            }
            /* hardDecision, decodeCL.c:64-86 */
            for (int col = 0; col < ldpcN; ++col) {
                float tmp0 = priorP0[col], tmp1 = priorP1[col];
                for (int ptr = t->hColFirstPtr[col]; ptr != -1; ptr = t->hColNextPtr[ptr]) {
                    tmp0 *= r0[ptr];
                    tmp1 *= r1[ptr];
                }
                if (tmp0 > tmp1) src[col] = 0;
                else if (tmp0 < tmp1) src[col] = 1;
                pt0[col] = tmp0; pt1[col] = tmp1;
            }
            /* checkResult, decodeCL.c:88-108 */
            unsigned char flag = 0;
            for (int row = 0; row < ldpcM && !flag; ++row) {
                unsigned char result = 0;
                for (int ptr = t->hRowFirstPtr[row]; ptr != -1; ptr = t->hRowNextPtr[ptr])
                    if (src[t->hCols[ptr]]) result ^= 1;
                if (result) flag = 1;
            }
            ++time; /* MyLdpc.cpp:1035-1039 */
            if (!flag) break;
            if (time == j->times) break;
            /* refreshQ, decodeCL.c:43-62 */
            for (int nodeInd = 0; nodeInd < nonZeros; ++nodeInd) {
                int hCol = t->hCols[nodeInd];
                float tmp0 = priorP0[hCol], tmp1 = priorP1[hCol];
                for (int ptr = t->hColFirstPtr[hCol]; ptr != -1; ptr = t->hColNextPtr[ptr]) {
                    if (nodeInd == ptr) continue;
                    tmp0 *= r0[ptr];
                    tmp1 *= r1[ptr];
                }
                q0[nodeInd] = tmp0 / (tmp0 + tmp1);
                q1[nodeInd] = tmp1 / (tmp0 + tmp1);
            }
        }
        if (j->info) { /* toChar, decodeCL.c:188-199 */
            uint8_t *o = j->info + (size_t)b * KB;
            memset(o, 0, (size_t)KB);
            for (int i = 0; i < K; ++i)
                if (src[i]) o[i >> 3] |= (uint8_t)(1u << (i & 7));
        }
        if (j->iters) j->iters[b] = time;
        if (j->hard) memcpy(j->hard + (size_t)b * ldpcN, src, (size_t)ldpcN);
        if (j->post0) memcpy(j->post0 + (size_t)b * ldpcN, pt0, sizeof(float) * (size_t)ldpcN);
        if (j->post1) memcpy(j->post1 + (size_t)b * ldpcN, pt1, sizeof(float) * (size_t)ldpcN);
    }
    free(q0); free(q1); free(r0); free(r1); free(priorP0); free(priorP1); free(pt0); free(pt1); free(src);
    return NULL;
}

int oracle_decode_sp_batch(const oracle_tables *t, int K, int times, const float *llr, int64_t ncw,
                           uint8_t *info, int32_t *iters, uint8_t *hard, float *post0, float *post1,
                           int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > ncw) nthreads = ncw > 0 ? (int)ncw : 1;
    spjob_t jobs[256];
    pthread_t th[256];
    for (int i = 0; i < nthreads; ++i) {
        spjob_t *j = &jobs[i];
        j->t = t; j->K = K; j->times = times; j->llr = llr;
        j->b0 = ncw * i / nthreads; j->b1 = ncw * (i + 1) / nthreads;
        j->info = info; j->iters = iters; j->hard = hard; j->post0 = post0; j->post1 = post1;
    }
    if (nthreads == 1) { spjob_run(&jobs[0]); return 0; }
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, spjob_run, &jobs[i]);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    return 0;
}

/* ---- layered (TDMP) min-sum restatement (decodeCL.c:203-292, MyLdpc.cpp:889-976) ----------------------
 * The reference's intent, per codeword, with its three defects repaired (see ldpc_oracle.h):
 *   init      lPostP = y, lR = 0, bits = 0                                   (decodeInitTDMP, :203-222)
 *   per layer (z consecutive rows), in row order:
 *     refreshQTDMP     lQ_e   = lPostP[col] - lR_e                            (:275-283)
 *     refreshRTDMP     lR_e   = (xor of the row's other (lQ<0)) ? -b : b,  b = fmin(1000, min |lQ|)  (:224-249)
 *     refreshPostPTDMP lPostP[col] = lQ_e + lR_e                              (:251-259)
 *   after the last layer: hardDecisionTDMP (>0 -> 0, <0 -> 1, ==0 keeps, :261-281), checkResult, ++time,
 *   stop when clean or time == times.                                         (MyLdpc.cpp:925-950)        */
typedef struct {
    const oracle_tables *t;
    int K, times, z;
    const float *llr;
    int64_t b0, b1;
    uint8_t *info, *hard;
    int32_t *iters;
    float *post;
} tdjob_t;

static void *tdjob_run(void *arg) {
    tdjob_t *j = (tdjob_t *)arg;
    const oracle_tables *t = j->t;
    const int nonZeros = t->nnz, ldpcN = t->N, ldpcM = t->M, K = j->K, KB = (K + 7) / 8, z = j->z;
    float *lQ = (float *)malloc(sizeof(float) * (size_t)nonZeros), *lR = (float *)malloc(sizeof(float) * (size_t)nonZeros);
    float *lPostP = (float *)malloc(sizeof(float) * (size_t)ldpcN);
    unsigned char *src = (unsigned char *)malloc((size_t)ldpcN);
    for (int64_t b = j->b0; b < j->b1; ++b) {
        const float *codes = j->llr + (size_t)b * ldpcN;
        for (int n = 0; n < ldpcN; ++n) { lPostP[n] = codes[n]; src[n] = 0; }
        for (int e = 0; e < nonZeros; ++e) lR[e] = 0;
        int time = 0;
        while (1) {
            for (int blockRow = 0; blockRow < ldpcM / z; ++blockRow) {
                const int e0 = t->hRowRange[blockRow * z], e1 = t->hRowRange[(blockRow + 1) * z];
                for (int e = e0; e < e1; ++e) lQ[e] = lPostP[t->hCols[e]] - lR[e];
                for (int e = e0; e < e1; ++e) {
                    int hRow = t->hRows[e];
                    int a = 0;
                    float bb = 1000;
                    for (int ptr = t->hRowFirstPtr[hRow]; ptr != -1; ptr = t->hRowNextPtr[ptr]) {
                        if (e == ptr) continue;
                        if (lQ[ptr] < 0) a ^= 1;
                        bb = fminf(bb, fabsf(lQ[ptr]));
                    }
                    lR[e] = a ? -bb : bb;
                }
                for (int e = e0; e < e1; ++e) lPostP[t->hCols[e]] = lQ[e] + lR[e];
            }
            for (int n = 0; n < ldpcN; ++n) {
                float tmp = lPostP[n];
                if (tmp > 0) src[n] = 0;
                else if (tmp < 0) src[n] = 1;
            }
            unsigned char flag = 0;
            for (int row = 0; row < ldpcM && !flag; ++row) {
                unsigned char result = 0;
                for (int ptr = t->hRowFirstPtr[row]; ptr != -1; ptr = t->hRowNextPtr[ptr])
                    if (src[t->hCols[ptr]]) result ^= 1;
                if (result) flag = 1;
            }
            ++time;
            if (!flag) break;
            if (time == j->times) break;
        }
        if (j->info) {
            uint8_t *o = j->info + (size_t)b * KB;
            memset(o, 0, (size_t)KB);
            for (int i = 0; i < K; ++i)
                if (src[i]) o[i >> 3] |= (uint8_t)(1u << (i & 7));
        }
        if (j->iters) j->iters[b] = time;
        if (j->hard) memcpy(j->hard + (size_t)b * ldpcN, src, (size_t)ldpcN);
        if (j->post) memcpy(j->post + (size_t)b * ldpcN, lPostP, sizeof(float) * (size_t)ldpcN);
    }
    free(lQ); free(lR); free(lPostP); free(src);
    return NULL;
}

int oracle_tdmp_layering_ok(const oracle_tables *t, int z) {
    if (z < 1 || t->M % z) return 0;
    int *seen = (int *)malloc(sizeof(int) * (size_t)t->N);
    for (int n = 0; n < t->N; ++n) seen[n] = -1;
    int ok = 1;
    for (int e = 0; e < t->nnz && ok; ++e) {
        int layer = t->hRows[e] / z;
        if (seen[t->hCols[e]] == layer) ok = 0;
        seen[t->hCols[e]] = layer;
    }
    free(seen);
    return ok;
}

int oracle_decode_tdmp_batch(const oracle_tables *t, int K, int times, int z, const float *llr, int64_t ncw,
                             uint8_t *info, int32_t *iters, uint8_t *hard, float *post, int nthreads) {
    if (!oracle_tdmp_layering_ok(t, z)) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > ncw) nthreads = ncw > 0 ? (int)ncw : 1;
    tdjob_t jobs[256];
    pthread_t th[256];
    for (int i = 0; i < nthreads; ++i) {
        tdjob_t *j = &jobs[i];
        j->t = t; j->K = K; j->times = times; j->z = z; j->llr = llr;
        j->b0 = ncw * i / nthreads; j->b1 = ncw * (i + 1) / nthreads;
        j->info = info; j->iters = iters; j->hard = hard; j->post = post;
    }
    if (nthreads == 1) { tdjob_run(&jobs[0]); return 0; }
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, tdjob_run, &jobs[i]);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    return 0;
}


/* ---- the two FUSED OpenCL kernels restated (decodeCL.c:307-426 decodeOnceTDMP, :432-567 decodeOnceMS) -----------
 * Both keep, per check row, the sign of every Q in lR and multiply it by  a * (second minimum for the edge that holds
 * the minimum, else the minimum)  where  a = sign(product of all Q of the row)  (a float product, taken in row order;
 * an exact zero or an underflowed product makes every message of the row zero, an overflow to inf keeps the sign, and
 * inf * 0 gives NaN -> sign 0).  The minimum search is  `if (t <= b) {c = b; b = t; ind = num;} else if (t > b && t <= c) c = t;`
 * from b = 1000, c = 1001 (decodeCL.c:346-365, 481-500), the hard decision is  bit = (lP < 0)  (:386, :541) -- a
 * zero posterior decodes as 0, where decodeCPU says 1 -- and the iteration caps are the literals 40 and 120 (:344, :479).
 * mode 0 = decodeOnceMS (flooding: all rows from the previous posterior, then lP = y + sum of lR in ascending row order,
 * :517-537); mode 1 = decodeOnceTDMP (layered: rows of one block row, lP = Q, then lP += lR, :346-380).
 * Pinned against the executed kernels (oracle/_ref/libmyldpc_refcl.so) in tests/test_oracle_vs_refcl.py.           */
static float cl_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? x : 0.0f)); }

typedef struct {
    const oracle_tables *t;
    int K, times, z, mode;
    const float *llr;
    int64_t b0, b1;
    uint8_t *info, *hard;
    int32_t *iters;
    float *post;
} fjob_t;

/* one check row: Q from lP and the old lR, new lR in place; layered mode also rewrites lP (decodeCL.c:346-380) */
static void fused_row(const oracle_tables *t, int row, float *lP, float *lR, int layered) {
    const int e0 = t->hRowRange[row], e1 = t->hRowRange[row + 1];
    float a = 1, b = 1000, c = 1001;
    int bInd = -1;
    for (int e = e0; e < e1; ++e) {
        float tmp = lP[t->hCols[e]] - lR[e];
        lR[e] = cl_sign(tmp);
        a *= tmp;
        if (layered) lP[t->hCols[e]] = tmp;
        tmp = fabsf(tmp);
        if (tmp <= b) { c = b; b = tmp; bInd = e; }
        else if (tmp > b && tmp <= c) { c = tmp; }
    }
    a = cl_sign(a);
    for (int e = e0; e < e1; ++e) {
        if (e == bInd) lR[e] *= a * c;
        else lR[e] *= a * b;
    }
    if (layered)
        for (int e = e0; e < e1; ++e) lP[t->hCols[e]] += lR[e];
}

static void *fjob_run(void *arg) {
    fjob_t *j = (fjob_t *)arg;
    const oracle_tables *t = j->t;
    const int nonZeros = t->nnz, ldpcN = t->N, ldpcM = t->M, K = j->K, KB = (K + 7) / 8, z = j->z;
    float *lR = (float *)malloc(sizeof(float) * (size_t)nonZeros), *lP = (float *)malloc(sizeof(float) * (size_t)ldpcN);
    unsigned char *src = (unsigned char *)malloc((size_t)ldpcN);
    for (int64_t b = j->b0; b < j->b1; ++b) {
        const float *codes = j->llr + (size_t)b * ldpcN;
        for (int n = 0; n < ldpcN; ++n) lP[n] = codes[n];
        for (int e = 0; e < nonZeros; ++e) lR[e] = 0;
        int time = 0;
        while (1) {
            if (j->mode == 1) {
                for (int layer = 0; layer < ldpcM / z; ++layer)
                    for (int r = layer * z; r < (layer + 1) * z; ++r) fused_row(t, r, lP, lR, 1);
            } else {
                for (int r = 0; r < ldpcM; ++r) fused_row(t, r, lP, lR, 0);
                for (int n = 0; n < ldpcN; ++n) { /* decodeCL.c:517-537: ascending seed row = ascending edge id */
                    float tmp = codes[n];
                    for (int ptr = t->hColFirstPtr[n]; ptr != -1; ptr = t->hColNextPtr[ptr]) tmp += lR[ptr];
                    lP[n] = tmp;
                }
            }
            for (int n = 0; n < ldpcN; ++n) src[n] = lP[n] < 0;
            unsigned char flag = 0;
            for (int row = 0; row < ldpcM && !flag; ++row) {
                unsigned char result = 0;
                for (int e = t->hRowRange[row]; e < t->hRowRange[row + 1]; ++e) result ^= src[t->hCols[e]];
                if (result) flag = 1;
            }
            ++time;
            if (!flag) break;
            if (time == j->times) break;
        }
        if (j->info) {
            uint8_t *o = j->info + (size_t)b * KB;
            memset(o, 0, (size_t)KB);
            for (int i = 0; i < K; ++i)
                if (src[i]) o[i >> 3] |= (uint8_t)(1u << (i & 7));
        }
        if (j->iters) j->iters[b] = time;
        if (j->hard) memcpy(j->hard + (size_t)b * ldpcN, src, (size_t)ldpcN);
        if (j->post) memcpy(j->post + (size_t)b * ldpcN, lP, sizeof(float) * (size_t)ldpcN);
    }
    free(lR); free(lP); free(src);
    return NULL;
}

int oracle_decode_fused_batch(const oracle_tables *t, int K, int times, int z, int mode, const float *llr, int64_t ncw,
                              uint8_t *info, int32_t *iters, uint8_t *hard, float *post, int nthreads) {
    if (mode == 1 && !oracle_tdmp_layering_ok(t, z)) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > ncw) nthreads = ncw > 0 ? (int)ncw : 1;
    fjob_t jobs[256];
    pthread_t th[256];
    for (int i = 0; i < nthreads; ++i) {
        fjob_t *j = &jobs[i];
        j->t = t; j->K = K; j->times = times; j->z = z; j->mode = mode; j->llr = llr;
        j->b0 = ncw * i / nthreads; j->b1 = ncw * (i + 1) / nthreads;
        j->info = info; j->iters = iters; j->hard = hard; j->post = post;
    }
    if (nthreads == 1) { fjob_run(&jobs[0]); return 0; }
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, fjob_run, &jobs[i]);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    return 0;
}

/* reference MyLdpc.cpp:1063-1072: bit (LSB first) 1 -> -1.0, 0 -> +1.0 */
void oracle_bpsk(const uint8_t *bytes, int nbytes, float *out) {
    for (int charOffset = 0; charOffset < nbytes; ++charOffset) {
        uint8_t tmp = bytes[charOffset];
        for (int bitOffset = 0; bitOffset < 8; ++bitOffset) {
            out[charOffset * 8 + bitOffset] = (tmp & (1 << bitOffset)) ? -1.0f : 1.0f;
        }
    }
}
