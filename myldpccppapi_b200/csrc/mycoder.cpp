// mycoder.cpp -- `class Coder` (include/MyLdpc.h) over the C-ABI of libldpc_b200.so.
//
// Host-side mirror of the reference's Coder (MyLdpc.cpp): same call order, same size rules, same
// stream layout.  What the reference did with cl::Context / cl::Buffer / cl::Kernel in forDecoder,
// addDecodeType and decodeOnce* is one ldpc_b200_* call here; the decode itself runs on the GPU(s).
#include "MyLdpc.h"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <string>
#include <thread>

#include "ldpc_b200.h"

struct Coder::Impl {
    int K = 0, N = 0, M = 0;
    int rate = -1;       // rate_type, or -1 for a CSR-constructed code
    int times = 40;      // reference MyLdpc.cpp:24
    int timesMSCL = 120; // the cap decodeOnceMS hard-codes (decodeCL.c:479); setMaxIter() overrides both
    bool strict = false; // setStrictDecodeType
    bool fusedExact = false;  // setFusedKernelArithmetic
    bool registerHost = false;  // setRegisterHostBuffers
    int lastAlg = -1;
    double stepTimes[4] = {0, 0, 0, 0};
    int batchSize = 0;
    bool early = true;
    std::vector<int> devices{0};
    std::vector<ldpc_b200_handle> handles;
    // result staging of decode(): page-locked when the runtime grants it (true DMA for the read-back), grown on demand
    int32_t *iters = nullptr;
    uint8_t *info = nullptr;
    size_t itersCap = 0, infoCap = 0;
    bool itersPinned = false, infoPinned = false;
    int lastCodeSize = 0;
    std::string err;
    // systematic encoder: parity = X u over GF(2); X stored row-wise, 64 info bits per word
    bool encoderReady = false;
    int kWords = 0;
    std::vector<uint64_t> X;  // [M][kWords]

    int fail(int code) {
        err = ldpc_b200_last_error();
        return code;
    }
    template <class T>
    static bool grow(T *&p, size_t &cap, bool &pinned, size_t n) {
        if (n <= cap) return true;
        if (p) { if (pinned) ldpc_b200_host_free(p); else std::free(p); p = nullptr; cap = 0; }
        void *q = ldpc_b200_host_alloc(n * sizeof(T));
        pinned = q != nullptr;
        if (!q) q = std::malloc(n * sizeof(T));
        if (!q) return false;
        std::memset(q, 0, n * sizeof(T));  // touched here, not by the first decode()
        p = static_cast<T *>(q);
        cap = n;
        return true;
    }
    void freeStaging() {
        if (iters) { if (itersPinned) ldpc_b200_host_free(iters); else std::free(iters); iters = nullptr; itersCap = 0; }
        if (info) { if (infoPinned) ldpc_b200_host_free(info); else std::free(info); info = nullptr; infoCap = 0; }
    }
    void destroyHandles() {
        for (ldpc_b200_handle h : handles) ldpc_b200_destroy(h);
        handles.clear();
    }
};

// ---- construction -----------------------------------------------------------------------------
Coder::Coder(int ldpcK, int ldpcN, enum rate_type rate) : kernelSourceCode(nullptr), impl(new Impl) {
    impl->K = ldpcK;
    impl->N = ldpcN;
    impl->M = ldpcN - ldpcK;
    impl->rate = (int)rate;
    int M = 0, nnz = 0;
    if (ldpc_b200_wimax_csr(ldpcK, ldpcN, (int)rate, nullptr, nullptr, &M, &nnz) != LDPC_B200_OK) {
        // the reference does not validate its arguments either; keep the object usable for lastError()
        impl->err = ldpc_b200_last_error();
        return;
    }
    checkMatrix.rows_ = M;
    checkMatrix.cols_ = ldpcN;
    checkMatrix.ptr_.resize(M + 1);
    checkMatrix.col_.resize(nnz);
    ldpc_b200_wimax_csr(ldpcK, ldpcN, (int)rate, checkMatrix.ptr_.data(), checkMatrix.col_.data(), &M, &nnz);
}

Coder::Coder(int ldpcM, int ldpcN, int ldpcK, const int *rowPtr, const int *colIdx)
    : kernelSourceCode(nullptr), impl(new Impl) {
    impl->K = ldpcK;
    impl->N = ldpcN;
    impl->M = ldpcM;
    checkMatrix.rows_ = ldpcM;
    checkMatrix.cols_ = ldpcN;
    checkMatrix.ptr_.assign(rowPtr, rowPtr + ldpcM + 1);
    checkMatrix.col_.assign(colIdx, colIdx + rowPtr[ldpcM]);
}

Coder::~Coder() {
    impl->destroyHandles();
    impl->freeStaging();
    delete impl;
}

// ---- [B200] configuration ---------------------------------------------------------------------
int Coder::setMaxIter(int times) {
    if (times < 1) return LDPC_B200_ERR_ARG;
    impl->times = times;
    impl->timesMSCL = times;
    for (ldpc_b200_handle h : impl->handles)
        if (ldpc_b200_set_max_iter(h, times) != LDPC_B200_OK) return impl->fail(LDPC_B200_ERR_ARG);
    return LDPC_SUCCESS;
}

int Coder::setEarlyTermination(bool on) {
    impl->early = on;
    for (ldpc_b200_handle h : impl->handles) ldpc_b200_set_early_termination(h, on ? 1 : 0);
    return LDPC_SUCCESS;
}

int Coder::setDevices(const int *deviceIds, int count) {
    if (!deviceIds || count < 1) return LDPC_B200_ERR_ARG;
    if (!impl->handles.empty()) {
        impl->err = "setDevices must be called before forDecoder";
        return LDPC_B200_ERR_ARG;
    }
    impl->devices.assign(deviceIds, deviceIds + count);
    return LDPC_SUCCESS;
}

int Coder::setStrictDecodeType(bool strict) {
    impl->strict = strict;
    return LDPC_SUCCESS;
}

int Coder::setFusedKernelArithmetic(bool exact) {
    impl->fusedExact = exact;
    return LDPC_SUCCESS;
}

int Coder::setRegisterHostBuffers(bool on) {
    impl->registerHost = on;
    for (ldpc_b200_handle h : impl->handles) ldpc_b200_set_option(h, "register_host", on ? 1 : 0);
    return LDPC_SUCCESS;
}

int Coder::lastAlgorithm() const { return impl->lastAlg; }

int Coder::lastStepTimes(double *seconds, int n) const {
    if (!seconds || n < 1) return 0;
    const int m = n < 4 ? n : 4;
    for (int i = 0; i < m; ++i) seconds[i] = impl->stepTimes[i];
    return m;
}

const int *Coder::lastIterations() const { return impl->iters; }
int Coder::lastCodeSize() const { return impl->lastCodeSize; }
const char *Coder::lastError() const { return impl->err.c_str(); }

// ---- decoder setup (reference MyLdpc.cpp:167-305, 307-552) ------------------------------------
int Coder::forDecoder(int batchSize) {
    impl->batchSize = batchSize;
    impl->destroyHandles();
    if (checkMatrix.rows_ == 0) return LDPC_B200_ERR_ARG;
    for (int dev : impl->devices) {
        ldpc_b200_handle h = nullptr;
        int rc = ldpc_b200_create(&h, impl->M, impl->N, impl->K, checkMatrix.ptr_.data(), checkMatrix.col_.data(), dev);
        if (rc != LDPC_B200_OK) {
            impl->fail(rc);
            impl->destroyHandles();
            return rc;
        }
        ldpc_b200_set_max_iter(h, impl->times);
        ldpc_b200_set_early_termination(h, impl->early ? 1 : 0);
        ldpc_b200_set_option(h, "register_host", impl->registerHost ? 1 : 0);
        if (impl->devices.size() > 1) {
            // every GPU's handle stages its shard of a pageable postCode with its own host threads: share the cores
            const long long hw = std::max(1u, std::thread::hardware_concurrency());
            ldpc_b200_set_option(h, "stage_threads", std::max(2LL, std::min(6LL, hw / (long long)impl->devices.size())));
        }
        if (impl->rate >= 0) ldpc_b200_set_layer_height(h, impl->N / 24);  // z, reference MyLdpc.cpp:22
        impl->handles.push_back(h);
    }
    return LDPC_SUCCESS;
}

int Coder::addDecodeType(enum decodeType deType) {
    // every kernel is precompiled and decode() picks one per call from its own deType; what happens here is what the
    // reference does here (it builds and allocates, MyLdpc.cpp:387-437): launch buffers for batchSize words, and one
    // launch of this deType's kernels so that the first decode() finds their code loaded on the device
    if (impl->handles.empty()) {
        impl->err = "forDecoder must be called before addDecodeType";
        return LDPC_B200_ERR_ARG;
    }
    if (impl->batchSize > 0) {
        int alg = deType == DecodeSP ? LDPC_B200_ALG_SUM_PRODUCT
                  : (deType == DecodeTDMP || deType == DecodeTDMPCL) ? LDPC_B200_ALG_LAYERED_MIN_SUM
                                                                     : LDPC_B200_ALG_MIN_SUM;
        if (impl->fusedExact && deType == DecodeMSCL) alg = LDPC_B200_ALG_FUSED_MIN_SUM;
        if (impl->fusedExact && deType == DecodeTDMPCL) alg = LDPC_B200_ALG_FUSED_LAYERED;
        auto setup = [&](ldpc_b200_handle h) -> int {
            int rc = ldpc_b200_reserve(h, impl->batchSize);  // min-sum: the host-buffer pipeline's buffers too
            if (rc == LDPC_B200_OK && alg != LDPC_B200_ALG_MIN_SUM) {
                if (ldpc_b200_set_algorithm(h, alg) == LDPC_B200_OK) rc = ldpc_b200_reserve(h, impl->batchSize);
                ldpc_b200_set_algorithm(h, LDPC_B200_ALG_MIN_SUM);  // (a code the kernel cannot hold: decode() reports it)
            }
            return rc;
        };
        // one host thread per device: a CUDA context loads kernels for its own device only
        std::vector<int> rcs(impl->handles.size(), LDPC_B200_OK);
        if (impl->handles.size() == 1) rcs[0] = setup(impl->handles[0]);
        else {
            std::vector<std::thread> th;
            for (size_t g = 0; g < impl->handles.size(); ++g) th.emplace_back([&, g]() { rcs[g] = setup(impl->handles[g]); });
            for (auto &t : th) t.join();
        }
        for (int rc : rcs)
            if (rc != LDPC_B200_OK) return impl->fail(rc);
    }
    if (impl->batchSize > 0) {  // result staging for a call of batchSize words: the first decode() allocates nothing
        Impl::grow(impl->iters, impl->itersCap, impl->itersPinned, (size_t)impl->batchSize);
        Impl::grow(impl->info, impl->infoCap, impl->infoPinned, (size_t)impl->batchSize * ((impl->K + 7) / 8));
    }
    return LDPC_SUCCESS;
}

int Coder::forTest() { return LDPC_SUCCESS; }  // declared but never defined by the reference (MyLdpc.h:112)

// ---- size helpers (reference MyLdpc.cpp:620-631) ----------------------------------------------
int Coder::getPriorCodeLength(int srcLength) {
    return (srcLength + (impl->K / 8) - 1) / (impl->K / 8) * (impl->N / 8);
}
int Coder::getPostCodeLength(int srcLength) { return (srcLength + (impl->K / 8) - 1) / (impl->K / 8) * impl->N; }
int Coder::getCodeSize(int srcLength) { return (srcLength + (impl->K / 8) - 1) / (impl->K / 8); }

// ---- decode (reference MyLdpc.cpp:571-618; semantics of :684-784) -----------------------------
int Coder::decode(float *postCode, char *srcCode, int srcLength, enum decodeType deType) {
    if (impl->handles.empty()) {
        impl->err = "forDecoder must be called before decode";
        return LDPC_B200_ERR_ARG;
    }
    const int codeSize = getCodeSize(srcLength);
    const int KB = (impl->K + 7) / 8;
    const int G = (int)impl->handles.size();
    // DecodeSP runs the probability-domain sum-product kernel (decodeCL.c:3-108 semantics) and DecodeTDMP /
    // DecodeTDMPCL the layered min-sum kernel (the schedule of decodeCL.c:203-292) where the code fits their
    // on-chip layouts; every other decodeType -- and those two on codes that do not fit -- runs flooding
    // min-sum with Coder::decodeCPU semantics.
    int alg = deType == DecodeSP ? LDPC_B200_ALG_SUM_PRODUCT
              : (deType == DecodeTDMP || deType == DecodeTDMPCL) ? LDPC_B200_ALG_LAYERED_MIN_SUM
                                                                 : LDPC_B200_ALG_MIN_SUM;
    if (impl->fusedExact && deType == DecodeMSCL) alg = LDPC_B200_ALG_FUSED_MIN_SUM;      // decodeOnceMS's own arithmetic
    if (impl->fusedExact && deType == DecodeTDMPCL) alg = LDPC_B200_ALG_FUSED_LAYERED;    // decodeOnceTDMP's own arithmetic
    // DecodeMSCL: the reference's fused kernel iterates up to 120 times whatever Coder::times says (decodeCL.c:479)
    const int cap = deType == DecodeMSCL ? impl->timesMSCL : impl->times;
    const char *fallback_note = "requested decodeType does not fit its on-chip layout for this code, decoded with flooding min-sum";
    bool fell_back = false;
    for (ldpc_b200_handle h : impl->handles) {
        ldpc_b200_set_max_iter(h, cap);
        if (ldpc_b200_set_algorithm(h, alg) != LDPC_B200_OK) fell_back = true;
    }
    if (fell_back) {
        if (impl->strict) {
            impl->fail(LDPC_B200_ERR_UNSUPPORTED);
            for (ldpc_b200_handle h : impl->handles) ldpc_b200_set_algorithm(h, LDPC_B200_ALG_MIN_SUM);
            return LDPC_B200_ERR_UNSUPPORTED;
        }
        alg = LDPC_B200_ALG_MIN_SUM;
        for (ldpc_b200_handle h : impl->handles) ldpc_b200_set_algorithm(h, alg);
    }
    if (!Impl::grow(impl->iters, impl->itersCap, impl->itersPinned, (size_t)codeSize) ||
        !Impl::grow(impl->info, impl->infoCap, impl->infoPinned, (size_t)codeSize * KB)) {
        impl->err = "out of host memory";
        return LDPC_B200_ERR_NOMEM;
    }
    impl->lastCodeSize = codeSize;
    std::vector<int> rcs(G, LDPC_B200_OK);
    std::vector<std::string> msgs(G);
    for (ldpc_b200_handle h : impl->handles) ldpc_b200_reset_timing(h);
    const auto wall0 = std::chrono::steady_clock::now();
    auto work = [&](int g) {
        // contiguous shard [b, e) of the codewords for device g: no data-path collective
        const int64_t b = (int64_t)codeSize * g / G, e = (int64_t)codeSize * (g + 1) / G;
        if (e <= b) return;
        rcs[g] = ldpc_b200_decode_host(impl->handles[g], postCode + (size_t)b * impl->N, e - b,
                                       impl->info + (size_t)b * KB, nullptr, impl->iters + b, nullptr);
        if (rcs[g] != LDPC_B200_OK) msgs[g] = ldpc_b200_last_error();
    };
    auto run_all = [&]() {
        if (G == 1) { work(0); return; }
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) th.emplace_back(work, g);
        for (auto &t : th) t.join();
    };
    run_all();
    bool unsupported = false;
    for (int g = 0; g < G; ++g)
        if (rcs[g] == LDPC_B200_ERR_UNSUPPORTED && alg != LDPC_B200_ALG_MIN_SUM) unsupported = true;
    if (unsupported && !impl->strict) {  // the kernel of this decodeType cannot hold the code: decode with min-sum instead, and say so
        fell_back = true;
        alg = LDPC_B200_ALG_MIN_SUM;
        for (ldpc_b200_handle h : impl->handles) ldpc_b200_set_algorithm(h, LDPC_B200_ALG_MIN_SUM);
        std::fill(rcs.begin(), rcs.end(), LDPC_B200_OK);
        run_all();
    }
    impl->stepTimes[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    impl->stepTimes[1] = impl->stepTimes[2] = impl->stepTimes[3] = 0.0;
    for (ldpc_b200_handle h : impl->handles) {  // the busiest device per phase
        ldpc_b200_timing tm;
        if (ldpc_b200_get_timing(h, &tm) != LDPC_B200_OK) continue;
        impl->stepTimes[1] = std::max(impl->stepTimes[1], tm.h2d_s);
        impl->stepTimes[2] = std::max(impl->stepTimes[2], tm.kernel_s);
        impl->stepTimes[3] = std::max(impl->stepTimes[3], tm.d2h_s);
    }
    impl->lastAlg = alg;
    if (fell_back) impl->err = fallback_note;
    for (int g = 0; g < G; ++g)
        if (rcs[g] != LDPC_B200_OK) {
            impl->err = msgs[g];
            return rcs[g];
        }
    // the stream holds the first srcLength bytes (the last codeword may be partly padding)
    std::memcpy(srcCode, impl->info, std::min<size_t>((size_t)srcLength, (size_t)codeSize * KB));
    return LDPC_SUCCESS;
}

// ---- encoder (reference MyLdpc.cpp:137-165, 554-569, 633-682) ---------------------------------
// The reference computes the parity bits with a Richardson-Urbanke split and dense GF(2) inverses
// through Eigen.  The parity part of H is invertible, so the parity bits are uniquely determined by
// H c = 0; we solve [B | A] once by Gauss-Jordan over GF(2) (B = last M columns) and keep X = B^-1 A.
int Coder::forEncoder() {
    const int M = impl->M, N = impl->N, K = impl->K;
    if (N - K != M || checkMatrix.rows_ != M) {
        impl->err = "forEncoder needs a square parity part (N - K == M)";
        return LDPC_B200_ERR_ARG;
    }
    const int W = (N + 63) / 64;
    std::vector<uint64_t> a((size_t)M * W, 0);  // row r: columns permuted to [parity (M) | info (K)]
    auto setbit = [&](int r, int c) { a[(size_t)r * W + (c >> 6)] |= 1ull << (c & 63); };
    auto getbit = [&](int r, int c) { return (a[(size_t)r * W + (c >> 6)] >> (c & 63)) & 1ull; };
    for (int r = 0; r < M; ++r)
        for (int e = checkMatrix.ptr_[r]; e < checkMatrix.ptr_[r + 1]; ++e) {
            const int c = checkMatrix.col_[e];
            setbit(r, c >= K ? c - K : M + c);
        }
    for (int c = 0; c < M; ++c) {
        int p = -1;
        for (int r = c; r < M; ++r)
            if (getbit(r, c)) { p = r; break; }
        if (p < 0) {
            impl->err = "parity part of H is singular";
            return LDPC_B200_ERR_ARG;
        }
        if (p != c)
            for (int w = 0; w < W; ++w) std::swap(a[(size_t)p * W + w], a[(size_t)c * W + w]);
        for (int r = 0; r < M; ++r)
            if (r != c && getbit(r, c))
                for (int w = 0; w < W; ++w) a[(size_t)r * W + w] ^= a[(size_t)c * W + w];
    }
    impl->kWords = (K + 63) / 64;
    impl->X.assign((size_t)M * impl->kWords, 0);
    for (int r = 0; r < M; ++r)
        for (int k = 0; k < K; ++k)
            if (getbit(r, M + k)) impl->X[(size_t)r * impl->kWords + (k >> 6)] |= 1ull << (k & 63);
    impl->encoderReady = true;
    return LDPC_SUCCESS;
}

int Coder::encode(char *srcCode, char *priorCode, int srcLength) {
    if (!impl->encoderReady) {
        impl->err = "forEncoder must be called before encode";
        return LDPC_B200_ERR_ARG;
    }
    const int K = impl->K, N = impl->N, M = impl->M, kb = K / 8, nb = N / 8;
    const int codeSize = getCodeSize(srcLength);
    std::vector<uint64_t> u(impl->kWords);
    for (int cw = 0; cw < codeSize; ++cw) {
        const int srcL = std::min(kb, srcLength - cw * kb);  // the last word may be short (MyLdpc.cpp:561-564)
        unsigned char *out = reinterpret_cast<unsigned char *>(priorCode) + (size_t)cw * nb;
        std::memset(out, 0, nb);
        std::memcpy(out, srcCode + (size_t)cw * kb, srcL);     // systematic part, LSB-first bits
        std::fill(u.begin(), u.end(), 0);
        for (int i = 0; i < srcL; ++i) u[i >> 3] |= (uint64_t)out[i] << ((i & 7) * 8);
        for (int r = 0; r < M; ++r) {
            uint64_t acc = 0;
            const uint64_t *x = &impl->X[(size_t)r * impl->kWords];
            for (int w = 0; w < impl->kWords; ++w) acc ^= x[w] & u[w];
            if (__builtin_popcountll(acc) & 1) {
                const int bit = K + r;
                out[bit >> 3] |= (unsigned char)(1u << (bit & 7));
            }
        }
    }
    return LDPC_SUCCESS;
}

// ---- channel (reference MyLdpc.cpp:1061-1078, 1093-1105) --------------------------------------
int Coder::test(char *priorCode, float *postCode, int priorCodeLength, float rate) {
    for (int charOffset = 0; charOffset < priorCodeLength; ++charOffset) {
        const char tmp = priorCode[charOffset];
        for (int bitOffset = 0; bitOffset < 8; ++bitOffset)
            postCode[charOffset * 8 + bitOffset] = (tmp & (1 << bitOffset)) ? -1.0f : 1.0f;
    }
    for (int i = 0; i < priorCodeLength * 8; ++i) postCode[i] += gaussian(0, rate);
    return LDPC_SUCCESS;
}

// Box-Muller, cosine branch only, two rand() draws per sample -- the reference's generator.
float gaussian(float ave, float sd) {
    const float pi = 3.1415926f;
    const float s1 = (float)((1.0 + rand()) / (RAND_MAX + 1.0));
    const float s2 = (float)((1.0 + rand()) / (RAND_MAX + 1.0));
    const float r = std::sqrt(-2 * std::log(s2));
    const float t = 2 * pi * s1;
    const float z = r * std::cos(t);
    return ave + z * sd;
}

// Kept for link compatibility: the reference read its OpenCL kernel file with this (MyLdpc.cpp:1079-1091).
char *load_program_source(const char *filename) {
    FILE *fh = std::fopen(filename, "r");
    if (!fh) return nullptr;
    std::fseek(fh, 0, SEEK_END);
    const long n = std::ftell(fh);
    std::fseek(fh, 0, SEEK_SET);
    char *source = (char *)std::malloc((size_t)n + 1);
    const size_t got = std::fread(source, 1, (size_t)n, fh);
    source[got] = '\0';
    std::fclose(fh);
    return source;
}

// ---- C doorway (include/MyLdpc_c.h): the same Coder for callers without a C++ compiler -- the Python binding in
// myldpccppapi_b200/decoder.py and the parity tests use it, so that the class above is the ONE implementation.
#include "MyLdpc_c.h"

extern "C" {
myldpc_coder *myldpc_coder_new(int ldpcK, int ldpcN, int rate) { return reinterpret_cast<myldpc_coder *>(new Coder(ldpcK, ldpcN, (enum rate_type)rate)); }
myldpc_coder *myldpc_coder_new_csr(int ldpcM, int ldpcN, int ldpcK, const int *rowPtr, const int *colIdx) {
    return reinterpret_cast<myldpc_coder *>(new Coder(ldpcM, ldpcN, ldpcK, rowPtr, colIdx));
}
void myldpc_coder_free(myldpc_coder *c) { delete reinterpret_cast<Coder *>(c); }
#define CODER(c) reinterpret_cast<Coder *>(c)
int myldpc_forEncoder(myldpc_coder *c) { return CODER(c)->forEncoder(); }
int myldpc_forDecoder(myldpc_coder *c, int batchSize) { return CODER(c)->forDecoder(batchSize); }
int myldpc_addDecodeType(myldpc_coder *c, int deType) { return CODER(c)->addDecodeType((enum decodeType)deType); }
int myldpc_encode(myldpc_coder *c, char *srcCode, char *priorCode, int srcLength) { return CODER(c)->encode(srcCode, priorCode, srcLength); }
int myldpc_decode(myldpc_coder *c, float *postCode, char *srcCode, int srcLength, int deType) {
    return CODER(c)->decode(postCode, srcCode, srcLength, (enum decodeType)deType);
}
int myldpc_test(myldpc_coder *c, char *priorCode, float *postCode, int priorCodeLength, float sd) { return CODER(c)->test(priorCode, postCode, priorCodeLength, sd); }
int myldpc_getPriorCodeLength(myldpc_coder *c, int srcLength) { return CODER(c)->getPriorCodeLength(srcLength); }
int myldpc_getPostCodeLength(myldpc_coder *c, int srcLength) { return CODER(c)->getPostCodeLength(srcLength); }
int myldpc_getCodeSize(myldpc_coder *c, int srcLength) { return CODER(c)->getCodeSize(srcLength); }
int myldpc_checkMatrix(myldpc_coder *c, int *rows, int *cols, int *nnz, int *rowPtr, int *colIdx) {
    const CheckMatrixCSR &m = CODER(c)->checkMatrix;
    if (rows) *rows = m.rows();
    if (cols) *cols = m.cols();
    if (nnz) *nnz = m.nonZeros();
    if (rowPtr) std::memcpy(rowPtr, m.outerIndexPtr(), sizeof(int) * (size_t)(m.rows() + 1));
    if (colIdx) std::memcpy(colIdx, m.innerIndexPtr(), sizeof(int) * (size_t)m.nonZeros());
    return LDPC_SUCCESS;
}
int myldpc_setMaxIter(myldpc_coder *c, int times) { return CODER(c)->setMaxIter(times); }
int myldpc_setDevices(myldpc_coder *c, const int *deviceIds, int count) { return CODER(c)->setDevices(deviceIds, count); }
int myldpc_setEarlyTermination(myldpc_coder *c, int on) { return CODER(c)->setEarlyTermination(on != 0); }
int myldpc_setStrictDecodeType(myldpc_coder *c, int strict) { return CODER(c)->setStrictDecodeType(strict != 0); }
int myldpc_setFusedKernelArithmetic(myldpc_coder *c, int exact) { return CODER(c)->setFusedKernelArithmetic(exact != 0); }
int myldpc_setRegisterHostBuffers(myldpc_coder *c, int on) { return CODER(c)->setRegisterHostBuffers(on != 0); }
int myldpc_lastAlgorithm(myldpc_coder *c) { return CODER(c)->lastAlgorithm(); }
const int *myldpc_lastIterations(myldpc_coder *c) { return CODER(c)->lastIterations(); }
int myldpc_lastCodeSize(myldpc_coder *c) { return CODER(c)->lastCodeSize(); }
const char *myldpc_lastError(myldpc_coder *c) { return CODER(c)->lastError(); }
int myldpc_lastStepTimes(myldpc_coder *c, double *seconds, int n) { return CODER(c)->lastStepTimes(seconds, n); }
#undef CODER
}
