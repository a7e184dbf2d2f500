// ldpc_b200.cu -- C-ABI (include/ldpc_b200.h) over the sm_100a kernels in ldpc_kernels.cuh.
// Host side of what the reference does with cl::Context/CommandQueue/Buffer/Kernel in
// Coder::forDecoder, addDecodeType and decodeOnceMS (MyLdpc.cpp:224-305, 387-437, 786-848).
#include "../../include/ldpc_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include "ldpc_kernels.cuh"
#include "ldpc_cluster.cuh"
#include "ldpc_qc.cuh"
#include "ldpc_qcg.cuh"
#include "ldpc_qcw.cuh"
#include "ldpc_qcm.cuh"
#include "ldpc_warp.cuh"
#include "ldpc_sp.cuh"
#include "ldpc_tdmp.cuh"
#include "ldpc_big.cuh"
#include "ldpc_stream.cuh"
#include "ldpc_encode.cuh"
#include "ldpc_tables.h"
#include "ldpc_launch.h"

using namespace ldpc_b200;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            return fail(LDPC_B200_ERR_CUDA,                                                       \
                        std::string(#expr) + ": " + cudaGetErrorName(e__) + " (" + cudaGetErrorString(e__) + ")"); \
        }                                                                                         \
    } while (0)

constexpr int kSlots = 3;          // rotating streams of the host-buffer pipeline
constexpr int kCounterRing = 256;  // one work-queue head per in-flight launch

// Experiment switches.  Seeded ONCE per handle from the LDPC_B200_* environment variables in
// ldpc_b200_create; the host-pipeline ones can be changed afterwards with ldpc_b200_set_option.  Nothing on the
// decode path calls getenv.
struct Options {
    // kernel choice: read when a plan is made
    bool no_qc = false, no_qcg = false, no_qcm = false, qcm_always = false, qc_generic = false, qc_ring = false;
    bool grp_no_profile = false, grp_no_ysmem = false, grp_prefer_16 = false, grp_t16 = false, grp_no_t16 = false;
    int sp_qc = -1;   // quasi-cyclic sum-product kernel (ldpc_spq.cuh): -1 = where the on-chip kernel does not fit, 0 never, 1 wherever it can run
    bool debug_placement = false, sp_big = false;  // sp_big: sum-product through the any-size kernel even where the on-chip one fits (tests)
    int grp_g = 0, grp_warps = 0, l16_warps = 0, tdmp_g = 0, stream_threads = 0, qc_prefer_g = 0;
    long long place_effort = 12;
    // launch / host pipeline: read per call
    int refill_wait = 1;              // measured: profiles/r01_refill_sweep.txt
    bool no_streamed = false, streamed_pageable = false, no_staged = false;
    bool avail_memcpy = false;        // announce streamed chunks by an 8-byte copy instead of a stream memory operation
    bool no_warm = false;             // skip the first-launch warm-up of ldpc_b200_reserve (measurements of the cold first call)
    bool register_host = false;       // page-lock a pageable input buffer on first sight (cudaHostRegister) and keep it registered
    long long staged_min_kb = 8 << 10;  // pageable input of 8 MB or more is staged by host threads
    long long stream_chunk = 0;       // words per input chunk (0 = 1, 2, then 4 MB)
    long long stream_batch_kb = 0;    // channel values per launch (0 = default)
    long long wait_timeout_ms = 4000; // bound of the kernel's wait for streamed input
    int stage_threads = 6;            // host threads that stage a pageable input buffer through the pinned ring (6-12 measured alike
                                      // with streaming stores, profiles/r02_stage_nt.txt; 4 was the optimum of the plain memcpy)
    int chunk_stage = 1;              // chunked pipeline: a pageable caller buffer is staged by host threads too (0: the driver stages it)
    int stage_nt = 1;                 // ... with non-temporal stores (stage_copy_nt, ldpc_tables.h); 0 = plain memcpy
    // early-termination kernel of the quasi-cyclic path (ldpc_qcw.cuh, a warp per codeword): -1 = chosen per launch from the
    // mean iteration count of the handle's previous launches, 0 = never, 1 = whenever it can run
    int qc_et = -1;
    int qc_et_every = 4;              // auto: the iteration counts are sampled after every n-th launch once the regime is known (13 us each)
    int qcw_warps = 0;                // warp-per-codeword kernel: fewer codewords in flight per SM than fit (occupancy experiments)
    int qcm_pack = -1;                // group-of-warps kernel, codewords per group: -1 = the measured best of the block size and rate, 0 / 1 = one, 2 / 3 = that many
    int qcm_multi_pct = -1;           // ... used while the mean iteration count is above this share of the cap (-1 = measured crossover, 0 = always)
    int qc_et_pct = 0;                // auto: used while the mean iteration count is at most this share of the cap; 0 = the measured
                                      // crossover of the code (profiles/r02_et_kernel.md: z = 24 97 %, z = 32 every regime)
};

struct OptionName { const char* name; int kind; size_t off; bool runtime; };  // kind 0 bool, 1 int, 2 long long
#define OPT(n, k) {#n, k, offsetof(Options, n), false}   // shapes the plan: environment only, read at create
#define OPTR(n, k) {#n, k, offsetof(Options, n), true}   // may change afterwards (ldpc_b200_set_option)
const OptionName kOptionNames[] = {
    OPT(no_qc, 0), OPT(no_qcg, 0), OPT(no_qcm, 0), OPT(qcm_always, 0), OPT(qc_generic, 0), OPT(qc_ring, 0), OPT(grp_no_profile, 0), OPT(grp_no_ysmem, 0), OPT(grp_prefer_16, 0),
    OPT(grp_t16, 0), OPT(grp_no_t16, 0), OPT(debug_placement, 0), OPT(sp_big, 0), OPT(grp_g, 1), OPT(grp_warps, 1), OPT(l16_warps, 1),
    OPT(tdmp_g, 1), OPT(stream_threads, 1), OPT(qc_prefer_g, 1), OPT(qcm_pack, 1), OPT(qcw_warps, 1), OPT(sp_qc, 1), OPT(place_effort, 2),
    OPTR(refill_wait, 1), OPTR(no_streamed, 0), OPTR(streamed_pageable, 0), OPTR(no_staged, 0), OPTR(no_warm, 0), OPTR(avail_memcpy, 0), OPTR(register_host, 0),
    OPTR(staged_min_kb, 2), OPTR(stream_chunk, 2), OPTR(stream_batch_kb, 2), OPTR(wait_timeout_ms, 2),
    OPTR(qc_et, 1), OPTR(qc_et_pct, 1), OPTR(qc_et_every, 1), OPTR(qcm_multi_pct, 1), OPTR(stage_threads, 1), OPTR(stage_nt, 1), OPTR(chunk_stage, 1),
};
#undef OPT
#undef OPTR

void option_store(Options* o, const OptionName& n, long long v) {
    char* p = reinterpret_cast<char*>(o) + n.off;
    if (n.kind == 0) *reinterpret_cast<bool*>(p) = v != 0;
    else if (n.kind == 1) *reinterpret_cast<int*>(p) = (int)v;
    else *reinterpret_cast<long long*>(p) = v;
}

Options options_from_env() {
    Options o;
    for (const OptionName& n : kOptionNames) {
        std::string env = "LDPC_B200_";
        for (const char* c = n.name; *c; ++c) env += (char)std::toupper((unsigned char)*c);
        if (const char* v = std::getenv(env.c_str())) option_store(&o, n, n.kind == 0 ? 1 : std::atoll(v));
    }
    return o;
}

struct Plan {
    int path = LDPC_B200_PATH_LANE_SMEM;
    int threads = 0;
    int ctas = 0;
    int cw_per_cta = 32;
    size_t smem = 0;
    size_t ws_stride = 0;  // floats per CTA (LANE_GLOBAL)
    int dcp = 0;           // LANE16: padded check degree of the per-warp tables
    int W = 0, CS = 0, VS = 0;  // LANE16 / GROUP: warps, check slots and variable slots per warp
    int G = 0, dmax = 0;        // GROUP: codewords per CTA group, unroll bound of the check pass
    bool tab_smem = true;       // GROUP: index tables in shared memory (else read from global/L2)
    bool y_smem = false;        // GROUP: channel values in shared memory (else registers)
    bool t16 = false;           // GROUP: 16-bit table entries (static-profile kernel only)
    int cn_stride = 0, vn_stride = 0, r_rows = 0;
};

// Layered (TDMP) decoder: one layer = z consecutive rows = one sweep of the CTA's W warps.
struct TdmpPlan {
    bool ok = false;
    int G = 0, W = 0, L = 0, VS = 0, z = 0;
    int cn_stride = 0, r_rows = 0, threads = 0, ctas_per_sm = 1;
    size_t smem = 0;
    uint8_t ldeg[kTdmpMaxLayers] = {0};
};

}  // namespace

struct ldpc_b200_decoder {
    HostTables host;
    int K = 0;
    int max_iter = 40;  // reference MyLdpc.cpp:24
    int early = 1;
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    int forced_path = -1;
    int algorithm = LDPC_B200_ALG_MIN_SUM;
    int plan_alg = -1;  // the flooding algorithm `plan` (and the group tables) were built for
    Options opt;
    bool tables_keep_edge_order = false;  // group tables were built in CSR edge order (needed by sum-product)
    Plan plan;
    bool planned = false;

    int32_t* d_row_ptr = nullptr;
    uint32_t* d_cn_col = nullptr;
    int32_t* d_col_ptr = nullptr;
    uint32_t* d_vn_edge = nullptr;
    size_t table_bytes = 0;
    // LANE16 tables (byte-offset form, copied into shared memory by the kernel)
    uint32_t* d16_cn_tab = nullptr;
    uint2* d16_vn_tab = nullptr;
    uint32_t* d16_var_of_pos = nullptr;
    uint32_t* d16_pos_of_var = nullptr;
    bool lane16_ready = false;
    // GROUP tables
    uint32_t* dg_cn_tab = nullptr;
    uint32_t* dg_vn_tab = nullptr;
    uint32_t* dg_var_of_pos = nullptr;
    uint32_t* dg_pos_of_var = nullptr;
    bool group_ready = false;
    // QC tables (warp-uniform, one slot of the __constant__ bank)
    QcParams qc;
    int qc_kind = -1;  // index into the compiled-profile list (qc_profiles())
    int qc_state = 0;  // 0 = not tried, 1 = tables built and uploaded, -1 = no match / no free slot
    int qc_slot = -1;
    size_t qc_smem = 0;
    size_t qc_ring_smem = 0;  // dynamic shared memory of the ring kernel (0: not usable for this profile)
    int qc_ring_per_sm = 0;
    // the early-termination kernel (ldpc_qcw.cuh: a warp per codeword), prepared next to the main one
    QcwParams qcw;
    int qcw_kind = -1, qcw_state = 0, qcw_warps = 0;
    uint32_t* d_syn_tab = nullptr;
    // a group of warps per codeword, any z (ldpc_qcm.cuh): the block sizes without a compiled lockstep profile
    QcmParams qcm;
    int qcm_kind = -1, qcm_state = 0, qcm_slot = -1, qcm_groups = 0;
    std::vector<unsigned char> qcm_tab;      // its table block (also uploaded to the sum-product unit's bank, ldpc_spq.cuh)
    int spq_state = 0;                       // quasi-cyclic sum-product kernel: 0 not tried, 1 ready, -1 not for this code
    QcmParams spq;                           // its geometry: the min-sum kernel's with a byte-wide hard-decision array in T's place
    int spq_groups = 0;
    QcmParams qcm_multi;                     // the same with several codewords per group (ldpc_ms_qcm_multi_kernel); 0 groups = not used
    int qcm_multi_groups = 0;
    // kernel choice per launch: mean iteration count of the previous launches, sampled on the device
    int32_t* d_iters_own = nullptr;          // iteration counts when the caller does not ask for them
    int64_t iters_own_cap = 0;
    unsigned long long* h_stats = nullptr;   // pinned, written by the device after every launch: [0] sum of the sampled counts, [1] words sampled
    unsigned stat_tick = 0;
    unsigned warmed = 0;                     // bit a: the kernels of algorithm a have been launched once (warm_kernels_locked)
    int last_kernel = 0;                     // ldpc_b200_info.kernel_variant of the most recent launch
    // QC tables with a run-time profile (ldpc_qcg.cuh)
    QcgParams qcg;
    QcgWarpTab qcg_tab[kQcgMaxW];
    QcgWarpTab* dq_tabs = nullptr;
    int qcg_state = 0, qcg_G = 0, qcg_ctas_per_sm = 0;
    size_t qcg_smem = 0;
    // WARP tables (sub-warp per check)
    uint32_t* dw_cn_col = nullptr;
    uint32_t* dw_vn_pos = nullptr;
    int w_sw = 0;
    bool warp_ready = false;
    // CLUSTER tables
    uint32_t* dc_cn_tab = nullptr;
    uint32_t* dc_vn_tab = nullptr;
    uint32_t* dc_var_of_pos = nullptr;
    uint32_t* dc_out_addr = nullptr;
    bool cluster_ready = false;
    // STREAM tables (long codes, global workspace)
    uint4* ds_cn_tab = nullptr;
    uint4* ds_vn_tab = nullptr;
    uint32_t* ds_var_of_pos = nullptr;
    uint32_t* ds_pos_of_var = nullptr;
    int s_np = 0, s_nb8 = 0, s_nb4 = 0, s_nb2 = 0;
    bool stream_ready = false;
    // encoder (X transposed, see ldpc_encode.cuh)
    uint32_t* de_xt = nullptr;
    int enc_kw = 0, enc_mw = 0;
    bool encoder_ready = false;
    // TDMP (layered) tables
    int layer_z = 0;  // rows per layer; 0 = unknown (layered decoding unavailable)
    TdmpPlan tdmp;
    uint32_t* dt_cn_tab = nullptr;
    bool tdmp_ready = false;
    bool tdmp_big = false;       // the code does not fit the on-chip layered layout: ldpc_tdmp_big_kernel
    // sum-product / layered decoding of codes of any size (ldpc_big.cuh): CTA-private global workspace
    float* d_ws_big = nullptr;
    size_t ws_big_bytes = 0;
    uint8_t g_vdeg[kGrpMaxVS] = {0};
    uint8_t g_cdeg[kGrpMaxCS] = {0};
    int l16_vn_stride = 0;
    uint8_t l16_vdeg[kL16MaxVS] = {0};
    uint8_t l16_cdeg[kL16MaxCS] = {0};

    unsigned long long* d_counters = nullptr;  // ring of work-queue heads (64-bit; lane kernels use the low word)
    int counter_next = 0;
    float* d_ws = nullptr;
    size_t ws_bytes = 0;
    cudaEvent_t ws_event = nullptr;
    bool ws_event_valid = false;

    // host-buffer pipeline
    int64_t reserved = 0;
    cudaStream_t streams[kSlots] = {nullptr, nullptr, nullptr};
    float* s_llr[kSlots] = {nullptr, nullptr, nullptr};
    uint8_t* s_info[kSlots] = {nullptr, nullptr, nullptr};
    uint8_t* s_hard[kSlots] = {nullptr, nullptr, nullptr};
    int32_t* s_iters[kSlots] = {nullptr, nullptr, nullptr};
    float* s_post[kSlots] = {nullptr, nullptr, nullptr};
    void* s_pack[kSlots] = {nullptr, nullptr, nullptr};   // packed channel values of a chunk (ldpc_b200_decode_host_packed)
    size_t s_pack_bytes = 0;

    // streamed host pipeline (one persistent launch per batch, input chunks announced through *d_avail)
    float* st_llr = nullptr;
    uint8_t* st_info = nullptr;
    uint8_t* st_hard = nullptr;
    int32_t* st_iters = nullptr;
    float* st_post = nullptr;
    int64_t st_cap = 0;                         // words the st_* buffers hold
    bool st_has_hard = false, st_has_post = false;
    unsigned long long* d_avail = nullptr;      // [0] words landed; [1] (as int) timeout status
    int* h_status = nullptr;                    // pinned: the kernel's timeout status, read back after each batch
    unsigned long long* h_avail_vals = nullptr; // pinned: the values the copy stream writes to *d_avail
    int64_t h_avail_cap = 0;
    cudaEvent_t st_event = nullptr;
    const unsigned long long* cur_avail = nullptr;  // set around a streamed launch (under mu)
    int cur_fmt = 0;                                // ... whose input buffer holds packed channel values (LDPC_B200_LLR_*)
    float cur_scale = 1.0f;
    // pageable input: ring of pinned staging buffers filled by host threads
    static constexpr int kStageSlots = 8;
    float* st_pin[kStageSlots] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t st_pin_ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t st_pin_bytes = 0;

    int64_t launches = 0;
    // run-time phase timers of the host-buffer decode (ldpc_b200_get_timing): CUDA events around the copies and the
    // kernels, accumulated per call -- what the reference keeps in stepTime[] (MyLdpc.cpp:26-28, 987-1056)
    // option register_host: input ranges this handle page-locked (released in destroy)
    std::vector<std::pair<const void*, size_t>> registered;
    ldpc_b200_timing timing{};
    std::vector<cudaEvent_t> tev;  // timing events, created on demand
    std::mutex mu;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Pick the warp count that splits checks and variables of a 32-word group most evenly.
int pick_lane_warps(int M, int N, int nnz) {
    int best_w = 8;
    double best = -1.0;
    for (int w = 8; w <= 32; ++w) {
        double cn = (double)M / w / ((M + w - 1) / w);
        double vn = (double)N / w / ((N + w - 1) / w);
        // issue-rate proxy: useful fraction of warp-slots (CN ~ 60 % of the work), scaled by the
        // latency hiding more warps give
        double eff = (0.6 * cn + 0.4 * vn) * (0.75 + 0.25 * w / 32.0);
        if (eff > best + 1e-9) { best = eff; best_w = w; }
    }
    (void)nnz;
    return best_w;
}

// ---- LANE16 layout: degree-sorted ownership, per-warp flat tables (see ldpc_kernels.cuh) ----
struct L16Shape {
    int W = 0, CS = 0, VS = 0, DCP = 0;
    long long vn_stride = 0;   // uint2 entries per warp
    double padded_work = 0.0;  // issue-slot proxy used to pick W
    size_t smem = 0;
};

// degrees sorted descending; slot degree = degree of the first (largest) member of the slot
L16Shape lane16_shape(const HostTables& t, int W) {
    L16Shape sh;
    sh.W = W;
    sh.CS = (t.M + W - 1) / W;
    sh.VS = (t.N + W - 1) / W;
    sh.DCP = (t.max_row_weight + 3) & ~3;
    std::vector<int> cd(t.M), vd(t.N);
    for (int r = 0; r < t.M; ++r) cd[r] = t.row_ptr[r + 1] - t.row_ptr[r];
    for (int c = 0; c < t.N; ++c) vd[c] = t.col_ptr[c + 1] - t.col_ptr[c];
    std::sort(cd.begin(), cd.end(), std::greater<int>());
    std::sort(vd.begin(), vd.end(), std::greater<int>());
    double cn_edges = 0, vn_edges = 0;
    for (int cs = 0; cs < sh.CS; ++cs) cn_edges += (double)cd[(size_t)cs * W] * W;
    for (int s = 0; s < sh.VS; ++s) { vn_edges += (double)vd[(size_t)s * W] * W; sh.vn_stride += vd[(size_t)s * W]; }
    sh.padded_work = 13.0 * cn_edges + 8.0 * vn_edges + 20.0 * sh.CS * W + 10.0 * sh.VS * W;
    sh.smem = (size_t)(sh.CS * W + 1) * kLanes * 16 + (size_t)(sh.VS * W + 1) * kLanes * 4 +
              (size_t)W * sh.CS * sh.DCP * 4 + (size_t)W * sh.vn_stride * 8;
    return sh;
}

bool lane16_pick(const HostTables& t, size_t smem_limit, const Options& opt, L16Shape* best) {
    if (t.max_row_weight > 32 || t.max_row_weight < 1 || t.nnz < 1) return false;
    bool found = false;
    double best_cost = 0.0;
    int w_lo = 8, w_hi = 32;
    if (opt.l16_warps >= 1 && opt.l16_warps <= 32) w_lo = w_hi = opt.l16_warps;  // tuning aid: pin the warp count
    for (int W = w_lo; W <= w_hi; ++W) {
        L16Shape sh = lane16_shape(t, W);
        if (sh.VS > kL16MaxVS || sh.CS > kL16MaxCS) continue;
        if (sh.smem + 1024 > smem_limit) continue;
        // measured on cfg2 (profiles/r01_warp_sweep.txt): the unpadded split wins over extra warps;
        // keep only a very mild preference for more warps among near-equal splits
        const double cost = sh.padded_work / (0.97 + 0.03 * W / 32.0);
        if (!found || cost < best_cost - 1e-9 || (std::abs(cost - best_cost) <= 1e-9 && W > best->W)) {
            found = true; best_cost = cost; *best = sh;
        }
    }
    return found;
}

int upload_lane16_tables(ldpc_b200_decoder* h) {
    if (h->lane16_ready) return LDPC_B200_OK;
    const HostTables& t = h->host;
    const Plan& pl = h->plan;
    const int W = pl.W, CS = pl.CS, VS = pl.VS, DCP = pl.dcp;
    // rank variables / checks by degree (descending, stable in index): position = rank = slot*W + warp
    std::vector<int> vorder(t.N), corder(t.M);
    for (int i = 0; i < t.N; ++i) vorder[i] = i;
    for (int i = 0; i < t.M; ++i) corder[i] = i;
    auto vdegf = [&](int c) { return t.col_ptr[c + 1] - t.col_ptr[c]; };
    auto cdegf = [&](int r) { return t.row_ptr[r + 1] - t.row_ptr[r]; };
    std::stable_sort(vorder.begin(), vorder.end(), [&](int a, int b) { return vdegf(a) > vdegf(b); });
    std::stable_sort(corder.begin(), corder.end(), [&](int a, int b) { return cdegf(a) > cdegf(b); });
    const int PD = VS * W, CPD = CS * W;
    std::vector<uint32_t> var_of_pos((size_t)PD, 0xffffffffu), pos_of_var(t.N), cpos_of_chk(t.M);
    for (int i = 0; i < t.N; ++i) { var_of_pos[i] = (uint32_t)vorder[i]; pos_of_var[vorder[i]] = (uint32_t)i; }
    for (int i = 0; i < t.M; ++i) cpos_of_chk[corder[i]] = (uint32_t)i;
    std::memset(h->l16_vdeg, 0, sizeof(h->l16_vdeg));
    std::memset(h->l16_cdeg, 0, sizeof(h->l16_cdeg));
    for (int s = 0; s < VS; ++s) h->l16_vdeg[s] = (uint8_t)vdegf(vorder[(size_t)s * W]);
    for (int cs = 0; cs < CS; ++cs) h->l16_cdeg[cs] = (uint8_t)cdegf(corder[(size_t)cs * W]);
    long long vn_stride = 0;
    std::vector<long long> voff(VS + 1, 0);
    for (int s = 0; s < VS; ++s) { voff[s] = vn_stride; vn_stride += h->l16_vdeg[s]; }
    h->l16_vn_stride = (int)vn_stride;

    // check-major per-warp table: real edges first (position j fixes the sign-bit slot), then dummies
    std::vector<uint32_t> cn_tab((size_t)W * CS * DCP, (uint32_t)PD * 128u);
    for (int w = 0; w < W; ++w)
        for (int cs = 0; cs < CS; ++cs) {
            const int rank = cs * W + w;
            if (rank >= t.M) continue;
            const int r = corder[rank], e0 = t.row_ptr[r], dc = cdegf(r);
            for (int j = 0; j < dc; ++j)
                cn_tab[((size_t)w * CS + cs) * DCP + j] = pos_of_var[t.col_idx[e0 + j]] * 128u;
        }
    // variable-major per-warp table: ascending-row edge order kept; dummies pad to the slot degree
    std::vector<uint2> vn_tab((size_t)W * std::max<long long>(vn_stride, 1), make_uint2((uint32_t)CPD * 512u, 0u));
    for (int w = 0; w < W; ++w)
        for (int s = 0; s < VS; ++s) {
            const int rank = s * W + w;
            if (rank >= t.N) continue;
            const int v = vorder[rank];
            for (int k = t.col_ptr[v]; k < t.col_ptr[v + 1]; ++k) {
                const uint32_t chk = t.vn_edge[k] >> kPosBits, pos = t.vn_edge[k] & ((1u << kPosBits) - 1u);
                const uint32_t cp = cpos_of_chk[chk];
                const int slot_deg = h->l16_cdeg[cp / W];
                const int shift = 32 - slot_deg + (int)pos;  // bit (slot_deg-1-pos) -> bit 31
                vn_tab[(size_t)w * vn_stride + voff[s] + (k - t.col_ptr[v])] = make_uint2(cp * 512u, 1u << shift);
            }
        }
    CU_TRY(cudaMalloc(&h->d16_cn_tab, cn_tab.size() * 4));
    CU_TRY(cudaMalloc(&h->d16_vn_tab, vn_tab.size() * 8));
    CU_TRY(cudaMalloc(&h->d16_var_of_pos, var_of_pos.size() * 4));
    CU_TRY(cudaMalloc(&h->d16_pos_of_var, pos_of_var.size() * 4));
    CU_TRY(cudaMemcpy(h->d16_cn_tab, cn_tab.data(), cn_tab.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d16_vn_tab, vn_tab.data(), vn_tab.size() * 8, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d16_var_of_pos, var_of_pos.data(), var_of_pos.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d16_pos_of_var, pos_of_var.data(), pos_of_var.size() * 4, cudaMemcpyHostToDevice));
    h->table_bytes += cn_tab.size() * 4 + vn_tab.size() * 8 + var_of_pos.size() * 4 + pos_of_var.size() * 4;
    h->lane16_ready = true;
    return LDPC_B200_OK;
}

// ---- GROUP layout (explicit per-edge messages on chip; see ldpc_kernels.cuh) ----------------
struct GrpShape {
    int G = 0, W = 0, CS = 0, VS = 0, dmax = 0;
    int cn_stride = 0, vn_stride = 0, r_rows = 0;
    bool tab_smem = true;
    bool y_smem = false;
    size_t smem = 0;
    double cost = 0.0;
};

bool group_shape(const HostTables& t, int G, int W, size_t smem_limit_in, bool sp, const Options& opt, GrpShape* out) {
    // G = 8 runs two CTAs per SM (phase-shifted check / variable passes overlap): half the budget each
    const size_t smem_limit = G == 8 ? (smem_limit_in + 1024) / 2 - 1024 : (G == 4 ? (smem_limit_in + 1024) / 3 - 1024 : smem_limit_in);
    const int SUB = 32 / G, NL = W * SUB;
    GrpShape sh;
    sh.G = G; sh.W = W;
    sh.CS = (t.M + NL - 1) / NL;
    sh.VS = (t.N + NL - 1) / NL;
    // the sum-product kernel is instantiated to check degree 20 (802.16e rate 5/6) and variable degree 8
    if (sh.CS > kGrpMaxCS || sh.VS > kGrpMaxVS || t.max_row_weight > (sp ? 20 : 16) || t.max_row_weight < 1) return false;
    if (sp && (t.max_col_weight > 8 || G == 1)) return false;
    sh.dmax = t.max_row_weight <= 8 ? 8 : (t.max_row_weight <= 16 ? 16 : 20);
    std::vector<int> cd(t.M), vd(t.N);
    for (int r = 0; r < t.M; ++r) cd[r] = t.row_ptr[r + 1] - t.row_ptr[r];
    for (int c = 0; c < t.N; ++c) vd[c] = t.col_ptr[c + 1] - t.col_ptr[c];
    std::sort(cd.begin(), cd.end(), std::greater<int>());
    std::sort(vd.begin(), vd.end(), std::greater<int>());
    long long quads = 0, rrows = 0, ventries = 0, vquads = 0;
    for (int cs = 0; cs < sh.CS; ++cs) { const int d = cd[(size_t)cs * NL]; rrows += d; quads += (d + 3) / 4; }
    for (int s = 0; s < sh.VS; ++s) { const int d = vd[(size_t)s * NL]; ventries += d; vquads += (d + 3) / 4; }
    sh.r_rows = (int)rrows;
    sh.cn_stride = (int)(quads * SUB * 4);
    sh.vn_stride = (int)(vquads * SUB * 4);
    const size_t tbytes = ((size_t)(sh.VS * NL + 1) * G * 4 + 127) & ~(size_t)127;
    const size_t core = tbytes + ((size_t)W * rrows + 1) * 128;
    const size_t tabs = ((size_t)sh.cn_stride + sh.vn_stride) * 4 * W;
    if (core + 512 > smem_limit) return false;
    sh.tab_smem = core + tabs + 512 <= smem_limit;
    sh.smem = core + (sh.tab_smem ? tabs : 0);
    // channel values on chip too when they fit (dynamic slot loop, two slots in flight)
    sh.y_smem = !sp && t.max_col_weight <= 12 && sh.smem + tbytes + 512 <= smem_limit && !opt.grp_no_ysmem;
    if (sh.y_smem) sh.smem += tbytes;
    // issue-slot proxy of the padded work, as for LANE16
    sh.cost = (14.0 * rrows + 4.0 * ventries + 12.0 * sh.CS + 6.0 * sh.VS) * NL / (0.97 + 0.03 * W / 32.0);
    if (!sh.tab_smem) sh.cost *= 1.15;
    *out = sh;
    return true;
}

bool group_pick(const HostTables& t, size_t smem_limit, bool sp, const Options& opt, GrpShape* best) {
    int g_lo = 1, g_hi = 16, w_lo = 8, w_hi = 32;
    if (opt.grp_g == 1 || opt.grp_g == 16 || opt.grp_g == 8 || opt.grp_g == 4) g_lo = g_hi = opt.grp_g;
    if (opt.grp_warps >= 1 && opt.grp_warps <= 32) w_lo = w_hi = opt.grp_warps;
    // instantiated: G = 16 (short codes, tables on chip) and G = 1 (one codeword per CTA)
    if (g_lo == 4) {  // experiment: three CTAs of 4 words per SM (LDPC_B200_GRP_G=4)
        bool found = false;
        GrpShape b;
        for (int W = w_lo; W <= std::min(w_hi, 9); ++W) {
            GrpShape sh;
            if (!group_shape(t, 4, W, smem_limit, sp, opt, &sh) || !sh.tab_smem) continue;
            if (!found || sh.cost < b.cost - 1e-9 || (std::abs(sh.cost - b.cost) <= 1e-9 && W > b.W)) { found = true; b = sh; }
        }
        if (found) { *best = b; return true; }
        return false;
    }
    for (int G : {16, 8, 1}) {
        if (G < g_lo || G > g_hi) continue;
        if (G == 16 && g_lo != 16 && !opt.grp_prefer_16) {
            // measured (profiles/): two phase-shifted CTAs of 8 words beat one CTA of 16 -- try G = 8 first
            bool found8 = false;
            GrpShape b8;
            for (int W = w_lo; W <= std::min(w_hi, 12); ++W) {
                GrpShape sh;
                if (!group_shape(t, 8, W, smem_limit, sp, opt, &sh) || !sh.tab_smem) continue;
                if (!found8 || sh.cost < b8.cost - 1e-9 || (std::abs(sh.cost - b8.cost) <= 1e-9 && W > b8.W)) { found8 = true; b8 = sh; }
            }
            if (found8) { *best = b8; return true; }
        }
        if (G == 8 && g_lo != 8) continue;
        bool found = false;
        GrpShape b;
        for (int W = w_lo; W <= w_hi; ++W) {
            GrpShape sh;
            if (!group_shape(t, G, W, smem_limit, sp, opt, &sh)) continue;
            if (G == 16 && !sh.tab_smem) continue;
            if (G == 8 && (!sh.tab_smem || W > 12)) continue;
            if (!found || sh.cost < b.cost - 1e-9 || (std::abs(sh.cost - b.cost) <= 1e-9 && W > b.W)) { found = true; b = sh; }
        }
        if (found) { *best = b; return true; }
    }
    return false;
}

int upload_group_tables(ldpc_b200_decoder* h) {
    if (h->group_ready) return LDPC_B200_OK;
    const HostTables& t = h->host;
    const Plan& pl = h->plan;
    const int G = pl.G, SUB = 32 / G, W = pl.W, NL = W * SUB, CS = pl.CS, VS = pl.VS;
    std::vector<int> vorder(t.N), corder(t.M);
    for (int i = 0; i < t.N; ++i) vorder[i] = i;
    for (int i = 0; i < t.M; ++i) corder[i] = i;
    auto vdegf = [&](int c) { return t.col_ptr[c + 1] - t.col_ptr[c]; };
    auto cdegf = [&](int r) { return t.row_ptr[r + 1] - t.row_ptr[r]; };
    std::stable_sort(vorder.begin(), vorder.end(), [&](int a, int b) { return vdegf(a) > vdegf(b); });
    std::stable_sort(corder.begin(), corder.end(), [&](int a, int b) { return cdegf(a) > cdegf(b); });
    const int PD = VS * NL, RD = W * pl.r_rows;
    std::vector<uint32_t> var_of_pos((size_t)PD, 0xffffffffu), pos_of_var(t.N), crank_of_chk(t.M);
    for (int i = 0; i < t.N; ++i) { var_of_pos[i] = (uint32_t)vorder[i]; pos_of_var[vorder[i]] = (uint32_t)i; }
    for (int i = 0; i < t.M; ++i) crank_of_chk[corder[i]] = (uint32_t)i;
    std::memset(h->g_vdeg, 0, sizeof(h->g_vdeg));
    std::memset(h->g_cdeg, 0, sizeof(h->g_cdeg));
    std::vector<int> coff(CS + 1, 0), qoff(CS + 1, 0), voff(VS + 1, 0);
    for (int cs = 0; cs < CS; ++cs) {
        h->g_cdeg[cs] = (uint8_t)cdegf(corder[(size_t)cs * NL]);
        coff[cs + 1] = coff[cs] + h->g_cdeg[cs];
        qoff[cs + 1] = qoff[cs] + (h->g_cdeg[cs] + 3) / 4;
    }
    for (int s = 0; s < VS; ++s) {
        h->g_vdeg[s] = (uint8_t)vdegf(vorder[(size_t)s * NL]);
        voff[s + 1] = voff[s] + (h->g_vdeg[s] + 3) / 4;  // in quads
    }
    // ---- bank-conflict-aware placement (SUB == 2: two nodes share a warp instruction) --------------
    // slot_of_edge[e] = row position j that edge e (CSR id) takes inside its check.  The order of a
    // check's edges is free for min/xor, so for each co-processed check pair (A: h=0, B: h=1) we match
    // edges whose T rows have opposite parity (different bank halves) to the same j.
    std::vector<int> slot_of_edge(t.nnz);
    for (int r = 0; r < t.M; ++r)
        for (int e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) slot_of_edge[e] = e - t.row_ptr[r];
    const bool keep_order = h->plan_alg == LDPC_B200_ALG_SUM_PRODUCT;  // products are taken in CSR edge order
    h->tables_keep_edge_order = keep_order;
    if (SUB == 2 || SUB == 4) {
        // Lanes of one warp instruction touch SUB different rows; a row occupies 32/SUB banks chosen by
        // (row index mod SUB) for T and by the owning check's node lane for R.  Equal classes collide.
        // (1) node lane of every check inside its co-processed group (SUB consecutive ranks): local
        //     search over swaps, minimising the collisions of the variable pass, where the k-th messages
        //     of SUB co-processed variables are read together.
        const int ngrp = (t.M + SUB - 1) / SUB;
        auto lane_of_chk = [&](int chk) { return (int)crank_of_chk[chk] % SUB; };
        auto vgroup_cost = [&](int vg) {
            long long cost = 0;
            int maxd = 0;
            for (int i = 0; i < SUB; ++i) { const int rk = vg * SUB + i; if (rk < t.N) maxd = std::max(maxd, vdegf(vorder[rk])); }
            for (int k = 0; k < maxd; ++k) {
                int cnt[4] = {0, 0, 0, 0};
                for (int i = 0; i < SUB; ++i) {
                    const int rk = vg * SUB + i;
                    if (rk >= t.N) continue;
                    const int v = vorder[rk];
                    if (k >= vdegf(v)) continue;
                    cnt[lane_of_chk((int)(t.vn_edge[t.col_ptr[v] + k] >> kPosBits))]++;
                }
                cost += std::max(std::max(cnt[0], cnt[1]), std::max(cnt[2], cnt[3]));
            }
            return cost;
        };
        std::vector<std::vector<int>> touched(ngrp);  // variable groups reading messages of each check group
        for (int v = 0; v < t.N; ++v)
            for (int k = t.col_ptr[v]; k < t.col_ptr[v + 1]; ++k)
                touched[(int)crank_of_chk[t.vn_edge[k] >> kPosBits] / SUB].push_back((int)pos_of_var[v] / SUB);
        for (auto& v : touched) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
        auto touched_cost = [&](int cg) { long long c = 0; for (int vg : touched[cg]) c += vgroup_cost(vg); return c; };
        auto swap_ranks = [&](int ra, int rb) {
            std::swap(corder[ra], corder[rb]);
            crank_of_chk[corder[ra]] = (uint32_t)ra;
            crank_of_chk[corder[rb]] = (uint32_t)rb;
        };
        for (int pass = 0; pass < 8; ++pass) {
            bool changed = false;
            for (int cg = 0; cg < ngrp; ++cg)
                for (int i = 0; i < SUB; ++i)
                    for (int j = i + 1; j < SUB; ++j) {
                        const int ra = cg * SUB + i, rb = cg * SUB + j;
                        if (rb >= t.M) continue;
                        const long long before = touched_cost(cg);
                        swap_ranks(ra, rb);
                        if (touched_cost(cg) < before) changed = true; else swap_ranks(ra, rb);
                    }
            if (!changed) break;
        }
        // (2) edge order inside each check group: at equal j the SUB T rows should fall in different bank
        //     classes (row index mod SUB).  Greedy column by column, preferring each check's fullest class.
        for (int cg = 0; cg < ngrp && !keep_order; ++cg) {
            std::vector<int> byclass[4][4];  // [member][class] -> edges
            int deg[4] = {0, 0, 0, 0}, nmem = 0;
            for (int i = 0; i < SUB; ++i) {
                const int rk = cg * SUB + i;
                if (rk >= t.M) break;
                nmem = i + 1;
                const int r = corder[rk];
                deg[i] = cdegf(r);
                for (int e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) byclass[i][pos_of_var[t.col_idx[e]] % SUB].push_back(e);
            }
            const int maxd = std::max(std::max(deg[0], deg[1]), std::max(deg[2], deg[3]));
            int placed[4] = {0, 0, 0, 0};
            for (int j = 0; j < maxd; ++j) {
                bool used[4] = {false, false, false, false};
                // members with fewer remaining choices go first
                int order[4] = {0, 1, 2, 3};
                std::sort(order, order + nmem, [&](int x, int y) {
                    int cx = 0, cy = 0;
                    for (int c = 0; c < SUB; ++c) { cx += !byclass[x][c].empty(); cy += !byclass[y][c].empty(); }
                    return cx < cy;
                });
                for (int oi = 0; oi < nmem; ++oi) {
                    const int i = order[oi];
                    if (placed[i] >= deg[i]) continue;
                    int best = -1;
                    for (int c = 0; c < SUB; ++c)
                        if (!byclass[i][c].empty() && !used[c] && (best < 0 || byclass[i][c].size() > byclass[i][best].size())) best = c;
                    if (best < 0)
                        for (int c = 0; c < SUB; ++c)
                            if (!byclass[i][c].empty() && (best < 0 || byclass[i][c].size() > byclass[i][best].size())) best = c;
                    const int e = byclass[i][best].back();
                    byclass[i][best].pop_back();
                    used[best] = true;
                    slot_of_edge[e] = placed[i]++;
                }
            }
        }
    }
    if (SUB >= 8) {
        // ---- bank placement for wide node groups (G <= 4; G = 1: 32 different checks / variables per
        // instruction).  T[pos] sits in bank (pos mod 32) = the owning variable's lane; R elements sit in
        // the bank of the owning check's lane.  Random codes collide ~3.4-way in both gathers.  Lanes
        // inside a 32-node group are free to permute, so:
        //   (a) permute variable lanes so that every check group sees each T bank equally often
        //       (then its edges split into conflict-free columns -- Koenig);
        //   (b) permute check lanes so that, for every k, the k-th messages of a variable group come
        //       from distinct R banks (the per-variable order k is fixed by the summation order);
        //   (c) order each check group's edges column by column with a bipartite matching.
        const int LN = SUB;  // lanes per group (bank classes); rows are 4*G bytes wide
        const int nvg = (t.N + LN - 1) / LN, ncg = (t.M + LN - 1) / LN;
        uint64_t rng = 0x9E3779B97F4A7C15ull;
        auto rnd = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 32); };
        // Order of work: (b) first, because it is the hard one (the message order inside a variable is fixed by
        // the summation order, so a variable group needs, for every k, 32 k-th checks in 32 distinct lanes):
        //   (b1) swap check lanes until every lane serves each k equally often (otherwise no grouping can work),
        //   (b2) regroup the variables with targeted swaps: take a variable whose k-th check collides inside its
        //        group, look for a partner among the variables whose k-th check sits in a lane the group lacks,
        //        swap if the collision count does not grow.
        // Then (a) without touching (b): variable lanes inside a group and the group membership of checks that
        // share a lane are still free.  Only nodes of equal degree trade places (a slot has one degree).
        const long long effort = h->opt.place_effort;
        const int maxdv = t.max_col_weight;
        auto kth_check = [&](int v, int k) { return (int)(t.vn_edge[t.col_ptr[v] + k] >> kPosBits); };
        if (LN == 32 && !keep_order && maxdv <= 16 && effort > 0) {
            // (b1) lane balance: cnt[k][lane] = number of variables whose k-th check has that lane
            std::vector<int> clane(t.M);
            for (int i = 0; i < t.M; ++i) clane[corder[i]] = i % LN;
            std::vector<int> ck((size_t)t.M * maxdv, 0), cnt((size_t)maxdv * LN, 0), tot(maxdv, 0);
            for (int v = 0; v < t.N; ++v)
                for (int k = 0; k < vdegf(v); ++k) { ck[(size_t)kth_check(v, k) * maxdv + k]++; tot[k]++; }
            for (int c = 0; c < t.M; ++c)
                for (int k = 0; k < maxdv; ++k) cnt[(size_t)k * LN + clane[c]] += ck[(size_t)c * maxdv + k];
            auto dev2 = [&](int k, int lane) { const long long d = (long long)cnt[(size_t)k * LN + lane] * LN - tot[k]; return d * d; };
            for (long long it = 0; it < (long long)t.M * 200; ++it) {
                const int ra = (int)(rnd() % (uint32_t)t.M), rb = (int)(rnd() % (uint32_t)t.M);
                const int a = corder[ra], b = corder[rb], la = ra % LN, lb = rb % LN;
                if (la == lb || cdegf(a) != cdegf(b)) continue;
                long long before = 0, after = 0;
                for (int k = 0; k < maxdv; ++k) before += dev2(k, la) + dev2(k, lb);
                for (int k = 0; k < maxdv; ++k) {
                    const int d = ck[(size_t)b * maxdv + k] - ck[(size_t)a * maxdv + k];
                    cnt[(size_t)k * LN + la] += d; cnt[(size_t)k * LN + lb] -= d;
                }
                for (int k = 0; k < maxdv; ++k) after += dev2(k, la) + dev2(k, lb);
                if (after <= before) { std::swap(corder[ra], corder[rb]); clane[a] = lb; clane[b] = la; }
                else
                    for (int k = 0; k < maxdv; ++k) {
                        const int d = ck[(size_t)b * maxdv + k] - ck[(size_t)a * maxdv + k];
                        cnt[(size_t)k * LN + la] -= d; cnt[(size_t)k * LN + lb] += d;
                    }
            }
            // (b2) variable groups
            std::vector<std::vector<int>> bylane((size_t)maxdv * LN);  // variables by (k, lane of the k-th check)
            for (int v = 0; v < t.N; ++v)
                for (int k = 0; k < vdegf(v); ++k) bylane[(size_t)k * LN + clane[kth_check(v, k)]].push_back(v);
            std::vector<int> rank_of_var(t.N);
            for (int r = 0; r < t.N; ++r) rank_of_var[vorder[r]] = r;
            std::vector<uint8_t> occ((size_t)nvg * maxdv * LN, 0);  // [group][k][lane]
            auto cell = [&](int g, int k, int lane) -> uint8_t& { return occ[((size_t)g * maxdv + k) * LN + lane]; };
            for (int v = 0; v < t.N; ++v)
                for (int k = 0; k < vdegf(v); ++k) cell(rank_of_var[v] / LN, k, clane[kth_check(v, k)])++;
            auto take = [&](int v, int g) { int d = 0; for (int k = 0; k < vdegf(v); ++k) { uint8_t& x = cell(g, k, clane[kth_check(v, k)]); if (x > 1) --d; --x; } return d; };
            auto put = [&](int v, int g) { int d = 0; for (int k = 0; k < vdegf(v); ++k) { uint8_t& x = cell(g, k, clane[kth_check(v, k)]); if (x >= 1) ++d; ++x; } return d; };
            for (long long it = 0; it < (long long)t.N * 100 * effort; ++it) {
                const int v = (int)(rnd() % (uint32_t)t.N), g = rank_of_var[v] / LN, dv = vdegf(v);
                int kk = -1;
                for (int k = 0; k < dv; ++k) if (cell(g, k, clane[kth_check(v, k)]) > 1) { kk = k; break; }
                if (kk < 0) continue;
                int miss[32], nm = 0;
                for (int x = 0; x < LN; ++x) if (cell(g, kk, x) == 0) miss[nm++] = x;
                if (!nm) continue;
                int best = -1, bestd = 1 << 30, ties = 0;
                for (int tr = 0; tr < 24; ++tr) {
                    const std::vector<int>& lst = bylane[(size_t)kk * LN + miss[rnd() % (uint32_t)nm]];
                    if (lst.empty()) continue;
                    const int w = lst[rnd() % (uint32_t)lst.size()], hgrp = rank_of_var[w] / LN;
                    if (hgrp == g || vdegf(w) != dv) continue;
                    const int d = take(v, g) + take(w, hgrp) + put(v, hgrp) + put(w, g);
                    take(v, hgrp); take(w, g); put(v, g); put(w, hgrp);
                    if (d < bestd) { bestd = d; best = w; ties = 1; }
                    else if (d == bestd && rnd() % (uint32_t)(++ties) == 0) best = w;
                }
                if (best < 0 || bestd > 0 || (bestd == 0 && (rnd() & 3))) continue;
                const int w = best, hgrp = rank_of_var[w] / LN;
                take(v, g); take(w, hgrp); put(v, hgrp); put(w, g);
                const int rv = rank_of_var[v], rw = rank_of_var[w];
                std::swap(vorder[rv], vorder[rw]);
                rank_of_var[v] = rw; rank_of_var[w] = rv;
            }
            for (int i = 0; i < t.M; ++i) crank_of_chk[corder[i]] = (uint32_t)i;
        }
        // (a) variable lanes inside their group, and the group of checks that share a lane
        {
            std::vector<int> cnt((size_t)ncg * LN, 0);  // [check group][bank]
            auto bump = [&](int v, int bank, int dlt, long long& cost) {
                for (int k = t.col_ptr[v]; k < t.col_ptr[v + 1]; ++k) {
                    const int cg = (int)crank_of_chk[t.vn_edge[k] >> kPosBits] / LN;
                    int& x = cnt[(size_t)cg * LN + bank];
                    cost += dlt > 0 ? 2 * x + 1 : -(2 * x - 1);
                    x += dlt;
                }
            };
            std::vector<int> lane_of_var(t.N);
            for (int r = 0; r < t.N; ++r) lane_of_var[vorder[r]] = r % LN;
            auto bump_chk = [&](int chk, int cg, int dlt, long long& cost) {  // check `chk` joins / leaves check group cg
                for (int e = t.row_ptr[chk]; e < t.row_ptr[chk + 1]; ++e) {
                    int& x = cnt[(size_t)cg * LN + lane_of_var[t.col_idx[e]]];
                    cost += dlt > 0 ? 2 * x + 1 : -(2 * x - 1);
                    x += dlt;
                }
            };
            std::vector<long long> cg_edges(ncg, 0);  // edges of a check group: its fair share per bank is 1/LN of them
            for (int r = 0; r < t.M; ++r) cg_edges[r / LN] += cdegf(corder[r]);
            long long cost = 0;
            for (int r = 0; r < t.N; ++r) bump(vorder[r], r % LN, +1, cost);
            const long long iters = (long long)t.N * 60 * std::max<long long>(1, effort / 2);
            for (long long it = 0; it < iters; ++it) {
                if (rnd() & 1) {
                    // a variable sitting in a bank that one of its check groups sees too often trades lanes with the
                    // best of a few partners from its own group
                    const int a = (int)(rnd() % (uint32_t)t.N), va = vorder[a];
                    bool over = false;
                    for (int k = t.col_ptr[va]; k < t.col_ptr[va + 1] && !over; ++k) {
                        const int chk = (int)(t.vn_edge[k] >> kPosBits);
                        const int cg = (int)crank_of_chk[chk] / LN;
                        over = cnt[(size_t)cg * LN + a % LN] > (int)((cg_edges[cg] + LN - 1) / LN);
                    }
                    if (!over) continue;
                    int best = -1;
                    long long bestd = 1;
                    for (int tr = 0; tr < 6; ++tr) {
                        const int b = a / LN * LN + (int)(rnd() % LN);
                        if (b == a || b >= t.N) continue;
                        long long d = 0;
                        bump(vorder[a], a % LN, -1, d); bump(vorder[b], b % LN, -1, d);
                        bump(vorder[a], b % LN, +1, d); bump(vorder[b], a % LN, +1, d);
                        long long u = 0;
                        bump(vorder[a], b % LN, -1, u); bump(vorder[b], a % LN, -1, u);
                        bump(vorder[a], a % LN, +1, u); bump(vorder[b], b % LN, +1, u);
                        if (d < bestd) { bestd = d; best = b; }
                    }
                    if (best < 0 || bestd > 0) continue;
                    const int b = best;
                    long long d = 0;
                    bump(vorder[a], a % LN, -1, d); bump(vorder[b], b % LN, -1, d);
                    bump(vorder[a], b % LN, +1, d); bump(vorder[b], a % LN, +1, d);
                    std::swap(vorder[a], vorder[b]);
                    lane_of_var[vorder[a]] = a % LN; lane_of_var[vorder[b]] = b % LN;
                    cost += d;
                } else {  // two checks in the same lane trade groups: (b) is untouched
                    const int ra = (int)(rnd() % (uint32_t)t.M), a = corder[ra], ga = ra / LN;
                    bool over = false;  // does this check feed a bank its group sees too often?
                    for (int e = t.row_ptr[a]; e < t.row_ptr[a + 1] && !over; ++e)
                        over = cnt[(size_t)ga * LN + lane_of_var[t.col_idx[e]]] > (int)((cg_edges[ga] + LN - 1) / LN);
                    if (!over) continue;
                    int best = -1;
                    long long bestd = 1;
                    for (int tr = 0; tr < 6; ++tr) {
                        const int rb = (int)(rnd() % (uint32_t)ncg) * LN + ra % LN;
                        if (rb == ra || rb >= t.M) continue;
                        const int b = corder[rb];
                        if (cdegf(a) != cdegf(b)) continue;
                        long long d = 0, u = 0;
                        bump_chk(a, ga, -1, d); bump_chk(b, rb / LN, -1, d);
                        bump_chk(a, rb / LN, +1, d); bump_chk(b, ga, +1, d);
                        bump_chk(a, rb / LN, -1, u); bump_chk(b, ga, -1, u);
                        bump_chk(a, ga, +1, u); bump_chk(b, rb / LN, +1, u);
                        if (d < bestd) { bestd = d; best = rb; }
                    }
                    if (best < 0 || bestd > 0) continue;
                    const int rb = best, b = corder[rb];
                    long long d = 0;
                    bump_chk(a, ga, -1, d); bump_chk(b, rb / LN, -1, d);
                    bump_chk(a, rb / LN, +1, d); bump_chk(b, ga, +1, d);
                    std::swap(corder[ra], corder[rb]);
                    crank_of_chk[a] = (uint32_t)rb; crank_of_chk[b] = (uint32_t)ra;
                    cost += d;
                }
            }
            for (int i = 0; i < t.N; ++i) { var_of_pos[i] = (uint32_t)vorder[i]; pos_of_var[vorder[i]] = (uint32_t)i; }
        }
        // (b) check lanes by plain swaps inside a group -- the fallback when (b1)/(b2) did not run
        if (!(LN == 32 && !keep_order && maxdv <= 16 && effort > 0)) {
            std::vector<int> cnt((size_t)nvg * maxdv * LN, 0);  // [variable group][k][bank]
            auto bump = [&](int chk, int bank, int dlt, long long& cost) {
                for (int e = t.row_ptr[chk]; e < t.row_ptr[chk + 1]; ++e) {
                    const int v = t.col_idx[e];
                    // k = index of this check in v's ascending-row list
                    int k = 0;
                    for (int q = t.col_ptr[v]; q < t.col_ptr[v + 1]; ++q, ++k)
                        if ((int)(t.vn_edge[q] >> kPosBits) == chk) break;
                    const int vg = (int)pos_of_var[v] / LN;
                    int& x = cnt[((size_t)vg * maxdv + k) * LN + bank];
                    cost += dlt > 0 ? 2 * x + 1 : -(2 * x - 1);
                    x += dlt;
                }
            };
            long long cost = 0;
            for (int r = 0; r < t.M; ++r) bump(corder[r], r % LN, +1, cost);
            const long long iters = (long long)t.M * 120;
            for (long long it = 0; it < iters; ++it) {
                const int cg = (int)(rnd() % (uint32_t)ncg);
                const int a = cg * LN + (int)(rnd() % LN), b = cg * LN + (int)(rnd() % LN);
                if (a == b || a >= t.M || b >= t.M) continue;
                long long d = 0;
                bump(corder[a], a % LN, -1, d); bump(corder[b], b % LN, -1, d);
                bump(corder[a], b % LN, +1, d); bump(corder[b], a % LN, +1, d);
                if (d <= 0) {
                    std::swap(corder[a], corder[b]);
                    cost += d;
                } else {
                    long long u = 0;
                    bump(corder[a], b % LN, -1, u); bump(corder[b], a % LN, -1, u);
                    bump(corder[a], a % LN, +1, u); bump(corder[b], b % LN, +1, u);
                }
            }
            for (int i = 0; i < t.M; ++i) crank_of_chk[corder[i]] = (uint32_t)i;
        }
        // (c) edge columns inside each check group: maximum bipartite matching (checks x banks) per column
        for (int cg = 0; cg < ncg && !keep_order; ++cg) {
            const int n = std::min(LN, t.M - cg * LN);
            std::vector<std::vector<int>> rem(n);  // remaining edges per member
            int maxd = 0;
            for (int i = 0; i < n; ++i) {
                const int r = corder[cg * LN + i];
                for (int e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) rem[i].push_back(e);
                maxd = std::max(maxd, (int)rem[i].size());
            }
            std::vector<int> placed(n, 0);
            for (int j = 0; j < maxd; ++j) {
                std::vector<int> bank_owner(LN, -1), pick(n, -1);
                std::vector<char> seen;
                std::function<bool(int)> aug = [&](int i) {
                    for (size_t x = 0; x < rem[i].size(); ++x) {
                        const int bnk = (int)(pos_of_var[t.col_idx[rem[i][x]]] % LN);
                        if (seen[bnk]) continue;
                        seen[bnk] = 1;
                        if (bank_owner[bnk] < 0 || aug(bank_owner[bnk])) { bank_owner[bnk] = i; pick[i] = (int)x; return true; }
                    }
                    return false;
                };
                for (int i = 0; i < n; ++i) {
                    if (rem[i].empty()) continue;
                    seen.assign(LN, 0);
                    aug(i);
                }
                // pick[] may be stale for members re-routed during augmentation: rebuild from bank_owner
                std::fill(pick.begin(), pick.end(), -1);
                for (int bnk = 0; bnk < LN; ++bnk) {
                    const int i = bank_owner[bnk];
                    if (i < 0) continue;
                    for (size_t x = 0; x < rem[i].size(); ++x)
                        if ((int)(pos_of_var[t.col_idx[rem[i][x]]] % LN) == bnk) { pick[i] = (int)x; break; }
                }
                for (int i = 0; i < n; ++i) {
                    if (rem[i].empty()) continue;
                    const int x = pick[i] >= 0 ? pick[i] : 0;  // unmatched: take any (a conflict)
                    slot_of_edge[rem[i][x]] = placed[i]++;
                    rem[i].erase(rem[i].begin() + x);
                }
            }
        }
    }
    if (SUB == 32 && h->opt.debug_placement) {
        // extra shared-memory wavefronts per iteration caused by bank conflicts in the two gathers
        long long exA = 0, exB = 0, lbA = 0;
        const int ncg = (t.M + 31) / 32, nvg = (t.N + 31) / 32;
        for (int cg = 0; cg < ncg; ++cg) {
            int cnt[32][32] = {};
            int maxd = 0;
            for (int i = 0; i < 32 && cg * 32 + i < t.M; ++i) {
                const int r = corder[cg * 32 + i];
                for (int e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) { cnt[slot_of_edge[e]][pos_of_var[t.col_idx[e]] % 32]++; maxd = std::max(maxd, slot_of_edge[e] + 1); }
            }
            for (int j = 0; j < maxd; ++j) { int mx = 1; for (int b = 0; b < 32; ++b) mx = std::max(mx, cnt[j][b]); exA += mx - 1; }
            for (int b = 0; b < 32; ++b) { int tot = 0; for (int j = 0; j < maxd; ++j) tot += cnt[j][b]; lbA += std::max(0, tot - maxd); }
        }
        for (int vg = 0; vg < nvg; ++vg) {
            int cnt[32][32] = {};
            int maxd = 0;
            for (int i = 0; i < 32 && vg * 32 + i < t.N; ++i) {
                const int v = vorder[vg * 32 + i];
                for (int k = t.col_ptr[v]; k < t.col_ptr[v + 1]; ++k) {
                    if (k - t.col_ptr[v] >= 32) break;
                    cnt[k - t.col_ptr[v]][crank_of_chk[t.vn_edge[k] >> kPosBits] % 32]++;
                    maxd = std::max(maxd, k - t.col_ptr[v] + 1);
                }
            }
            for (int k = 0; k < maxd; ++k) { int mx = 1; for (int b = 0; b < 32; ++b) mx = std::max(mx, cnt[k][b]); exB += mx - 1; }
        }
        std::fprintf(stderr, "[ldpc_b200] placement: extra wavefronts per iteration: check-pass gathers %lld (bank imbalance alone: %lld), variable-pass gathers %lld\n", exA, lbA, exB);
    }
    // check pass: T-row byte offset of every edge, [warp][slot][quad][h][4]; padding -> dummy row PD
    std::vector<uint32_t> cn_tab((size_t)W * pl.cn_stride, (uint32_t)PD * G * 4u);
    for (int rank = 0; rank < t.M; ++rank) {
        const int cs = rank / NL, nl = rank % NL, w = nl / SUB, hh = nl % SUB;
        const int r = corder[rank], e0 = t.row_ptr[r], dc = cdegf(r);
        for (int e = e0; e < e0 + dc; ++e) {
            const int j = slot_of_edge[e];
            cn_tab[(size_t)w * pl.cn_stride + ((size_t)(qoff[cs] + j / 4) * SUB + hh) * 4 + (j & 3)] =
                pos_of_var[t.col_idx[e]] * (uint32_t)(G * 4);
        }
    }
    // variable pass: byte offset of the R element of every edge in ascending-row order, [warp][slot][quad][h][4]
    std::vector<uint32_t> vn_tab((size_t)W * pl.vn_stride, (uint32_t)RD * 128u);
    for (int rank = 0; rank < t.N; ++rank) {
        const int s = rank / NL, nl = rank % NL, w = nl / SUB, hh = nl % SUB;
        const int v = vorder[rank];
        for (int k = t.col_ptr[v]; k < t.col_ptr[v + 1]; ++k) {
            const uint32_t chk = t.vn_edge[k] >> kPosBits, pos = t.vn_edge[k] & ((1u << kPosBits) - 1u);
            const int crank = (int)crank_of_chk[chk];
            const int ccs = crank / NL, cnl = crank % NL, cw = cnl / SUB, ch = cnl % SUB;
            const uint32_t row = (uint32_t)cw * pl.r_rows + (uint32_t)coff[ccs] + (uint32_t)slot_of_edge[t.row_ptr[chk] + pos];
            const int kk = k - t.col_ptr[v];
            vn_tab[(size_t)w * pl.vn_stride + ((size_t)(voff[s] + kk / 4) * SUB + hh) * 4 + (kk & 3)] =
                row * 128u + (uint32_t)(ch * G * 4);
        }
    }
    // 16-bit tables for the static-profile kernel (Test.cpp's code, G = 8): entries are row indices,
    // [slot][j/8][lane][8] for the check pass, [slot][lane][2|4|8] for the variable pass.  They cut the
    // table wavefronts by ~0.6 per edge but cost one unpack op per edge; measured 4.40 ms vs 4.26 ms with
    // 32-bit tables on cfg2 (profiles/r01_t16_experiment.txt), so they are off unless LDPC_B200_GRP_T16 is set.
    {
        // Test.cpp's code (G = 8, tables on chip): opt-in, measured slower.  Regular (3,6) N = 8192 (G = 1, tables
        // read through L1 from L2): the kernel is bound by L1 data-stage wavefronts, of which the 32-bit tables are
        // a third, so 16-bit entries are the default there (LDPC_B200_GRP_NO_T16 turns them off).
        auto prof_match = [&](auto prof) {
            using P = decltype(prof);
            if (CS != P::CS || VS != P::VS) return false;
            for (int i = 0; i < CS; ++i) if (h->g_cdeg[i] != P::cdeg(i)) return false;
            for (int i = 0; i < VS; ++i) if (h->g_vdeg[i] != P::vdeg(i)) return false;
            return true;
        };
        bool match = false;
        if (!keep_order && !h->opt.grp_no_profile) {
            if (G == 8 && pl.tab_smem && !pl.y_smem && h->opt.grp_t16) match = prof_match(ProfileWimax34B576{});
            if (G == 1 && !pl.tab_smem && !h->opt.grp_no_t16) match = prof_match(ProfileRegular36N8192{});
        }
        if (match) {
            std::vector<int> ooff(CS + 1, 0), vbyte(VS + 1, 0);
            for (int cs = 0; cs < CS; ++cs) ooff[cs + 1] = ooff[cs] + (h->g_cdeg[cs] + 7) / 8;
            auto sd_of = [](int d) { return d <= 2 ? 4 : (d <= 4 ? 8 : 16); };
            for (int sidx = 0; sidx < VS; ++sidx) vbyte[sidx + 1] = vbyte[sidx] + SUB * sd_of(h->g_vdeg[sidx]);
            const int cn_words = ooff[CS] * SUB * 4, vn_words = (vbyte[VS] + 15) / 16 * 4;
            std::vector<uint16_t> c16((size_t)W * cn_words * 2, (uint16_t)PD), v16((size_t)W * vn_words * 2, (uint16_t)(RD * SUB));
            for (int rank = 0; rank < t.M; ++rank) {
                const int cs = rank / NL, nl = rank % NL, w = nl / SUB, hh = nl % SUB;
                const int r = corder[rank], e0 = t.row_ptr[r], dc = cdegf(r);
                for (int e = e0; e < e0 + dc; ++e) {
                    const int j = slot_of_edge[e];
                    c16[(size_t)w * cn_words * 2 + ((size_t)(ooff[cs] + j / 8) * SUB + hh) * 8 + (j & 7)] = (uint16_t)pos_of_var[t.col_idx[e]];
                }
            }
            for (int rank = 0; rank < t.N; ++rank) {
                const int sidx = rank / NL, nl = rank % NL, w = nl / SUB, hh = nl % SUB;
                const int v = vorder[rank];
                const int sd = sd_of(h->g_vdeg[sidx]);
                for (int k = t.col_ptr[v]; k < t.col_ptr[v + 1]; ++k) {
                    const uint32_t chk = t.vn_edge[k] >> kPosBits, pos = t.vn_edge[k] & ((1u << kPosBits) - 1u);
                    const int crank = (int)crank_of_chk[chk];
                    const int ccs = crank / NL, cnl = crank % NL, cw = cnl / SUB, ch = cnl % SUB;
                    const uint32_t row = (uint32_t)cw * pl.r_rows + (uint32_t)coff[ccs] + (uint32_t)slot_of_edge[t.row_ptr[chk] + pos];
                    v16[(size_t)w * vn_words * 2 + (size_t)(vbyte[sidx] + hh * sd) / 2 + (k - t.col_ptr[v])] = (uint16_t)(row * SUB + ch);
                }
            }
            if ((size_t)(PD + 1) < 65536 && (size_t)(RD + 1) * SUB < 65536) {
                cn_tab.assign((size_t)W * cn_words, 0u);
                vn_tab.assign((size_t)W * vn_words, 0u);
                std::memcpy(cn_tab.data(), c16.data(), cn_tab.size() * 4);
                std::memcpy(vn_tab.data(), v16.data(), vn_tab.size() * 4);
                h->plan.t16 = true;
                h->plan.cn_stride = cn_words;
                h->plan.vn_stride = vn_words;
            }
        }
    }
    CU_TRY(cudaMalloc(&h->dg_cn_tab, cn_tab.size() * 4));
    CU_TRY(cudaMalloc(&h->dg_vn_tab, vn_tab.size() * 4));
    CU_TRY(cudaMalloc(&h->dg_var_of_pos, var_of_pos.size() * 4));
    CU_TRY(cudaMalloc(&h->dg_pos_of_var, pos_of_var.size() * 4));
    CU_TRY(cudaMemcpy(h->dg_cn_tab, cn_tab.data(), cn_tab.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dg_vn_tab, vn_tab.data(), vn_tab.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dg_var_of_pos, var_of_pos.data(), var_of_pos.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dg_pos_of_var, pos_of_var.data(), pos_of_var.size() * 4, cudaMemcpyHostToDevice));
    h->table_bytes += (cn_tab.size() + vn_tab.size() + var_of_pos.size() + pos_of_var.size()) * 4;
    h->group_ready = true;
    return LDPC_B200_OK;
}

// 0 / cudaError_t / kNoKernel from a launcher in k_*.cu  ->  C-ABI status
int launch_status(int rc, const char* what) {
    if (rc == 0) return LDPC_B200_OK;
    if (rc == kNoKernel) return fail(LDPC_B200_ERR_UNSUPPORTED, std::string("no ") + what + " kernel instantiated for this shape");
    if (rc == kNoCluster) return fail(LDPC_B200_ERR_UNSUPPORTED, "no 8-CTA cluster of this size can be resident");
    return fail(LDPC_B200_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorName((cudaError_t)rc) + " (" + cudaGetErrorString((cudaError_t)rc) + ")");
}

int launch_group(const Plan& pl, const GroupParams& q, int grid, bool allow_profile, cudaStream_t stream) {
    GroupSel sel;
    sel.G = pl.G; sel.dmax = pl.dmax; sel.threads = pl.threads; sel.CS = pl.CS; sel.VS = pl.VS;
    sel.tab_smem = pl.tab_smem; sel.y_smem = pl.y_smem; sel.t16 = pl.t16;
    sel.allow_profile = allow_profile;
    return launch_status(k_launch_group(sel, q, grid, pl.smem, stream), "group");
}

// ---- QC layout (see ldpc_qc.cuh) ---------------------------------------------------------------------
// The QC builders recognise H as a grid of z x z blocks (zero or one cyclically shifted identity each), deal groups
// of SUB consecutive rows / columns of a block to (slot, warp) in degree order and write one warp-uniform base per
// (warp, slot, edge).
// Block structure of H for block size z: rows[br] = the circulants (block column, shift) of block row br in
// ascending column order.  false when H is not a grid of z x z blocks that are zero or one shifted identity.
bool qc_blocks(const HostTables& t, int z, std::vector<std::vector<QcBlk>>* rows_out) {
    if (z < 1 || t.M % z || t.N % z) return false;
    const int MB = t.M / z, NB = t.N / z;
    std::vector<std::vector<QcBlk>> rows(MB);
    for (int br = 0; br < MB; ++br) {
        const int r0 = br * z;
        std::vector<bool> used(NB, false);
        for (int e = t.row_ptr[r0]; e < t.row_ptr[r0 + 1]; ++e) {
            const int bc = t.col_idx[e] / z;
            if (used[bc]) return false;
            used[bc] = true;
            rows[br].push_back({bc, t.col_idx[e] % z});
        }
        std::vector<int> want(rows[br].size());
        for (int r = 0; r < z; ++r) {
            if (t.row_ptr[r0 + r + 1] - t.row_ptr[r0 + r] != (int)rows[br].size()) return false;
            for (size_t j = 0; j < rows[br].size(); ++j) want[j] = rows[br][j].bc * z + (r + rows[br][j].s) % z;
            std::sort(want.begin(), want.end());
            for (size_t j = 0; j < want.size(); ++j)
                if (t.col_idx[t.row_ptr[r0 + r] + (int)j] != want[j]) return false;
        }
    }
    *rows_out = std::move(rows);
    return true;
}

// Run-time-profile QC tables for G codewords per CTA (see ldpc_qcg.cuh).  false = this (z, G) does not work.
bool qcg_build(const HostTables& t, int z, int G, const std::vector<std::vector<QcBlk>>& rows, QcgParams* out, QcgWarpTab* tabs,
               size_t* smem_out) {
    const int SUB = 32 / G;
    if (z % SUB) return false;
    const uint32_t ROWB = (uint32_t)G * 4u, RS = (uint32_t)(z + SUB) * ROWB;
    const int MB = t.M / z, NB = t.N / z, gpb = z / SUB;
    const int cgroups = MB * gpb, vgroups = NB * gpb;
    // warps per CTA: the count that leaves the fewest empty (warp, slot) places -- groups need not divide evenly, the
    // last slot of some warps stays empty
    int W = 0;
    double best_u = 0.0;
    for (int w = kQcgMaxW; w >= 4; --w) {
        const int cs = (cgroups + w - 1) / w, vs = (vgroups + w - 1) / w;
        if (cs > kQcgMaxCS || vs > kQcgMaxVS) continue;
        const double u = (0.6 * cgroups / (double)(cs * w) + 0.4 * vgroups / (double)(vs * w)) * (0.9 + 0.1 * w / kQcgMaxW);
        if (u > best_u + 1e-9) { best_u = u; W = w; }
    }
    if (!W) return false;
    const int CS = (cgroups + W - 1) / W, VS = (vgroups + W - 1) / W;
    struct Col { int br, j, s; };
    std::vector<std::vector<Col>> cols(NB);
    for (int br = 0; br < MB; ++br)
        for (int j = 0; j < (int)rows[br].size(); ++j) cols[rows[br][j].bc].push_back({br, j, rows[br][j].s});
    std::vector<int> border(MB), corder(NB);
    for (int i = 0; i < MB; ++i) border[i] = i;
    for (int i = 0; i < NB; ++i) corder[i] = i;
    std::stable_sort(border.begin(), border.end(), [&](int a, int b) { return rows[a].size() > rows[b].size(); });
    std::stable_sort(corder.begin(), corder.end(), [&](int a, int b) { return cols[a].size() > cols[b].size(); });
    QcgParams& q = *out;
    std::memset(&q, 0, sizeof(q));
    std::memset(tabs, 0, sizeof(QcgWarpTab) * kQcgMaxW);
    // slot degrees = the largest degree among the slot's groups (groups are in descending degree order)
    for (int p = 0; p < cgroups; ++p) q.cdeg[p / W] = std::max<uint8_t>(q.cdeg[p / W], (uint8_t)rows[border[p / gpb]].size());
    for (int p = 0; p < vgroups; ++p) q.vdeg[p / W] = std::max<uint8_t>(q.vdeg[p / W], (uint8_t)cols[corder[p / gpb]].size());
    int ce = 0, ve = 0;
    std::vector<int> coff(CS + 1, 0), voff(VS + 1, 0);
    for (int cs = 0; cs < CS; ++cs) { if (q.cdeg[cs] < 1 || q.cdeg[cs] > kQcgMaxCD) return false; coff[cs] = ce; ce += (q.cdeg[cs] + 3) & ~3; }
    for (int s = 0; s < VS; ++s) { if (q.vdeg[s] < 1 || q.vdeg[s] > kQcgMaxVD) return false; voff[s] = ve; ve += (q.vdeg[s] + 3) & ~3; }
    if (ce > kQcgMaxCE || ve > kQcgMaxVE) return false;
    // R blocks of a block row: as many as the largest slot degree any of its groups is processed with
    std::vector<int> dpad(MB, 0), eb0(MB + 1, 0);
    for (int p = 0; p < cgroups; ++p) dpad[border[p / gpb]] = std::max(dpad[border[p / gpb]], (int)q.cdeg[p / W]);
    for (int br = 0; br < MB; ++br) eb0[br + 1] = eb0[br] + dpad[br];
    const uint32_t t_bytes = (uint32_t)NB * (z + SUB) * ROWB, r_bytes = (uint32_t)eb0[MB] * RS;
    const uint32_t zero_row = t_bytes + r_bytes, inf_row = zero_row + 128u;
    for (int p = 0; p < cgroups; ++p) {
        const int br = border[p / gpb], g = p % gpb, slot = p / W, w = p % W, r0 = g * SUB;
        QcgWarpTab& tb = tabs[w];
        tb.cn_r[slot] = t_bytes + (uint32_t)eb0[br] * RS + (uint32_t)(SUB + r0) * ROWB;
        tb.cact |= 1u << slot;
        if (g == gpb - 1) tb.cdup |= 1u << slot;
        for (int j = 0; j < q.cdeg[slot]; ++j)
            tb.cn_t[coff[slot] + j] = j < (int)rows[br].size()
                ? (uint32_t)(rows[br][j].bc * (z + SUB) + (r0 + rows[br][j].s) % z) * ROWB
                : inf_row;  // padded edge: T = -inf is neutral for the minima, the sign parity and the syndrome
    }
    for (int p = 0; p < vgroups; ++p) {
        const int bc = corder[p / gpb], g = p % gpb, slot = p / W, w = p % W, i0 = g * SUB;
        const int d = (int)cols[bc].size();
        QcgWarpTab& tb = tabs[w];
        tb.vn_t[slot] = (uint32_t)(bc * (z + SUB) + i0) * ROWB;
        tb.var0[slot] = (uint32_t)(bc * z + i0);
        tb.vact |= 1u << slot;
        if (g == 0) tb.vdup |= 1u << slot;
        for (int k = 0; k < q.vdeg[slot]; ++k) {
            if (k >= d) { tb.vn_r[voff[slot] + k] = zero_row; continue; }
            const Col& cd = cols[bc][k];
            int m = ((i0 - cd.s) % z + z) % z;
            if (m > z - SUB) m -= z;
            tb.vn_r[voff[slot] + k] = t_bytes + (uint32_t)(eb0[cd.br] + cd.j) * RS + (uint32_t)(SUB + m) * ROWB;
        }
    }
    q.N = t.N; q.Z = z; q.W = W; q.CS = CS; q.VS = VS;
    q.RS = RS; q.WRAP = (uint32_t)z * ROWB;
    q.t_bytes = t_bytes; q.r_bytes = r_bytes;
    *smem_out = (size_t)t_bytes + r_bytes + 256 + (size_t)W * sizeof(QcgWarpTab);
    return true;
}

// Slots of the __constant__ table banks, per device: which handle owns each.
constexpr int kQcMaxDevices = 64;
std::mutex g_qc_mu;
const void* g_qc_owner[kQcMaxDevices][kQcTabSlots] = {};

int qc_acquire_slot(const void* owner, int device) {
    if (device < 0 || device >= kQcMaxDevices) return -1;
    std::lock_guard<std::mutex> lk(g_qc_mu);
    for (int s = 0; s < kQcTabSlots; ++s)
        if (!g_qc_owner[device][s]) { g_qc_owner[device][s] = owner; return s; }
    return -1;
}

void qc_release_slot(const void* owner, int device, int slot) {
    if (device < 0 || device >= kQcMaxDevices || slot < 0 || slot >= kQcTabSlots) return;
    std::lock_guard<std::mutex> lk(g_qc_mu);
    if (g_qc_owner[device][slot] == owner) g_qc_owner[device][slot] = nullptr;
}

// The compiled profiles: rate x (z, G, W), one table per translation unit (k_qc.cu built once per rate).
const std::vector<QcProfileEntry>& qc_profiles() {
    static const std::vector<QcProfileEntry> all = [] {
        std::vector<QcProfileEntry> v;
        for (auto fn : {&qc_profiles_34B, &qc_profiles_34A, &qc_profiles_23B, &qc_profiles_23A, &qc_profiles_12, &qc_profiles_56}) {
            int n = 0;
            const QcProfileEntry* e = fn(&n);
            v.insert(v.end(), e, e + n);
        }
        return v;
    }();
    return all;
}

// Finds the compiled profile the code matches, builds its tables, takes a slot of the constant bank and uploads them
// (current device = the handle's).  false = use another path.
bool qc_prepare(ldpc_b200_decoder* h) {
    const HostTables& t = h->host;
    std::vector<std::vector<QcBlk>> rows;
    int rows_z = 0;
    std::vector<unsigned char> tab;
    const std::vector<QcProfileEntry>& profiles = qc_profiles();
    for (int k = 0; k < (int)profiles.size(); ++k) {
        const QcProfileEntry& pe = profiles[k];
        if (t.M % pe.z || t.N % pe.z) continue;
        if (h->opt.qc_prefer_g ? pe.G != h->opt.qc_prefer_g : (pe.z == 24 && pe.G == 4)) continue;  // (24, 4, 6) only on request
        if (rows_z != pe.z) { rows.clear(); rows_z = pe.z; if (!qc_blocks(t, pe.z, &rows)) rows.clear(); }
        if (rows.empty()) continue;
        if (!pe.build(t, rows, &h->qc, &tab, &h->qc_smem)) continue;
        if (tab.size() > (size_t)kQcBankBytes) continue;
        DeviceGuard guard(h->device);
        if (!guard.ok) return false;
        const int slot = qc_acquire_slot(h, h->device);
        if (slot < 0) return false;
        if (pe.upload(slot, tab.data(), tab.size()) != 0) {
            (void)cudaGetLastError();
            qc_release_slot(h, h->device, slot);
            return false;
        }
        h->qc_slot = slot;
        h->qc.tab_slot = slot;
        h->qc_kind = k;
        h->table_bytes += tab.size();
        return true;
    }
    return false;
}

// The warp-per-codeword kernel (ldpc_qcw.cuh; 802.16e codes with z = 24 or 32, their circulants compiled in): finds the
// instantiation whose code this handle's H is.  false = there is none (the lockstep kernel serves every regime).
bool qcw_prepare(ldpc_b200_decoder* h) {
    const HostTables& t = h->host;
    if (h->qc_state != 1) return false;
    int n = 0;
    const QcwProfileEntry* profiles = qcw_profiles(&n);
    std::vector<std::vector<QcBlk>> rows;
    int rows_z = 0;
    for (int k = 0; k < n; ++k) {
        const QcwProfileEntry& pe = profiles[k];
        if (t.N != 24 * pe.z) continue;
        if (rows_z != pe.z) { rows.clear(); rows_z = pe.z; if (!qc_blocks(t, pe.z, &rows)) rows.clear(); }
        if (rows.empty()) continue;
        std::vector<uint32_t> syn;
        if (!pe.build(t, rows, &syn)) continue;
        // a multiple of four warps (one register-file partition each); 16 at z = 24, 12 at z = 32
        int warps = std::min<int>(kQcwMaxWarps, (int)(h->smem_optin / (size_t)pe.warp_bytes)) & ~3;
        if (h->opt.qcw_warps >= 4) warps = std::min(warps, h->opt.qcw_warps & ~3);   // (occupancy experiments)
        if (warps < 8 && h->opt.qcw_warps < 4) return false;
        DeviceGuard guard(h->device);
        if (!guard.ok) return false;
        if (cudaMalloc(&h->d_syn_tab, syn.size() * 4) != cudaSuccess ||
            cudaMemcpy(h->d_syn_tab, syn.data(), syn.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            (void)cudaGetLastError();
            return false;
        }
        std::memset(&h->qcw, 0, sizeof(h->qcw));
        h->qcw.N = t.N;
        h->qcw.syn_tab = h->d_syn_tab;
        h->qcw_kind = k;
        h->qcw_warps = warps;
        h->table_bytes += syn.size() * 4;
        return true;
    }
    return false;
}

// The group-of-warps-per-codeword kernel (ldpc_qcm.cuh): any quasi-cyclic code with 24 block columns and the degree
// sequences of an 802.16e rate, z <= 96.  Builds the tables, takes a slot of that unit's constant bank.
std::mutex g_qcm_mu;
const void* g_qcm_owner[kQcMaxDevices][kQcTabSlots] = {};

bool qcm_prepare(ldpc_b200_decoder* h) {
    const HostTables& t = h->host;
    if (t.N % 24) return false;
    const int z = t.N / 24;
    std::vector<std::vector<QcBlk>> rows;
    if (z > 96 || !qc_blocks(t, z, &rows)) return false;
    int n = 0;
    const QcmProfileEntry* profiles = qcm_profiles(&n);
    for (int k = 0; k < n; ++k) {
        const QcmProfileEntry& pe = profiles[k];
        std::vector<unsigned char> tab;
        int groups = 0;
        if (!pe.build(t, z, rows, h->smem_optin, &h->qcm, &tab, &groups)) continue;
        DeviceGuard guard(h->device);
        if (!guard.ok || h->device < 0 || h->device >= kQcMaxDevices) return false;
        int slot = -1;
        {
            std::lock_guard<std::mutex> lk(g_qcm_mu);
            for (int s2 = 0; s2 < kQcTabSlots && slot < 0; ++s2)
                if (!g_qcm_owner[h->device][s2]) { g_qcm_owner[h->device][s2] = h; slot = s2; }
        }
        if (slot < 0) return false;
        if (pe.upload(slot, tab.data(), tab.size()) != 0) {
            (void)cudaGetLastError();
            std::lock_guard<std::mutex> lk(g_qcm_mu);
            g_qcm_owner[h->device][slot] = nullptr;
            return false;
        }
        h->qcm.tab_slot = slot;
        h->qcm_slot = slot;
        h->qcm_kind = k;
        h->qcm_groups = groups;
        h->qcm_tab = tab;
        // several codewords per group (ldpc_ms_qcm_multi_kernel) where one leaves most lanes of the last warp idle and
        // shared memory has room for more: the measured best of the reference's family (profiles/r02_qcm_pack.txt:
        // z = 36 3.6-5.5 -> 4.9-7.0 Gbit/s, z = 44 4.5-6.8 -> 5.2-7.4, z = 68 / 72 / 76 +11-15 % where listed; every other
        // block size is fastest with one codeword per group)
        h->qcm_multi_groups = 0;
        {
            static const int kRateOfKind[6] = {4, 3, 2, 1, 0, 5};   // qcm_profiles() order: 3/4B, 3/4A, 2/3B, 2/3A, 1/2, 5/6
            struct Best { int z; int pack[6]; };                    // by rate: 1/2, 2/3A, 2/3B, 3/4A, 3/4B, 5/6
            static const Best kBest[] = {{36, {3, 3, 3, 3, 3, 3}}, {44, {2, 2, 2, 2, 2, 2}}, {68, {3, 3, 3, 3, 3, 1}},
                                         {72, {3, 3, 3, 1, 1, 1}}, {76, {2, 1, 1, 1, 1, 1}}};
            int pack = h->opt.qcm_pack;
            if (pack < 0) {
                pack = 1;
                for (const Best& b : kBest)
                    if (b.z == z && k < 6) pack = b.pack[kRateOfKind[k]];
            }
            QcmParams qm;
            int gm = 0;
            if (pack >= 2 && pe.multi_geometry(h->qcm, pack, h->smem_optin, &qm, &gm)) { h->qcm_multi = qm; h->qcm_multi_groups = gm; }
        }
        h->table_bytes += tab.size();
        return true;
    }
    return false;
}

// Run-time-profile QC path: picks G (most codewords per CTA with two CTAs per SM, else one CTA), builds and uploads.
bool qcg_prepare(ldpc_b200_decoder* h) {
    const HostTables& t = h->host;
    std::vector<int> zs;
    if (h->layer_z > 0) zs.push_back(h->layer_z);
    if (t.N % 24 == 0 && (zs.empty() || zs[0] != t.N / 24)) zs.push_back(t.N / 24);  // 802.16e: 24 block columns
    for (int z : zs) {
        std::vector<std::vector<QcBlk>> rows;
        if (!qc_blocks(t, z, &rows)) continue;
        for (int per_sm : {2, 1}) {
            for (int G : {8, 4, 2}) {
                size_t smem = 0;
                if (!qcg_build(t, z, G, rows, &h->qcg, h->qcg_tab, &smem)) continue;
                if ((size_t)per_sm * (smem + 2048) > h->smem_optin + 1024 || smem + 1024 > h->smem_optin) continue;
                DeviceGuard guard(h->device);
                if (!guard.ok) return false;
                if (!h->dq_tabs && cudaMalloc(&h->dq_tabs, sizeof(h->qcg_tab)) != cudaSuccess) { (void)cudaGetLastError(); return false; }
                if (cudaMemcpy(h->dq_tabs, h->qcg_tab, sizeof(h->qcg_tab), cudaMemcpyHostToDevice) != cudaSuccess) { (void)cudaGetLastError(); return false; }
                h->qcg.tabs = h->dq_tabs;
                h->qcg_G = G; h->qcg_ctas_per_sm = per_sm; h->qcg_smem = smem;
                h->table_bytes += sizeof(h->qcg_tab);
                return true;
            }
        }
    }
    return false;
}

// ---- WARP layout (see ldpc_warp.cuh) ------------------------------------------------------------------
int warp_sub_width(const HostTables& t) { return t.max_row_weight <= 8 ? 8 : (t.max_row_weight <= 16 ? 16 : (t.max_row_weight <= 32 ? 32 : 0)); }

int upload_warp_tables(ldpc_b200_decoder* h) {
    if (h->warp_ready) return LDPC_B200_OK;
    const HostTables& t = h->host;
    const int SW = warp_sub_width(t);
    std::vector<uint32_t> cn_col((size_t)t.M * SW, 0xffffffffu), vn_pos((size_t)std::max(t.nnz, 1));
    for (int r = 0; r < t.M; ++r)
        for (int e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) cn_col[(size_t)r * SW + (e - t.row_ptr[r])] = (uint32_t)t.col_idx[e];
    for (int k = 0; k < t.nnz; ++k) vn_pos[k] = (t.vn_edge[k] >> kPosBits) * (uint32_t)SW + (t.vn_edge[k] & ((1u << kPosBits) - 1u));
    CU_TRY(cudaMalloc(&h->dw_cn_col, cn_col.size() * 4));
    CU_TRY(cudaMalloc(&h->dw_vn_pos, vn_pos.size() * 4));
    CU_TRY(cudaMemcpy(h->dw_cn_col, cn_col.data(), cn_col.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dw_vn_pos, vn_pos.data(), vn_pos.size() * 4, cudaMemcpyHostToDevice));
    h->table_bytes += (cn_col.size() + vn_pos.size()) * 4;
    h->w_sw = SW;
    h->warp_ready = true;
    return LDPC_B200_OK;
}

// ---- CLUSTER layout (one codeword per 8-CTA cluster, DSMEM gathers; see ldpc_cluster.cuh) ------------
struct ClShape {
    int W = 32, CS = 0, VS = 0, dmax = 8;
    int cn_stride = 0, vn_stride = 0, r_rows = 0;
    size_t smem = 0;
};

// Partition: checks in contiguous row blocks; a variable goes to the CTA holding most of its checks
// (capacity-bounded), which keeps e.g. a dual-diagonal parity part entirely local.
void cluster_partition(const HostTables& t, std::vector<int>* chk_rank, std::vector<int>* var_rank) {
    const int CL = kClusterSize;
    chk_rank->resize(t.M);
    var_rank->assign(t.N, -1);
    for (int r = 0; r < t.M; ++r) (*chk_rank)[r] = (int)((long long)r * CL / t.M);
    const int cap = (t.N + CL - 1) / CL;
    std::vector<int> fill(CL, 0);
    std::vector<std::pair<int, int>> order;  // (-best count, variable)
    std::vector<std::array<int, kClusterSize>> hist(t.N);
    for (int v = 0; v < t.N; ++v) {
        hist[v].fill(0);
        for (int k = t.col_ptr[v]; k < t.col_ptr[v + 1]; ++k) hist[v][(*chk_rank)[t.vn_edge[k] >> kPosBits]]++;
        order.emplace_back(-*std::max_element(hist[v].begin(), hist[v].end()), v);
    }
    std::stable_sort(order.begin(), order.end());
    for (auto& ov : order) {
        const int v = ov.second;
        int best = -1;
        for (int r = 0; r < CL; ++r)
            if (fill[r] < cap && (best < 0 || hist[v][r] > hist[v][best] || (hist[v][r] == hist[v][best] && fill[r] < fill[best]))) best = r;
        (*var_rank)[v] = best;
        fill[best]++;
    }
}

bool cluster_shape(const HostTables& t, size_t smem_limit, ClShape* out, std::vector<int>* chk_rank, std::vector<int>* var_rank) {
    if (t.max_row_weight > 16 || t.max_row_weight < 1 || t.M < kClusterSize || t.N < kClusterSize) return false;
    cluster_partition(t, chk_rank, var_rank);
    const int CL = kClusterSize;
    ClShape sh;
    sh.dmax = t.max_row_weight <= 8 ? 8 : 16;
    const int NL = sh.W * 32;
    std::vector<std::vector<int>> cd(CL), vd(CL);
    for (int r = 0; r < t.M; ++r) cd[(*chk_rank)[r]].push_back(t.row_ptr[r + 1] - t.row_ptr[r]);
    for (int v = 0; v < t.N; ++v) vd[(*var_rank)[v]].push_back(t.col_ptr[v + 1] - t.col_ptr[v]);
    size_t cmax = 0, vmax = 0;
    for (int r = 0; r < CL; ++r) {
        std::sort(cd[r].begin(), cd[r].end(), std::greater<int>());
        std::sort(vd[r].begin(), vd[r].end(), std::greater<int>());
        cmax = std::max(cmax, cd[r].size()); vmax = std::max(vmax, vd[r].size());
    }
    sh.CS = (int)((cmax + NL - 1) / NL);
    sh.VS = (int)((vmax + NL - 1) / NL);
    if (sh.CS > kGrpMaxCS || sh.VS > kGrpMaxVS) return false;
    long long rrows = 0, cquads = 0, vquads = 0;
    for (int cs = 0; cs < sh.CS; ++cs) {
        int d = 0;
        for (int r = 0; r < CL; ++r) if ((size_t)cs * NL < cd[r].size()) d = std::max(d, cd[r][(size_t)cs * NL]);
        rrows += d; cquads += (d + 3) / 4;
    }
    for (int sidx = 0; sidx < sh.VS; ++sidx) {
        int d = 0;
        for (int r = 0; r < CL; ++r) if ((size_t)sidx * NL < vd[r].size()) d = std::max(d, vd[r][(size_t)sidx * NL]);
        vquads += (d + 3) / 4;
    }
    sh.r_rows = (int)rrows;
    sh.cn_stride = (int)(cquads * 32 * 4);
    sh.vn_stride = (int)(vquads * 32 * 4);
    sh.smem = (((size_t)(sh.VS * NL + 1) * 4 + 127) & ~(size_t)127) + ((size_t)sh.W * rrows + 1) * 128;
    if (sh.smem + 1024 > smem_limit) return false;
    if ((size_t)(sh.VS * NL + 1) * 4 >= (1u << 24) || ((size_t)sh.W * rrows + 1) * 128 >= (1u << 24)) return false;
    *out = sh;
    return true;
}

int upload_cluster_tables(ldpc_b200_decoder* h) {
    if (h->cluster_ready) return LDPC_B200_OK;
    const HostTables& t = h->host;
    const Plan& pl = h->plan;
    const int CL = kClusterSize, W = pl.W, NL = W * 32, CS = pl.CS, VS = pl.VS;
    std::vector<int> chk_rank, var_rank;
    cluster_partition(t, &chk_rank, &var_rank);
    auto vdegf = [&](int c) { return t.col_ptr[c + 1] - t.col_ptr[c]; };
    auto cdegf = [&](int r) { return t.row_ptr[r + 1] - t.row_ptr[r]; };
    std::vector<std::vector<int>> corder(CL), vorder(CL);
    for (int r = 0; r < t.M; ++r) corder[chk_rank[r]].push_back(r);
    for (int v = 0; v < t.N; ++v) vorder[var_rank[v]].push_back(v);
    for (int r = 0; r < CL; ++r) {
        std::stable_sort(corder[r].begin(), corder[r].end(), [&](int a, int b) { return cdegf(a) > cdegf(b); });
        std::stable_sort(vorder[r].begin(), vorder[r].end(), [&](int a, int b) { return vdegf(a) > vdegf(b); });
    }
    std::memset(h->g_vdeg, 0, sizeof(h->g_vdeg));
    std::memset(h->g_cdeg, 0, sizeof(h->g_cdeg));
    std::vector<int> coff(CS + 1, 0), qoff(CS + 1, 0), voff(VS + 1, 0);
    for (int cs = 0; cs < CS; ++cs) {
        int d = 0;
        for (int r = 0; r < CL; ++r) if ((size_t)cs * NL < corder[r].size()) d = std::max(d, cdegf(corder[r][(size_t)cs * NL]));
        h->g_cdeg[cs] = (uint8_t)d;
        coff[cs + 1] = coff[cs] + d;
        qoff[cs + 1] = qoff[cs] + (d + 3) / 4;
    }
    for (int sidx = 0; sidx < VS; ++sidx) {
        int d = 0;
        for (int r = 0; r < CL; ++r) if ((size_t)sidx * NL < vorder[r].size()) d = std::max(d, vdegf(vorder[r][(size_t)sidx * NL]));
        h->g_vdeg[sidx] = (uint8_t)d;
        voff[sidx + 1] = voff[sidx] + (d + 3) / 4;
    }
    const int PD = VS * NL, RD = W * pl.r_rows;
    // local rank (position) of every node inside its CTA
    std::vector<uint32_t> lpos_of_var(t.N), lrank_of_chk(t.M), var_of_pos((size_t)CL * PD, 0xffffffffu), out_addr(t.N);
    for (int r = 0; r < CL; ++r) {
        for (size_t i = 0; i < vorder[r].size(); ++i) {
            lpos_of_var[vorder[r][i]] = (uint32_t)i;
            var_of_pos[(size_t)r * PD + i] = (uint32_t)vorder[r][i];
            out_addr[vorder[r][i]] = ((uint32_t)r << 24) + (uint32_t)i * 4u;
        }
        for (size_t i = 0; i < corder[r].size(); ++i) lrank_of_chk[corder[r][i]] = (uint32_t)i;
    }
    std::vector<uint32_t> cn_tab((size_t)CL * W * pl.cn_stride), vn_tab((size_t)CL * W * pl.vn_stride);
    for (int r = 0; r < CL; ++r) {
        std::fill(cn_tab.begin() + (size_t)r * W * pl.cn_stride, cn_tab.begin() + (size_t)(r + 1) * W * pl.cn_stride,
                  ((uint32_t)r << 24) + (uint32_t)PD * 4u);      // padding -> this CTA's dummy T entry
        std::fill(vn_tab.begin() + (size_t)r * W * pl.vn_stride, vn_tab.begin() + (size_t)(r + 1) * W * pl.vn_stride,
                  ((uint32_t)r << 24) + (uint32_t)RD * 128u);    // padding -> this CTA's dummy R row
    }
    for (int c = 0; c < t.M; ++c) {
        const int r = chk_rank[c], lr = (int)lrank_of_chk[c];
        const int cs = lr / NL, nl = lr % NL, w = nl / 32, hh = nl % 32;
        const int e0 = t.row_ptr[c], dc = cdegf(c);
        for (int j = 0; j < dc; ++j) {
            const int v = t.col_idx[e0 + j];
            cn_tab[((size_t)r * W + w) * pl.cn_stride + ((size_t)(qoff[cs] + j / 4) * 32 + hh) * 4 + (j & 3)] =
                ((uint32_t)var_rank[v] << 24) + lpos_of_var[v] * 4u;
        }
    }
    for (int v = 0; v < t.N; ++v) {
        const int r = var_rank[v], lp = (int)lpos_of_var[v];
        const int sidx = lp / NL, nl = lp % NL, w = nl / 32, hh = nl % 32;
        for (int k = t.col_ptr[v]; k < t.col_ptr[v + 1]; ++k) {
            const uint32_t chk = t.vn_edge[k] >> kPosBits, pos = t.vn_edge[k] & ((1u << kPosBits) - 1u);
            const int cr = chk_rank[chk], clr = (int)lrank_of_chk[chk];
            const int ccs = clr / NL, cnl = clr % NL, cw = cnl / 32, ch = cnl % 32;
            const uint32_t row = (uint32_t)cw * pl.r_rows + (uint32_t)coff[ccs] + pos;
            const int kk = k - t.col_ptr[v];
            vn_tab[((size_t)r * W + w) * pl.vn_stride + ((size_t)(voff[sidx] + kk / 4) * 32 + hh) * 4 + (kk & 3)] =
                ((uint32_t)cr << 24) + row * 128u + (uint32_t)ch * 4u;
        }
    }
    CU_TRY(cudaMalloc(&h->dc_cn_tab, cn_tab.size() * 4));
    CU_TRY(cudaMalloc(&h->dc_vn_tab, vn_tab.size() * 4));
    CU_TRY(cudaMalloc(&h->dc_var_of_pos, var_of_pos.size() * 4));
    CU_TRY(cudaMalloc(&h->dc_out_addr, out_addr.size() * 4));
    CU_TRY(cudaMemcpy(h->dc_cn_tab, cn_tab.data(), cn_tab.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dc_vn_tab, vn_tab.data(), vn_tab.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dc_var_of_pos, var_of_pos.data(), var_of_pos.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dc_out_addr, out_addr.data(), out_addr.size() * 4, cudaMemcpyHostToDevice));
    h->table_bytes += (cn_tab.size() + vn_tab.size() + var_of_pos.size() + out_addr.size()) * 4;
    h->cluster_ready = true;
    return LDPC_B200_OK;
}

// ---- STREAM layout (see ldpc_stream.cuh) --------------------------------------------------------------
struct StreamShape { int np = 0, nb8 = 0, nb4 = 0, nb2 = 0; };

StreamShape stream_shape(const HostTables& t) {
    StreamShape sh;
    int n8 = 0, n4 = 0, n2 = 0;
    for (int v = 0; v < t.N; ++v) {
        const int d = t.col_ptr[v + 1] - t.col_ptr[v];
        if (d > 4) ++n8; else if (d > 2) ++n4; else ++n2;
    }
    sh.nb8 = n8; sh.nb4 = (n4 + 1) / 2; sh.nb2 = (n2 + 3) / 4;
    sh.np = sh.nb8 + 2 * sh.nb4 + 4 * sh.nb2;
    return sh;
}

int upload_stream_tables(ldpc_b200_decoder* h) {
    if (h->stream_ready) return LDPC_B200_OK;
    const HostTables& t = h->host;
    const StreamShape sh = stream_shape(t);
    std::vector<uint32_t> var_of_pos((size_t)sh.np, kStreamNone), pos_of_var(t.N);
    {
        int p8 = 0, p4 = sh.nb8, p2 = sh.nb8 + 2 * sh.nb4;
        for (int v = 0; v < t.N; ++v) {
            const int d = t.col_ptr[v + 1] - t.col_ptr[v];
            int& slot = d > 4 ? p8 : (d > 2 ? p4 : p2);
            var_of_pos[slot] = (uint32_t)v;
            pos_of_var[v] = (uint32_t)slot;
            ++slot;
        }
    }
    std::vector<uint32_t> cn_tab((size_t)t.M * 8, kStreamNone), vn_tab((size_t)(sh.nb8 + sh.nb4 + sh.nb2) * 8, kStreamNone);
    for (int r = 0; r < t.M; ++r)
        for (int e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) cn_tab[(size_t)r * 8 + (e - t.row_ptr[r])] = pos_of_var[t.col_idx[e]];
    for (int pos = 0; pos < sh.np; ++pos) {
        const uint32_t v = var_of_pos[pos];
        if (v == kStreamNone) continue;
        size_t base;  // first table entry of this variable
        if (pos < sh.nb8) base = (size_t)pos * 8;
        else if (pos < sh.nb8 + 2 * sh.nb4) base = (size_t)sh.nb8 * 8 + (size_t)(pos - sh.nb8) * 4;
        else base = (size_t)(sh.nb8 + sh.nb4) * 8 + (size_t)(pos - sh.nb8 - 2 * sh.nb4) * 2;
        for (int k = t.col_ptr[v]; k < t.col_ptr[v + 1]; ++k) {
            const uint32_t chk = t.vn_edge[k] >> kPosBits, j = t.vn_edge[k] & ((1u << kPosBits) - 1u);
            vn_tab[base + (size_t)(k - t.col_ptr[v])] = chk * 8u + j;  // ascending row = the reference's summation order
        }
    }
    CU_TRY(cudaMalloc(&h->ds_cn_tab, cn_tab.size() * 4));
    CU_TRY(cudaMalloc(&h->ds_vn_tab, vn_tab.size() * 4));
    CU_TRY(cudaMalloc(&h->ds_var_of_pos, var_of_pos.size() * 4));
    CU_TRY(cudaMalloc(&h->ds_pos_of_var, pos_of_var.size() * 4));
    CU_TRY(cudaMemcpy(h->ds_cn_tab, cn_tab.data(), cn_tab.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->ds_vn_tab, vn_tab.data(), vn_tab.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->ds_var_of_pos, var_of_pos.data(), var_of_pos.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->ds_pos_of_var, pos_of_var.data(), pos_of_var.size() * 4, cudaMemcpyHostToDevice));
    h->table_bytes += (cn_tab.size() + vn_tab.size() + var_of_pos.size() + pos_of_var.size()) * 4;
    h->s_np = sh.np; h->s_nb8 = sh.nb8; h->s_nb4 = sh.nb4; h->s_nb2 = sh.nb2;
    h->stream_ready = true;
    return LDPC_B200_OK;
}

// ---- TDMP layout (see ldpc_tdmp.cuh) ------------------------------------------------------------------
int tdmp_plan(ldpc_b200_decoder* h) {
    const HostTables& t = h->host;
    const int z = h->layer_z;
    TdmpPlan best;
    if (z < 1) return fail(LDPC_B200_ERR_UNSUPPORTED, "layered decoding needs the layer height (ldpc_b200_set_layer_height)");
    if (t.M % z || t.N % z) return fail(LDPC_B200_ERR_UNSUPPORTED, "layer height must divide M and N");
    const int L = t.M / z, VS = t.N / z;
    h->tdmp_big = false;
    {   // no column may appear twice inside a layer (the rows of a layer are updated concurrently)
        std::vector<int> seen(t.N, -1);
        for (int r = 0; r < t.M; ++r)
            for (int e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) {
                if (seen[t.col_idx[e]] == r / z) return fail(LDPC_B200_ERR_UNSUPPORTED, "layered decoding: a column appears twice inside a layer");
                seen[t.col_idx[e]] = r / z;
            }
    }
    const int force_g = h->opt.tdmp_g;
    const bool onchip_shape = L <= kTdmpMaxLayers && VS <= 32 && t.max_row_weight <= 20 && force_g != 32;  // (tdmp_g = 32 forces the any-size kernel: tests)
    for (int G : {4, 8, 16}) {
        if (!onchip_shape) break;
        if (force_g && G != force_g) continue;
        const int SUB = 32 / G;
        if (z % SUB) continue;
        TdmpPlan pl;
        pl.G = G; pl.W = z / SUB; pl.L = L; pl.VS = VS; pl.z = z;
        if (pl.W > 32) continue;
        pl.threads = pl.W * 32;
        long long quads = 0;
        for (int l = 0; l < L; ++l) {
            int d = 0;
            for (int r = l * z; r < (l + 1) * z; ++r) d = std::max(d, t.row_ptr[r + 1] - t.row_ptr[r]);
            pl.ldeg[l] = (uint8_t)d;
            pl.r_rows += d;
            quads += (d + 3) / 4;
        }
        pl.cn_stride = (int)(quads * SUB * 4);
        const size_t tbytes = ((size_t)(t.N + SUB) * G * 4 + 127) & ~(size_t)127;
        pl.smem = tbytes + (size_t)pl.W * pl.r_rows * 128 + (size_t)pl.W * pl.cn_stride * 4;
        if (pl.smem + 512 > h->smem_optin) continue;
        const size_t sm_total = h->smem_optin + 1024;  // per-SM capacity; every resident CTA reserves 1 KB
        pl.ctas_per_sm = (int)std::min<size_t>({sm_total / (pl.smem + 1024), (size_t)(2048 / pl.threads), (size_t)32});
        if (pl.threads > 384) pl.ctas_per_sm = std::min(pl.ctas_per_sm, 1);  // register budget of the 1024-thread variant
        else pl.ctas_per_sm = std::min(pl.ctas_per_sm, 384 / pl.threads);    // 168 registers x 384 threads per SM
        if (pl.ctas_per_sm < 1) continue;
        pl.ok = true;
        // more resident warps first, then more (smaller) CTAs: their layer barriers interleave
        const int warps = pl.ctas_per_sm * pl.W, bw = best.ok ? best.ctas_per_sm * best.W : -1;
        if (!best.ok || warps > bw || (warps == bw && pl.ctas_per_sm > best.ctas_per_sm)) best = pl;
    }
    if (!best.ok) {  // does not fit on chip: the any-size kernel (messages in a global workspace) takes it
        h->tdmp_big = true;
        h->tdmp = TdmpPlan{};
        h->tdmp.z = z;
        return LDPC_B200_OK;
    }
    h->tdmp = best;
    return LDPC_B200_OK;
}

int upload_tdmp_tables(ldpc_b200_decoder* h) {
    if (h->tdmp_ready) return LDPC_B200_OK;
    int rc = tdmp_plan(h);
    if (rc) return rc;
    const HostTables& t = h->host;
    const TdmpPlan& pl = h->tdmp;
    const int G = pl.G, SUB = 32 / G, z = pl.z;
    std::vector<uint32_t> cn_tab((size_t)pl.W * pl.cn_stride);
    std::vector<int> qoff(pl.L + 1, 0);
    for (int l = 0; l < pl.L; ++l) qoff[l + 1] = qoff[l] + (pl.ldeg[l] + 3) / 4;
    for (int w = 0; w < pl.W; ++w)
        for (int l = 0; l < pl.L; ++l)
            for (int hh = 0; hh < SUB; ++hh) {
                const int r = l * z + w * SUB + hh;
                const int dc = t.row_ptr[r + 1] - t.row_ptr[r];
                for (int j = 0; j < ((pl.ldeg[l] + 3) / 4) * 4; ++j) {
                    const uint32_t col = j < dc ? (uint32_t)t.col_idx[t.row_ptr[r] + j] : (uint32_t)(t.N + hh);  // per-h dummy row
                    cn_tab[(size_t)w * pl.cn_stride + ((size_t)(qoff[l] + j / 4) * SUB + hh) * 4 + (j & 3)] = col * (uint32_t)(G * 4);
                }
            }
    CU_TRY(cudaMalloc(&h->dt_cn_tab, cn_tab.size() * 4));
    CU_TRY(cudaMemcpy(h->dt_cn_tab, cn_tab.data(), cn_tab.size() * 4, cudaMemcpyHostToDevice));
    h->table_bytes += cn_tab.size() * 4;
    h->tdmp_ready = true;
    return LDPC_B200_OK;
}

// Sum-product (layered = false) or layered min-sum of a code of any size: 32 codewords per CTA, lane = codeword, every
// message in a CTA-private slice of a global workspace (ldpc_big.cuh).
// does the handle's current algorithm run an any-size kernel (global workspace, launches serialise on it)?
bool sp_group_fits(const ldpc_b200_decoder* h) {   // the on-chip sum-product kernel of ldpc_sp.cuh can hold the code
    const Plan& pl = h->plan;
    const HostTables& t = h->host;
    return pl.path == LDPC_B200_PATH_GROUP && pl.tab_smem && !pl.t16 && (pl.G == 8 || pl.G == 16) &&
           t.max_col_weight <= 8 && t.max_row_weight <= 20 && !h->opt.sp_big;
}

// Sum-product on the quasi-cyclic layout (ldpc_spq.cuh): tables and geometry of the group-of-warps min-sum kernel,
// uploaded to the sum-product unit's own __constant__ bank.  (h->mu held)
bool spq_prepare(ldpc_b200_decoder* h) {
    if (h->spq_state != 0) return h->spq_state == 1;
    h->spq_state = -1;
    if (h->opt.sp_qc == 0 || h->opt.sp_big) return false;
    // (the min-sum plan may have skipped the group-of-warps tables -- option qc_et = 0 --: they are built here then)
    if (h->qcm_state != 1 && h->qcm_tab.empty()) h->qcm_state = qcm_prepare(h) ? 1 : -1;
    if (h->qcm_state != 1 || h->qcm_tab.empty()) return false;
    int n = 0;
    const SpqProfileEntry* profiles = spq_profiles(&n);
    if (h->qcm_kind < 0 || h->qcm_kind >= n) return false;
    DeviceGuard guard(h->device);
    if (!guard.ok) return false;
    if (profiles[h->qcm_kind].upload(h->qcm_slot, h->qcm_tab.data(), h->qcm_tab.size()) != 0) { (void)cudaGetLastError(); return false; }
    // slice of a codeword: [HB bytes | E | 128 B slack | bit buffer] -- E, slack and bit buffer as in the min-sum layout
    QcmParams q = h->qcm;
    const uint32_t tail = h->qcm.word_bytes - h->qcm.t_bytes;                 // everything behind T
    q.hb_bytes = ((uint32_t)(24 * 2 * q.z) + 15u) & ~15u;
    q.bits_off = h->qcm.bits_off - h->qcm.t_bytes + q.hb_bytes;
    q.word_bytes = (q.hb_bytes + tail + 15u) & ~15u;
    int groups = (int)((h->smem_optin - 1024) / q.word_bytes);                // (the kernel's static shared memory counts too)
    groups = std::min(groups, kQcmMaxWarps / q.NW);
    if (q.NW > 1) groups = std::min(groups, kQcmMaxGroups);
    if (groups < 1) return false;
    h->spq = q;
    h->spq_groups = groups;
    h->spq_state = 1;
    return true;
}
// is it the kernel a sum-product decode of this handle runs?  (after spq_prepare; the plan is the sum-product plan)
bool uses_spq(const ldpc_b200_decoder* h) {
    return h->algorithm == LDPC_B200_ALG_SUM_PRODUCT && h->spq_state == 1 && (h->opt.sp_qc > 0 || !sp_group_fits(h));
}

bool uses_big_kernel(const ldpc_b200_decoder* h) {
    const Plan& pl = h->plan;
    const HostTables& t = h->host;
    switch (h->algorithm) {
        case LDPC_B200_ALG_FUSED_MIN_SUM: case LDPC_B200_ALG_FUSED_LAYERED: return true;
        case LDPC_B200_ALG_LAYERED_MIN_SUM: return h->tdmp_big;
        case LDPC_B200_ALG_SUM_PRODUCT:
            if (uses_spq(h)) return false;
            return !(pl.path == LDPC_B200_PATH_GROUP && pl.tab_smem && !pl.t16 && (pl.G == 8 || pl.G == 16) &&
                     t.max_col_weight <= 8 && t.max_row_weight <= 20 && !h->opt.sp_big);
        default: return false;
    }
}

enum BigKind { kBigSumProduct, kBigLayered, kBigFusedFlooding, kBigFusedLayered };
int launch_big(ldpc_b200_decoder* h, BigKind kind, const float* d_llr, int64_t ncw, uint8_t* d_info, uint8_t* d_hard,
               int32_t* d_iters, float* d_post, cudaStream_t stream) {
    const HostTables& t = h->host;
    const bool layered = kind != kBigSumProduct;  // (workspace shape: P + R + bits; sum-product keeps four arrays)
    if (kind == kBigSumProduct && d_post) return fail(LDPC_B200_ERR_UNSUPPORTED, "sum-product mode has no posterior output");
    const int64_t ngroups = (ncw + kLanes - 1) / kLanes;
    if (ngroups > 0x7fffffff) return fail(LDPC_B200_ERR_ARG, "too many codewords in one call");
    const int grid = (int)std::min<int64_t>(ngroups, h->sm_count);
    const size_t stride = layered ? ((size_t)t.N + t.nnz) * kLanes + (size_t)t.N + 32
                                  : ((size_t)2 * t.nnz + (size_t)2 * t.N) * kLanes + (size_t)t.N + 32;
    const size_t need = stride * sizeof(float) * (size_t)h->sm_count;
    if (h->ws_big_bytes < need) {
        CU_TRY(cudaDeviceSynchronize());
        if (h->d_ws_big) { cudaFree(h->d_ws_big); h->d_ws_big = nullptr; h->ws_big_bytes = 0; }
        CU_TRY(cudaMalloc(&h->d_ws_big, need));
        h->ws_big_bytes = need;
    }
    unsigned long long* ctr64 = h->d_counters + h->counter_next;
    h->counter_next = (h->counter_next + 1) % kCounterRing;
    CU_TRY(cudaMemsetAsync(ctr64, 0, sizeof(unsigned long long), stream));
    BigParams q;
    q.row_ptr = h->d_row_ptr; q.cn_col = h->d_cn_col; q.col_ptr = h->d_col_ptr; q.vn_edge = h->d_vn_edge;
    q.M = t.M; q.N = t.N; q.K = h->K; q.nnz = t.nnz;
    q.max_iter = h->max_iter; q.early_term = h->early; q.z = h->layer_z;
    q.fused_layered = kind == kBigFusedLayered ? 1 : 0;
    q.llr = d_llr; q.ncw = ncw;
    q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = d_post;
    q.ws = h->d_ws_big; q.ws_stride = stride;
    q.counter = reinterpret_cast<unsigned int*>(ctr64); q.ngroups = (int)ngroups;
    // the workspace is shared by every launch of this handle: serialise launches on it
    if (h->ws_event_valid) CU_TRY(cudaStreamWaitEvent(stream, h->ws_event, 0));
    int rc = kind == kBigSumProduct ? launch_status(k_launch_sp_big(q, grid, stream), "sum-product (any size)")
           : kind == kBigLayered    ? launch_status(k_launch_tdmp_big(q, grid, stream), "layered (any size)")
                                    : launch_status(k_launch_fused_big(q, grid, stream), "fused-kernel arithmetic");
    if (rc) return rc;
    CU_TRY(cudaEventRecord(h->ws_event, stream));
    h->ws_event_valid = true;
    h->launches += 1;
    return LDPC_B200_OK;
}

int launch_tdmp(ldpc_b200_decoder* h, const float* d_llr, int64_t ncw, uint8_t* d_info, uint8_t* d_hard, int32_t* d_iters,
                float* d_post, cudaStream_t stream) {
    if (h->tdmp_big) return launch_big(h, kBigLayered, d_llr, ncw, d_info, d_hard, d_iters, d_post, stream);
    int rc = upload_tdmp_tables(h);
    if (rc) return rc;
    const HostTables& t = h->host;
    const TdmpPlan& pl = h->tdmp;
    unsigned long long* ctr64 = h->d_counters + h->counter_next;
    h->counter_next = (h->counter_next + 1) % kCounterRing;
    CU_TRY(cudaMemsetAsync(ctr64, 0, sizeof(unsigned long long), stream));
    TdmpParams q;
    q.cn_tab = h->dt_cn_tab;
    q.M = t.M; q.N = t.N; q.K = h->K; q.W = pl.W; q.L = pl.L; q.VS = pl.VS;
    q.cn_stride = pl.cn_stride; q.r_rows_per_warp = pl.r_rows;
    q.max_iter = h->max_iter; q.early_term = h->early;
    q.llr = d_llr; q.ncw = ncw;
    q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = d_post;
    q.counter64 = ctr64;
    std::memcpy(q.ldeg, pl.ldeg, sizeof(q.ldeg));
    const int64_t ngroups = (ncw + pl.G - 1) / pl.G;
    const int grid = (int)std::min<int64_t>(ngroups, (int64_t)h->sm_count * pl.ctas_per_sm);
    rc = launch_status(k_launch_tdmp(pl.G, q, grid, pl.threads, pl.smem, stream), "layered");
    if (rc) return rc;
    h->launches += 1;
    return LDPC_B200_OK;
}

// Plans the FLOODING decoder for `flood_alg` (min-sum or sum-product); the layered decoder has its own plan (tdmp_plan).
int make_plan_for(ldpc_b200_decoder* h, const int flood_alg) {
    const HostTables& t = h->host;
    Plan pl;
    {   // quasi-cyclic code matching a compiled profile: warp-uniform tables (min-sum only)
        const bool want = flood_alg == LDPC_B200_ALG_MIN_SUM &&
                          (h->forced_path == LDPC_B200_PATH_QC || (h->forced_path < 0 && !h->opt.no_qc));
        const bool compiled = want && !h->opt.qc_generic && !h->opt.qcm_always;  // (the switches force the run-time profile / the group-of-warps kernel: tests)
        if (compiled && h->qc_state == 0) h->qc_state = qc_prepare(h) ? 1 : -1;
        const bool fits = compiled && h->qc_state == 1 && 2 * (h->qc_smem + 2048) <= h->smem_optin + 1024;
        // (the kernel's launch bounds allow three CTAs per SM up to 288 threads)
        const int qc_per_sm = fits && qc_profiles()[h->qc_kind].W * 32 <= 288 && 3 * (h->qc_smem + 2048) <= h->smem_optin + 1024 ? 3 : 2;
        if (fits) {
            // the ring kernel (refill off the critical path) needs a staging ring of G/2 codewords on top: used when as
            // many CTAs stay resident
            const QcProfileEntry& pe = qc_profiles()[h->qc_kind];
            const uint32_t ring_off = (uint32_t)((h->qc_smem + 127) & ~(size_t)127);
            const uint32_t ring_bytes = (uint32_t)std::max(1, pe.G / 2) * (uint32_t)t.N * 4u;
            h->qc.ring_off = ring_off;
            h->qc_ring_smem = 0;
            if (h->opt.qc_ring && pe.launch_ring) {  // opt-in: measured slower than the loop-top refill (profiles/r02_ring_kernel.md)
                DeviceGuard guard(h->device);
                const size_t need = (size_t)ring_off + ring_bytes;
                const int n = guard.ok ? pe.ring_ctas_per_sm(need) : 0;
                if (n >= qc_per_sm) { h->qc_ring_smem = need; h->qc_ring_per_sm = n; }
            }
            if (h->qcw_state == 0) h->qcw_state = (h->opt.qc_et != 0 && qcw_prepare(h)) ? 1 : -1;
            if (h->qcw_state != 1 && h->qcm_state == 0) h->qcm_state = (h->opt.qc_et != 0 && !h->opt.no_qcm && qcm_prepare(h)) ? 1 : -1;
            if ((h->qcw_state == 1 || h->qcm_state == 1) && !h->h_stats) {
                DeviceGuard guard(h->device);
                if (!guard.ok || cudaMallocHost(&h->h_stats, 2 * sizeof(unsigned long long)) != cudaSuccess) {
                    (void)cudaGetLastError();
                    h->h_stats = nullptr; h->qcw_state = -1; h->qcm_state = -1;
                } else {
                    h->h_stats[0] = h->h_stats[1] = 0ull;
                }
            }
            pl.path = LDPC_B200_PATH_QC;
            pl.dmax = 1;  // marks the compiled profile
            pl.threads = 32 * qc_profiles()[h->qc_kind].W;
            pl.smem = h->qc_smem;
            pl.ctas = h->sm_count * qc_per_sm;
            pl.cw_per_cta = qc_profiles()[h->qc_kind].G;
            pl.W = qc_profiles()[h->qc_kind].W; pl.G = qc_profiles()[h->qc_kind].G;
            h->plan = pl;
            h->planned = true;
            return LDPC_B200_OK;
        }
        // any other quasi-cyclic code: the same kernel with a run-time profile -- when the generic on-chip kernel would
        // have to run one codeword per CTA (z > 24: measured 1.6-2.2x faster); with 8 or 16 words per CTA the group
        // kernel is as fast (z = 24 rates: 1.31-1.67 ms against 1.49-1.56 ms per 16,384 words) and stays the choice
        // no compiled lockstep profile (z = 28, 36, 44, 52, 56, 60, 68, 72, 76, 84, 88, 92 of the reference's family): a group
        // of ceil(z / 32) warps per codeword (measured 2-3x the kernels below: profiles/r02_wimax_family.txt)
        if (want && !h->opt.no_qcm && !h->opt.qc_generic && h->qcm_state == 0) h->qcm_state = qcm_prepare(h) ? 1 : -1;
        if (want && !h->opt.no_qcm && !h->opt.qc_generic && h->qcm_state == 1) {
            if (h->qcm_multi_groups > 0 && !h->h_stats) {   // the regime decides between one and several codewords per group
                DeviceGuard guard(h->device);
                if (guard.ok && cudaMallocHost(&h->h_stats, 2 * sizeof(unsigned long long)) == cudaSuccess) h->h_stats[0] = h->h_stats[1] = 0ull;
                else { (void)cudaGetLastError(); h->h_stats = nullptr; }
            }
            pl.path = LDPC_B200_PATH_QC;
            pl.dmax = 2;  // marks the group-of-warps kernel
            pl.threads = 32 * h->qcm_groups * h->qcm.NW;
            pl.smem = (size_t)h->qcm_groups * h->qcm.word_bytes;
            pl.ctas = h->sm_count;
            pl.cw_per_cta = h->qcm_groups;
            pl.W = h->qcm_groups * h->qcm.NW; pl.G = h->qcm_groups;
            h->plan = pl;
            h->planned = true;
            return LDPC_B200_OK;
        }
        bool generic = want && !h->opt.no_qcg;
        if (generic && h->forced_path < 0 && !h->opt.qc_generic) {
            GrpShape sh;
            if (group_pick(t, h->smem_optin, false, h->opt, &sh) && sh.G >= 8 && sh.tab_smem) generic = false;
        }
        if (generic && h->qcg_state == 0) h->qcg_state = qcg_prepare(h) ? 1 : -1;
        if (generic && h->qcg_state == 1) {
            pl.path = LDPC_B200_PATH_QC;
            pl.threads = 32 * h->qcg.W;
            pl.smem = h->qcg_smem;
            pl.ctas = h->sm_count * h->qcg_ctas_per_sm;
            pl.cw_per_cta = h->qcg_G;
            pl.W = h->qcg.W; pl.CS = h->qcg.CS; pl.VS = h->qcg.VS; pl.G = h->qcg_G;
            pl.dmax = 0;  // marks the run-time profile
            h->plan = pl;
            h->planned = true;
            return LDPC_B200_OK;
        }
        if (h->forced_path == LDPC_B200_PATH_QC)
            return fail(LDPC_B200_ERR_UNSUPPORTED, "code is not quasi-cyclic in a supported shape (or the algorithm is not min-sum)");
    }
    if (h->forced_path == LDPC_B200_PATH_WARP) {  // opt-in: sub-warp per check, shuffle reductions
        const int SW = warp_sub_width(t);
        const size_t smem = ((size_t)2 * t.N + (size_t)t.M * SW) * sizeof(float);
        if (flood_alg != LDPC_B200_ALG_MIN_SUM || SW == 0 || smem + 1024 > h->smem_optin)
            return fail(LDPC_B200_ERR_UNSUPPORTED, "the warp-per-check path needs min-sum, check degree <= 32 and a codeword that fits one SM's shared memory");
        pl.path = LDPC_B200_PATH_WARP;
        pl.threads = std::min(1024, std::max(128, (t.N + 31) / 32 * 32));
        pl.smem = smem;
        const size_t per_sm = std::min<size_t>({(h->smem_optin + 1024) / (smem + 1024), (size_t)(2048 / pl.threads), (size_t)16});
        pl.ctas = h->sm_count * (int)std::max<size_t>(per_sm, 1);
        pl.cw_per_cta = 1;
        h->plan = pl;
        h->planned = true;
        return LDPC_B200_OK;
    }
    {   // explicit per-edge messages on chip (G codewords per CTA)
        GrpShape sh;
        const bool fits = group_pick(t, h->smem_optin, flood_alg == LDPC_B200_ALG_SUM_PRODUCT, h->opt, &sh);
        if (h->forced_path == LDPC_B200_PATH_GROUP && !fits)
            return fail(LDPC_B200_ERR_UNSUPPORTED, "code does not fit the group shared-memory path");
        if ((h->forced_path < 0 || h->forced_path == LDPC_B200_PATH_GROUP) && fits) {
            if (h->group_ready && (h->plan.W != sh.W || h->plan.G != sh.G)) return fail(LDPC_B200_ERR_ARG, "group layout changed after upload");
            pl.path = LDPC_B200_PATH_GROUP;
            pl.threads = 32 * sh.W;
            pl.smem = sh.smem;
            pl.ctas = h->sm_count * (sh.G == 8 ? 2 : (sh.G == 4 ? 3 : 1));
            pl.cw_per_cta = sh.G;
            pl.W = sh.W; pl.CS = sh.CS; pl.VS = sh.VS; pl.G = sh.G; pl.dmax = sh.dmax; pl.tab_smem = sh.tab_smem;
            pl.cn_stride = sh.cn_stride; pl.vn_stride = sh.vn_stride; pl.r_rows = sh.r_rows;
            pl.y_smem = sh.y_smem;
            h->plan = pl;
            h->planned = true;
            return LDPC_B200_OK;
        }
    }
    {   // tuned short-code path: channel values in registers, 16-byte check state, tables in smem
        L16Shape sh;
        const bool fits = (uint64_t)t.M * 512u < (1ull << 31) && lane16_pick(t, h->smem_optin, h->opt, &sh);
        if (h->forced_path == LDPC_B200_PATH_LANE16 && !fits)
            return fail(LDPC_B200_ERR_UNSUPPORTED, "code does not fit the lane16 shared-memory path");
        if ((h->forced_path < 0 || h->forced_path == LDPC_B200_PATH_LANE16) && fits) {
            if (h->lane16_ready && (h->plan.W != sh.W)) return fail(LDPC_B200_ERR_ARG, "lane16 layout changed after upload");
            pl.path = LDPC_B200_PATH_LANE16;
            pl.threads = 32 * sh.W;
            pl.smem = sh.smem;
            pl.ctas = h->sm_count;
            pl.dcp = sh.DCP; pl.W = sh.W; pl.CS = sh.CS; pl.VS = sh.VS;
            h->plan = pl;
            h->planned = true;
            return LDPC_B200_OK;
        }
    }
    {   // long codes: one codeword per 8-CTA cluster, state in distributed shared memory
        ClShape sh;
        std::vector<int> cr, vr;
        // Opt-in only: measured on cfg5 (profiles/r01_cluster_dsmem.txt) the 4-byte DSMEM gathers run at
        // ~0.5 per clock per SM, which leaves this path at half the speed of the global-workspace path.
        const bool fits = h->forced_path == LDPC_B200_PATH_CLUSTER && cluster_shape(t, h->smem_optin, &sh, &cr, &vr);
        if (h->forced_path == LDPC_B200_PATH_CLUSTER && !fits)
            return fail(LDPC_B200_ERR_UNSUPPORTED, "code does not fit the cluster path");
        if (fits) {
            pl.path = LDPC_B200_PATH_CLUSTER;
            pl.threads = 32 * sh.W;
            pl.smem = sh.smem;
            pl.ctas = (h->sm_count / kClusterSize) * kClusterSize;
            pl.cw_per_cta = 1;
            pl.W = sh.W; pl.CS = sh.CS; pl.VS = sh.VS; pl.G = 1; pl.dmax = sh.dmax;
            pl.cn_stride = sh.cn_stride; pl.vn_stride = sh.vn_stride; pl.r_rows = sh.r_rows;
            h->plan = pl;
            h->planned = true;
            return LDPC_B200_OK;
        }
    }
    if (t.max_row_weight > kMaxCheckDegree)
        return fail(LDPC_B200_ERR_UNSUPPORTED, "check degree above 27 is only supported on the lane16 path");
    const size_t state_bytes = ((size_t)2 * t.N + (size_t)3 * t.M) * kLanes * sizeof(float);
    const size_t static_smem = 1024;  // s_group, s_flag + slack
    int path = h->forced_path;
    const bool stream_ok = t.max_row_weight <= 8 && t.max_col_weight <= 8 && (uint64_t)t.M * 8 * kLanes < (1ull << 32);
    if (path == LDPC_B200_PATH_STREAM && !stream_ok)
        return fail(LDPC_B200_ERR_UNSUPPORTED, "the streamed global path needs check and variable degree <= 8");
    if (path < 0) path = (state_bytes + static_smem <= h->smem_optin) ? LDPC_B200_PATH_LANE_SMEM
                         : (stream_ok ? LDPC_B200_PATH_STREAM : LDPC_B200_PATH_LANE_GLOBAL);
    if (path == LDPC_B200_PATH_LANE_SMEM) {
        if (state_bytes + static_smem > h->smem_optin)
            return fail(LDPC_B200_ERR_UNSUPPORTED, "code too large for the shared-memory lane path");
        pl.path = path;
        pl.threads = 32 * pick_lane_warps(t.M, t.N, t.nnz);
        pl.smem = state_bytes;
        pl.ctas = h->sm_count;  // one persistent CTA per SM
        pl.ws_stride = 0;
    } else if (path == LDPC_B200_PATH_LANE_GLOBAL) {
        pl.path = path;
        pl.threads = 1024;
        pl.smem = 0;
        pl.ctas = h->sm_count;
        pl.ws_stride = state_bytes / sizeof(float);
    } else if (path == LDPC_B200_PATH_STREAM) {
        pl.path = path;
        pl.threads = 1024;
        { const int th = h->opt.stream_threads; if (th >= 32 && th <= 1024 && th % 32 == 0) pl.threads = th; }
        pl.smem = 0;
        pl.ctas = h->sm_count;
        pl.ws_stride = ((size_t)2 * stream_shape(t).np + (size_t)8 * t.M) * kLanes;
    } else {
        return fail(LDPC_B200_ERR_UNSUPPORTED, "unknown kernel path");
    }
    h->plan = pl;
    h->planned = true;
    return LDPC_B200_OK;
}

int make_plan(ldpc_b200_decoder* h) {
    const int flood_alg = h->algorithm == LDPC_B200_ALG_SUM_PRODUCT ? LDPC_B200_ALG_SUM_PRODUCT : LDPC_B200_ALG_MIN_SUM;
    const int rc = make_plan_for(h, flood_alg);
    if (rc == LDPC_B200_OK) h->plan_alg = flood_alg;
    return rc;
}

int ensure_workspace(ldpc_b200_decoder* h) {
    if (h->plan.path != LDPC_B200_PATH_LANE_GLOBAL && h->plan.path != LDPC_B200_PATH_STREAM) return LDPC_B200_OK;
    const size_t need = h->plan.ws_stride * sizeof(float) * (size_t)h->plan.ctas;
    if (h->ws_bytes >= need) return LDPC_B200_OK;
    if (h->d_ws) { cudaFree(h->d_ws); h->d_ws = nullptr; h->ws_bytes = 0; }
    CU_TRY(cudaMalloc(&h->d_ws, need));
    h->ws_bytes = need;
    return LDPC_B200_OK;
}

void free_slots(ldpc_b200_decoder* h) {
    for (int s = 0; s < kSlots; ++s) { cudaFree(h->s_pack[s]); h->s_pack[s] = nullptr; }
    h->s_pack_bytes = 0;
    for (int s = 0; s < kSlots; ++s) {
        cudaFree(h->s_llr[s]); h->s_llr[s] = nullptr;
        cudaFree(h->s_info[s]); h->s_info[s] = nullptr;
        cudaFree(h->s_hard[s]); h->s_hard[s] = nullptr;
        cudaFree(h->s_iters[s]); h->s_iters[s] = nullptr;
        cudaFree(h->s_post[s]); h->s_post[s] = nullptr;
    }
    h->reserved = 0;
}

int launch_decode(ldpc_b200_decoder* h, const float* d_llr, int64_t ncw, uint8_t* d_info, uint8_t* d_hard,
                  int32_t* d_iters, float* d_post, cudaStream_t stream) {
    if (ncw == 0) return LDPC_B200_OK;
    if (h->algorithm == LDPC_B200_ALG_LAYERED_MIN_SUM) return launch_tdmp(h, d_llr, ncw, d_info, d_hard, d_iters, d_post, stream);
    if (h->algorithm == LDPC_B200_ALG_FUSED_MIN_SUM) return launch_big(h, kBigFusedFlooding, d_llr, ncw, d_info, d_hard, d_iters, d_post, stream);
    if (h->algorithm == LDPC_B200_ALG_FUSED_LAYERED) return launch_big(h, kBigFusedLayered, d_llr, ncw, d_info, d_hard, d_iters, d_post, stream);
    if (!h->planned) {
        int rc = make_plan(h);
        if (rc) return rc;
    }
    const Plan& pl = h->plan;
    const HostTables& t = h->host;
    if (h->algorithm == LDPC_B200_ALG_SUM_PRODUCT) {
        // the on-chip kernel takes short codes (group layout with 8 or 16 words per CTA, variable degree <= 8, check degree
        // <= 20); every other code goes to the any-size kernel -- DecodeSP never decodes with another algorithm
        (void)spq_prepare(h);   // (quasi-cyclic codes: ldpc_spq.cuh, launched below)
        if (uses_big_kernel(h)) return launch_big(h, kBigSumProduct, d_llr, ncw, d_info, d_hard, d_iters, d_post, stream);
    }
    int rc = uses_spq(h) ? LDPC_B200_OK : ensure_workspace(h);
    if (rc) return rc;
    const int per_group = (pl.path == LDPC_B200_PATH_GROUP || pl.path == LDPC_B200_PATH_QC) ? pl.G : (pl.path == LDPC_B200_PATH_WARP ? 1 : kLanes);
    const int64_t ngroups = (ncw + per_group - 1) / per_group;
    if (ngroups > 0x7fffffff) return fail(LDPC_B200_ERR_ARG, "too many codewords in one call");
    unsigned long long* ctr64 = h->d_counters + h->counter_next;
    unsigned int* ctr = reinterpret_cast<unsigned int*>(ctr64);
    h->counter_next = (h->counter_next + 1) % kCounterRing;
    CU_TRY(cudaMemsetAsync(ctr64, 0, sizeof(unsigned long long), stream));
    const int grid = (int)std::min<int64_t>(ngroups, pl.ctas);

    if (pl.path == LDPC_B200_PATH_WARP) {
        rc = upload_warp_tables(h);
        if (rc) return rc;
        WarpParams q;
        q.cn_col = h->dw_cn_col; q.col_ptr = h->d_col_ptr; q.vn_pos = h->dw_vn_pos;
        q.M = t.M; q.N = t.N; q.K = h->K; q.SW = h->w_sw;
        q.max_iter = h->max_iter; q.early_term = h->early;
        q.llr = d_llr; q.ncw = ncw;
        q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = d_post;
        q.counter64 = ctr64;
        rc = launch_status(k_launch_warp(h->w_sw, q, grid, pl.threads, pl.smem, stream), "warp-per-check");
        h->launches += 1;
        return LDPC_B200_OK;
    }

    // the group-of-warps kernel (ldpc_qcm.cuh): the plan's kernel for block sizes without a compiled lockstep profile, and
    // the early-termination alternative of the profiled sizes that have no warp-per-codeword kernel
    auto launch_qcm = [&](int32_t* iters_out, bool multi = false) -> int {
        QcmParams& q = multi ? h->qcm_multi : h->qcm;  // tables and geometry filled by qcm_build; per-launch fields below
        q.tab_slot = h->qcm.tab_slot;
        q.K = h->K;
        q.max_iter = h->max_iter; q.early_term = h->early;
        q.llr = d_llr; q.ncw = ncw;
        q.info = d_info; q.hard = d_hard; q.iters = iters_out; q.post = d_post;
        q.counter64 = ctr64;
        q.avail = h->cur_avail;
        q.status = h->cur_avail ? reinterpret_cast<int*>(h->d_avail + 1) : nullptr;
        q.wait_ns = (unsigned long long)std::max<long long>(1, h->opt.wait_timeout_ms) * 1000000ull;
        q.fmt = h->cur_fmt; q.scale = h->cur_scale;
        int np = 0;
        if (multi) {
            const int64_t per_cta = (int64_t)h->qcm_multi_groups * q.PK;
            const int g = (int)std::min<int64_t>((ncw + per_cta - 1) / per_cta, (int64_t)h->sm_count);
            return launch_status(qcm_profiles(&np)[h->qcm_kind].launch_multi(q, g, h->qcm_multi_groups, stream), "quasi-cyclic (warps per group of codewords)");
        }
        const int g = (int)std::min<int64_t>((ncw + h->qcm_groups - 1) / h->qcm_groups, (int64_t)h->sm_count);
        return launch_status(qcm_profiles(&np)[h->qcm_kind].launch(q, g, h->qcm_groups, stream), "quasi-cyclic (warps per codeword)");
    };
    if (uses_spq(h)) {   // sum-product, quasi-cyclic layout: the group-of-warps geometry with the arithmetic of ldpc_sp.cuh
        if (d_post) return fail(LDPC_B200_ERR_UNSUPPORTED, "sum-product mode has no posterior output");
        QcmParams q = h->spq;
        q.tab_slot = h->qcm.tab_slot;
        q.K = h->K;
        q.max_iter = h->max_iter; q.early_term = h->early;
        q.llr = d_llr; q.ncw = ncw;
        q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = nullptr;
        q.counter64 = ctr64;
        q.avail = nullptr; q.status = nullptr; q.wait_ns = 0ull;
        q.fmt = LDPC_B200_LLR_F32; q.scale = 1.0f;
        int np = 0;
        const int g = (int)std::min<int64_t>((ncw + h->spq_groups - 1) / h->spq_groups, (int64_t)h->sm_count);
        rc = launch_status(spq_profiles(&np)[h->qcm_kind].launch(q, g, h->spq_groups, stream), "sum-product (quasi-cyclic)");
        if (rc) return rc;
        h->last_kernel = 5;
        h->launches += 1;
        return LDPC_B200_OK;
    }
    if (pl.path == LDPC_B200_PATH_QC && pl.dmax == 2) {
        // several codewords per group while the words run long; one per group (nothing waits for a slower neighbour) once
        // the previous launches' words stopped early on average -- the regime is sampled as for the lockstep kernel below
        const bool can_multi = h->qcm_multi_groups > 0 && h->cur_fmt == LDPC_B200_LLR_F32;
        const int pct = h->opt.qcm_multi_pct < 0 ? 80 : h->opt.qcm_multi_pct;   // (measured: 2-5 % slower at 70 % of the cap)
        const bool track = can_multi && h->early && h->h_stats && pct > 0;
        bool use_multi = can_multi;
        int32_t* it_out = d_iters;
        if (track) {
            if (h->h_stats[1] > 0) use_multi = (double)h->h_stats[0] * 100.0 > (double)pct * (double)h->max_iter * (double)h->h_stats[1];
            if (!it_out) {
                if (h->iters_own_cap < ncw) {
                    cudaFree(h->d_iters_own); h->d_iters_own = nullptr; h->iters_own_cap = 0;
                    CU_TRY(cudaMalloc(&h->d_iters_own, (size_t)ncw * sizeof(int32_t)));
                    h->iters_own_cap = ncw;
                }
                it_out = h->d_iters_own;
            }
        }
        rc = launch_qcm(it_out, use_multi);
        if (rc) return rc;
        h->last_kernel = use_multi ? 4 : 3;
        h->launches += 1;
        if (track && (h->h_stats[1] == 0 || h->stat_tick++ % (unsigned)std::max(1, h->opt.qc_et_every) == 0)) {
            rc = launch_status(k_launch_iter_stats(it_out, ncw, h->h_stats, stream), "iteration statistics");
            if (rc) return rc;
            h->launches += 1;
        }
        return LDPC_B200_OK;
    }

    if (pl.path == LDPC_B200_PATH_QC && pl.dmax == 0) {
        QcgParams& q = h->qcg;  // tables filled by qcg_build; per-launch fields below
        q.K = h->K;
        q.max_iter = h->max_iter; q.early_term = h->early;
        q.refill_wait = h->opt.refill_wait;
        q.llr = d_llr; q.ncw = ncw;
        q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = d_post;
        q.counter64 = ctr64;
        q.avail = h->cur_avail;
        q.status = h->cur_avail ? reinterpret_cast<int*>(h->d_avail + 1) : nullptr;
        q.wait_ns = (unsigned long long)std::max<long long>(1, h->opt.wait_timeout_ms) * 1000000ull;
        rc = launch_status(k_launch_qcg(h->qcg_G, q, grid, pl.smem, stream), "quasi-cyclic (run-time profile)");
        if (rc) return rc;
        h->launches += 1;
        return LDPC_B200_OK;
    }

    if (pl.path == LDPC_B200_PATH_QC) {
        auto& q = h->qc;  // tables filled by qc_build; per-launch fields below
        q.K = h->K;
        q.max_iter = h->max_iter; q.early_term = h->early;
        q.refill_wait = h->opt.refill_wait;
        q.llr = d_llr; q.ncw = ncw;
        q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = d_post;
        q.counter64 = ctr64;
        q.avail = h->cur_avail;
        q.status = h->cur_avail ? reinterpret_cast<int*>(h->d_avail + 1) : nullptr;
        q.wait_ns = (unsigned long long)std::max<long long>(1, h->opt.wait_timeout_ms) * 1000000ull;
        // Which kernel: the lockstep kernel, or -- when the handle's recent words stopped early on average -- the kernel
        // that decodes every codeword on its own warp(s): ldpc_qcw.cuh where the code has one (z = 24 / 32), else
        // ldpc_qcm.cuh.  The regime is the mean iteration count of the previous launches, sampled on the device and read
        // from pinned memory (no synchronisation); unknown = lockstep.  Measured crossovers (share of the cap):
        // qcw z = 24 97 % (38.99 of 40: 1.02x, 39.76: 0.95x), z = 32 every regime; qcm 25 % (profiles/r02_qcw_*.txt, r02_qcm_vs_lockstep.txt).
        const bool aligned = (reinterpret_cast<uintptr_t>(d_llr) & 15u) == 0;  // bulk copies of the ring kernel
        const bool have_w = h->qcw_state == 1, have_m = !have_w && h->qcm_state == 1;
        const bool et_ready = (have_w || have_m) && h->early && !h->qc_ring_smem;
        const bool track = et_ready && h->opt.qc_et < 0;   // the regime is tracked through the iteration counts
        bool use_et = et_ready && h->opt.qc_et > 0;
        if (track) {
            int np0 = 0;
            const int pct = h->opt.qc_et_pct > 0 ? h->opt.qc_et_pct : (have_m ? 25 : (qcw_profiles(&np0)[h->qcw_kind].z >= 32 ? 100 : 97));
            if (pct >= 100) use_et = true;
            else if (h->h_stats[1] > 0) use_et = (double)h->h_stats[0] * 100.0 <= (double)pct * (double)h->max_iter * (double)h->h_stats[1];
        }
        if (h->cur_fmt != LDPC_B200_LLR_F32) {   // packed input is widened at the load: only the per-codeword kernels read it
            if (!have_w && !have_m) return fail(LDPC_B200_ERR_UNSUPPORTED, "packed channel values need a per-codeword kernel");
            use_et = true;
        }
        if (track && !q.iters) {
            if (h->iters_own_cap < ncw) {
                cudaFree(h->d_iters_own); h->d_iters_own = nullptr; h->iters_own_cap = 0;
                CU_TRY(cudaMalloc(&h->d_iters_own, (size_t)ncw * sizeof(int32_t)));
                h->iters_own_cap = ncw;
            }
            q.iters = h->d_iters_own;
        }
        if (use_et && have_m) {
            rc = launch_qcm(q.iters);
            h->last_kernel = 3;
        } else         if (use_et) {
            QcwParams& e = h->qcw;
            e.K = q.K; e.max_iter = q.max_iter; e.early_term = q.early_term;
            e.llr = q.llr; e.ncw = q.ncw; e.info = q.info; e.hard = q.hard; e.iters = q.iters; e.post = q.post;
            e.counter64 = q.counter64; e.avail = q.avail; e.status = q.status; e.wait_ns = q.wait_ns;
            e.fmt = h->cur_fmt; e.scale = h->cur_scale;
            int np = 0;
            const QcwProfileEntry& pe = qcw_profiles(&np)[h->qcw_kind];
            const int et_grid = (int)std::min<int64_t>((ncw + h->qcw_warps - 1) / h->qcw_warps, (int64_t)h->sm_count);
            rc = launch_status(pe.launch(e, et_grid, h->qcw_warps, stream), "quasi-cyclic (warp per codeword)");
            h->last_kernel = 1;
        } else if (h->qc_ring_smem && aligned) {  // ring kernel unless its bulk copies cannot be used
            rc = launch_status(qc_profiles()[h->qc_kind].launch_ring(q, grid, h->qc_ring_smem, stream), "quasi-cyclic (ring)");
            h->last_kernel = 2;
        } else {
            rc = launch_status(qc_profiles()[h->qc_kind].launch(q, grid, pl.smem, stream), "quasi-cyclic");
            h->last_kernel = 0;
        }
        if (rc) return rc;
        h->launches += 1;
        if (track && (h->h_stats[1] == 0 || h->stat_tick++ % (unsigned)std::max(1, h->opt.qc_et_every) == 0)) {
            // (written straight into pinned host memory: no copy operation between two decodes of the stream)
            rc = launch_status(k_launch_iter_stats(q.iters, ncw, h->h_stats, stream), "iteration statistics");
            if (rc) return rc;
            h->launches += 1;
        }
        return LDPC_B200_OK;
    }

    if (pl.path == LDPC_B200_PATH_CLUSTER) {
        rc = upload_cluster_tables(h);
        if (rc) return rc;
        ClusterParams q;
        q.cn_tab = h->dc_cn_tab; q.vn_tab = h->dc_vn_tab; q.var_of_pos = h->dc_var_of_pos; q.out_addr = h->dc_out_addr;
        q.M = t.M; q.N = t.N; q.K = h->K; q.W = pl.W; q.CS = pl.CS; q.VS = pl.VS;
        q.cn_stride = pl.cn_stride; q.vn_stride = pl.vn_stride; q.r_rows_per_warp = pl.r_rows;
        q.max_iter = h->max_iter; q.early_term = h->early;
        q.llr = d_llr; q.ncw = ncw;
        q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = d_post;
        q.counter64 = ctr64;
        std::memcpy(q.vdeg, h->g_vdeg, sizeof(q.vdeg));
        std::memcpy(q.cdeg, h->g_cdeg, sizeof(q.cdeg));
        const int want = (int)std::min<int64_t>(ncw, 1 << 20);
        rc = launch_status(k_launch_cluster(pl.dmax, q, want, pl.threads, pl.smem, stream), "cluster");
        if (rc) return rc;
        h->launches += 1;
        return LDPC_B200_OK;
    }

    if (pl.path == LDPC_B200_PATH_GROUP) {
        rc = upload_group_tables(h);
        if (rc) return rc;
        GroupParams q;
        q.cn_tab = h->dg_cn_tab; q.vn_tab = h->dg_vn_tab;
        q.var_of_pos = h->dg_var_of_pos; q.pos_of_var = h->dg_pos_of_var;
        q.M = t.M; q.N = t.N; q.K = h->K; q.W = pl.W; q.CS = pl.CS; q.VS = pl.VS;
        q.cn_stride = pl.cn_stride; q.vn_stride = pl.vn_stride; q.r_rows_per_warp = pl.r_rows;
        q.max_iter = h->max_iter; q.early_term = h->early;
        q.llr = d_llr; q.ncw = ncw;
        q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = d_post;
        q.counter = ctr; q.counter64 = ctr64; q.ngroups = (int)ngroups;
        q.refill_wait = h->opt.refill_wait;
        std::memcpy(q.vdeg, h->g_vdeg, sizeof(q.vdeg));
        std::memcpy(q.cdeg, h->g_cdeg, sizeof(q.cdeg));
        q.n_vclass = 0;
        std::memset(q.vclass_deg, 0, sizeof(q.vclass_deg));
        std::memset(q.vclass_cnt, 0, sizeof(q.vclass_cnt));
        for (int sidx = 0; sidx < pl.VS; ++sidx) {  // runs of equal slot degree
            if (q.n_vclass && q.vclass_deg[q.n_vclass - 1] == h->g_vdeg[sidx]) q.vclass_cnt[q.n_vclass - 1]++;
            else { q.vclass_deg[q.n_vclass] = h->g_vdeg[sidx]; q.vclass_cnt[q.n_vclass] = 1; q.n_vclass++; }
        }
        if (h->algorithm == LDPC_B200_ALG_SUM_PRODUCT) {
            if (!pl.tab_smem || pl.t16 || (pl.G != 8 && pl.G != 16) || t.max_col_weight > 8 || t.max_row_weight > 20)
                return fail(LDPC_B200_ERR_UNSUPPORTED, "sum-product needs the on-chip group layout (short codes, variable degree <= 8, check degree <= 20)");
            if (d_post) return fail(LDPC_B200_ERR_UNSUPPORTED, "sum-product mode has no posterior output");
            rc = launch_status(k_launch_sp(pl.G, pl.threads, q, grid, pl.smem, stream), "sum-product");
            if (rc) return rc;
            h->launches += 1;
            return LDPC_B200_OK;
        }
        rc = launch_group(pl, q, grid, !h->opt.grp_no_profile, stream);
        if (rc) return rc;
        h->launches += 1;
        return LDPC_B200_OK;
    }
    if (h->algorithm == LDPC_B200_ALG_SUM_PRODUCT)
        return fail(LDPC_B200_ERR_UNSUPPORTED, "sum-product needs the on-chip group layout (short codes)");

    if (pl.path == LDPC_B200_PATH_LANE16) {
        rc = upload_lane16_tables(h);
        if (rc) return rc;
        Lane16Params q;
        q.cn_tab = h->d16_cn_tab; q.vn_tab = h->d16_vn_tab;
        q.var_of_pos = h->d16_var_of_pos; q.pos_of_var = h->d16_pos_of_var;
        q.M = t.M; q.N = t.N; q.K = h->K; q.W = pl.W; q.CS = pl.CS; q.VS = pl.VS; q.DCP = pl.dcp;
        q.vn_stride = h->l16_vn_stride;
        q.max_iter = h->max_iter; q.early_term = h->early;
        q.llr = d_llr; q.ncw = ncw;
        q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = d_post;
        q.counter = ctr; q.ngroups = (int)ngroups;
        std::memcpy(q.vdeg, h->l16_vdeg, sizeof(q.vdeg));
        std::memcpy(q.cdeg, h->l16_cdeg, sizeof(q.cdeg));
        rc = launch_status(k_launch_lane16(q, grid, pl.threads, pl.smem, stream), "lane16");
        if (rc) return rc;
        h->launches += 1;
        return LDPC_B200_OK;
    }

    if (pl.path == LDPC_B200_PATH_STREAM) {
        rc = upload_stream_tables(h);
        if (rc) return rc;
        StreamParams q;
        q.cn_tab = h->ds_cn_tab; q.vn_tab = h->ds_vn_tab; q.var_of_pos = h->ds_var_of_pos; q.pos_of_var = h->ds_pos_of_var;
        q.M = t.M; q.N = t.N; q.K = h->K; q.NP = h->s_np;
        q.nb8 = h->s_nb8; q.nb4 = h->s_nb4; q.nb2 = h->s_nb2;
        q.max_iter = h->max_iter; q.early_term = h->early;
        q.llr = d_llr; q.ncw = ncw;
        q.info = d_info; q.hard = d_hard; q.iters = d_iters; q.post = d_post;
        q.ws = h->d_ws; q.ws_stride = pl.ws_stride;
        q.counter = ctr; q.ngroups = (int)ngroups;
        // the workspace is shared by every launch of this handle: serialise launches on it
        if (h->ws_event_valid) CU_TRY(cudaStreamWaitEvent(stream, h->ws_event, 0));
        rc = launch_status(k_launch_stream(q, grid, pl.threads, stream), "stream");
        if (rc) return rc;
        CU_TRY(cudaEventRecord(h->ws_event, stream));
        h->ws_event_valid = true;
        h->launches += 1;
        return LDPC_B200_OK;
    }

    DecodeParams p;
    p.row_ptr = h->d_row_ptr;
    p.cn_col = h->d_cn_col;
    p.col_ptr = h->d_col_ptr;
    p.vn_edge = h->d_vn_edge;
    p.M = t.M; p.N = t.N; p.K = h->K;
    p.max_iter = h->max_iter; p.early_term = h->early;
    p.llr = d_llr; p.ncw = ncw;
    p.info = d_info; p.hard = d_hard; p.iters = d_iters; p.post = d_post;
    p.ws = h->d_ws; p.ws_stride = pl.ws_stride;
    p.ngroups = (int)ngroups;
    p.counter = ctr;
    if (pl.path == LDPC_B200_PATH_LANE_SMEM) {
        CU_TRY(cudaFuncSetAttribute(ldpc_ms_lane_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        ldpc_ms_lane_kernel<true><<<grid, pl.threads, pl.smem, stream>>>(p);
    } else {
        // the workspace is shared by every launch of this handle: serialise launches on it
        if (h->ws_event_valid) CU_TRY(cudaStreamWaitEvent(stream, h->ws_event, 0));
        ldpc_ms_lane_kernel<false><<<grid, pl.threads, 0, stream>>>(p);
        CU_TRY(cudaEventRecord(h->ws_event, stream));
        h->ws_event_valid = true;
    }
    CU_TRY(cudaGetLastError());
    h->launches += 1;
    return LDPC_B200_OK;
}

}  // namespace

extern "C" {

const char* ldpc_b200_last_error(void) { return g_err.c_str(); }
const char* ldpc_b200_version(void) { return "ldpc_b200 0.2 (sm_100a)"; }

int ldpc_b200_create(ldpc_b200_handle* out, int M, int N, int K, const int32_t* row_ptr, const int32_t* col_idx,
                     int device) {
    if (!out) return fail(LDPC_B200_ERR_ARG, "out is null");
    *out = nullptr;
    if (K <= 0 || K > N) return fail(LDPC_B200_ERR_ARG, "K must be in 1..N");
    ldpc_b200_decoder* h = new (std::nothrow) ldpc_b200_decoder();
    if (!h) return fail(LDPC_B200_ERR_NOMEM, "out of host memory");
    std::string msg = build_tables(M, N, row_ptr, col_idx, &h->host);
    if (!msg.empty()) { delete h; return fail(LDPC_B200_ERR_ARG, msg); }
    if (h->host.max_row_weight > 32) {
        delete h;
        return fail(LDPC_B200_ERR_UNSUPPORTED, "check degree above 32 is not supported by the packed sign word");
    }
    h->K = K;
    h->device = device;
    h->opt = options_from_env();  // the only place the experiment switches are read from the environment

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        delete h;
        return fail(LDPC_B200_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                            " (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) { delete h; return fail(LDPC_B200_ERR_ARG, "device ordinal out of range"); }
    DeviceGuard guard(device);
    if (!guard.ok) { delete h; return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed"); }

    auto cleanup_fail = [&](int code, const std::string& m) {
        ldpc_b200_destroy(h);
        return fail(code, m);
    };
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
        return cleanup_fail(LDPC_B200_ERR_CUDA, "cudaDeviceGetAttribute(SM count) failed");
    h->sm_count = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess)
        return cleanup_fail(LDPC_B200_ERR_CUDA, "cudaDeviceGetAttribute(smem optin) failed");
    h->smem_optin = (size_t)v;

    const HostTables& t = h->host;
    const size_t nnz1 = (size_t)std::max(t.nnz, 1);
    bool ok = true;
    ok = ok && cudaMalloc(&h->d_row_ptr, sizeof(int32_t) * (t.M + 1)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_cn_col, sizeof(uint32_t) * nnz1) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_col_ptr, sizeof(int32_t) * (t.N + 1)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_vn_edge, sizeof(uint32_t) * nnz1) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_counters, sizeof(unsigned long long) * kCounterRing) == cudaSuccess;
    if (!ok) return cleanup_fail(LDPC_B200_ERR_CUDA, std::string("cudaMalloc(tables): ") + cudaGetErrorString(cudaGetLastError()));
    h->table_bytes = sizeof(int32_t) * (t.M + 1 + t.N + 1) + sizeof(uint32_t) * 2 * nnz1;
    ok = ok && cudaMemcpy(h->d_row_ptr, t.row_ptr.data(), sizeof(int32_t) * (t.M + 1), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_cn_col, t.col_idx.data(), sizeof(uint32_t) * t.nnz, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_col_ptr, t.col_ptr.data(), sizeof(int32_t) * (t.N + 1), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_vn_edge, t.vn_edge.data(), sizeof(uint32_t) * t.nnz, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ws_event, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return cleanup_fail(LDPC_B200_ERR_CUDA, std::string("table upload: ") + cudaGetErrorString(cudaGetLastError()));
    int rc = make_plan(h);
    if (rc) { std::string m = g_err; return cleanup_fail(rc, m); }
    *out = h;
    return LDPC_B200_OK;
}

int ldpc_b200_create_wimax(ldpc_b200_handle* out, int K, int N, int rate, int device) {
    std::vector<int32_t> rp, ci;
    int M = 0;
    std::string msg = wimax_csr(K, N, rate, &rp, &ci, &M);
    if (!msg.empty()) return fail(LDPC_B200_ERR_ARG, msg);
    int rc = ldpc_b200_create(out, M, N, K, rp.data(), ci.data(), device);
    if (rc == LDPC_B200_OK) (*out)->layer_z = N / 24;  // one block row of the 802.16e base matrix
    return rc;
}

int ldpc_b200_destroy(ldpc_b200_handle h) {
    if (!h) return LDPC_B200_OK;
    {
        DeviceGuard guard(h->device);
        if (guard.ok) {
            cudaDeviceSynchronize();
            free_slots(h);
            if (h->qc_slot >= 0) qc_release_slot(h, h->device, h->qc_slot);
            if (h->qcm_slot >= 0 && h->device >= 0 && h->device < kQcMaxDevices) {
                std::lock_guard<std::mutex> lk(g_qcm_mu);
                if (g_qcm_owner[h->device][h->qcm_slot] == h) g_qcm_owner[h->device][h->qcm_slot] = nullptr;
            }
            cudaFree(h->d_syn_tab); cudaFree(h->d_iters_own);
            if (h->h_stats) cudaFreeHost(h->h_stats);
            cudaFree(h->dq_tabs);
            cudaFree(h->st_llr); cudaFree(h->st_info); cudaFree(h->st_hard); cudaFree(h->st_iters); cudaFree(h->st_post);
            cudaFree(h->d_avail);
            for (int i = 0; i < ldpc_b200_decoder::kStageSlots; ++i) {
                if (h->st_pin[i]) cudaFreeHost(h->st_pin[i]);
                if (h->st_pin_ev[i]) cudaEventDestroy(h->st_pin_ev[i]);
            }
            if (h->h_avail_vals) cudaFreeHost(h->h_avail_vals);
            if (h->h_status) cudaFreeHost(h->h_status);
            if (h->st_event) cudaEventDestroy(h->st_event);
            for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
            for (auto& r : h->registered) cudaHostUnregister(const_cast<void*>(r.first));
            for (int s = 0; s < kSlots; ++s)
                if (h->streams[s]) cudaStreamDestroy(h->streams[s]);
            cudaFree(h->d_row_ptr); cudaFree(h->d_cn_col); cudaFree(h->d_col_ptr); cudaFree(h->d_vn_edge);
            cudaFree(h->dc_cn_tab); cudaFree(h->dc_vn_tab); cudaFree(h->dc_var_of_pos); cudaFree(h->dc_out_addr);
            cudaFree(h->dg_cn_tab); cudaFree(h->dg_vn_tab); cudaFree(h->dg_var_of_pos); cudaFree(h->dg_pos_of_var);
            cudaFree(h->d16_cn_tab); cudaFree(h->d16_vn_tab); cudaFree(h->d16_var_of_pos); cudaFree(h->d16_pos_of_var);
            cudaFree(h->dt_cn_tab); cudaFree(h->de_xt);
            cudaFree(h->dw_cn_col); cudaFree(h->dw_vn_pos);
            cudaFree(h->ds_cn_tab); cudaFree(h->ds_vn_tab); cudaFree(h->ds_var_of_pos); cudaFree(h->ds_pos_of_var);
            cudaFree(h->d_counters); cudaFree(h->d_ws); cudaFree(h->d_ws_big);
            if (h->ws_event) cudaEventDestroy(h->ws_event);
        }
    }
    delete h;
    return LDPC_B200_OK;
}

int ldpc_b200_set_max_iter(ldpc_b200_handle h, int max_iter) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    if (max_iter < 1 || max_iter > 65535) return fail(LDPC_B200_ERR_ARG, "max_iter must be in 1..65535");
    std::lock_guard<std::mutex> lk(h->mu);
    h->max_iter = max_iter;
    return LDPC_B200_OK;
}

int ldpc_b200_set_early_termination(ldpc_b200_handle h, int on) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->early = on ? 1 : 0;
    return LDPC_B200_OK;
}

int ldpc_b200_set_algorithm(ldpc_b200_handle h, int algorithm) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    if (algorithm < LDPC_B200_ALG_MIN_SUM || algorithm > LDPC_B200_ALG_FUSED_LAYERED)
        return fail(LDPC_B200_ERR_ARG, "unknown algorithm");
    std::lock_guard<std::mutex> lk(h->mu);
    if (algorithm == h->algorithm) return LDPC_B200_OK;
    if (algorithm == LDPC_B200_ALG_FUSED_MIN_SUM) {  // any-size kernel, its own workspace: nothing to plan
        h->algorithm = algorithm;
        return LDPC_B200_OK;
    }
    if (algorithm == LDPC_B200_ALG_LAYERED_MIN_SUM || algorithm == LDPC_B200_ALG_FUSED_LAYERED) {
        // its own layout and tables; nothing of the flooding plan changes.  Fails here (not at the first decode)
        // when the code cannot be layered, so a caller can fall back.
        if (!h->tdmp_ready) {
            int rc = tdmp_plan(h);
            if (rc) return rc;
        }
        h->algorithm = algorithm;
        return LDPC_B200_OK;
    }
    const int prev = h->algorithm;
    h->algorithm = algorithm;
    // the plan and the group tables were built for one flooding algorithm (sum-product multiplies in CSR edge order and
    // admits check degree 20; min-sum reorders edges for bank placement): keep them only if that is the one asked for
    if (h->planned && h->plan_alg == algorithm) return LDPC_B200_OK;
    if (h->group_ready) {
        DeviceGuard guard(h->device);
        if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(h->dg_cn_tab); cudaFree(h->dg_vn_tab); cudaFree(h->dg_var_of_pos); cudaFree(h->dg_pos_of_var);
        h->dg_cn_tab = h->dg_vn_tab = h->dg_var_of_pos = h->dg_pos_of_var = nullptr;
        h->group_ready = false;
    }
    const int rc = make_plan(h);
    if (rc) {  // leave the handle as it was
        const std::string keep = g_err;
        h->algorithm = prev;
        (void)make_plan(h);
        g_err = keep;
    } else if (algorithm == LDPC_B200_ALG_SUM_PRODUCT) {
        (void)spq_prepare(h);   // quasi-cyclic codes: the kernel of ldpc_spq.cuh (known before the first decode sizes its chunks)
    }
    return rc;
}

int ldpc_b200_set_option(ldpc_b200_handle h, const char* name, long long value) {
    if (!h || !name) return fail(LDPC_B200_ERR_ARG, "null argument");
    const int n = (int)(sizeof(kOptionNames) / sizeof(kOptionNames[0]));
    for (int i = 0; i < n; ++i) {
        if (std::strcmp(kOptionNames[i].name, name) != 0) continue;
        if (!kOptionNames[i].runtime)
            return fail(LDPC_B200_ERR_UNSUPPORTED, std::string("option '") + name + "' shapes the plan: set LDPC_B200_<NAME> before ldpc_b200_create");
        std::lock_guard<std::mutex> lk(h->mu);
        option_store(&h->opt, kOptionNames[i], value);
        return LDPC_B200_OK;
    }
    return fail(LDPC_B200_ERR_ARG, std::string("unknown option '") + name + "'");
}

int ldpc_b200_set_layer_height(ldpc_b200_handle h, int z) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    if (z < 1) return fail(LDPC_B200_ERR_ARG, "layer height must be positive");
    std::lock_guard<std::mutex> lk(h->mu);
    if (z == h->layer_z) return LDPC_B200_OK;
    if (h->tdmp_ready) {
        DeviceGuard guard(h->device);
        if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(h->dt_cn_tab);
        h->dt_cn_tab = nullptr;
        h->tdmp_ready = false;
    }
    const int prev = h->layer_z;
    h->layer_z = z;
    if (h->algorithm == LDPC_B200_ALG_LAYERED_MIN_SUM) {
        int rc = tdmp_plan(h);
        if (rc) { h->layer_z = prev; return rc; }
    }
    return LDPC_B200_OK;
}

int ldpc_b200_encoder_init(ldpc_b200_handle h) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->encoder_ready) return LDPC_B200_OK;
    const HostTables& t = h->host;
    if (h->K % 8 || t.N % 8) return fail(LDPC_B200_ERR_UNSUPPORTED, "the encoder needs K and N to be multiples of 8");
    if (h->K > 32 * 32 * kEncMaxKW) return fail(LDPC_B200_ERR_UNSUPPORTED, "the encoder handles K <= 4096");
    std::vector<uint32_t> xt;
    const std::string msg = build_encoder(t, h->K, &xt, &h->enc_kw, &h->enc_mw);
    if (!msg.empty()) return fail(LDPC_B200_ERR_UNSUPPORTED, msg);
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
    CU_TRY(cudaMalloc(&h->de_xt, xt.size() * 4));
    CU_TRY(cudaMemcpy(h->de_xt, xt.data(), xt.size() * 4, cudaMemcpyHostToDevice));
    h->table_bytes += xt.size() * 4;
    h->encoder_ready = true;
    return LDPC_B200_OK;
}

int ldpc_b200_encode_device(ldpc_b200_handle h, const uint8_t* d_info, int64_t ncw, uint8_t* d_codewords, void* stream) {
    if (!h || (ncw > 0 && (!d_info || !d_codewords)) || ncw < 0) return fail(LDPC_B200_ERR_ARG, "bad argument");
    if (!h->encoder_ready) {
        int rc = ldpc_b200_encoder_init(h);
        if (rc) return rc;
    }
    if (ncw == 0) return LDPC_B200_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
    EncodeParams q;
    q.xt = h->de_xt; q.K = h->K; q.N = h->host.N; q.M = h->host.M; q.KW = h->enc_kw; q.MW = h->enc_mw;
    q.info = d_info; q.out = d_codewords; q.ncw = ncw;
    const int64_t blocks = std::min<int64_t>((ncw + 7) / 8, (int64_t)h->sm_count * 8);
    ldpc_encode_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(q);
    CU_TRY(cudaGetLastError());
    h->launches += 1;
    return LDPC_B200_OK;
}

int ldpc_b200_encode_host(ldpc_b200_handle h, const uint8_t* info, int64_t ncw, uint8_t* codewords) {
    if (!h || (ncw > 0 && (!info || !codewords)) || ncw < 0) return fail(LDPC_B200_ERR_ARG, "bad argument");
    if (!h->encoder_ready) {
        int rc = ldpc_b200_encoder_init(h);
        if (rc) return rc;
    }
    if (ncw == 0) return LDPC_B200_OK;
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
    const size_t KB = (size_t)h->K / 8, NB = (size_t)h->host.N / 8;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    CU_TRY(cudaMalloc(&d_in, KB * (size_t)ncw));
    if (cudaMalloc(&d_out, NB * (size_t)ncw) != cudaSuccess) { cudaFree(d_in); return fail(LDPC_B200_ERR_CUDA, "cudaMalloc(codewords)"); }
    int rc = LDPC_B200_OK;
    if (cudaMemcpy(d_in, info, KB * (size_t)ncw, cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(LDPC_B200_ERR_CUDA, "cudaMemcpy(info)");
    if (!rc) rc = ldpc_b200_encode_device(h, d_in, ncw, d_out, nullptr);
    if (!rc && cudaMemcpy(codewords, d_out, NB * (size_t)ncw, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(LDPC_B200_ERR_CUDA, "cudaMemcpy(codewords)");
    cudaFree(d_in); cudaFree(d_out);
    return rc;
}

int ldpc_b200_set_path(ldpc_b200_handle h, int path) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    const int prev = h->forced_path;
    h->forced_path = path;
    int rc = make_plan(h);
    if (rc) { h->forced_path = prev; make_plan(h); return rc; }
    return LDPC_B200_OK;
}

int ldpc_b200_get_info(ldpc_b200_handle h, ldpc_b200_info* info) {
    if (!h || !info) return fail(LDPC_B200_ERR_ARG, "null argument");
    const HostTables& t = h->host;
    info->M = t.M; info->N = t.N; info->K = h->K; info->nnz = t.nnz;
    info->max_row_weight = t.max_row_weight; info->max_col_weight = t.max_col_weight;
    info->max_iter = h->max_iter; info->early_termination = h->early;
    info->device = h->device; info->sm_count = h->sm_count;
    info->path = h->plan.path; info->threads_per_cta = h->plan.threads; info->ctas = h->plan.ctas;
    info->codewords_per_cta = h->plan.cw_per_cta;
    info->smem_bytes = h->plan.smem;
    info->workspace_bytes = h->plan.ws_stride * sizeof(float) * (size_t)h->plan.ctas;
    info->table_bytes = h->table_bytes;
    info->kernel_variant = h->last_kernel;
    info->et_available = h->qcw_state == 1 ? 1 : (h->planned && h->plan.path == LDPC_B200_PATH_QC && h->plan.dmax == 1 && h->qcm_state == 1 ? 2 : 0);
    return LDPC_B200_OK;
}

int ldpc_b200_get_csr(ldpc_b200_handle h, int32_t* row_ptr, int32_t* col_idx) {
    if (!h || !row_ptr || !col_idx) return fail(LDPC_B200_ERR_ARG, "null argument");
    std::memcpy(row_ptr, h->host.row_ptr.data(), sizeof(int32_t) * (h->host.M + 1));
    std::memcpy(col_idx, h->host.col_idx.data(), sizeof(int32_t) * h->host.nnz);
    return LDPC_B200_OK;
}

}  // extern "C"

namespace {
// n timing-enabled events of the handle (created on demand); false if the runtime refuses
bool timing_events(ldpc_b200_decoder* h, size_t n) {
    while (h->tev.size() < n) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        h->tev.push_back(e);
    }
    return true;
}
double event_seconds(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { (void)cudaGetLastError(); return 0.0; }
    return (double)ms * 1e-3;
}

// Device buffers of the streamed host pipeline for up to `want` words per launch (grown, never shrunk).
// (h->mu held, current device = the handle's)
int ensure_streamed_buffers(ldpc_b200_decoder* h, int64_t want, bool hard, bool post) {
    const HostTables& t = h->host;
    const size_t KB = (h->K + 7) / 8, NB = (t.N + 7) / 8;
    if (h->st_cap < want || (hard && !h->st_has_hard) || (post && !h->st_has_post)) {
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(h->st_llr); cudaFree(h->st_info); cudaFree(h->st_hard); cudaFree(h->st_iters); cudaFree(h->st_post);
        h->st_llr = nullptr; h->st_info = nullptr; h->st_hard = nullptr; h->st_iters = nullptr; h->st_post = nullptr;
        const int64_t cap = std::max(want, h->st_cap);  // never shrink: alternating call sizes must not reallocate
        h->st_cap = 0;
        h->st_has_hard = h->st_has_hard || hard;
        h->st_has_post = h->st_has_post || post;
        CU_TRY(cudaMalloc(&h->st_llr, sizeof(float) * (size_t)cap * t.N));
        CU_TRY(cudaMalloc(&h->st_info, (size_t)cap * KB));
        CU_TRY(cudaMalloc(&h->st_iters, sizeof(int32_t) * (size_t)cap));
        if (h->st_has_hard) CU_TRY(cudaMalloc(&h->st_hard, (size_t)cap * NB));
        if (h->st_has_post) CU_TRY(cudaMalloc(&h->st_post, sizeof(float) * (size_t)cap * t.N));
        h->st_cap = cap;
    }
    return LDPC_B200_OK;
}

// CUDA loads a kernel's code at its first launch.  Loading needs the device idle, so a kernel first launched behind
// the persistent launch of the host-buffer path -- which is spinning on input the host has not queued yet -- sat
// out the kernel's whole wait bound: the first Coder::decode of a process took 4 s (tools/setdevices_first_call.py).
// Every kernel a decode with the handle's current algorithm can launch is therefore run once here, on a few all-zero
// words, where the reference compiles its OpenCL kernels (addDecodeType, MyLdpc.cpp:387-437).
// (h->mu held, current device = the handle's)
int warm_kernels_locked(ldpc_b200_decoder* h) {
    const unsigned bit = 1u << (unsigned)h->algorithm;
    if ((h->warmed & bit) || h->opt.no_warm) return LDPC_B200_OK;
    if (!h->planned) { const int rc = make_plan(h); if (rc) return rc; }
    const HostTables& t = h->host;
    const size_t KB = (h->K + 7) / 8, NB = (t.N + 7) / 8;
    const int64_t n = std::max<int64_t>(1, h->plan.cw_per_cta);
    // scratch of its own: [channel values | info bytes | hard bytes | iteration counts]
    const size_t o_info = sizeof(float) * (size_t)n * t.N, o_hard = o_info + (((size_t)n * KB + 15) & ~(size_t)15),
                 o_iters = o_hard + (((size_t)n * NB + 15) & ~(size_t)15), bytes = o_iters + sizeof(int32_t) * (size_t)n;
    char* d = nullptr;
    CU_TRY(cudaMalloc(&d, bytes));
    float* llr = reinterpret_cast<float*>(d);
    uint8_t* info = reinterpret_cast<uint8_t*>(d + o_info);
    uint8_t* hard = reinterpret_cast<uint8_t*>(d + o_hard);
    int32_t* iters = reinterpret_cast<int32_t*>(d + o_iters);
    cudaStream_t s = nullptr;  // (the legacy stream: the handle's own streams may not exist yet)
    const int64_t launches = h->launches;
    const int last_kernel = h->last_kernel, qc_et = h->opt.qc_et;
    int rc = cudaMemsetAsync(d, 0, bytes, s) == cudaSuccess ? LDPC_B200_OK : fail(LDPC_B200_ERR_CUDA, "kernel warm-up: memset failed");
    const bool qc_min_sum = h->algorithm == LDPC_B200_ALG_MIN_SUM && h->plan.path == LDPC_B200_PATH_QC;
    if (rc == LDPC_B200_OK && qc_min_sum && qc_et < 0) {
        // lockstep kernel, per-codeword kernel (fp32, float16 and int8 channel values), iteration statistics
        const int fmts[3] = {LDPC_B200_LLR_F32, LDPC_B200_LLR_F16, LDPC_B200_LLR_I8};
        for (int et = 0; et <= 1 && rc == LDPC_B200_OK; ++et) {
            h->opt.qc_et = et;
            rc = launch_decode(h, llr, n, info, hard, iters, nullptr, s);
        }
        const bool per_word = h->qcw_state == 1 || h->qcm_state == 1;
        for (int f = 1; f < 3 && per_word && h->early && rc == LDPC_B200_OK; ++f) {
            h->cur_fmt = fmts[f];
            rc = launch_decode(h, llr, n, info, hard, iters, nullptr, s);
        }
        h->cur_fmt = LDPC_B200_LLR_F32;
        h->opt.qc_et = -1;
        if (rc == LDPC_B200_OK) rc = launch_decode(h, llr, n, info, hard, nullptr, nullptr, s);
        h->opt.qc_et = qc_et;
    } else if (rc == LDPC_B200_OK) {
        rc = launch_decode(h, llr, n, info, hard, iters, nullptr, s);
    }
    const cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(d);
    h->launches = launches;          // the caller's launch count and kernel choice start from a clean state
    h->last_kernel = last_kernel;
    h->stat_tick = 0;
    if (h->h_stats) h->h_stats[0] = h->h_stats[1] = 0ull;
    if (rc) return rc;
    if (e != cudaSuccess) return fail(LDPC_B200_ERR_CUDA, std::string("kernel warm-up: ") + cudaGetErrorString(e));
    h->warmed |= bit;
    return LDPC_B200_OK;
}

// (h->mu held)
int reserve_locked(ldpc_b200_decoder* h, int64_t batch) {
    if (batch <= h->reserved) {
        if (h->reserved < 1) return LDPC_B200_OK;
        DeviceGuard guard(h->device);
        if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
        return warm_kernels_locked(h);   // (a new algorithm since the last reservation)
    }
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
    CU_TRY(cudaDeviceSynchronize());
    free_slots(h);
    const HostTables& t = h->host;
    const size_t KB = (h->K + 7) / 8, NB = (t.N + 7) / 8;
    for (int s = 0; s < kSlots; ++s) {
        if (!h->streams[s]) CU_TRY(cudaStreamCreateWithFlags(&h->streams[s], cudaStreamNonBlocking));
        CU_TRY(cudaMalloc(&h->s_llr[s], sizeof(float) * (size_t)batch * t.N));
        CU_TRY(cudaMalloc(&h->s_info[s], (size_t)batch * KB));
        CU_TRY(cudaMalloc(&h->s_iters[s], sizeof(int32_t) * (size_t)batch));
        CU_TRY(cudaMalloc(&h->s_hard[s], (size_t)batch * NB));
    }
    h->reserved = batch;
    // quasi-cyclic path: host buffers go through the persistent launch; set up what does not depend on the call's size
    // (streams, counters, the pinned staging ring for pageable input) here, outside a caller's timed decode
    if (h->planned && h->plan.path == LDPC_B200_PATH_QC && h->algorithm == LDPC_B200_ALG_MIN_SUM) {
        if (!h->st_event) CU_TRY(cudaEventCreateWithFlags(&h->st_event, cudaEventDisableTiming));
        if (!h->d_avail) CU_TRY(cudaMalloc(&h->d_avail, 2 * sizeof(unsigned long long)));
        const int64_t g = h->plan.cw_per_cta;
        const int64_t chunk = std::max<int64_t>(g, ((((int64_t)4 << 20)) / ((int64_t)t.N * 4)) / g * g);
        const size_t need = sizeof(float) * (size_t)chunk * t.N;
        if (!h->h_status) CU_TRY(cudaMallocHost(&h->h_status, sizeof(int)));
        for (int s = 0; s < 2; ++s)
            if (!h->streams[s]) CU_TRY(cudaStreamCreateWithFlags(&h->streams[s], cudaStreamNonBlocking));
        {   // the launch buffers for a call of `batch` words (what the reference allocates in addDecodeType,
            // MyLdpc.cpp:387-437): the first decode() then costs what the tenth does
            int64_t batch_bytes = (int64_t)512 << 20;
            if (h->opt.stream_batch_kb >= 1) batch_bytes = h->opt.stream_batch_kb << 10;
            const int64_t batch_cap = std::max<int64_t>(g, (batch_bytes / ((int64_t)t.N * 4)) / g * g);
            const int rc = ensure_streamed_buffers(h, std::min(batch, batch_cap), false, false);
            if (rc) return rc;
            const int64_t nchunks_max = (std::min(batch, batch_cap) + chunk - 1) / chunk + 8;
            if (h->h_avail_cap < nchunks_max) {
                if (h->h_avail_vals) cudaFreeHost(h->h_avail_vals);
                h->h_avail_vals = nullptr; h->h_avail_cap = 0;
                CU_TRY(cudaMallocHost(&h->h_avail_vals, sizeof(unsigned long long) * (size_t)nchunks_max));
                h->h_avail_cap = nchunks_max;
            }
            (void)timing_events(h, 5);
        }
        if (h->st_pin_bytes < need && !h->opt.no_staged) {
            for (int i = 0; i < ldpc_b200_decoder::kStageSlots; ++i) {
                if (h->st_pin[i]) { cudaFreeHost(h->st_pin[i]); h->st_pin[i] = nullptr; }
            }
            h->st_pin_bytes = 0;
            for (int i = 0; i < ldpc_b200_decoder::kStageSlots; ++i) {
                CU_TRY(cudaMallocHost(&h->st_pin[i], need));
                std::memset(h->st_pin[i], 0, need);  // (the host threads' first copy into it should not be the one that maps it)
                if (!h->st_pin_ev[i]) CU_TRY(cudaEventCreateWithFlags(&h->st_pin_ev[i], cudaEventDisableTiming));
            }
            h->st_pin_bytes = need;
        }
    }
    return warm_kernels_locked(h);
}
}  // namespace

extern "C" {

int ldpc_b200_reserve(ldpc_b200_handle h, int64_t batch) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    if (batch < 1) return fail(LDPC_B200_ERR_ARG, "batch must be positive");
    std::lock_guard<std::mutex> lk(h->mu);
    return reserve_locked(h, batch);
}

int ldpc_b200_decode_device(ldpc_b200_handle h, const float* d_llr, int64_t ncw, uint8_t* d_info, uint8_t* d_hard,
                            int32_t* d_iters, float* d_post, void* stream) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    if (ncw < 0) return fail(LDPC_B200_ERR_ARG, "ncw must be >= 0");
    if (ncw > 0 && !d_llr) return fail(LDPC_B200_ERR_ARG, "d_llr is null");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
    if (ncw > 0) {   // a handle nobody reserved on: its kernels load here, once, not at the launch that first switches kernels
        const int wrc = warm_kernels_locked(h);   // (3.5 ms in the middle of a stream of decodes: tools/auto_regime_probe.py)
        if (wrc) return wrc;
    }
    return launch_decode(h, d_llr, ncw, d_info, d_hard, d_iters, d_post, (cudaStream_t)stream);
}

namespace {

// Host buffers through ONE persistent launch per batch: all input chunks are queued on a copy stream, each
// followed by an 8-byte write that advances *d_avail; the kernel, launched on a second stream, takes words from
// its work queue as usual and only waits when it is ahead of the copies.  No kernel boundaries, no tail per chunk,
// and the device-to-host copy of the bits follows the kernel.  Everything is queued before the launch, so a
// serialising tool (profiler, CUDA_LAUNCH_BLOCKING) degrades to copy-then-decode instead of deadlocking.
//
// staged = true (pageable input): the driver would stage such copies synchronously at ~7.5 GB/s; instead four host
// threads copy the chunks into a ring of pinned buffers and queue the DMA of each (in chunk order), while the kernel
// -- launched first in this mode -- already decodes.  If nothing can feed the kernel (a tool that serialises kernel
// launches and blocks the other threads' API calls), its 4 s bound expires and the caller falls back to the chunked
// pipeline: returns kStreamedRetry.
constexpr int kStreamedRetry = 1;
// (h->mu held, current device = the handle's; the caller drains the streams when this returns non-zero)
// The copy stream announces a landed chunk by advancing the device counter *avail.  A stream memory operation
// (cuStreamWriteValue64, looked up through the runtime: no link against libcuda) does that without a DMA descriptor of
// its own; the 8-byte copy from pinned memory is the fallback (option "avail_memcpy" forces it).
typedef CUresult (*StreamWriteValue64Fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
StreamWriteValue64Fn stream_write_value64() {
    static StreamWriteValue64Fn fn = []() -> StreamWriteValue64Fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<StreamWriteValue64Fn>(p);
    }();
    return fn;
}
cudaError_t announce_chunk(ldpc_b200_decoder* h, int64_t j, unsigned long long value, cudaStream_t cs) {
    StreamWriteValue64Fn w = h->opt.avail_memcpy ? nullptr : stream_write_value64();
    if (w && w(cs, reinterpret_cast<CUdeviceptr>(h->d_avail), (cuuint64_t)value, CU_STREAM_WRITE_VALUE_DEFAULT) == CUDA_SUCCESS) return cudaSuccess;
    h->h_avail_vals[j] = value;
    return cudaMemcpyAsync(h->d_avail, h->h_avail_vals + j, sizeof(unsigned long long), cudaMemcpyHostToDevice, cs);
}

int decode_host_streamed_body(ldpc_b200_decoder* h, const float* llr, int64_t ncw, uint8_t* info, uint8_t* hard,
                              int32_t* iters, float* post, bool staged, const void* packed = nullptr, int format = LDPC_B200_LLR_F32,
                              float scale = 1.0f) {
    // packed (pinned only): the chunks are copied as they are and the per-codeword kernels widen them at the load
    const size_t esz = format == LDPC_B200_LLR_F16 ? 2 : (format == LDPC_B200_LLR_I8 ? 1 : 4);
    const char* src_bytes = packed ? static_cast<const char*>(packed) : reinterpret_cast<const char*>(llr);
    const HostTables& t = h->host;
    const size_t KB = (h->K + 7) / 8, NB = (t.N + 7) / 8;
    const int64_t g = h->plan.cw_per_cta;
    int64_t batch_bytes = (int64_t)512 << 20;  // <= 512 MB of channel values per launch
    if (h->opt.stream_batch_kb >= 1) batch_bytes = h->opt.stream_batch_kb << 10;
    const int64_t batch_cap = std::max<int64_t>(g, (batch_bytes / ((int64_t)t.N * 4)) / g * g);
    const int64_t want = std::min(ncw, batch_cap);
    for (int s = 0; s < 2; ++s)
        if (!h->streams[s]) CU_TRY(cudaStreamCreateWithFlags(&h->streams[s], cudaStreamNonBlocking));
    if (!h->st_event) CU_TRY(cudaEventCreateWithFlags(&h->st_event, cudaEventDisableTiming));
    if (!h->d_avail) CU_TRY(cudaMalloc(&h->d_avail, 2 * sizeof(unsigned long long)));
    if (!h->h_status) CU_TRY(cudaMallocHost(&h->h_status, sizeof(int)));
    { const int rc = ensure_streamed_buffers(h, want, hard != nullptr, post != nullptr); if (rc) return rc; }
    // input chunks of ~1 MB, 2 MB, then ~4 MB: the first words land after a few tens of microseconds and PCIe stays
    // efficient.  Larger chunks lose: the kernel consumes words in order at 70 % of the PCIe rate, so it keeps
    // running into the end of the announced range and waits for a whole chunk (measured, cfg2 end to end:
    // 4.08 ms with 4 MB chunks, 4.11 with 8 MB, 4.27 with 16 MB).
    auto words_of = [&](int64_t bytes) { return std::max<int64_t>(g, (bytes / ((int64_t)t.N * (int64_t)esz)) / g * g); };
    int64_t chunk0 = words_of((int64_t)1 << 20), chunk_max = words_of((int64_t)4 << 20);
    // Words that stop early (the previous launches' mean iteration count is at most a quarter of the cap): the kernel is
    // several times faster than the copy and only the copy's own rate matters -- chunks grow to 16 MB (measured at
    // 3.5 dB, 65,536 words of Test.cpp's code: 4 MB 3.28 ms, 8 MB 3.13, 16 MB 3.09, 32 MB 3.05 against 2.72 for one
    // plain copy; the kernel's tail after the last chunk grows from 0.10 to 0.17 ms)
    if (!staged && h->h_stats && h->h_stats[1] > 0 && h->h_stats[0] * 4ull <= (unsigned long long)h->max_iter * h->h_stats[1])
        chunk_max = words_of(((int64_t)4 << 20) * (int64_t)esz);   // (packed values: the same words per chunk -- int8 copies are as fast as the kernel)
    if (h->opt.stream_chunk >= 1) chunk0 = chunk_max = (h->opt.stream_chunk + g - 1) / g * g;
    if (staged) {
        chunk0 = chunk_max;  // fixed-size chunks = ring slots
        const size_t need = sizeof(float) * (size_t)chunk_max * t.N;
        if (h->st_pin_bytes < need) {
            CU_TRY(cudaDeviceSynchronize());
            for (int i = 0; i < ldpc_b200_decoder::kStageSlots; ++i) {
                if (h->st_pin[i]) { cudaFreeHost(h->st_pin[i]); h->st_pin[i] = nullptr; }
            }
            h->st_pin_bytes = 0;
            for (int i = 0; i < ldpc_b200_decoder::kStageSlots; ++i) {
                CU_TRY(cudaMallocHost(&h->st_pin[i], need));
                if (!h->st_pin_ev[i]) CU_TRY(cudaEventCreateWithFlags(&h->st_pin_ev[i], cudaEventDisableTiming));
            }
            h->st_pin_bytes = need;
        }
    }
    const int64_t nchunks_max = (want + chunk_max - 1) / chunk_max + 8;
    if (h->h_avail_cap < nchunks_max) {
        if (h->h_avail_vals) cudaFreeHost(h->h_avail_vals);
        h->h_avail_vals = nullptr; h->h_avail_cap = 0;
        CU_TRY(cudaMallocHost(&h->h_avail_vals, sizeof(unsigned long long) * (size_t)nchunks_max));
        h->h_avail_cap = nchunks_max;
    }
    cudaStream_t cs = h->streams[0], ks = h->streams[1];
    const bool timed = timing_events(h, 5);  // [0] copies start, [1] copies end, [2] kernel start, [3] kernel end, [4] read-back end
    for (int64_t off = 0; off < ncw; off += batch_cap) {
        const int64_t n = std::min(batch_cap, ncw - off);
        // (a later batch reuses the device buffers: its copies wait for the previous batch's kernel and read-back)
        if (off > 0) { CU_TRY(cudaEventRecord(h->st_event, ks)); CU_TRY(cudaStreamWaitEvent(cs, h->st_event, 0)); CU_TRY(cudaStreamSynchronize(cs)); }
        CU_TRY(cudaMemsetAsync(h->d_avail, 0, 2 * sizeof(unsigned long long), cs));
        CU_TRY(cudaEventRecord(h->st_event, cs));
        CU_TRY(cudaStreamWaitEvent(ks, h->st_event, 0));  // the kernel must not see a stale count
        if (timed) { CU_TRY(cudaEventRecord(h->tev[0], cs)); CU_TRY(cudaEventRecord(h->tev[2], ks)); }
        int rc = LDPC_B200_OK;
        if (!staged) {
            int64_t j = 0, chunk = chunk0;
            for (int64_t c0 = 0; c0 < n; c0 += chunk, chunk = std::min(chunk * 2, chunk_max), ++j) {
                const int64_t m = std::min(chunk, n - c0);
                CU_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(h->st_llr) + esz * (size_t)c0 * t.N, src_bytes + esz * (size_t)(off + c0) * t.N,
                                       esz * (size_t)m * t.N, cudaMemcpyHostToDevice, cs));
                CU_TRY(announce_chunk(h, j, (unsigned long long)(c0 + m), cs));
            }
            h->cur_avail = h->d_avail;
            h->cur_fmt = packed ? format : LDPC_B200_LLR_F32; h->cur_scale = scale;
            rc = launch_decode(h, h->st_llr, n, info ? h->st_info : nullptr, hard ? h->st_hard : nullptr,
                               iters ? h->st_iters : nullptr, post ? h->st_post : nullptr, ks);
            h->cur_avail = nullptr;
            h->cur_fmt = LDPC_B200_LLR_F32; h->cur_scale = 1.0f;
        } else {
            h->cur_avail = h->d_avail;
            rc = launch_decode(h, h->st_llr, n, info ? h->st_info : nullptr, hard ? h->st_hard : nullptr,
                               iters ? h->st_iters : nullptr, post ? h->st_post : nullptr, ks);
            h->cur_avail = nullptr;
            if (rc == LDPC_B200_OK) {
                constexpr int S = ldpc_b200_decoder::kStageSlots;
                const int64_t cw = chunk_max, nch = (n + cw - 1) / cw;
                std::atomic<int64_t> next_chunk{0}, next_enq{0};
                std::atomic<int> err{0};
                const float* src = llr + (size_t)off * t.N;
                auto work = [&]() {
                    if (cudaSetDevice(h->device) != cudaSuccess) { err.store(1); return; }
                    for (;;) {
                        const int64_t j = next_chunk.fetch_add(1);
                        if (j >= nch || err.load()) break;
                        const int slot = (int)(j % S);
                        if (j >= S) {  // the slot's previous chunk must have been queued and its DMA finished
                            while (next_enq.load() <= j - S && !err.load()) std::this_thread::yield();
                            if (err.load()) break;
                            if (cudaEventSynchronize(h->st_pin_ev[slot]) != cudaSuccess) { err.store(1); break; }
                        }
                        const int64_t c0 = j * cw, m = std::min(cw, n - c0);
                        if (h->opt.stage_nt) stage_copy_nt(h->st_pin[slot], src + (size_t)c0 * t.N, sizeof(float) * (size_t)m * t.N);
                        else std::memcpy(h->st_pin[slot], src + (size_t)c0 * t.N, sizeof(float) * (size_t)m * t.N);
                        while (next_enq.load() != j && !err.load()) std::this_thread::yield();  // queue in chunk order
                        if (err.load()) break;
                        if (cudaMemcpyAsync(h->st_llr + (size_t)c0 * t.N, h->st_pin[slot], sizeof(float) * (size_t)m * t.N, cudaMemcpyHostToDevice, cs) != cudaSuccess ||
                            announce_chunk(h, j, (unsigned long long)(c0 + m), cs) != cudaSuccess ||
                            cudaEventRecord(h->st_pin_ev[slot], cs) != cudaSuccess) { err.store(1); break; }
                        next_enq.store(j + 1);
                    }
                };
                const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::max(1, h->opt.stage_threads), nch, (int64_t)std::max(1u, std::thread::hardware_concurrency())}));
                std::vector<std::thread> pool;
                for (int i = 1; i < nthreads; ++i) pool.emplace_back(work);
                work();
                for (auto& th : pool) th.join();
                if (err.load()) return fail(LDPC_B200_ERR_CUDA, std::string("staged input copy: ") + cudaGetErrorString(cudaGetLastError()));
            }
        }
        if (rc) return rc;
        if (timed) { CU_TRY(cudaEventRecord(h->tev[1], cs)); CU_TRY(cudaEventRecord(h->tev[3], ks)); }
        if (info) CU_TRY(cudaMemcpyAsync(info + (size_t)off * KB, h->st_info, (size_t)n * KB, cudaMemcpyDeviceToHost, ks));
        if (hard) CU_TRY(cudaMemcpyAsync(hard + (size_t)off * NB, h->st_hard, (size_t)n * NB, cudaMemcpyDeviceToHost, ks));
        if (iters) CU_TRY(cudaMemcpyAsync(iters + off, h->st_iters, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ks));
        if (post) CU_TRY(cudaMemcpyAsync(post + (size_t)off * t.N, h->st_post, sizeof(float) * (size_t)n * t.N, cudaMemcpyDeviceToHost, ks));
        *h->h_status = 0;
        CU_TRY(cudaMemcpyAsync(h->h_status, reinterpret_cast<int*>(h->d_avail + 1), sizeof(int), cudaMemcpyDeviceToHost, ks));
        if (timed) CU_TRY(cudaEventRecord(h->tev[4], ks));
        CU_TRY(cudaStreamSynchronize(cs));
        CU_TRY(cudaStreamSynchronize(ks));
        if (timed) {
            h->timing.h2d_s += event_seconds(h->tev[0], h->tev[1]);
            h->timing.kernel_s += event_seconds(h->tev[2], h->tev[3]);  // (waits for streamed input included)
            h->timing.d2h_s += event_seconds(h->tev[3], h->tev[4]);
        }
        // the kernel gave up waiting for its input (a tool that serialises launches, a copy engine busy elsewhere,
        // time-slicing): nothing is wrong with the data -- the caller reruns the call through the chunked pipeline
        if (*h->h_status) return kStreamedRetry;
    }
    return LDPC_B200_OK;
}

int decode_host_streamed(ldpc_b200_decoder* h, const float* llr, int64_t ncw, uint8_t* info, uint8_t* hard,
                         int32_t* iters, float* post, bool staged, const void* packed = nullptr, int format = LDPC_B200_LLR_F32,
                         float scale = 1.0f) {
    const int rc = decode_host_streamed_body(h, llr, ncw, info, hard, iters, post, staged, packed, format, scale);
    h->cur_fmt = LDPC_B200_LLR_F32; h->cur_scale = 1.0f; h->cur_avail = nullptr;
    if (rc != LDPC_B200_OK) {  // no copy into the caller's buffers may still be in flight when an error returns
        const std::string keep = g_err;
        for (int s = 0; s < 2; ++s) if (h->streams[s]) cudaStreamSynchronize(h->streams[s]);
        (void)cudaGetLastError();
        g_err = keep;
    }
    return rc;
}

// The chunked pipeline below with a PAGEABLE caller buffer.  cudaMemcpyAsync from pageable memory is staged by the driver,
// synchronously, at ~7.5 GB/s; a long code's kernel wants its input several times faster (regular (3,6) N = 8192:
// 4.3 GB per 131,072 words in 86 ms).  Here host threads copy pieces of at most one ring slot (>= 4 MB) into the pinned
// ring with streaming stores, and the thread that holds the next piece IN ORDER queues its DMA on the chunk's stream and,
// behind a chunk's last piece, the chunk's launch and read-back: streams, device buffers and results are exactly those
// of the plain loop.  (h->mu held, current device = the handle's; fp32 input only)
int decode_host_chunked_staged(ldpc_b200_decoder* h, const float* llr, int64_t ncw, uint8_t* info, uint8_t* hard,
                               int32_t* iters, float* post, int64_t chunk) {
    const HostTables& t = h->host;
    const size_t KB = (h->K + 7) / 8, NB = (t.N + 7) / 8;
    constexpr int S = ldpc_b200_decoder::kStageSlots;
    const size_t wbytes = sizeof(float) * (size_t)t.N;
    const size_t need = std::max<size_t>((size_t)4 << 20, wbytes);
    if (h->st_pin_bytes < need) {
        CU_TRY(cudaDeviceSynchronize());
        for (int i = 0; i < S; ++i) {
            if (h->st_pin[i]) { cudaFreeHost(h->st_pin[i]); h->st_pin[i] = nullptr; }
        }
        h->st_pin_bytes = 0;
        for (int i = 0; i < S; ++i) {
            CU_TRY(cudaMallocHost(&h->st_pin[i], need));
            std::memset(h->st_pin[i], 0, need);
            if (!h->st_pin_ev[i]) CU_TRY(cudaEventCreateWithFlags(&h->st_pin_ev[i], cudaEventDisableTiming));
        }
        h->st_pin_bytes = need;
    }
    const int64_t pw = std::max<int64_t>(1, (int64_t)(h->st_pin_bytes / wbytes));   // words per piece
    const int64_t nchunks = (ncw + chunk - 1) / chunk, ppc = (chunk + pw - 1) / pw;  // pieces per full chunk
    const int64_t n_last = ncw - (nchunks - 1) * chunk;
    const int64_t npieces = (nchunks - 1) * ppc + (n_last + pw - 1) / pw;
    constexpr int64_t kTimedChunks = 64;
    const int64_t ntimed = std::min(nchunks, kTimedChunks);
    const bool timed = timing_events(h, (size_t)(4 * ntimed));
    std::atomic<int64_t> next_piece{0}, next_enq{0};
    std::atomic<int> err{0};
    int err_rc = LDPC_B200_ERR_CUDA;
    std::string err_msg;
    auto cuda_failed = [&](const char* what) {   // (called by the thread whose turn it is, or before any queueing: no race on err_msg)
        if (!err.exchange(1)) { err_rc = LDPC_B200_ERR_CUDA; err_msg = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError()); }
    };
    auto work = [&]() {
        if (cudaSetDevice(h->device) != cudaSuccess) { cuda_failed("cudaSetDevice"); return; }
        for (;;) {
            const int64_t j = next_piece.fetch_add(1);
            if (j >= npieces || err.load()) break;
            const int rs = (int)(j % S);
            const int64_t c = j / ppc, k = j % ppc;
            const int64_t c_first = c * chunk, c_n = std::min(chunk, ncw - c_first);
            const int64_t w0 = k * pw, m = std::min(pw, c_n - w0);            // words [w0, w0 + m) of chunk c
            const bool last_of_chunk = w0 + m == c_n;
            if (j >= S) {  // the ring slot's previous piece must have been queued and its DMA finished
                while (next_enq.load() <= j - S && !err.load()) std::this_thread::yield();
                if (err.load()) break;
                if (cudaEventSynchronize(h->st_pin_ev[rs]) != cudaSuccess) { cuda_failed("staging ring"); break; }
            }
            const float* src = llr + (size_t)(c_first + w0) * t.N;
            if (h->opt.stage_nt) stage_copy_nt(h->st_pin[rs], src, wbytes * (size_t)m);
            else std::memcpy(h->st_pin[rs], src, wbytes * (size_t)m);
            while (next_enq.load() != j && !err.load()) std::this_thread::yield();  // queue in piece order
            if (err.load()) break;
            const int slot = (int)(c % kSlots);
            cudaStream_t st = h->streams[slot];
            const bool tc = timed && c < ntimed;
            bool ok = true;
            if (k == 0 && tc) ok = cudaEventRecord(h->tev[4 * c], st) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(h->s_llr[slot] + (size_t)w0 * t.N, h->st_pin[rs], wbytes * (size_t)m, cudaMemcpyHostToDevice, st) == cudaSuccess;
            ok = ok && cudaEventRecord(h->st_pin_ev[rs], st) == cudaSuccess;
            if (ok && last_of_chunk) {
                if (tc) ok = cudaEventRecord(h->tev[4 * c + 1], st) == cudaSuccess;
                if (ok) {
                    const int rc = launch_decode(h, h->s_llr[slot], c_n, info ? h->s_info[slot] : nullptr, hard ? h->s_hard[slot] : nullptr,
                                                 iters ? h->s_iters[slot] : nullptr, post ? h->s_post[slot] : nullptr, st);
                    if (rc) { if (!err.exchange(1)) { err_rc = rc; err_msg = g_err; } break; }
                }
                if (ok && tc) ok = cudaEventRecord(h->tev[4 * c + 2], st) == cudaSuccess;
                if (ok && info) ok = cudaMemcpyAsync(info + (size_t)c_first * KB, h->s_info[slot], (size_t)c_n * KB, cudaMemcpyDeviceToHost, st) == cudaSuccess;
                if (ok && hard) ok = cudaMemcpyAsync(hard + (size_t)c_first * NB, h->s_hard[slot], (size_t)c_n * NB, cudaMemcpyDeviceToHost, st) == cudaSuccess;
                if (ok && iters) ok = cudaMemcpyAsync(iters + c_first, h->s_iters[slot], sizeof(int32_t) * (size_t)c_n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
                if (ok && post) ok = cudaMemcpyAsync(post + (size_t)c_first * t.N, h->s_post[slot], wbytes * (size_t)c_n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
                if (ok && tc) ok = cudaEventRecord(h->tev[4 * c + 3], st) == cudaSuccess;
            }
            if (!ok) { cuda_failed("staged chunk"); break; }
            next_enq.store(j + 1);
        }
    };
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::max(1, h->opt.stage_threads), npieces, (int64_t)std::max(1u, std::thread::hardware_concurrency())}));
    std::vector<std::thread> pool;
    for (int i = 1; i < nthreads; ++i) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    if (err.load()) return fail(err_rc, err_msg);
    for (int s2 = 0; s2 < kSlots; ++s2) CU_TRY(cudaStreamSynchronize(h->streams[s2]));
    if (timed) {
        const double scale = (double)nchunks / (double)ntimed;
        for (int64_t c = 0; c < ntimed; ++c) {
            h->timing.h2d_s += scale * event_seconds(h->tev[4 * c], h->tev[4 * c + 1]);
            h->timing.kernel_s += scale * event_seconds(h->tev[4 * c + 1], h->tev[4 * c + 2]);
            h->timing.d2h_s += scale * event_seconds(h->tev[4 * c + 2], h->tev[4 * c + 3]);
        }
    }
    return LDPC_B200_OK;
}

// The chunked 3-stream pipeline: H2D, kernel and D2H of consecutive chunks overlap across kSlots streams.
// (h->mu held, current device = the handle's)
int decode_host_chunked_body(ldpc_b200_decoder* h, const float* llr, int64_t ncw, uint8_t* info, uint8_t* hard,
                             int32_t* iters, float* post, const void* packed = nullptr, int format = LDPC_B200_LLR_F32, float scale = 1.0f,
                             int64_t chunk_words = 0) {
    const HostTables& t = h->host;
    const size_t KB = (h->K + 7) / 8, NB = (t.N + 7) / 8;
    const int64_t chunk = chunk_words > 0 ? std::min(chunk_words, h->reserved) : h->reserved;
    const size_t esz = format == LDPC_B200_LLR_F16 ? 2 : (format == LDPC_B200_LLR_I8 ? 1 : 4);
    if (packed && h->s_pack_bytes < esz * (size_t)chunk * t.N) {   // packed chunks land here and are widened into s_llr
        for (int s = 0; s < kSlots; ++s) {
            if (h->streams[s]) CU_TRY(cudaStreamSynchronize(h->streams[s]));
            cudaFree(h->s_pack[s]); h->s_pack[s] = nullptr;
        }
        h->s_pack_bytes = 0;
        for (int s = 0; s < kSlots; ++s) CU_TRY(cudaMalloc(&h->s_pack[s], esz * (size_t)chunk * t.N));
        h->s_pack_bytes = esz * (size_t)chunk * t.N;
    }
    if (post) {
        for (int s = 0; s < kSlots; ++s)
            if (!h->s_post[s]) CU_TRY(cudaMalloc(&h->s_post[s], sizeof(float) * (size_t)chunk * t.N));
    }
    if (!packed && h->opt.chunk_stage && !h->opt.no_staged && (int64_t)ncw * t.N * 4 >= (h->opt.staged_min_kb << 10)) {
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, llr) == cudaSuccess &&
                            (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        if (!pinned) {  // pageable input of some size: staged by host threads instead of by the driver
            (void)cudaGetLastError();
            return decode_host_chunked_staged(h, llr, ncw, info, hard, iters, post, chunk);
        }
    }
    // phase timers: four events per chunk for the first kTimedChunks chunks (a sample when there are more)
    constexpr int64_t kTimedChunks = 64;
    const int64_t nchunks = (ncw + chunk - 1) / chunk, ntimed = std::min(nchunks, kTimedChunks);
    const bool timed = timing_events(h, (size_t)(4 * ntimed));
    int slot = 0;
    int64_t ci = 0;
    for (int64_t off = 0; off < ncw; off += chunk, slot = (slot + 1) % kSlots, ++ci) {
        const int64_t n = std::min(chunk, ncw - off);
        cudaStream_t st = h->streams[slot];
        const bool tc = timed && ci < ntimed;
        if (tc) CU_TRY(cudaEventRecord(h->tev[4 * ci], st));
        if (packed) {
            CU_TRY(cudaMemcpyAsync(h->s_pack[slot], static_cast<const char*>(packed) + esz * (size_t)off * t.N, esz * (size_t)n * t.N,
                                   cudaMemcpyHostToDevice, st));
            const int wrc = launch_status(k_launch_widen(format, h->s_pack[slot], h->s_llr[slot], (long long)n * t.N, scale, st), "widen");
            if (wrc) return wrc;
            h->launches += 1;
        } else {
            CU_TRY(cudaMemcpyAsync(h->s_llr[slot], llr + (size_t)off * t.N, sizeof(float) * (size_t)n * t.N,
                                   cudaMemcpyHostToDevice, st));
        }
        if (tc) CU_TRY(cudaEventRecord(h->tev[4 * ci + 1], st));
        int rc = launch_decode(h, h->s_llr[slot], n, info ? h->s_info[slot] : nullptr, hard ? h->s_hard[slot] : nullptr,
                               iters ? h->s_iters[slot] : nullptr, post ? h->s_post[slot] : nullptr, st);
        if (rc) return rc;
        if (tc) CU_TRY(cudaEventRecord(h->tev[4 * ci + 2], st));
        if (info) CU_TRY(cudaMemcpyAsync(info + (size_t)off * KB, h->s_info[slot], (size_t)n * KB, cudaMemcpyDeviceToHost, st));
        if (hard) CU_TRY(cudaMemcpyAsync(hard + (size_t)off * NB, h->s_hard[slot], (size_t)n * NB, cudaMemcpyDeviceToHost, st));
        if (iters) CU_TRY(cudaMemcpyAsync(iters + off, h->s_iters[slot], sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (post) CU_TRY(cudaMemcpyAsync(post + (size_t)off * t.N, h->s_post[slot], sizeof(float) * (size_t)n * t.N, cudaMemcpyDeviceToHost, st));
        if (tc) CU_TRY(cudaEventRecord(h->tev[4 * ci + 3], st));
    }
    for (int s = 0; s < kSlots; ++s) CU_TRY(cudaStreamSynchronize(h->streams[s]));
    if (timed) {  // per-stream sums (chunks on different streams overlap), scaled up when only a sample was timed
        const double scale = (double)nchunks / (double)ntimed;
        for (int64_t c = 0; c < ntimed; ++c) {
            h->timing.h2d_s += scale * event_seconds(h->tev[4 * c], h->tev[4 * c + 1]);
            h->timing.kernel_s += scale * event_seconds(h->tev[4 * c + 1], h->tev[4 * c + 2]);
            h->timing.d2h_s += scale * event_seconds(h->tev[4 * c + 2], h->tev[4 * c + 3]);
        }
    }
    return LDPC_B200_OK;
}

}  // namespace

int ldpc_b200_decode_host(ldpc_b200_handle h, const float* llr, int64_t ncw, uint8_t* info, uint8_t* hard,
                          int32_t* iters, float* post) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    if (ncw < 0) return fail(LDPC_B200_ERR_ARG, "ncw must be >= 0");
    if (ncw == 0) return LDPC_B200_OK;
    if (!llr) return fail(LDPC_B200_ERR_ARG, "llr is null");
    std::lock_guard<std::mutex> lk(h->mu);  // one lock for the whole call: plan, reservation and buffers stay put
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
    struct WallTimer {  // accumulates the call into the handle's phase timers on every exit path
        ldpc_b200_decoder* h; int64_t n; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ~WallTimer() {
            h->timing.wall_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            h->timing.calls += 1; h->timing.codewords += n;
        }
    } wall_timer{h, ncw};
    const HostTables& t = h->host;
    // (also after ldpc_b200_reserve / Coder::forDecoder(batchSize): the reference's batch size is only its internal
    // chunking, MyLdpc.cpp:577-616, and kernels of this path do not speed each other up across streams)
    if (h->algorithm == LDPC_B200_ALG_MIN_SUM && h->planned && h->plan.path == LDPC_B200_PATH_QC && !h->opt.no_streamed) {
        // pinned (or managed) input: every chunk copy is queued before the launch.  Copies from pageable memory are
        // staged synchronously by the driver, so queueing them all first would serialise copy and decode.
        cudaPointerAttributes attr;
        bool pinned = cudaPointerGetAttributes(&attr, llr) == cudaSuccess &&
                      (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        if (!pinned) (void)cudaGetLastError();
        if (!pinned && h->opt.register_host) {
            // the caller promised that the buffer outlives the handle (or is decoded from again): page-lock it once, then
            // every call DMAs straight out of it -- no staging copy on the host (a malloc'd postCode decoded repeatedly,
            // as in Test.cpp's loop, then costs what a pinned one does)
            const size_t bytes = sizeof(float) * (size_t)ncw * t.N;
            if (cudaHostRegister(const_cast<float*>(llr), bytes, cudaHostRegisterDefault) == cudaSuccess) {
                h->registered.emplace_back(llr, bytes);
                pinned = true;
            } else {
                (void)cudaGetLastError();  // (overlaps an older registration, or the driver refused: staged path below)
            }
        }
        int rc = warm_kernels_locked(h);  // (a handle nobody called ldpc_b200_reserve on: the kernels load here, once)
        if (rc) return rc;
        rc = kStreamedRetry;
        if (pinned || h->opt.streamed_pageable) {
            rc = decode_host_streamed(h, llr, ncw, info, hard, iters, post, false);
        } else if ((int64_t)ncw * t.N * 4 >= (h->opt.staged_min_kb << 10) && !h->opt.no_staged) {
            // pageable input of some size: four host threads stage it through pinned buffers under the running kernel.
            // Repeated decodes of 65,536 words from malloc'd memory: 4.8-6.1 ms against 21.6 ms for the chunked
            // pipeline below (the driver stages pageable copies itself at ~7.5 GB/s).
            rc = decode_host_streamed(h, llr, ncw, info, hard, iters, post, true);
        }
        if (rc != kStreamedRetry) return rc;
        // kStreamedRetry: the persistent kernel's bounded wait for input expired (or the mode does not apply);
        // every output is rewritten by the chunked pipeline below
    }
    if (h->reserved == 0) {
        // default chunk.  Global-workspace paths: launches serialise on the workspace, so whole waves of the
        // persistent grid (bounded to ~256 MB of channel values).  On-chip paths: kernels of consecutive chunks run
        // concurrently and refill each other's tail, so small chunks win -- 3/4 of a wave measured best
        // (tools/e2e_chunk_sweep.py: 9472 -> 4.69 ms, 1776 -> 4.42 ms per 65,536 words of Test.cpp's code).
        int64_t wave = (int64_t)h->plan.ctas * h->plan.cw_per_cta;
        int64_t chunk = wave * 4;
        const int64_t cap = std::max<int64_t>(wave, ((int64_t)256 << 20) / ((int64_t)t.N * 4) / wave * wave);
        chunk = std::min(chunk, cap);
        chunk = std::min(chunk, (ncw + wave - 1) / wave * wave);
        const bool workspace = uses_big_kernel(h) ||
                               (h->algorithm == LDPC_B200_ALG_MIN_SUM && (h->plan.path == LDPC_B200_PATH_LANE_GLOBAL || h->plan.path == LDPC_B200_PATH_STREAM));
        if (uses_big_kernel(h)) { wave = (int64_t)h->sm_count * kLanes; chunk = std::min(wave * 4, (ncw + wave - 1) / wave * wave); }
        if (!workspace) {
            const int64_t g = h->plan.cw_per_cta;
            chunk = std::min(chunk, std::max<int64_t>(g, wave * 3 / 4 / g * g));
        }
        int rc = reserve_locked(h, chunk);
        if (rc) return rc;
    }
    const int rc = decode_host_chunked_body(h, llr, ncw, info, hard, iters, post);
    if (rc != LDPC_B200_OK) {  // drain the streams: no copy into the caller's buffers may outlive an error return
        const std::string keep = g_err;
        for (int s = 0; s < kSlots; ++s) if (h->streams[s]) cudaStreamSynchronize(h->streams[s]);
        (void)cudaGetLastError();
        g_err = keep;
    }
    return rc;
}

int ldpc_b200_decode_host_packed(ldpc_b200_handle h, const void* llr, int format, float scale, int64_t ncw, uint8_t* info,
                                 uint8_t* hard, int32_t* iters, float* post) {
    if (format == LDPC_B200_LLR_F32 && scale == 1.0f) return ldpc_b200_decode_host(h, static_cast<const float*>(llr), ncw, info, hard, iters, post);
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    if (format != LDPC_B200_LLR_F16 && format != LDPC_B200_LLR_I8) return fail(LDPC_B200_ERR_ARG, "format must be LDPC_B200_LLR_F16 or LDPC_B200_LLR_I8 (fp32 takes scale 1)");
    if (ncw < 0) return fail(LDPC_B200_ERR_ARG, "ncw must be >= 0");
    if (ncw == 0) return LDPC_B200_OK;
    if (!llr) return fail(LDPC_B200_ERR_ARG, "llr is null");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed");
    struct WallTimer {
        ldpc_b200_decoder* h; int64_t n; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ~WallTimer() {
            h->timing.wall_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            h->timing.calls += 1; h->timing.codewords += n;
        }
    } wall_timer{h, ncw};
    if (!h->planned) {
        const int prc = make_plan(h);
        if (prc) return prc;
    }
    // Quasi-cyclic codes with a per-codeword kernel, pinned input: ONE persistent launch fed by the copy stream, exactly as
    // the fp32 call -- the chunks are copied packed and the kernel widens each value where it loads it.
    if (h->algorithm == LDPC_B200_ALG_MIN_SUM && h->plan.path == LDPC_B200_PATH_QC && !h->opt.no_streamed && !h->qc_ring_smem &&
        (h->plan.dmax == 2 || (h->plan.dmax == 1 && (h->qcw_state == 1 || h->qcm_state == 1)))) {
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, llr) == cudaSuccess &&
                            (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        if (!pinned) (void)cudaGetLastError();
        if (pinned) {
            const int wrc = warm_kernels_locked(h);
            if (wrc) return wrc;
            const int src = decode_host_streamed(h, nullptr, ncw, info, hard, iters, post, false, llr, format, scale);
            if (src != kStreamedRetry) return src;
        }
    }
    // chunks of 3/4 of a wave of the persistent grid on three streams: copy, widen + decode, copy back overlap, and the
    // kernels of consecutive chunks fill each other's tails (the chunk rule of the fp32 pipeline, tools/e2e_chunk_sweep.py)
    const int64_t wave = std::max<int64_t>(1, (int64_t)h->plan.ctas * h->plan.cw_per_cta);
    const int64_t g = std::max<int64_t>(1, h->plan.cw_per_cta);
    const int64_t chunk = std::max<int64_t>(g, wave * 3 / 4 / g * g);   // (4 MB chunks measured slower in both regimes)
    if (h->reserved == 0) {
        const int rrc = reserve_locked(h, chunk);
        if (rrc) return rrc;
    }
    const int rc = decode_host_chunked_body(h, nullptr, ncw, info, hard, iters, post, llr, format, scale, chunk);
    if (rc != LDPC_B200_OK) {
        const std::string keep = g_err;
        for (int s = 0; s < kSlots; ++s) if (h->streams[s]) cudaStreamSynchronize(h->streams[s]);
        (void)cudaGetLastError();
        g_err = keep;
    }
    return rc;
}

int ldpc_b200_synth_llr(float* d_llr, int64_t ncw, int N, float sigma, uint64_t seed, const uint8_t* d_bits, int device,
                        void* stream) {
    if (ncw < 0 || N <= 0) return fail(LDPC_B200_ERR_ARG, "bad size");
    if (ncw == 0) return LDPC_B200_OK;
    if (!d_llr) return fail(LDPC_B200_ERR_ARG, "d_llr is null");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed (no CUDA device?)");
    const long long total = (long long)ncw * N;
    int sms = 0;
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
    ldpc_synth_llr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_llr, total, N, sigma, seed, d_bits, 0);
    CU_TRY(cudaGetLastError());
    return LDPC_B200_OK;
}

void* ldpc_b200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        fail(LDPC_B200_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

int ldpc_b200_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) return fail(LDPC_B200_ERR_CUDA, std::string("cudaFreeHost: ") + cudaGetErrorString(cudaGetLastError()));
    return LDPC_B200_OK;
}

int64_t ldpc_b200_launch_count(ldpc_b200_handle h) { return h ? h->launches : 0; }

int ldpc_b200_get_timing(ldpc_b200_handle h, ldpc_b200_timing* out) {
    if (!h || !out) return fail(LDPC_B200_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    *out = h->timing;
    return LDPC_B200_OK;
}

int ldpc_b200_reset_timing(ldpc_b200_handle h) {
    if (!h) return fail(LDPC_B200_ERR_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->timing = ldpc_b200_timing{};
    return LDPC_B200_OK;
}

int ldpc_b200_probe_smem_bandwidth(int device, double* gbytes_per_s) {
    if (!gbytes_per_s) return fail(LDPC_B200_ERR_ARG, "null argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(LDPC_B200_ERR_CUDA, "cudaSetDevice failed (no CUDA device?)");
    int sms = 0;
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    uint32_t* sink = nullptr;
    CU_TRY(cudaMalloc(&sink, sizeof(uint32_t) * sms));
    const int smem = 131072, loops = 4096;
    cudaEvent_t a = nullptr, b = nullptr;
    int rc = LDPC_B200_OK;
    double best = 0.0;
    do {
        if (cudaFuncSetAttribute(ldpc_smem_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
            cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
            rc = fail(LDPC_B200_ERR_CUDA, "probe setup failed");
            break;
        }
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(a, 0);
            ldpc_smem_probe_kernel<<<sms, 1024, smem, 0>>>(sink, loops);
            cudaEventRecord(b, 0);
            if (cudaEventSynchronize(b) != cudaSuccess) { rc = fail(LDPC_B200_ERR_CUDA, "probe kernel failed"); break; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, a, b);
            const double bytes = (double)sms * 1024.0 * 16.0 * 8.0 * loops;
            if (rep > 0 && ms > 0.f) best = std::max(best, bytes / (ms * 1e-3) / 1e9);
        }
    } while (0);
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    cudaFree(sink);
    *gbytes_per_s = best;
    return rc;
}

int ldpc_b200_wimax_csr(int K, int N, int rate, int32_t* row_ptr, int32_t* col_idx, int* M_out, int* nnz_out) {
    std::vector<int32_t> rp, ci;
    int M = 0;
    std::string msg = wimax_csr(K, N, rate, &rp, &ci, &M);
    if (!msg.empty()) return fail(LDPC_B200_ERR_ARG, msg);
    if (M_out) *M_out = M;
    if (nnz_out) *nnz_out = (int)ci.size();
    if (row_ptr) std::memcpy(row_ptr, rp.data(), sizeof(int32_t) * rp.size());
    if (col_idx) std::memcpy(col_idx, ci.data(), sizeof(int32_t) * ci.size());
    return LDPC_B200_OK;
}

int ldpc_b200_edge_tables(int M, int N, const int32_t* row_ptr, const int32_t* col_idx, int32_t* col_ptr,
                          uint32_t* vn_edge, int* max_row_weight, int* max_col_weight) {
    HostTables t;
    std::string msg = build_tables(M, N, row_ptr, col_idx, &t);
    if (!msg.empty()) return fail(LDPC_B200_ERR_ARG, msg);
    if (col_ptr) std::memcpy(col_ptr, t.col_ptr.data(), sizeof(int32_t) * (N + 1));
    if (vn_edge) std::memcpy(vn_edge, t.vn_edge.data(), sizeof(uint32_t) * t.nnz);
    if (max_row_weight) *max_row_weight = t.max_row_weight;
    if (max_col_weight) *max_col_weight = t.max_col_weight;
    return LDPC_B200_OK;
}

}  // extern "C"
