// oracle/shim/cl_kernels.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the reference's OpenCL C device code (/root/reference/decodeCL.c, 18 kernels) UNMODIFIED as C++ and
// runs it on the CPU for oracle/_ref/libmyldpc_refcl.so (see cl_exec.h).  The source is #included where it lies
// (oracle/Makefile passes -I/root/reference); nothing of it is copied into this repository.
//
// What stands in for the OpenCL C language and built-ins:
//   kernel / global / local / constant     address-space qualifiers: dropped (constant -> const)
//   get_global_id, get_local_id, get_group_id   the work-item the executor is running
//   barrier(CLK_LOCAL_MEM_FENCE)           switch to the next fiber of the work-group
//   fmin, fabs                             exact IEEE operations (fminf / fabsf)
//   sign                                   OpenCL 1.2 s6.12.4: +1, -1, +-0 for +-0, 0 for NaN
//   exp                                    an OpenCL built-in with implementation-defined rounding (<= 3 ulp): there is
//                                          no canonical value, so the fixed fp32 routine the oracle and the CUDA
//                                          kernel share (oracle_sp_expf) is used; every other operation of the
//                                          sum-product path is an IEEE +,-,*,/ and is therefore the reference's own
// Arithmetic is plain C++ float on x86-64 SSE (-ffp-contract=off, no fast-math): one IEEE binary32 rounding per
// operation, which is what an OpenCL device without -cl-mad-enable / -cl-fast-relaxed-math computes
// (the reference builds its program with no options, MyLdpc.cpp:242).
#include "cl_exec.h"

#include <ucontext.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

extern "C" float oracle_sp_expf(float x);  // oracle/ldpc_oracle.c

namespace clexec {
thread_local WorkItem g_wi;
void lockstep_point();  // see lockstep_prologue below
}

namespace refcl {
using clexec::g_wi;
inline size_t get_global_id(unsigned d) { return g_wi.gid[d]; }
inline size_t get_local_id(unsigned d) { return g_wi.lid[d]; }
inline size_t get_group_id(unsigned d) { return g_wi.grp[d]; }
inline size_t get_global_size(unsigned d) { return g_wi.gsz[d]; }
inline size_t get_local_size(unsigned d) { return g_wi.lsz[d]; }
inline float exp(float x) { return oracle_sp_expf(x); }
inline float fmin(float a, float b) { return ::fminf(a, b); }
inline float fabs(float a) { return ::fabsf(a); }
inline float sign(float x) { clexec::lockstep_point(); return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? x : 0.0f)); }
enum { CLK_LOCAL_MEM_FENCE = 1, CLK_GLOBAL_MEM_FENCE = 2 };
inline void barrier(int) { clexec::barrier_yield(); }

#define kernel
#define global
#define local
#define constant const
#include "decodeCL.c"  // the reference's device code, where it lies
#undef kernel
#undef global
#undef local
#undef constant
}  // namespace refcl

namespace clexec {
namespace {

template <class F> struct Invoker;
template <class... A>
struct Invoker<void (*)(A...)> {
    static constexpr int N = sizeof...(A);
    template <size_t... I>
    static void call(void (*f)(A...), void **a, std::index_sequence<I...>) {
        f(*reinterpret_cast<typename std::remove_cv<A>::type *>(a[I])...);
    }
};

#define REG(name)                                                                                        \
    {#name,                                                                                              \
     [](void **a) {                                                                                      \
         using Inv = Invoker<decltype(&refcl::name)>;                                                    \
         Inv::call(&refcl::name, a, std::make_index_sequence<Inv::N>());                                 \
     },                                                                                                  \
     Invoker<decltype(&refcl::name)>::N}

const KernelInfo kKernels[] = {
    REG(decodeInit), REG(refreshR), REG(refreshQ), REG(hardDecision), REG(checkResult),
    REG(decodeInitMS), REG(refreshRMS), REG(refreshPostPMS), REG(refreshQMS), REG(toChar),
    REG(decodeInitTDMP), REG(refreshRTDMP), REG(refreshPostPTDMP), REG(hardDecisionTDMP), REG(refreshQTDMP),
    REG(checkDones), REG(decodeOnceTDMP), REG(decodeOnceMS),
};
#undef REG

// ---- fibers: the work-items of one work-group, switched at barrier() ----------------------------------
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    WorkItem wi;
    bool done = false;
    bool passed_lockstep_point = false;
};
struct GroupRun {
    ucontext_t main;
    std::vector<Fiber> fibers;
    size_t current = 0;
    const KernelInfo *k = nullptr;
    void **argv = nullptr;
    bool lockstep = false;  // K17: see lockstep_prologue
};
thread_local GroupRun *g_run = nullptr;

void fiber_entry() {
    GroupRun *r = g_run;
    Fiber &f = r->fibers[r->current];
    g_wi = f.wi;
    r->k->thunk(r->argv);
    f.done = true;
    swapcontext(&f.ctx, &r->main);
}

}  // namespace

void barrier_yield() {
    GroupRun *r = g_run;
    if (!r) return;  // no explicit work-group: a barrier among one work-item
    Fiber &f = r->fibers[r->current];
    swapcontext(&f.ctx, &r->main);
    g_wi = f.wi;
}

// decodeOnceTDMP (K17) loads its codeword into __local lP with one strided loop per work-item (decodeCL.c:333-335)
// and starts the first layer WITHOUT a barrier: every work-item reads lP entries that other work-items load.  On
// the GPU the reference was written on, a work-group of z <= 32 items is one SIMD wavefront running in lockstep, so
// the load is complete before any item proceeds; a serial schedule (also a legal OpenCL execution) reads unloaded
// entries and lets late loads overwrite updated posteriors.  The executor reproduces the lockstep outcome for this
// kernel: (1) before a work-group starts, lP already holds what its items are about to load (their own stores
// rewrite the same values); (2) every work-item yields once at its first sign() call -- after its load loop and
// before its first store to lP (decodeCL.c:351-355) -- so all loads are done before any posterior changes.
void lockstep_point() {
    GroupRun *r = g_run;
    if (!r || !r->lockstep) return;
    Fiber &f = r->fibers[r->current];
    if (f.passed_lockstep_point) return;
    f.passed_lockstep_point = true;
    swapcontext(&f.ctx, &r->main);
    g_wi = f.wi;
}

void lockstep_prologue(const KernelInfo *k, const std::vector<Arg> &args, std::vector<std::vector<char>> &local_mem, size_t group) {
    if (std::strcmp(k->name, "decodeOnceTDMP") != 0) return;
    const char z = args[2].value[0];
    const size_t N = (size_t)24 * (size_t)z;
    if (local_mem[5].size() < N * sizeof(float)) return;
    std::memcpy(local_mem[5].data(), static_cast<const float *>(args[0].ptr) + group * N, N * sizeof(float));
}

const KernelInfo *find_kernel(const char *name) {
    for (const KernelInfo &k : kKernels)
        if (std::strcmp(k.name, name) == 0) return &k;
    return nullptr;
}

void run_ndrange(const KernelInfo *k, const std::vector<Arg> &args, int dims, const size_t *global, const size_t *local) {
    const int n = k->nargs;
    std::vector<void *> ptr_slot(n, nullptr);  // storage for pointer-valued arguments
    std::vector<void *> argv(n, nullptr);
    std::vector<std::vector<char>> local_mem(n);
    for (int i = 0; i < n; ++i) {
        const Arg &a = args[i];
        switch (a.kind) {
            case Arg::BUFFER: ptr_slot[i] = a.ptr; argv[i] = &ptr_slot[i]; break;
            case Arg::VALUE: argv[i] = const_cast<char *>(a.value.data()); break;
            case Arg::LOCAL: local_mem[i].assign(a.local_bytes + 64, 0); ptr_slot[i] = local_mem[i].data(); argv[i] = &ptr_slot[i]; break;
            default: std::fprintf(stderr, "cl_exec: kernel %s launched with argument %d unset\n", k->name, i); std::abort();
        }
    }
    size_t g[3] = {1, 1, 1}, l[3] = {1, 1, 1};
    for (int d = 0; d < dims; ++d) { g[d] = global[d]; if (local) l[d] = local[d]; }
    WorkItem wi;
    for (int d = 0; d < 3; ++d) { wi.gsz[d] = g[d]; wi.lsz[d] = l[d]; }
    if (!local) {  // every work-item is its own group; dimension 0 fastest
        for (size_t i2 = 0; i2 < g[2]; ++i2)
            for (size_t i1 = 0; i1 < g[1]; ++i1)
                for (size_t i0 = 0; i0 < g[0]; ++i0) {
                    wi.gid[0] = i0; wi.gid[1] = i1; wi.gid[2] = i2;
                    wi.lid[0] = wi.lid[1] = wi.lid[2] = 0;
                    wi.grp[0] = i0; wi.grp[1] = i1; wi.grp[2] = i2;
                    g_wi = wi;
                    k->thunk(argv.data());
                }
        return;
    }
    const size_t ng[3] = {g[0] / l[0], g[1] / l[1], g[2] / l[2]};
    const size_t nl = l[0] * l[1] * l[2];
    GroupRun run;
    run.k = k;
    run.lockstep = std::strcmp(k->name, "decodeOnceTDMP") == 0;
    run.argv = argv.data();
    run.fibers.resize(nl);
    for (Fiber &f : run.fibers) f.stack.resize(64 << 10);  // the fused kernels keep < 2 KB of private arrays
    for (size_t b2 = 0; b2 < ng[2]; ++b2)
        for (size_t b1 = 0; b1 < ng[1]; ++b1)
            for (size_t b0 = 0; b0 < ng[0]; ++b0) {
                for (int i = 0; i < n; ++i)  // __local memory is per work-group (contents undefined at start: zeroed here)
                    if (args[i].kind == Arg::LOCAL) std::fill(local_mem[i].begin(), local_mem[i].end(), 0);
                lockstep_prologue(k, args, local_mem, b0);
                size_t idx = 0;
                for (size_t j2 = 0; j2 < l[2]; ++j2)
                    for (size_t j1 = 0; j1 < l[1]; ++j1)
                        for (size_t j0 = 0; j0 < l[0]; ++j0, ++idx) {
                            Fiber &f = run.fibers[idx];
                            f.done = false;
                            f.passed_lockstep_point = false;
                            f.wi = wi;
                            f.wi.lid[0] = j0; f.wi.lid[1] = j1; f.wi.lid[2] = j2;
                            f.wi.grp[0] = b0; f.wi.grp[1] = b1; f.wi.grp[2] = b2;
                            f.wi.gid[0] = b0 * l[0] + j0; f.wi.gid[1] = b1 * l[1] + j1; f.wi.gid[2] = b2 * l[2] + j2;
                            getcontext(&f.ctx);
                            f.ctx.uc_stack.ss_sp = f.stack.data();
                            f.ctx.uc_stack.ss_size = f.stack.size();
                            f.ctx.uc_link = &run.main;
                            makecontext(&f.ctx, fiber_entry, 0);
                        }
                g_run = &run;
                for (size_t left = nl; left > 0;) {  // rounds: every live fiber runs to its next barrier (or to the end)
                    left = 0;
                    for (size_t i = 0; i < nl; ++i) {
                        if (run.fibers[i].done) continue;
                        run.current = i;
                        swapcontext(&run.main, &run.fibers[i].ctx);
                        if (!run.fibers[i].done) ++left;
                    }
                }
                g_run = nullptr;
            }
}

}  // namespace clexec
