// oracle/shim/cl_exec.h -- TEST INFRASTRUCTURE ONLY.
// An EXECUTING stand-in for the handful of Khronos C++ wrapper classes the reference's `Coder` uses
// (MyLdpc.h:170-236, MyLdpc.cpp:224-552, 786-1059).  Force-included (with -DCL_HPP_, which blanks the
// reference's vendored cl.hpp) when oracle/_ref/libmyldpc_refcl.so is built, so that MyLdpc.cpp compiles
// UNMODIFIED and its OpenCL decode variants -- DecodeMS, DecodeSP, DecodeTDMP, DecodeTDMPCL, DecodeMSCL --
// actually run, on the CPU:
//   cl::Buffer                   host memory (zero-filled, with guard bands: the reference's TDMP host loop
//                                 indexes outside its buffers for codes with mixed row weights)
//   cl::Kernel(program, name)    looks the kernel up in the registry of oracle/shim/cl_kernels.cpp, which
//                                 compiles the reference's decodeCL.c unmodified as C++
//   setArg                       records buffers, by-value scalars and __local sizes
//   enqueueNDRangeKernel         runs the kernel function once per work-item, serially, dimension 0 fastest;
//                                 with an explicit local size each work-group's items run as fibers that
//                                 switch at barrier() (K17/K18, decodeCL.c:307-567)
//   enqueueRead/WriteBuffer      memcpy
// A serial schedule is one legal OpenCL execution; the reference's data races (flags[] OR, isDones latch,
// decodeCL.c:45-49,101-105,177-182) resolve the same way under any order.
#ifndef ORACLE_SHIM_CL_EXEC_H_
#define ORACLE_SHIM_CL_EXEC_H_
#include <CL/cl.h>

#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace clexec {

struct WorkItem {
    size_t gid[3], lid[3], grp[3], gsz[3], lsz[3];
};
extern thread_local WorkItem g_wi;      // the work-item being executed (read by get_global_id & co)
void barrier_yield();                   // barrier(): switch to the next fiber of the work-group

typedef void (*Thunk)(void **args);     // unpacks args[i] (pointer to the i-th argument's value) and calls the kernel
struct KernelInfo {
    const char *name;
    Thunk thunk;
    int nargs;
};
const KernelInfo *find_kernel(const char *name);

struct Arg {
    enum Kind { NONE, BUFFER, VALUE, LOCAL } kind = NONE;
    std::shared_ptr<std::vector<char>> store;  // BUFFER: keeps the allocation alive
    void *ptr = nullptr;                       // BUFFER: start of the user-visible region
    std::vector<char> value;                   // VALUE: the bytes passed to setArg
    size_t local_bytes = 0;                    // LOCAL
};

// Runs one NDRange.  dims: 1..3; local == nullptr means "left to the runtime" (no barriers possible).
void run_ndrange(const KernelInfo *k, const std::vector<Arg> &args, int dims, const size_t *global, const size_t *local);

constexpr size_t kGuardBytes = 1 << 20;  // on both sides of every buffer

}  // namespace clexec

namespace cl {

class Platform {
public:
    static cl_int get(std::vector<Platform> *v) { if (v) v->assign(1, Platform()); return CL_SUCCESS; }
    cl_platform_id operator()() const { return nullptr; }
};

class Device {
public:
    cl_device_id operator()() const { return nullptr; }
};

class Context {
public:
    Context() {}
    Context(cl_device_type, cl_context_properties * = nullptr, void * = nullptr, void * = nullptr, cl_int *err = nullptr) {
        if (err) *err = CL_SUCCESS;
    }
    template <cl_int name> std::vector<Device> getInfo(cl_int * = nullptr) const { return std::vector<Device>(1); }
};

class Event {};

class NDRange {
public:
    NDRange() : dims_(0) { s_[0] = s_[1] = s_[2] = 1; }
    NDRange(size_t a) : dims_(1) { s_[0] = a; s_[1] = s_[2] = 1; }
    NDRange(size_t a, size_t b) : dims_(2) { s_[0] = a; s_[1] = b; s_[2] = 1; }
    NDRange(size_t a, size_t b, size_t c) : dims_(3) { s_[0] = a; s_[1] = b; s_[2] = c; }
    int dimensions() const { return dims_; }
    const size_t *sizes() const { return s_; }
private:
    int dims_;
    size_t s_[3];
};
static const NDRange NullRange;

class Buffer {
public:
    Buffer() {}
    Buffer(const Context &, cl_mem_flags flags, size_t bytes, void *host = nullptr, cl_int *err = nullptr) : bytes_(bytes) {
        store_ = std::make_shared<std::vector<char>>(bytes + 2 * clexec::kGuardBytes, 0);
        if ((flags & CL_MEM_COPY_HOST_PTR) && host) std::memcpy(data(), host, bytes);
        if (err) *err = CL_SUCCESS;
    }
    char *data() const { return store_ ? store_->data() + clexec::kGuardBytes : nullptr; }
    size_t size() const { return bytes_; }
    const std::shared_ptr<std::vector<char>> &store() const { return store_; }
private:
    std::shared_ptr<std::vector<char>> store_;
    size_t bytes_ = 0;
};

class Program {
public:
    typedef std::vector<std::pair<const char *, size_t> > Sources;
    Program() {}
    Program(const Context &, const Sources &, cl_int *err = nullptr) { if (err) *err = CL_SUCCESS; }
    cl_program operator()() const { return nullptr; }
};

class Kernel {
public:
    Kernel() {}
    Kernel(const Program &, const char *name, cl_int *err = nullptr) : info_(clexec::find_kernel(name)) {
        if (info_) args_.resize(info_->nargs);
        if (err) *err = info_ ? CL_SUCCESS : CL_INVALID_KERNEL_NAME;
    }
    cl_int setArg(cl_uint i, const Buffer &b) {
        if (!info_ || (int)i >= info_->nargs) return CL_INVALID_ARG_INDEX;
        clexec::Arg &a = args_[i];
        a.kind = clexec::Arg::BUFFER; a.store = b.store(); a.ptr = b.data();
        return CL_SUCCESS;
    }
    cl_int setArg(cl_uint i, size_t bytes, const void *value) {
        if (!info_ || (int)i >= info_->nargs) return CL_INVALID_ARG_INDEX;
        clexec::Arg &a = args_[i];
        if (value) { a.kind = clexec::Arg::VALUE; a.value.assign((const char *)value, (const char *)value + bytes); }
        else { a.kind = clexec::Arg::LOCAL; a.local_bytes = bytes; }
        return CL_SUCCESS;
    }
    const clexec::KernelInfo *info() const { return info_; }
    const std::vector<clexec::Arg> &args() const { return args_; }
private:
    const clexec::KernelInfo *info_ = nullptr;
    std::vector<clexec::Arg> args_;
};

class CommandQueue {
public:
    CommandQueue() {}
    CommandQueue(const Context &, const Device &, cl_command_queue_properties = 0, cl_int *err = nullptr) {
        if (err) *err = CL_SUCCESS;
    }
    cl_int enqueueWriteBuffer(const Buffer &b, cl_bool, size_t off, size_t bytes, const void *ptr,
                              const std::vector<Event> * = nullptr, Event * = nullptr) const {
        if (!b.data() || off + bytes > b.size()) return CL_INVALID_VALUE;
        std::memcpy(b.data() + off, ptr, bytes);
        return CL_SUCCESS;
    }
    cl_int enqueueReadBuffer(const Buffer &b, cl_bool, size_t off, size_t bytes, void *ptr,
                             const std::vector<Event> * = nullptr, Event * = nullptr) const {
        if (!b.data() || off + bytes > b.size()) return CL_INVALID_VALUE;
        std::memcpy(ptr, b.data() + off, bytes);
        return CL_SUCCESS;
    }
    cl_int enqueueNDRangeKernel(const Kernel &k, const NDRange &, const NDRange &global, const NDRange &local,
                                const std::vector<Event> * = nullptr, Event * = nullptr) const {
        if (!k.info()) return CL_INVALID_KERNEL;
        clexec::run_ndrange(k.info(), k.args(), global.dimensions(), global.sizes(),
                            local.dimensions() ? local.sizes() : nullptr);
        return CL_SUCCESS;
    }
    cl_int finish() const { return CL_SUCCESS; }
};

}  // namespace cl
#endif
