// oracle/shim/cl_stub.h -- TEST INFRASTRUCTURE ONLY.
// Force-included (with -DCL_HPP_, which makes the reference's vendored cl.hpp an empty file) when
// oracle/_ref is built: inert stand-ins for the handful of Khronos C++ wrapper classes that are
// MEMBERS of the reference's `Coder`, so that MyLdpc.cpp compiles unmodified.  Nothing executes on
// a device; only Coder's host paths (initCheckMatrix, forDecoder's tables, decodeCPU, forEncoder,
// encode, test) are meaningful in oracle/_ref.  The OpenCL decode variants become no-ops.
#ifndef ORACLE_SHIM_CL_STUB_H_
#define ORACLE_SHIM_CL_STUB_H_
#include <CL/cl.h>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

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
    NDRange() {}
    NDRange(size_t) {}
    NDRange(size_t, size_t) {}
    NDRange(size_t, size_t, size_t) {}
};
static const NDRange NullRange;

class Buffer {
public:
    Buffer() {}
    Buffer(const Context &, cl_mem_flags, size_t, void * = nullptr, cl_int *err = nullptr) { if (err) *err = CL_SUCCESS; }
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
    Kernel(const Program &, const char *, cl_int *err = nullptr) { if (err) *err = CL_SUCCESS; }
    template <class T> cl_int setArg(cl_uint, const T &) { return CL_SUCCESS; }
    cl_int setArg(cl_uint, size_t, const void *) { return CL_SUCCESS; }
};

class CommandQueue {
public:
    CommandQueue() {}
    CommandQueue(const Context &, const Device &, cl_command_queue_properties = 0, cl_int *err = nullptr) {
        if (err) *err = CL_SUCCESS;
    }
    cl_int enqueueWriteBuffer(const Buffer &, cl_bool, size_t, size_t, const void *, const std::vector<Event> * = nullptr,
                              Event * = nullptr) const { return CL_SUCCESS; }
    cl_int enqueueReadBuffer(const Buffer &, cl_bool, size_t, size_t bytes, void *ptr, const std::vector<Event> * = nullptr,
                             Event * = nullptr) const {
        if (ptr) std::memset(ptr, 0, bytes);  // "all flags clear": the no-op device loops stop at once
        return CL_SUCCESS;
    }
    cl_int enqueueNDRangeKernel(const Kernel &, const NDRange &, const NDRange &, const NDRange &,
                                const std::vector<Event> * = nullptr, Event * = nullptr) const { return CL_SUCCESS; }
    cl_int finish() const { return CL_SUCCESS; }
};

}  // namespace cl
#endif
