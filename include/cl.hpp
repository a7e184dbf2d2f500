// cl.hpp -- intentionally empty.
// The reference's sources include "cl.hpp" (the Khronos OpenCL C++ wrapper) next to "MyLdpc.h"
// (reference Test.cpp:10, MyLdpc.h:21).  The B200 build has no OpenCL anywhere: the decode path is
// CUDA behind include/ldpc_b200.h.  This stub only keeps `#include "cl.hpp"` lines compiling.
#ifndef MYLDPC_B200_CL_HPP_STUB_
#define MYLDPC_B200_CL_HPP_STUB_
#endif
