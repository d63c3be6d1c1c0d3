// Host-side helpers shared by the C-ABI translation units: error reporting, device queries.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include "../../include/vitb200.h"

namespace vb {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int num_sms();         // SM count of the current device (cached per device)
int check_arch();      // VB_OK iff current device is sm_100

#define VB_CUDA_CHECK(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) return ::vb::fail(VB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,      \
                                                 cudaGetErrorString(_e), __FILE__, __LINE__);     \
    } while (0)

#define VB_REQUIRE(cond, ...)                                           \
    do {                                                                \
        if (!(cond)) return ::vb::fail(VB_ERR_ARG, __VA_ARGS__);        \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace vb
