// Host-side helpers shared by the C-ABI translation units: error reporting, device queries.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
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

// Per-device "already configured" flag for cudaFuncSetAttribute(MaxDynamicSharedMemorySize): the attribute belongs to the
// device's context, so a process that touches a second GPU must opt in there too.  The flag is set AFTER the attribute call
// (a concurrent thread at worst repeats it), so no launch can slip in ahead of the opt-in.
struct DeviceOnce {
    std::atomic<unsigned char> done[64];
    int device() const {
        int dev = 0;
        return (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) ? dev : -1;
    }
    bool is_set() const {
        const int dev = device();
        return dev >= 0 && done[dev].load(std::memory_order_acquire) != 0;
    }
    void set() {
        const int dev = device();
        if (dev >= 0) done[dev].store(1, std::memory_order_release);
    }
};
#define VB_ONCE_PER_DEVICE(flag, ...)   \
    do {                                \
        if (!(flag).is_set()) {         \
            __VA_ARGS__;                \
            (flag).set();               \
        }                               \
    } while (0)

}  // namespace vb
