// Library-level entry points of libvitb200.so: version, error text, device checks.
#include "common.h"
#include "pdl.cuh"
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace vb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::mutex g_dev_mutex;
static int g_sm_count[64];
static int g_cc_major[64];
static bool g_dev_known[64];

static int query_device(int dev) {
    if (dev < 0 || dev >= 64) return -1;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (!g_dev_known[dev]) {
        int sms = 0, major = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
        g_sm_count[dev] = sms;
        g_cc_major[dev] = major;
        g_dev_known[dev] = true;
    }
    return 0;
}

int pdl_level() {   // VITB200_PDL: 0 = off (default), 1 = every PDL-aware kernel, 2 = GEMM only, 3 = everything but the GEMM
    static int lvl = -1;
    if (lvl < 0) {
        const char* e = getenv("VITB200_PDL");
        lvl = e ? atoi(e) : 0;
    }
    return lvl;
}
bool pdl_enabled() { return pdl_level() == 1 || pdl_level() == 3; }
bool pdl_enabled_gemm() { return pdl_level() == 1 || pdl_level() == 2; }

int num_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || query_device(dev) != 0) return 0;
    return g_sm_count[dev];
}

int check_arch() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(VB_ERR_CUDA, "cudaGetDevice failed (no CUDA device?)");
    if (query_device(dev) != 0) return fail(VB_ERR_CUDA, "cannot query device %d", dev);
    if (g_cc_major[dev] != 10)
        return fail(VB_ERR_ARCH, "device %d has compute capability major %d; libvitb200 is sm_100a-only", dev,
                    g_cc_major[dev]);
    return VB_OK;
}

}  // namespace vb

extern "C" {

int vb_version(void) { return 1; }

const char* vb_last_error(void) { return vb::g_err; }

int vb_device_check(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return vb::fail(VB_ERR_CUDA, "no CUDA device visible");
    if (device < 0 || device >= count) return vb::fail(VB_ERR_ARG, "device %d out of range (%d visible)", device, count);
    if (vb::query_device(device) != 0) return vb::fail(VB_ERR_CUDA, "cannot query device %d", device);
    if (vb::g_cc_major[device] != 10)
        return vb::fail(VB_ERR_ARCH, "device %d is not sm_100 (major %d)", device, vb::g_cc_major[device]);
    return VB_OK;
}

int vb_sm_count(int device) {
    if (vb::query_device(device) != 0) return vb::fail(VB_ERR_CUDA, "cannot query device %d", device);
    return vb::g_sm_count[device];
}

}  // extern "C"
