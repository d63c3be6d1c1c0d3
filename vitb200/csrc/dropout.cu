// Element-wise dropout kernels (HBM-bound) for the hidden-state dropout sites of the ViT encoder
// (vanilla_vit.py:38 mlp.2, :42 mlp.4, :68/:78 EncoderBlock.dropout, :94/:104 Encoder.dropout).  Masks: dropout.cuh.
#include "common.h"
#include "dropout.cuh"
#include <cuda_bf16.h>

namespace vb {

// dst = keep ? src / (1 - p) : 0  (+ aux); written as fp32 and/or bf16.  4 elements per thread.
__global__ void __launch_bounds__(256) dropout_f32_kernel(const float* __restrict__ src, long long ldsrc, const float* __restrict__ aux,
                                                          long long ldaux, float* dst, long long lddst, __nv_bfloat16* dstb, long long lddstb,
                                                          int rows, int cols, uint32_t thresh, float inv_keep, const uint32_t* seed,
                                                          uint32_t stream_id) {
    const uint32_t key = dropout_key(*seed, stream_id);
    const int cols4 = cols >> 2;
    const long long total = (long long)rows * cols4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(t / cols4), c = (int)(t - (long long)r * cols4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * ldsrc + c);
        float o[4] = {v.x, v.y, v.z, v.w};
        const uint32_t idx = (uint32_t)r * (uint32_t)cols + (uint32_t)c;
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = dropout_keep(key, idx + i, thresh) ? o[i] * inv_keep : 0.f;
        if (aux != nullptr) {
            const float4 a = *reinterpret_cast<const float4*>(aux + r * ldaux + c);
            o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
        }
        if (dst != nullptr) *reinterpret_cast<float4*>(dst + r * lddst + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (dstb != nullptr) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
            uint2 w;
            w.x = *reinterpret_cast<uint32_t*>(&lo);
            w.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(dstb + r * lddstb + c) = w;
        }
    }
}

// In place on one or two bf16 matrices sharing the mask (the GELU output and the saved gelu' multiplier): 8 elements per thread.
__global__ void __launch_bounds__(256) dropout_bf16_pair_kernel(__nv_bfloat16* x1, __nv_bfloat16* x2, long long ld, int rows, int cols,
                                                                uint32_t thresh, float inv_keep, const uint32_t* seed, uint32_t stream_id) {
    const uint32_t key = dropout_key(*seed, stream_id);
    const int cols8 = cols >> 3;
    const long long total = (long long)rows * cols8;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(t / cols8), c = (int)(t - (long long)r * cols8) * 8;
        const uint32_t idx = (uint32_t)r * (uint32_t)cols + (uint32_t)c;
        bool keep[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) keep[i] = dropout_keep(key, idx + i, thresh);
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            __nv_bfloat16* x = which == 0 ? x1 : x2;
            if (x == nullptr) continue;
            uint4 v = *reinterpret_cast<uint4*>(x + r * ld + c);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float lo = keep[2 * i] ? __uint_as_float(w[i] << 16) * inv_keep : 0.f;
                const float hi = keep[2 * i + 1] ? __uint_as_float(w[i] & 0xFFFF0000u) * inv_keep : 0.f;
                __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
                w[i] = *reinterpret_cast<uint32_t*>(&pk);
            }
            *reinterpret_cast<uint4*>(x + r * ld + c) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

__global__ void dropout_mask_u8_kernel(uint8_t* out, long long n, uint32_t thresh, const uint32_t* seed, uint32_t stream_id) {
    const uint32_t key = dropout_key(*seed, stream_id);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = dropout_keep(key, (uint32_t)i, thresh) ? 1 : 0;
}

static unsigned ew_grid(long long work_items) {
    long long g = (work_items + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace vb

extern "C" int vb_dropout_f32(const float* src, int64_t ldsrc, const float* aux, int64_t ldaux, float* dst_f32, int64_t lddst, void* dst_bf16,
                              int64_t lddstb, int32_t rows, int32_t cols, float p, const uint32_t* seed_dev, uint32_t stream_id, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(src && seed_dev && (dst_f32 || dst_bf16), "dropout_f32: null pointer");
    VB_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0 && ldsrc % 4 == 0 && (!aux || ldaux % 4 == 0) && (!dst_f32 || lddst % 4 == 0) &&
                   (!dst_bf16 || lddstb % 4 == 0), "dropout_f32: cols and pitches must be multiples of 4");
    VB_REQUIRE(p >= 0.f && p < 1.f, "dropout_f32: p must be in [0, 1)");
    VB_REQUIRE((long long)rows * cols < (1ll << 32), "dropout_f32: more than 2^32 elements");
    dropout_f32_kernel<<<ew_grid((long long)rows * (cols / 4)), 256, 0, as_stream(stream)>>>(
        src, ldsrc, aux, ldaux, dst_f32, lddst, (__nv_bfloat16*)dst_bf16, lddstb, rows, cols, dropout_threshold(p), 1.0f / (1.0f - p), seed_dev,
        stream_id);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_dropout_bf16_pair(void* x1, void* x2, int64_t ld, int32_t rows, int32_t cols, float p, const uint32_t* seed_dev,
                                    uint32_t stream_id, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(x1 && seed_dev, "dropout_bf16_pair: null pointer");
    VB_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0, "dropout_bf16_pair: cols and pitch must be multiples of 8");
    VB_REQUIRE(p >= 0.f && p < 1.f, "dropout_bf16_pair: p must be in [0, 1)");
    VB_REQUIRE((long long)rows * cols < (1ll << 32), "dropout_bf16_pair: more than 2^32 elements");
    dropout_bf16_pair_kernel<<<ew_grid((long long)rows * (cols / 8)), 256, 0, as_stream(stream)>>>(
        (__nv_bfloat16*)x1, (__nv_bfloat16*)x2, ld, rows, cols, dropout_threshold(p), 1.0f / (1.0f - p), seed_dev, stream_id);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_dropout_mask_u8(uint8_t* out, int64_t n, float p, const uint32_t* seed_dev, uint32_t stream_id, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(out && seed_dev && n > 0 && n < (1ll << 32), "dropout_mask_u8: bad arguments");
    dropout_mask_u8_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(out, n, dropout_threshold(p), seed_dev, stream_id);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}
