// LayerNorm forward / backward for the fp32 residual stream (HBM-bound; one warp per row, 128-bit accesses,
// warp-shuffle reductions, fp32 statistics).
//
// Replaces nn.LayerNorm(eps=1e-6) at vanilla_vit.py:66,70,100 (ATen native_layer_norm / _backward) and the
// eps=1e-5 norms of the DETR encoder layer (transformer.py:201-202).  The backward kernel also folds in the
// residual-gradient add (x = x + f(LN(x)) => dx = dres + LN'(dy)) and the per-column sum of its output, which
// is the bias gradient of the linear layer that produced this LayerNorm's input stream.
#include "common.h"
#include "pdl.cuh"
#include <cuda_bf16.h>

namespace vb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&lo);
    r.y = *reinterpret_cast<uint32_t*>(&hi);
    return r;
}

// D is a multiple of 64 (head_dim 64 x any head count: 192 = the reference's deit_tiny, utils/args.py:43-45).  A lane owns
// NV = ceil(D / 128) float4 vectors; when D is an odd multiple of 64 the last vector exists for lanes < 16 only (ok(i)).
template <int D>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ y_bf16,
                                                     long long ldyb, float* __restrict__ y_f32, long long ldyf,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
                                                     float eps, const float* __restrict__ add, long long ldadd,
                                                     __nv_bfloat16* __restrict__ y2_bf16, long long ldy2) {
    constexpr int NV = (D + 127) / 128;
    constexpr bool kTail = (D % 128) != 0;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    pdl_launch_dependents();
    pdl_wait();
    if (row >= rows) return;
    auto ok = [&](int i) { return !kTail || i < NV - 1 || lane < 16; };
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * ldx);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = ok(i) ? __ldg(xr + lane + i * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        if (ok(i)) q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (!ok(i)) continue;
        const float4 g = __ldg(g4 + lane + i * 32), b = __ldg(b4 + lane + i * 32);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (y_bf16) *reinterpret_cast<uint2*>(y_bf16 + (long long)row * ldyb + (lane + i * 32) * 4) = pack4_bf16(o.x, o.y, o.z, o.w);
        if (y_f32) *reinterpret_cast<float4*>(y_f32 + (long long)row * ldyf + (lane + i * 32) * 4) = o;
        if (y2_bf16) {  // second bf16 output y + add (DETR: q = k = src + pos, transformer.py:218)
            const float4 a = __ldg(reinterpret_cast<const float4*>(add + (long long)row * ldadd) + lane + i * 32);
            *reinterpret_cast<uint2*>(y2_bf16 + (long long)row * ldy2 + (lane + i * 32) * 4) = pack4_bf16(o.x + a.x, o.y + a.y, o.z + a.z, o.w + a.w);
        }
    }
}

// Backward. Persistent over rows: warp w of block b walks rows (b * wpb + w), += gridDim.x * wpb, ...; each lane
// keeps per-column partial sums for dgamma / dbeta / colsum(dx) in a warp-private smem slab (registers stay free
// for two resident blocks per SM = more loads in flight), reduced across warps at the end and across blocks with
// one atomicAdd per column per block.
template <int D, bool DY_BF16>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(const void* __restrict__ dy_, long long lddy, const float* __restrict__ x,
                                                     long long ldx, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ dres, long long lddres,
                                                     float* __restrict__ dx, long long lddx, __nv_bfloat16* __restrict__ dx_bf16,
                                                     long long lddxb, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     float* __restrict__ dx_colsum, int rows,
                                                     const __nv_bfloat16* __restrict__ dy_add, long long lddya) {
    constexpr int NV = (D + 127) / 128;
    constexpr bool kTail = (D % 128) != 0;
    extern __shared__ float red[];  // [warps][3][D]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    auto ok = [&](int i) { return !kTail || i < NV - 1 || lane < 16; };
    float* mine = red + warp * 3 * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (!ok(i)) continue;
        const int c = (lane + i * 32) * 4;
        *reinterpret_cast<float4*>(mine + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(mine + D + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(mine + 2 * D + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    pdl_launch_dependents();
    pdl_wait();

    for (int row = blockIdx.x * wpb + warp; row < rows; row += gridDim.x * wpb) {
        const float mu = mean[row], rs = rstd[row];
        float4 xh[NV], dyv[NV], rv[NV];
        if (dres) {
#pragma unroll
            for (int i = 0; i < NV; ++i)
                rv[i] = ok(i) ? __ldg(reinterpret_cast<const float4*>(dres + (long long)row * lddres) + lane + i * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (!ok(i)) {   // columns beyond D: contribute nothing to the row statistics
                xh[i] = dyv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)row * ldx) + lane + i * 32);
            xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
            if (DY_BF16) {
                const uint2 w = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + (long long)row * lddy) + lane + i * 32);
                dyv[i] = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                                     __uint_as_float(w.y & 0xFFFF0000u));
            } else {
                dyv[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + (long long)row * lddy) + lane + i * 32);
            }
            if (dy_add) {  // post-norm blocks: the gradient reaching a LayerNorm output is a sum of two streams
                const uint2 w = __ldg(reinterpret_cast<const uint2*>(dy_add + (long long)row * lddya) + lane + i * 32);
                dyv[i].x += __uint_as_float(w.x << 16); dyv[i].y += __uint_as_float(w.x & 0xFFFF0000u);
                dyv[i].z += __uint_as_float(w.y << 16); dyv[i].w += __uint_as_float(w.y & 0xFFFF0000u);
            }
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (!ok(i)) continue;
            const float4 gi = __ldg(g4 + lane + i * 32);
            const float a = dyv[i].x * gi.x, b = dyv[i].y * gi.y, cc = dyv[i].z * gi.z, d = dyv[i].w * gi.w;
            s1 += (a + b) + (cc + d);
            s2 += (a * xh[i].x + b * xh[i].y) + (cc * xh[i].z + d * xh[i].w);
            const int c = (lane + i * 32) * 4;
            float4 ag = *reinterpret_cast<float4*>(mine + c), ab = *reinterpret_cast<float4*>(mine + D + c);
            ag.x += dyv[i].x * xh[i].x; ag.y += dyv[i].y * xh[i].y; ag.z += dyv[i].z * xh[i].z; ag.w += dyv[i].w * xh[i].w;
            ab.x += dyv[i].x; ab.y += dyv[i].y; ab.z += dyv[i].z; ab.w += dyv[i].w;
            *reinterpret_cast<float4*>(mine + c) = ag;
            *reinterpret_cast<float4*>(mine + D + c) = ab;
        }
        const float m1 = warp_sum(s1) * (1.0f / D), m2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (!ok(i)) continue;
            const float4 gi = __ldg(g4 + lane + i * 32);
            float4 o;
            o.x = rs * (dyv[i].x * gi.x - m1 - xh[i].x * m2) + rv[i].x;
            o.y = rs * (dyv[i].y * gi.y - m1 - xh[i].y * m2) + rv[i].y;
            o.z = rs * (dyv[i].z * gi.z - m1 - xh[i].z * m2) + rv[i].z;
            o.w = rs * (dyv[i].w * gi.w - m1 - xh[i].w * m2) + rv[i].w;
            float4 ac = *reinterpret_cast<float4*>(mine + 2 * D + (lane + i * 32) * 4);
            ac.x += o.x; ac.y += o.y; ac.z += o.z; ac.w += o.w;
            *reinterpret_cast<float4*>(mine + 2 * D + (lane + i * 32) * 4) = ac;
            if (dx) *reinterpret_cast<float4*>(dx + (long long)row * lddx + (lane + i * 32) * 4) = o;
            if (dx_bf16) *reinterpret_cast<uint2*>(dx_bf16 + (long long)row * lddxb + (lane + i * 32) * 4) = pack4_bf16(o.x, o.y, o.z, o.w);
        }
    }
    // cross-warp reduction of the three column sums
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * D; idx += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < wpb; ++w) s += red[w * 3 * D + idx];
        const int which = idx / D, c = idx - which * D;
        if (which == 0) { if (dgamma) atomicAdd(dgamma + c, s); }
        else if (which == 1) { if (dbeta) atomicAdd(dbeta + c, s); }
        else { if (dx_colsum) atomicAdd(dx_colsum + c, s); }
    }
}

template <int D>
static int ln_fwd_launch(const float* x, long long ldx, const float* gamma, const float* beta, void* y_bf16, long long ldyb,
                         float* y_f32, long long ldyf, float* mean, float* rstd, int rows, float eps, const float* add,
                         long long ldadd, void* y2, long long ldy2, cudaStream_t st) {
    const int wpb = 8;
    VB_CUDA_CHECK(launch_pdl(ln_fwd_kernel<D>, dim3((rows + wpb - 1) / wpb), dim3(wpb * 32), 0, st, x, ldx, gamma, beta,
                             reinterpret_cast<__nv_bfloat16*>(y_bf16), ldyb, y_f32, ldyf, mean, rstd, rows, eps, add, ldadd,
                             reinterpret_cast<__nv_bfloat16*>(y2), ldy2));
    return VB_OK;
}

template <int D, bool DYB>
static int ln_bwd_launch(const void* dy, long long lddy, const float* x, long long ldx, const float* mean, const float* rstd,
                         const float* gamma, const float* dres, long long lddres, float* dx, long long lddx, void* dxb,
                         long long lddxb, float* dgamma, float* dbeta, float* colsum, int rows, const void* dy_add, long long lddya,
                         cudaStream_t st) {
    const int wpb = 8;
    const size_t smem = size_t(wpb) * 3 * D * sizeof(float);
    auto kern = ln_bwd_kernel<D, DYB>;
    static DeviceOnce configured;   // the attribute is per device: every GPU a process touches opts in once
    VB_ONCE_PER_DEVICE(configured, VB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    int grid = num_sms() * 2;
    const int need = (rows + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    VB_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(wpb * 32), smem, st, dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx,
                             reinterpret_cast<__nv_bfloat16*>(dxb), lddxb, dgamma, dbeta, colsum, rows,
                             reinterpret_cast<const __nv_bfloat16*>(dy_add), lddya));
    return VB_OK;
}

}  // namespace vb

extern "C" int vb_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y_bf16, int64_t ldy_bf16,
                                float* y_f32, int64_t ldy_f32, float* mean, float* rstd, int32_t rows, int32_t dim, float eps,
                                const float* add, int64_t ldadd, void* y2_bf16, int64_t ldy2, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(x && gamma && beta && (y_bf16 || y_f32), "layernorm_fwd: null pointer");
    VB_REQUIRE((y2_bf16 == nullptr) || (add != nullptr && ldadd % 4 == 0 && ldy2 % 4 == 0), "layernorm_fwd: y2 needs `add` and 4-aligned pitches");
    VB_REQUIRE(rows >= 0 && dim > 0 && dim % 64 == 0 && dim <= 1024, "layernorm_fwd: dim %d must be a multiple of 64 and <= 1024", dim);
    VB_REQUIRE(ldx % 4 == 0 && ldy_bf16 % 4 == 0 && ldy_f32 % 4 == 0, "layernorm_fwd: row pitches must be multiples of 4 elements");
    if (rows == 0) return VB_OK;
    cudaStream_t st = as_stream(stream);
    switch (dim / 64) {
#define VB_CASE(N64) case N64: return ln_fwd_launch<N64 * 64>(x, ldx, gamma, beta, y_bf16, ldy_bf16, y_f32, ldy_f32, mean, rstd, rows, eps, add, ldadd, y2_bf16, ldy2, st);
        VB_CASE(1) VB_CASE(2) VB_CASE(3) VB_CASE(4) VB_CASE(5) VB_CASE(6) VB_CASE(7) VB_CASE(8)
        VB_CASE(9) VB_CASE(10) VB_CASE(11) VB_CASE(12) VB_CASE(13) VB_CASE(14) VB_CASE(15) VB_CASE(16)
#undef VB_CASE
    }
    return fail(VB_ERR_UNSUPPORTED, "layernorm_fwd: dim %d", dim);
}

extern "C" int vb_layernorm_bwd(const void* dy, int32_t dy_dtype, int64_t lddy, const float* x, int64_t ldx, const float* mean,
                                const float* rstd, const float* gamma, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                                void* dx_bf16, int64_t lddx_bf16, float* dgamma, float* dbeta, float* dx_colsum, int32_t rows,
                                int32_t dim, const void* dy_add_bf16, int64_t lddy_add, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(dy && x && mean && rstd && gamma && (dx || dx_bf16), "layernorm_bwd: null pointer");
    VB_REQUIRE(rows >= 0 && dim > 0 && dim % 64 == 0 && dim <= 1024, "layernorm_bwd: dim %d must be a multiple of 64 and <= 1024", dim);
    VB_REQUIRE(dy_dtype == VB_BF16 || dy_dtype == VB_F32, "layernorm_bwd: bad dy dtype");
    VB_REQUIRE(lddy % 4 == 0 && ldx % 4 == 0 && lddres % 4 == 0 && lddx % 4 == 0 && lddx_bf16 % 4 == 0 && lddy_add % 4 == 0,
               "layernorm_bwd: pitches must be multiples of 4");
    if (rows == 0) return VB_OK;
    cudaStream_t st = as_stream(stream);
    switch (dim / 64) {
#define VB_CASE(N64)                                                                                                             \
    case N64:                                                                                                                    \
        if (dy_dtype == VB_BF16)                                                                                                 \
            return ln_bwd_launch<N64 * 64, true>(dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dx_bf16, lddx_bf16, dgamma, \
                                           dbeta, dx_colsum, rows, dy_add_bf16, lddy_add, st);                                   \
        return ln_bwd_launch<N64 * 64, false>(dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dx_bf16, lddx_bf16, dgamma,    \
                                        dbeta, dx_colsum, rows, dy_add_bf16, lddy_add, st);
        VB_CASE(1) VB_CASE(2) VB_CASE(3) VB_CASE(4) VB_CASE(5) VB_CASE(6) VB_CASE(7) VB_CASE(8)
        VB_CASE(9) VB_CASE(10) VB_CASE(11) VB_CASE(12) VB_CASE(13) VB_CASE(14) VB_CASE(15) VB_CASE(16)
#undef VB_CASE
    }
    return fail(VB_ERR_UNSUPPORTED, "layernorm_bwd: dim %d", dim);
}
