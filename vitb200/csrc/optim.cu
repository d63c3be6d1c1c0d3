// Fused training-step tail: softmax cross-entropy forward+backward on the logits, and a flat multi-tensor Adam
// that also refreshes the bf16 parameter shadow.  (reference: nn.CrossEntropyLoss + optim.Adam(lr=1e-4)
// at vanilla_vit.py:220-221,237-239 / base.py:53-57.)
#include "common.h"
#include <cuda_bf16.h>

namespace vb {

// one warp per sample: loss_sum += w * (logsumexp(z) - z[y]); dz = scale * w * (softmax(z) - onehot(y)) as bf16 (+ fp32 optional)
__global__ void ce_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels, int B, int C,
                          float* __restrict__ loss_accum, float weight, __nv_bfloat16* __restrict__ dz_bf16, long long lddz,
                          float* __restrict__ dz_f32, long long lddzf, float grad_scale, int* __restrict__ correct_accum) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    const float* z = logits + (long long)row * ld;
    float mx = -INFINITY;
    int arg = 0;
    for (int c = lane; c < C; c += 32) {
        const float v = z[c];
        if (v > mx) { mx = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __expf(z[c] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    int y = (int)labels[row];
    const bool bad_label = y < 0 || y >= C;      // PyTorch raises a device assert here; an out-of-range label must not read out of bounds
    if (bad_label) y = 0;
    const float lse = mx + __logf(s);
    if (lane == 0) {
        atomicAdd(loss_accum, bad_label ? __int_as_float(0x7fc00000) : weight * (lse - z[y]));   // NaN loss: loud, not silent
        if (correct_accum && arg == y) atomicAdd(correct_accum, 1);
    }
    const float inv = 1.f / s;
    const float poison = bad_label ? __int_as_float(0x7fc00000) : 0.f;   // a bad label must not train the sample towards class 0: NaN gradient too
    for (int c = lane; c < C; c += 32) {
        const float g = grad_scale * weight * (__expf(z[c] - mx) * inv - (c == y ? 1.f : 0.f)) + poison;
        if (dz_bf16) dz_bf16[(long long)row * lddz + c] = __float2bfloat16_rn(g);
        if (dz_f32) dz_f32[(long long)row * lddzf + c] = g;
    }
}

// Knowledge-distillation loss of utils/distillation_loss.py:30-75, forward + both logits gradients in one pass, one warp per sample:
//   loss = (1-alpha) * CE(z, y) + alpha * D,   D = CE(z_kd, argmax teacher)                       (kind 2, 'hard', :70-71)
//                                              D = T^2/(B*C) * sum q (log q - log p), p = softmax(z_kd/T), q = softmax(t/T)  (kind 1, 'soft', :55-65)
// dz = (1-alpha)/B (softmax(z) - onehot(y));  dz_kd = alpha/B (softmax(z_kd) - onehot(argmax t))  or  alpha*T/(B*C) (p - q).
struct WarpStat { float mx; int arg; float sum; };
__device__ __forceinline__ WarpStat warp_softmax_stat(const float* __restrict__ z, int C, float inv_t, int lane) {
    WarpStat r{-INFINITY, 0, 0.f};
    for (int c = lane; c < C; c += 32) {
        const float v = z[c] * inv_t;
        if (v > r.mx) { r.mx = v; r.arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, r.mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, r.arg, o);
        if (om > r.mx || (om == r.mx && oa < r.arg)) { r.mx = om; r.arg = oa; }
    }
    for (int c = lane; c < C; c += 32) r.sum += __expf(z[c] * inv_t - r.mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r.sum += __shfl_xor_sync(0xffffffffu, r.sum, o);
    return r;
}

__global__ void distill_loss_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ logits_kd, long long ldkd,
                                    const float* __restrict__ teacher, long long ldt, const long long* __restrict__ labels, int B, int C,
                                    int kind, float alpha, float tau, float* __restrict__ loss_accum, __nv_bfloat16* __restrict__ dz,
                                    long long lddz, __nv_bfloat16* __restrict__ dzkd, long long lddzkd, float* __restrict__ dz_f32,
                                    long long lddzf, float* __restrict__ dzkd_f32, long long lddzkdf, float grad_scale,
                                    int* __restrict__ correct_accum) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    const float invB = 1.f / (float)B;
    const float* z = logits + (long long)row * ld;
    int y = (int)labels[row];
    const bool bad_label = y < 0 || y >= C;      // as in ce_kernel: no out-of-bounds read, NaN loss
    if (bad_label) y = 0;
    // base criterion on the class-token logits
    const WarpStat a = warp_softmax_stat(z, C, 1.f, lane);
    float loss = bad_label ? __int_as_float(0x7fc00000) : (1.f - alpha) * invB * (a.mx + __logf(a.sum) - z[y]);
    {
        const float w = grad_scale * (1.f - alpha) * invB, inv = 1.f / a.sum;
        const float poison = bad_label ? __int_as_float(0x7fc00000) : 0.f;   // as in ce_kernel: the gradient of a bad-label sample is NaN, not "class 0"
        for (int c = lane; c < C; c += 32) {
            const float g = w * (__expf(z[c] - a.mx) * inv - (c == y ? 1.f : 0.f)) + poison;
            if (dz) dz[(long long)row * lddz + c] = __float2bfloat16_rn(g);
            if (dz_f32) dz_f32[(long long)row * lddzf + c] = g;
        }
    }
    const float* zk = logits_kd + (long long)row * ldkd;
    const float* t = teacher + (long long)row * ldt;
    if (kind == 2) {
        const WarpStat ts = warp_softmax_stat(t, C, 1.f, lane);   // only the arg-max is used (first maximal index, as torch.argmax)
        const WarpStat k = warp_softmax_stat(zk, C, 1.f, lane);
        const int yt = ts.arg;
        loss += alpha * invB * (k.mx + __logf(k.sum) - zk[yt]);
        const float w = grad_scale * alpha * invB, inv = 1.f / k.sum;
        for (int c = lane; c < C; c += 32) {
            const float g = w * (__expf(zk[c] - k.mx) * inv - (c == yt ? 1.f : 0.f));
            if (dzkd) dzkd[(long long)row * lddzkd + c] = __float2bfloat16_rn(g);
            if (dzkd_f32) dzkd_f32[(long long)row * lddzkdf + c] = g;
        }
    } else {
        const float inv_t = 1.f / tau;
        const WarpStat ts = warp_softmax_stat(t, C, inv_t, lane);
        const WarpStat k = warp_softmax_stat(zk, C, inv_t, lane);
        const float lse_t = ts.mx + __logf(ts.sum), lse_k = k.mx + __logf(k.sum);
        const float w = grad_scale * alpha * tau * invB / (float)C;
        float kl = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float lq = t[c] * inv_t - lse_t, lp = zk[c] * inv_t - lse_k;
            const float q = __expf(lq), p = __expf(lp);
            kl += q * (lq - lp);
            const float g = w * (p - q);
            if (dzkd) dzkd[(long long)row * lddzkd + c] = __float2bfloat16_rn(g);
            if (dzkd_f32) dzkd_f32[(long long)row * lddzkdf + c] = g;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
        loss += alpha * tau * tau * invB / (float)C * kl;
    }
    if (lane == 0) {
        atomicAdd(loss_accum, loss);
        if (correct_accum && a.arg == y) atomicAdd(correct_accum, 1);   // deit.py:72-74 counts the class-token head's arg-max
    }
}

__global__ void step_inc_kernel(int* step) { *step += 1; }

__global__ void adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                            uint2* __restrict__ p_bf16, long long n4, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2_sqrt, float grad_scale, const int* __restrict__ step_dev) {
    if (step_dev) {  // device-side step counter: lets the whole training step live in one CUDA graph
        const float t = (float)(*step_dev);
        bc1 = 1.f - powf(b1, t);
        bc2_sqrt = sqrtf(1.f - powf(b2, t));
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = p[i], gv = g[i], mv = m[i], vv = v[i];
        float* pp = reinterpret_cast<float*>(&pv);
        float* gp = reinterpret_cast<float*>(&gv);
        float* mp = reinterpret_cast<float*>(&mv);
        float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gr = gp[j] * grad_scale + wd * pp[j];
            mp[j] = b1 * mp[j] + (1.f - b1) * gr;
            vp[j] = b2 * vp[j] + (1.f - b2) * gr * gr;
            const float denom = sqrtf(vp[j]) / bc2_sqrt + eps;
            pp[j] -= (lr / bc1) * (mp[j] / denom);
        }
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (p_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(pp[0], pp[1]), hi = __floats2bfloat162_rn(pp[2], pp[3]);
            uint2 r;
            r.x = *reinterpret_cast<uint32_t*>(&lo);
            r.y = *reinterpret_cast<uint32_t*>(&hi);
            p_bf16[i] = r;
        }
    }
}

}  // namespace vb

extern "C" int vb_cross_entropy(const float* logits, int64_t ld, const int64_t* labels, int32_t B, int32_t C, float* loss_accum,
                                float weight, void* dlogits_bf16, int64_t lddz, float* dlogits_f32, int64_t lddzf, float grad_scale,
                                int32_t* correct_accum, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(logits && labels && loss_accum && B > 0 && C > 0, "cross_entropy: bad arguments");
    ce_kernel<<<(B + 7) / 8, 256, 0, as_stream(stream)>>>(logits, ld, reinterpret_cast<const long long*>(labels), B, C, loss_accum, weight,
                                                          reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), lddz, dlogits_f32, lddzf,
                                                          grad_scale, correct_accum);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_distill_loss(const float* logits, int64_t ld, const float* logits_kd, int64_t ldkd, const float* teacher_logits,
                               int64_t ldt, const int64_t* labels, int32_t B, int32_t C, int32_t kind, float alpha, float tau,
                               float* loss_accum, void* dlogits_bf16, int64_t lddz, void* dlogits_kd_bf16, int64_t lddzkd,
                               float* dlogits_f32, int64_t lddzf, float* dlogits_kd_f32, int64_t lddzkdf, float grad_scale,
                               int32_t* correct_accum, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(logits && logits_kd && teacher_logits && labels && loss_accum && B > 0 && C > 0, "distill_loss: bad arguments");
    VB_REQUIRE(kind == 1 || kind == 2, "distill_loss: kind must be 1 (soft) or 2 (hard)");
    VB_REQUIRE(kind == 2 || tau > 0.f, "distill_loss: tau must be positive");
    distill_loss_kernel<<<(B + 7) / 8, 256, 0, as_stream(stream)>>>(
        logits, ld, logits_kd, ldkd, teacher_logits, ldt, reinterpret_cast<const long long*>(labels), B, C, kind, alpha, tau, loss_accum,
        reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), lddz, reinterpret_cast<__nv_bfloat16*>(dlogits_kd_bf16), lddzkd, dlogits_f32, lddzf,
        dlogits_kd_f32, lddzkdf, grad_scale, correct_accum);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16, int64_t n,
                            float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                            int32_t* step_counter_dev, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(params && grads && exp_avg && exp_avg_sq && n >= 0 && n % 4 == 0 && (step >= 1 || step_counter_dev), "adam_step: bad arguments");
    if (step_counter_dev) {
        step_inc_kernel<<<1, 1, 0, as_stream(stream)>>>(step_counter_dev);
        VB_CUDA_CHECK(cudaGetLastError());
        if (step < 1) step = 1;
    }
    if (n == 0) return VB_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    adam_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads),
                                                           reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq),
                                                           reinterpret_cast<uint2*>(params_bf16), n / 4, lr, beta1, beta2, eps,
                                                           weight_decay, bc1, bc2s, grad_scale, step_counter_dev);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}
