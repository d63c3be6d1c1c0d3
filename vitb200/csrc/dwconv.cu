// Conditional positional encoding (PEG): a depthwise 3x3 convolution over the patch-token grid of a token-major fp32 stream.
//
// Reference: ConditionalPositionalEncoding at models/image_classification/cpe_vit.py:16-30 == cpvt.py:16-30 — the class token is
// split off, the remaining S-1 tokens are viewed as a [D, G, G] image (G = sqrt(S-1)), nn.Conv2d(D, D, 3, padding=1, groups=D) is
// applied and the class token is re-attached.  Used once after the patch embedding (cpe_vit.py:143,197; cpvt.py:144,199) and, in
// CPVT, at the end of every encoder block (cpvt.py:80,93-96).
//
// Here the tokens never leave their [B, S, D] layout: channel c of token (h, w) sits at x[b, n_prefix + h*G + w, c], so the nine taps
// of a channel are nine token rows at the same column, consecutive threads own consecutive channel quads (16-byte accesses) and the
// 9x re-read is served by L1/L2.  HBM-bound: 4 B/elem in + 4 B/elem out (+ the fused operands below).
//
//   forward   out = [prefix rows: x | patch rows: conv(x) + bias] (+ pos[s, :] broadcast over the batch) (+ x - sub)
//             (+ pos:  Encoder.forward's input + pos_embedding, cpe_vit.py:112;  + x - sub: CPVT's block tail, cpvt.py:93-96:
//              x2 = x1 + y; return peg(x2) + y  with y recovered as x2 - x1)
//   dgrad     dx = [prefix rows: dy | patch rows: conv^T(dy)];  optional sum = dy + dx as fp32 and / or bf16
//   wgrad     dw[c, i, j] += sum_{b,h,w} dy[b,h,w,c] * x[b,h+i-1,w+j-1,c];   db[c] += sum_{b,h,w} dy[b,h,w,c]
#include "common.h"
#include <cuda_bf16.h>

namespace vb {

__device__ __forceinline__ void load_w4(const float* __restrict__ w, int c, float (&k)[9][4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int t = 0; t < 9; ++t) k[t][q] = __ldg(w + (long long)(c + q) * 9 + t);
}

// TRANSPOSED = false: correlation (forward); true: the adjoint (taps mirrored)
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256) dwconv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                     const float* __restrict__ pos, const float* __restrict__ sub, float* __restrict__ out,
                                                     float* __restrict__ sum_f32, __nv_bfloat16* __restrict__ sum_bf16, int B, int S, int D,
                                                     int n_prefix, int G) {
    const int d4 = D >> 2;
    const long long total = (long long)B * S * d4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % d4) * 4;
        const long long tok = i / d4;
        const int s = (int)(tok % S);
        const long long b = tok / S;
        const float* xb = x + b * S * D;
        const float4 self = *reinterpret_cast<const float4*>(xb + (long long)s * D + c);
        float4 r = self;
        if (s >= n_prefix) {
            float k[9][4];
            load_w4(w, c, k);
            const int p = s - n_prefix, h = p / G, ww = p - h * G;
            r = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int di = -1; di <= 1; ++di) {
#pragma unroll
                for (int dj = -1; dj <= 1; ++dj) {
                    const int hh = h + di, wj = ww + dj;
                    if (hh < 0 || hh >= G || wj < 0 || wj >= G) continue;
                    const float4 v = *reinterpret_cast<const float4*>(xb + (long long)(n_prefix + hh * G + wj) * D + c);
                    // forward: out[h,w] += w[di+1][dj+1] * x[h+di, w+dj];  adjoint: dx[h,w] += w[1-di][1-dj] * dy[h+di, w+dj]
                    const int t = TRANSPOSED ? (1 - di) * 3 + (1 - dj) : (di + 1) * 3 + (dj + 1);
                    r.x = fmaf(k[t][0], v.x, r.x); r.y = fmaf(k[t][1], v.y, r.y);
                    r.z = fmaf(k[t][2], v.z, r.z); r.w = fmaf(k[t][3], v.w, r.w);
                }
            }
        }
        if (pos) {
            const float4 pv = __ldg(reinterpret_cast<const float4*>(pos + (long long)s * D + c));
            r.x += pv.x; r.y += pv.y; r.z += pv.z; r.w += pv.w;
        }
        if (sub) {
            const float4 sv = *reinterpret_cast<const float4*>(sub + tok * D + c);
            r.x += self.x - sv.x; r.y += self.y - sv.y; r.z += self.z - sv.z; r.w += self.w - sv.w;
        }
        if (out) *reinterpret_cast<float4*>(out + tok * D + c) = r;
        if (sum_f32 || sum_bf16) {
            const float4 t = make_float4(self.x + r.x, self.y + r.y, self.z + r.z, self.w + r.w);
            if (sum_f32) *reinterpret_cast<float4*>(sum_f32 + tok * D + c) = t;
            if (sum_bf16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(sum_bf16 + tok * D + c) = pk;
            }
        }
    }
}

// grid (D / 128, chunks): thread = one channel, a block sums its chunk of images; ten atomics per thread at the end
__global__ void __launch_bounds__(128) dwconv_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                                           float* __restrict__ db, int B, int S, int D, int n_prefix, int G,
                                                           int images_per_chunk) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= D) return;
    const int b0 = blockIdx.y * images_per_chunk;
    const int b1 = min(B, b0 + images_per_chunk);
    float acc[9], accb = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    for (int b = b0; b < b1; ++b) {
        const float* xb = x + (long long)b * S * D + c;
        const float* gb = dy + (long long)b * S * D + c;
        for (int h = 0; h < G; ++h) {
            for (int ww = 0; ww < G; ++ww) {
                const float g = gb[(long long)(n_prefix + h * G + ww) * D];
                accb += g;
#pragma unroll
                for (int di = -1; di <= 1; ++di) {
#pragma unroll
                    for (int dj = -1; dj <= 1; ++dj) {
                        const int hh = h + di, wj = ww + dj;
                        if (hh < 0 || hh >= G || wj < 0 || wj >= G) continue;
                        acc[(di + 1) * 3 + (dj + 1)] = fmaf(g, xb[(long long)(n_prefix + hh * G + wj) * D], acc[(di + 1) * 3 + (dj + 1)]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(dw + (long long)c * 9 + t, acc[t]);
    if (db) atomicAdd(db + c, accb);
}

static int dw_grid(long long work_items) {
    long long blocks = (work_items + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static int check_dw(const void* x, const void* w, int B, int S, int D, int n_prefix, int G) {
    VB_REQUIRE(x && w && B > 0 && D > 0 && D % 4 == 0 && n_prefix >= 0 && G > 0, "dwconv: bad arguments");
    VB_REQUIRE(n_prefix + G * G == S, "dwconv: S = %d must be n_prefix + G*G (n_prefix %d, G %d) — cpe_vit.py:25", S, n_prefix, G);
    VB_REQUIRE(((uintptr_t)x & 15) == 0, "dwconv: tensors must be 16-byte aligned");
    return VB_OK;
}

}  // namespace vb

extern "C" int vb_dwconv3x3_fwd(const float* x, const float* w, const float* bias, const float* pos, const float* sub, float* out,
                                int32_t B, int32_t S, int32_t D, int32_t n_prefix, int32_t G, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    if (int rc = check_dw(x, w, B, S, D, n_prefix, G)) return rc;
    VB_REQUIRE(out && out != x, "dwconv_fwd: needs an output distinct from the input");
    dwconv_kernel<false><<<dw_grid((long long)B * S * (D / 4)), 256, 0, as_stream(stream)>>>(x, w, bias, pos, sub, out, nullptr, nullptr, B, S, D,
                                                                                         n_prefix, G);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_dwconv3x3_bwd_data(const float* dy, const float* w, float* dx, float* sum_f32, void* sum_bf16, int32_t B, int32_t S,
                                     int32_t D, int32_t n_prefix, int32_t G, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    if (int rc = check_dw(dy, w, B, S, D, n_prefix, G)) return rc;
    VB_REQUIRE((dx || sum_f32 || sum_bf16) && dx != dy && sum_f32 != dy, "dwconv_bwd_data: outputs must be distinct from dy");
    dwconv_kernel<true><<<dw_grid((long long)B * S * (D / 4)), 256, 0, as_stream(stream)>>>(
        dy, w, nullptr, nullptr, nullptr, dx, sum_f32, reinterpret_cast<__nv_bfloat16*>(sum_bf16), B, S, D, n_prefix, G);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_dwconv3x3_bwd_weight(const float* dy, const float* x, float* dw_accum, float* db_accum, int32_t B, int32_t S, int32_t D,
                                       int32_t n_prefix, int32_t G, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    if (int rc = check_dw(dy, x, B, S, D, n_prefix, G)) return rc;
    VB_REQUIRE(dw_accum, "dwconv_bwd_weight: null gradient buffer");
    const int cblocks = (D + 127) / 128;
    int chunks = (num_sms() * 4 + cblocks - 1) / cblocks;
    if (chunks > B) chunks = B;
    const int ipc = (B + chunks - 1) / chunks;
    chunks = (B + ipc - 1) / ipc;
    dwconv_wgrad_kernel<<<dim3(cblocks, chunks), 128, 0, as_stream(stream)>>>(dy, x, dw_accum, db_accum, B, S, D, n_prefix, G, ipc);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}
