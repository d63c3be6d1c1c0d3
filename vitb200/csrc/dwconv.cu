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

// weights of four consecutive channels: 36 consecutive floats (c % 4 == 0 -> 16-byte aligned) read as nine float4, de-interleaved into
// k[tap][channel]
__device__ __forceinline__ void load_w4(const float* __restrict__ w, int c, float (&k)[9][4]) {
    float flat[36];
    const float4* src = reinterpret_cast<const float4*>(w + (long long)c * 9);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float4 v = __ldg(src + i);
        flat[4 * i] = v.x; flat[4 * i + 1] = v.y; flat[4 * i + 2] = v.z; flat[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int t = 0; t < 9; ++t) k[t][q] = flat[q * 9 + t];
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void fma4(float4& r, const float (&k)[4], const float4& v) {
    r.x = fmaf(k[0], v.x, r.x); r.y = fmaf(k[1], v.y, r.y); r.z = fmaf(k[2], v.z, r.z); r.w = fmaf(k[3], v.w, r.w);
}

// One thread = four channels of one ROW of the token grid (or of one prefix token): the 36 weights are loaded once and a 3 x 3 window
// of float4 slides along the row, so a token costs three new 16-byte loads instead of nine (+ 36 scalar weight loads in the first
// version of this kernel, which ran at 0.8 TB/s: profiles/r1d_hbm_kernels_ncu_summary.txt).  Consecutive threads own consecutive
// channel quads: every access of a warp is 512 contiguous bytes.
// TRANSPOSED = false: correlation (forward); true: the adjoint (taps mirrored)
template <bool TRANSPOSED>
__global__ void __launch_bounds__(128) dwconv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                     const float* __restrict__ pos, const float* __restrict__ sub, float* __restrict__ out,
                                                     float* __restrict__ sum_f32, __nv_bfloat16* __restrict__ sum_bf16, int B, int S, int D,
                                                     int n_prefix, int G) {
    const int d4 = D >> 2;
    const int rows = n_prefix + G;                         // work rows per image: the prefix tokens, then the G grid rows
    const long long total = (long long)B * rows * d4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % d4) * 4;
        const long long br = i / d4;
        const int r = (int)(br % rows);
        const long long b = br / rows;
        const float* xb = x + b * S * D + c;
        const long long ob = b * S * D + c;

        auto emit = [&](int s, const float4& self, float4 res) {
            const long long o = ob + (long long)s * D;
            if (pos) {
                const float4 pv = __ldg(reinterpret_cast<const float4*>(pos + (long long)s * D + c));
                res.x += pv.x; res.y += pv.y; res.z += pv.z; res.w += pv.w;
            }
            if (sub) {
                const float4 sv = ld4(sub + o);
                res.x += self.x - sv.x; res.y += self.y - sv.y; res.z += self.z - sv.z; res.w += self.w - sv.w;
            }
            if (out) *reinterpret_cast<float4*>(out + o) = res;
            if (sum_f32 || sum_bf16) {
                const float4 t = make_float4(self.x + res.x, self.y + res.y, self.z + res.z, self.w + res.w);
                if (sum_f32) *reinterpret_cast<float4*>(sum_f32 + o) = t;
                if (sum_bf16) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(sum_bf16 + o) = pk;
                }
            }
        };

        if (r < n_prefix) {                                // class / distillation token: passes through
            const float4 self = ld4(xb + (long long)r * D);
            emit(r, self, self);
            continue;
        }
        const int h = r - n_prefix;
        float k[9][4];
        load_w4(w, c, k);
        const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : zero4();
        const bool up = h > 0, down = h + 1 < G;
        const float* row_m = xb + (long long)(n_prefix + (h - 1) * G) * D;   // only dereferenced when `up`
        const float* row_c = xb + (long long)(n_prefix + h * G) * D;
        const float* row_p = xb + (long long)(n_prefix + (h + 1) * G) * D;   // only dereferenced when `down`
        float4 win[3][3];                                  // [row -1, 0, +1][column -1, 0, +1]
#pragma unroll
        for (int a = 0; a < 3; ++a) win[a][0] = win[a][1] = zero4();
        win[0][2] = up ? ld4(row_m) : zero4();
        win[1][2] = ld4(row_c);
        win[2][2] = down ? ld4(row_p) : zero4();
        for (int ww = 0; ww < G; ++ww) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { win[a][0] = win[a][1]; win[a][1] = win[a][2]; }
            if (ww + 1 < G) {
                const long long off = (long long)(ww + 1) * D;
                win[0][2] = up ? ld4(row_m + off) : zero4();
                win[1][2] = ld4(row_c + off);
                win[2][2] = down ? ld4(row_p + off) : zero4();
            } else {
                win[0][2] = win[1][2] = win[2][2] = zero4();
            }
            float4 res = bv;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    // forward: out[h,w] += w[a][e] * x[h+a-1, w+e-1];  adjoint: dx[h,w] += w[2-a][2-e] * dy[h+a-1, w+e-1]
                    const int t = TRANSPOSED ? (2 - a) * 3 + (2 - e) : a * 3 + e;
                    fma4(res, k[t], win[a][e]);
                }
            emit(n_prefix + h * G + ww, win[1][1], res);
        }
    }
}

// Weight gradient: one thread = four channels, looping over (image, grid row) units strided by gridDim.y with the same sliding
// window over x; 40 fp32 accumulators per thread, 40 atomics at the end.  grid (ceil(D / 4 / 64), chunks), 64 threads.
__global__ void __launch_bounds__(64) dwconv_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                                          float* __restrict__ db, int B, int S, int D, int n_prefix, int G) {
    const int c4 = blockIdx.x * 64 + threadIdx.x;
    if (c4 * 4 >= D) return;
    const int c = c4 * 4;
    float4 acc[9], accb = zero4();
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = zero4();
    const long long units = (long long)B * G;
    for (long long u = blockIdx.y; u < units; u += gridDim.y) {
        const long long b = u / G;
        const int h = (int)(u - b * G);
        const float* xb = x + b * S * D + c;
        const float* gb = dy + b * S * D + c + (long long)(n_prefix + h * G) * D;
        const bool up = h > 0, down = h + 1 < G;
        const float* row_m = xb + (long long)(n_prefix + (h - 1) * G) * D;
        const float* row_c = xb + (long long)(n_prefix + h * G) * D;
        const float* row_p = xb + (long long)(n_prefix + (h + 1) * G) * D;
        float4 win[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a) win[a][0] = win[a][1] = zero4();
        win[0][2] = up ? ld4(row_m) : zero4();
        win[1][2] = ld4(row_c);
        win[2][2] = down ? ld4(row_p) : zero4();
        for (int ww = 0; ww < G; ++ww) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { win[a][0] = win[a][1]; win[a][1] = win[a][2]; }
            if (ww + 1 < G) {
                const long long off = (long long)(ww + 1) * D;
                win[0][2] = up ? ld4(row_m + off) : zero4();
                win[1][2] = ld4(row_c + off);
                win[2][2] = down ? ld4(row_p + off) : zero4();
            } else {
                win[0][2] = win[1][2] = win[2][2] = zero4();
            }
            const float4 g = ld4(gb + (long long)ww * D);
            accb.x += g.x; accb.y += g.y; accb.z += g.z; accb.w += g.w;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    float4& t = acc[a * 3 + e];
                    const float4& v = win[a][e];
                    t.x = fmaf(g.x, v.x, t.x); t.y = fmaf(g.y, v.y, t.y); t.z = fmaf(g.z, v.z, t.z); t.w = fmaf(g.w, v.w, t.w);
                }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        atomicAdd(dw + (long long)(c + 0) * 9 + t, acc[t].x);
        atomicAdd(dw + (long long)(c + 1) * 9 + t, acc[t].y);
        atomicAdd(dw + (long long)(c + 2) * 9 + t, acc[t].z);
        atomicAdd(dw + (long long)(c + 3) * 9 + t, acc[t].w);
    }
    if (db) {
        atomicAdd(db + c, accb.x); atomicAdd(db + c + 1, accb.y); atomicAdd(db + c + 2, accb.z); atomicAdd(db + c + 3, accb.w);
    }
}

static int dw_grid(long long work_items) {
    long long blocks = (work_items + 127) / 128;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static int check_dw(const void* x, const void* w, int B, int S, int D, int n_prefix, int G) {
    VB_REQUIRE(x && w && B > 0 && D > 0 && D % 4 == 0 && n_prefix >= 0 && G > 0, "dwconv: bad arguments");
    VB_REQUIRE(n_prefix + G * G == S, "dwconv: S = %d must be n_prefix + G*G (n_prefix %d, G %d) — cpe_vit.py:25", S, n_prefix, G);
    VB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, "dwconv: tensors must be 16-byte aligned");
    return VB_OK;
}

}  // namespace vb

extern "C" int vb_dwconv3x3_fwd(const float* x, const float* w, const float* bias, const float* pos, const float* sub, float* out,
                                int32_t B, int32_t S, int32_t D, int32_t n_prefix, int32_t G, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    if (int rc = check_dw(x, w, B, S, D, n_prefix, G)) return rc;
    VB_REQUIRE(out && out != x, "dwconv_fwd: needs an output distinct from the input");
    dwconv_kernel<false><<<dw_grid((long long)B * (n_prefix + G) * (D / 4)), 128, 0, as_stream(stream)>>>(x, w, bias, pos, sub, out, nullptr,
                                                                                                      nullptr, B, S, D, n_prefix, G);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_dwconv3x3_bwd_data(const float* dy, const float* w, float* dx, float* sum_f32, void* sum_bf16, int32_t B, int32_t S,
                                     int32_t D, int32_t n_prefix, int32_t G, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    if (int rc = check_dw(dy, w, B, S, D, n_prefix, G)) return rc;
    VB_REQUIRE((dx || sum_f32 || sum_bf16) && dx != dy && sum_f32 != dy, "dwconv_bwd_data: outputs must be distinct from dy");
    dwconv_kernel<true><<<dw_grid((long long)B * (n_prefix + G) * (D / 4)), 128, 0, as_stream(stream)>>>(
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
    const int cblocks = (D / 4 + 63) / 64;
    long long chunks = ((long long)num_sms() * 16 + cblocks - 1) / cblocks;     // ~16 small CTAs per SM keep enough loads in flight
    if (chunks > (long long)B * G) chunks = (long long)B * G;
    if (chunks < 1) chunks = 1;
    dwconv_wgrad_kernel<<<dim3(cblocks, (unsigned)chunks), 64, 0, as_stream(stream)>>>(dy, x, dw_accum, db_accum, B, S, D, n_prefix, G);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}
