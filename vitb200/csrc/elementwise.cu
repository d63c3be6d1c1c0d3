// HBM-bound helper kernels around the encoder: parameter cast, patch extraction, token rows, column sums,
// embedding backward.  All vectorised (128-bit where alignment allows) and grid-sized from the SM count.
#include "common.h"
#include <cuda_bf16.h>

namespace vb {

__device__ __forceinline__ uint2 pack4(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&lo);
    r.y = *reinterpret_cast<uint32_t*>(&hi);
    return r;
}

// ---- fp32 -> bf16 cast of a flat buffer (master parameters -> tensor-core operands) -------------------------
__global__ void cast_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src + i);
        dst[i] = pack4(v.x, v.y, v.z, v.w);
    }
}

// ---- images [B,C,H,W] fp32 -> patch matrix [B, (H/p)*(W/p), C*p*p] bf16, k = c*p*p + i*p + j ---------------
// (the K ordering of conv_proj.weight.reshape(D, -1): vanilla_vit.py:129,196)
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int C, int H, int W, int p) {
    // One thread per 16 bytes of OUTPUT (8 consecutive pixels of one patch row): stores are fully coalesced along k, loads are 32-byte
    // sectors of an image row (p % 8 == 0), or one thread per 8 bytes of output when only p % 4 == 0 holds.
    const int nw = W / p, np = (H / p) * nw, kdim = C * p * p;
    const int vec = (p % 8 == 0) ? 8 : 4;
    const long long total = (long long)B * C * H * W / vec;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long o = t * vec;                   // flat output index = (b * np + patch) * kdim + k
        const int k = (int)(o % kdim);
        const long long bp = o / kdim;
        const int patch = (int)(bp % np);
        const long long b = bp / np;
        const int c = k / (p * p), ij = k - c * p * p, i = ij / p, j = ij - i * p;
        const int ph = patch / nw, pw = patch - ph * nw;
        const float* src = img + ((b * C + c) * H + (ph * p + i)) * (long long)W + pw * p + j;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
        if (vec == 8) {
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
            const uint2 lo = pack4(v0.x, v0.y, v0.z, v0.w), hi = pack4(v1.x, v1.y, v1.z, v1.w);
            *reinterpret_cast<uint4*>(out + o) = make_uint4(lo.x, lo.y, hi.x, hi.y);
        } else {
            *reinterpret_cast<uint2*>(out + o) = pack4(v0.x, v0.y, v0.z, v0.w);
        }
    }
}

// ---- inverse of patchify: d images [B,C,H,W] fp32 <- d patches [B, (H/p)(W/p), C*p*p] bf16 (the input gradient of conv_proj) ----
__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ dpat, float* __restrict__ dimg, int B, int C, int H, int W, int p) {
    const int nw = W / p, np = (H / p) * nw, kdim = C * p * p;
    const long long total = (long long)B * C * H * W / 4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long o = t * 4;                     // flat patch-matrix index (b * np + patch) * kdim + k
        const int k = (int)(o % kdim);
        const long long bp = o / kdim;
        const int patch = (int)(bp % np);
        const long long b = bp / np;
        const int c = k / (p * p), ij = k - c * p * p, i = ij / p, j = ij - i * p;
        const int ph = patch / nw, pw = patch - ph * nw;
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(dpat + o));
        float* dst = dimg + ((b * C + c) * H + (ph * p + i)) * (long long)W + pw * p + j;
        *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xFFFF0000u),
                                                      __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xFFFF0000u));
    }
}

// ---- prefix token rows: x[b, t, :] = token_t + pos[t]  (vanilla_vit.py:202-203 + :104) ----------------------
__global__ void token_rows_kernel(float* __restrict__ x, const float* __restrict__ tok0, const float* __restrict__ tok1,
                                  const float* __restrict__ pos, int B, int S, int D, int n_prefix) {
    const long long total = (long long)B * n_prefix * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = i % D;
        const long long r = i / D;
        const int t = r % n_prefix;
        const long long b = r / n_prefix;
        const float* tok = t == 0 ? tok0 : tok1;
        x[(b * S + t) * D + c] = tok[c] + (pos ? pos[(long long)t * D + c] : 0.f);
    }
}

// ---- out[c] += sum_r x[r, c] for bf16 x (bias gradients) ----------------------------------------------------
// block = 256 threads = 8 warps; a warp covers 256 columns (8 per lane, one 16-byte load) of one row per step.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int rows, int cols,
                                                          float* __restrict__ out, int rows_per_block) {
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + lane * 8;
    const int r_begin = blockIdx.y * rows_per_block;
    const int r_end = min(rows, r_begin + rows_per_block);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c0 < cols) {
        // fp32 += bf16 with the mixed-precision add of sm_100 (FHADD.BF16 on half-register selectors: no unpacking)
        auto add8 = [&](const uint4& w) {
            const uint32_t v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}"
                    : "+f"(acc[2 * q]), "+f"(acc[2 * q + 1]) : "r"(v[q]));
        };
        int r = r_begin + warp;
        for (; r + 24 < r_end; r += 32) {   // four independent 16-byte loads in flight per lane
            const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(x + (long long)r * ld + c0));
            const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(x + (long long)(r + 8) * ld + c0));
            const uint4 w2 = __ldg(reinterpret_cast<const uint4*>(x + (long long)(r + 16) * ld + c0));
            const uint4 w3 = __ldg(reinterpret_cast<const uint4*>(x + (long long)(r + 24) * ld + c0));
            add8(w0); add8(w1); add8(w2); add8(w3);
        }
        for (; r < r_end; r += 8) add8(__ldg(reinterpret_cast<const uint4*>(x + (long long)r * ld + c0)));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < cols) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        atomicAdd(out + c, s);
    }
}

// ---- embedding backward, pass 1: per-position batch sums + compact bf16 copy of the patch-row gradients ------
// dx [B,S,D] fp32.  possum[s, c] += sum_{b in chunk} dx[b,s,c]   (possum zeroed by the caller)
// dxp[b*P + (s - n_prefix), c] = bf16(dx[b,s,c]) for s >= n_prefix   (A operand of the conv_proj wgrad GEMM)
__global__ void embed_bwd_reduce_kernel(const float* __restrict__ dx, float* __restrict__ possum, __nv_bfloat16* __restrict__ dxp,
                                        int B, int S, int D, int n_prefix, int b_per_chunk) {
    const int s = blockIdx.x;
    const int b0 = blockIdx.y * b_per_chunk, b1 = min(B, b0 + b_per_chunk);
    const int P = S - n_prefix;
    for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int b = b0; b < b1; ++b) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(dx + ((long long)b * S + s) * D + c));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            if (s >= n_prefix && dxp) *reinterpret_cast<uint2*>(dxp + ((long long)b * P + (s - n_prefix)) * D + c) = pack4(v.x, v.y, v.z, v.w);
        }
        float* o = possum + (long long)s * D + c;
        atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
    }
}
// pass 2: dpos += possum; dtok_t += possum[t] (t < n_prefix); dbias += sum_{s >= n_prefix} possum[s]
__global__ void embed_bwd_finalize_kernel(const float* __restrict__ possum, float* __restrict__ dpos, float* __restrict__ dtok0,
                                          float* __restrict__ dtok1, float* __restrict__ dbias, int S, int D, int n_prefix, int rows_per_block) {
    // grid (ceil(D / 128), ceil(S / rows_per_block)): a thread owns one column of a few positions (the serial 197-step walk of a
    // 6-block grid was latency-bound); the conv-bias partial sums of the position chunks meet in one atomicAdd per thread
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    const int s0 = blockIdx.y * rows_per_block, s1 = min(S, s0 + rows_per_block);
    float bsum = 0.f;
    for (int s = s0; s < s1; ++s) {
        const float v = possum[(long long)s * D + c];
        if (dpos) dpos[(long long)s * D + c] += v;
        if (s < n_prefix) {
            float* t = s == 0 ? dtok0 : dtok1;
            if (t) t[c] += v;
        } else {
            bsum += v;
        }
    }
    if (dbias && s1 > n_prefix) atomicAdd(dbias + c, bsum);
}

// ---- out[r, :] = x[r, :] + pos[r % period, :]  (Encoder.forward on caller-supplied tokens: input + pos_embedding, vanilla_vit.py:104)
__global__ void add_rows_bcast_kernel(const float4* __restrict__ x, const float4* __restrict__ pos, float4* __restrict__ out, long long rows,
                                      int period, int d4) {
    const long long total = rows * d4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d4;
        const int c = (int)(i - r * d4);
        float4 v = __ldg(x + i);
        const float4 w = __ldg(pos + (long long)(r % period) * d4 + c);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        out[i] = v;
    }
}
// ---- out_bf16 = bf16(a + b) (b optional): DETR q = k = src + pos operand and the bf16 copy of src -------------
__global__ void add_cast_kernel(const float4* __restrict__ a, const float4* __restrict__ b, uint2* __restrict__ out, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(a + i);
        if (b) { const float4 w = __ldg(b + i); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        out[i] = pack4(v.x, v.y, v.z, v.w);
    }
}
// ---- out_f32 = a_f32 + b_bf16 (+ c_bf16);  accum_f32 += b_bf16 (optional): DETR d_src / d_pos assembly ---------
__global__ void add3_kernel(const float4* __restrict__ a, const uint2* __restrict__ b, const uint2* __restrict__ c, float4* __restrict__ out,
                            float4* __restrict__ accum, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(a + i);
        const uint2 w = __ldg(b + i);
        const float4 bv = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                                      __uint_as_float(w.y & 0xFFFF0000u));
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        if (c) {
            const uint2 u = __ldg(c + i);
            v.x += __uint_as_float(u.x << 16); v.y += __uint_as_float(u.x & 0xFFFF0000u);
            v.z += __uint_as_float(u.y << 16); v.w += __uint_as_float(u.y & 0xFFFF0000u);
        }
        out[i] = v;
        if (accum) { float4 t = accum[i]; t.x += bv.x; t.y += bv.y; t.z += bv.z; t.w += bv.w; accum[i] = t; }
    }
}

// ---- dst_bf16[r, c] = bf16(src_f32[r, c]) for row-pitched matrices: the NCHW backbone feature map as the (K x M, M contiguous) bf16
// operand of the input_proj 1x1 convolution-as-GEMM (detr.py:125); the destination pitch is padded to a TMA-legal multiple of 8 ----
__global__ void cast_rows_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst, long long ldd, long long rows,
                                 int cols) {
    const int cp = (cols + 1) >> 1;   // pairs per row
    const long long total = rows * cp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cp;
        const int c = (int)(i - r * cp) * 2;
        const float a = src[r * lds + c];
        const float b = c + 1 < cols ? src[r * lds + c + 1] : 0.f;
        if (c + 1 < ldd) *reinterpret_cast<__nv_bfloat162*>(dst + r * ldd + c) = __floats2bfloat162_rn(a, b);
        else dst[r * ldd + c] = __float2bfloat16_rn(a);
    }
}

// ---- learned 2-D position embedding of DETR (AbsolutePositionalEncoding.forward, detr.py:49-63) written straight in the encoder's
// sequence-first layout: pos[(y * w + x), n, c] = c < pf ? col_embed[x, c] : row_embed[y, c - pf]  (the reference builds [N, 2 pf, h, w]
// with cat / permute / repeat and Transformer.forward flattens it back: transformer.py:49-52) ----
__global__ void pos_embed_2d_fwd_kernel(const float* __restrict__ row_embed, const float* __restrict__ col_embed, float* __restrict__ pos,
                                        int h, int w, int N, int pf) {
    const int C4 = pf / 2;   // float4 per embedding half... (2 * pf) / 4 vectors per token, pf / 4 per half
    const long long total = (long long)h * w * N * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % C4);
        const long long tok = i / C4;             // (y * w + x) * N + n
        const long long yx = tok / N;
        const int y = (int)(yx / w), x = (int)(yx - (long long)y * w);
        const int c = v * 4;
        const float4 e = c < pf ? __ldg(reinterpret_cast<const float4*>(col_embed + (long long)x * pf + c))
                                : __ldg(reinterpret_cast<const float4*>(row_embed + (long long)y * pf + (c - pf)));
        *reinterpret_cast<float4*>(pos + tok * (2 * pf) + c) = e;
    }
}
// d col_embed[x] += sum_{y, n} dpos[(y, x), n, :pf];  d row_embed[y] += sum_{x, n} dpos[(y, x), n, pf:]   (fp32 atomics: a few MB per step)
__global__ void pos_embed_2d_bwd_kernel(const float* __restrict__ dpos, float* __restrict__ drow, float* __restrict__ dcol, int h, int w, int N,
                                        int pf) {
    // one thread per (y or x, channel): the serial walk over the other grid axis and the batch keeps the sums deterministic
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int line = blockIdx.y;                 // 0 .. w - 1: column x;  w .. w + h - 1: row y
    if (c >= pf) return;
    float acc = 0.f;
    if (line < w) {
        const int x = line;
        for (int y = 0; y < h; ++y)
            for (int n = 0; n < N; ++n) acc += dpos[(((long long)y * w + x) * N + n) * (2 * pf) + c];
        dcol[(long long)x * pf + c] += acc;
    } else {
        const int y = line - w;
        for (int x = 0; x < w; ++x)
            for (int n = 0; n < N; ++n) acc += dpos[(((long long)y * w + x) * N + n) * (2 * pf) + pf + c];
        drow[(long long)y * pf + c] += acc;
    }
}

// ---- y[n] += sum_k x[k] * W[k, n]  (fp32, W row-major [K, N]);  x_accum[k] += x[k] (optional) -------------------------------------
// The value-bias gradient of the packed in-projection without touching dV: colsum(dV) = colsum(dO) (rows of P sum to one) and
// dO = d W_proj, so colsum(dO) = colsum(d) W_proj — colsum(d) is the out-proj bias gradient the LayerNorm backward has just summed.
__global__ void __launch_bounds__(128) vecmat_accum_kernel(const float* __restrict__ x, const float* __restrict__ W, long long ldw,
                                                           float* __restrict__ y, float* __restrict__ x_accum, int K, int N, int k_per_block) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int k0 = blockIdx.y * k_per_block, k1 = min(K, k0 + k_per_block);
    if (x_accum != nullptr && blockIdx.x == 0)
        for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x) x_accum[k] += x[k];
    if (n >= N) return;
    float acc = 0.f;
#pragma unroll 4
    for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(x + k), __ldg(W + (long long)k * ldw + n), acc);
    atomicAdd(y + n, acc);
}

static int grid_for(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace vb

extern "C" int vb_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(src && dst && n >= 0 && n % 4 == 0, "cast: n=%lld must be a multiple of 4", (long long)n);
    VB_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "cast: misaligned pointers");
    if (n == 0) return VB_OK;
    cast_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n / 4);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_patchify(const float* images, void* patches_bf16, int32_t B, int32_t C, int32_t H, int32_t W, int32_t patch,
                           void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(images && patches_bf16 && B > 0 && C > 0 && patch > 0, "patchify: bad arguments");
    VB_REQUIRE(H % patch == 0 && W % patch == 0 && patch % 4 == 0, "patchify: H, W must be multiples of patch and patch of 4");
    VB_REQUIRE(((uintptr_t)images & 15) == 0 && ((uintptr_t)patches_bf16 & 7) == 0, "patchify: misaligned pointers");
    const long long items = (long long)B * C * H * W / (patch % 8 == 0 ? 8 : 4);
    VB_REQUIRE(patch % 8 != 0 || ((uintptr_t)patches_bf16 & 15) == 0, "patchify: output must be 16-byte aligned");
    patchify_kernel<<<grid_for(items, 256), 256, 0, as_stream(stream)>>>(images, reinterpret_cast<__nv_bfloat16*>(patches_bf16), B, C, H, W, patch);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_unpatchify(const void* dpatches_bf16, float* dimages, int32_t B, int32_t C, int32_t H, int32_t W, int32_t patch,
                             void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(dpatches_bf16 && dimages && B > 0 && C > 0 && patch > 0, "unpatchify: bad arguments");
    VB_REQUIRE(H % patch == 0 && W % patch == 0 && patch % 4 == 0, "unpatchify: H, W must be multiples of patch and patch of 4");
    VB_REQUIRE(((uintptr_t)dimages & 15) == 0 && ((uintptr_t)dpatches_bf16 & 7) == 0, "unpatchify: misaligned pointers");
    unpatchify_kernel<<<grid_for((long long)B * C * H * W / 4, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(dpatches_bf16), dimages, B, C, H, W, patch);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_token_rows(float* x, const float* tok0, const float* tok1, const float* pos, int32_t B, int32_t S, int32_t D,
                             int32_t n_prefix, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(x && tok0 && n_prefix >= 1 && n_prefix <= 2 && (n_prefix == 1 || tok1), "token_rows: bad arguments");
    token_rows_kernel<<<grid_for((long long)B * n_prefix * D, 256), 256, 0, as_stream(stream)>>>(x, tok0, tok1, pos, B, S, D, n_prefix);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_colsum_bf16(const void* x, int64_t ld, int32_t rows, int32_t cols, float* out_accum, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(x && out_accum && rows >= 0 && cols > 0, "colsum: bad arguments");
    VB_REQUIRE(cols % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0, "colsum: cols and pitch must be multiples of 8, base 16-byte aligned");
    if (rows == 0) return VB_OK;
    const int col_blocks = (cols + 255) / 256;
    int row_blocks = (num_sms() * 4 + col_blocks - 1) / col_blocks;
    int rpb = (rows + row_blocks - 1) / row_blocks;
    if (rpb < 64) rpb = 64;
    row_blocks = (rows + rpb - 1) / rpb;
    dim3 grid(col_blocks, row_blocks);
    colsum_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, rows, cols, out_accum, rpb);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_embed_bwd(const float* dx, float* possum_scratch, void* dx_patches_bf16, float* dpos, float* dtok0, float* dtok1,
                            float* dbias, int32_t B, int32_t S, int32_t D, int32_t n_prefix, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(dx && possum_scratch && B > 0 && S > n_prefix && D % 4 == 0, "embed_bwd: bad arguments");
    cudaStream_t st = as_stream(stream);
    VB_CUDA_CHECK(cudaMemsetAsync(possum_scratch, 0, sizeof(float) * (size_t)S * D, st));
    int chunks = (num_sms() * 4 + S - 1) / S;
    if (chunks > B) chunks = B;
    if (chunks < 1) chunks = 1;
    const int bpc = (B + chunks - 1) / chunks;
    chunks = (B + bpc - 1) / bpc;
    int threads = D / 4;
    if (threads > 256) threads = 256;
    threads = (threads + 31) / 32 * 32;
    embed_bwd_reduce_kernel<<<dim3(S, chunks), threads, 0, st>>>(dx, possum_scratch, reinterpret_cast<__nv_bfloat16*>(dx_patches_bf16), B, S,
                                                                D, n_prefix, bpc);
    VB_CUDA_CHECK(cudaGetLastError());
    const int rpb = 8;
    embed_bwd_finalize_kernel<<<dim3((D + 127) / 128, (S + rpb - 1) / rpb), 128, 0, st>>>(possum_scratch, dpos, dtok0, dtok1, dbias, S, D, n_prefix, rpb);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_add_rows_bcast(const float* x, const float* pos, float* out, int64_t rows, int32_t period, int32_t D, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(x && pos && out && rows >= 0 && period > 0 && D > 0 && D % 4 == 0, "add_rows_bcast: bad arguments");
    if (rows == 0) return VB_OK;
    add_rows_bcast_kernel<<<grid_for(rows * (D / 4), 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(pos), reinterpret_cast<float4*>(out), rows, period, D / 4);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_add_cast_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(a && out_bf16 && n >= 0 && n % 4 == 0, "add_cast: n=%lld must be a multiple of 4", (long long)n);
    if (n == 0) return VB_OK;
    add_cast_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                                         reinterpret_cast<uint2*>(out_bf16), n / 4);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_add3(const float* a, const void* b_bf16, const void* c_bf16, float* out, float* accum, int64_t n, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(a && b_bf16 && out && n >= 0 && n % 4 == 0, "add3: n=%lld must be a multiple of 4", (long long)n);
    if (n == 0) return VB_OK;
    add3_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const uint2*>(b_bf16),
                                                                     reinterpret_cast<const uint2*>(c_bf16), reinterpret_cast<float4*>(out),
                                                                     reinterpret_cast<float4*>(accum), n / 4);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_cast_rows_bf16(const float* src, int64_t ldsrc, void* dst_bf16, int64_t lddst, int64_t rows, int32_t cols, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(src && dst_bf16 && rows >= 0 && cols > 0 && ldsrc >= cols && lddst >= cols, "cast_rows: bad arguments");
    VB_REQUIRE(lddst % 2 == 0 && ((uintptr_t)dst_bf16 & 3) == 0, "cast_rows: destination pitch must be even and the base 4-byte aligned");
    if (rows == 0) return VB_OK;
    cast_rows_kernel<<<grid_for(rows * ((cols + 1) / 2), 256), 256, 0, as_stream(stream)>>>(src, ldsrc, reinterpret_cast<__nv_bfloat16*>(dst_bf16),
                                                                                           lddst, rows, cols);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_pos_embed_2d_fwd(const float* row_embed, const float* col_embed, float* pos, int32_t h, int32_t w, int32_t N, int32_t pf,
                                   void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(row_embed && col_embed && pos && h > 0 && w > 0 && N > 0 && pf > 0 && pf % 4 == 0, "pos_embed_2d_fwd: bad arguments");
    VB_REQUIRE(((uintptr_t)row_embed & 15) == 0 && ((uintptr_t)col_embed & 15) == 0 && ((uintptr_t)pos & 15) == 0, "pos_embed_2d_fwd: misaligned");
    pos_embed_2d_fwd_kernel<<<grid_for((long long)h * w * N * (pf / 2), 256), 256, 0, as_stream(stream)>>>(row_embed, col_embed, pos, h, w, N, pf);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_pos_embed_2d_bwd(const float* dpos, float* drow_accum, float* dcol_accum, int32_t h, int32_t w, int32_t N, int32_t pf,
                                   void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(dpos && drow_accum && dcol_accum && h > 0 && w > 0 && N > 0 && pf > 0, "pos_embed_2d_bwd: bad arguments");
    pos_embed_2d_bwd_kernel<<<dim3((pf + 127) / 128, h + w), 128, 0, as_stream(stream)>>>(dpos, drow_accum, dcol_accum, h, w, N, pf);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_vecmat_accum(const float* x, const float* W, int64_t ldw, float* y_accum, float* x_accum, int32_t K, int32_t N, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    VB_REQUIRE(x && W && y_accum && K > 0 && N > 0 && ldw >= N, "vecmat_accum: bad arguments");
    const int kpb = 32;
    vecmat_accum_kernel<<<dim3((N + 127) / 128, (K + kpb - 1) / kpb), 128, 0, as_stream(stream)>>>(x, W, ldw, y_accum, x_accum, K, N, kpb);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}
