// tcgen05 / TMEM attention forward for short sequences (S <= 256: every ViT / DeiT config), head_dim 64.
//
// One persistent CTA per SM walks (batch, head) pairs.  Per 128-query tile ("item"):
//   S = Q K^T      tcgen05.mma 128 x Npad x 64 (SS), fp32 accumulator in a 256-column TMEM slot
//   softmax        128 threads, ONE THREAD PER QUERY ROW (TMEM lane): two passes over the row with tcgen05.ld,
//                  row max / sum need no shuffles; P is written back as packed bf16 INTO THE SAME TMEM COLUMNS
//   O = P V        tcgen05.mma 128 x 64 x Npad with the A operand read from TMEM, V as an MN-major smem operand;
//                  the O accumulator aliases the dead tail of the S columns
// Two TMEM slots / two softmax warpgroups ping-pong so the tensor pipe works on one tile while the other is in softmax.
// Q, K, V arrive by TMA (3-D maps: column, token, batch -> rows >= S are zero-filled) in a 2-stage ring.
//
// Replaces F.scaled_dot_product_attention reached from nn.MultiheadAttention (vanilla_vit.py:77,
// torch/nn/functional.py:6676-6688).  The mma.sync kernels in attention.cu remain for S > 256 and for backward.
#include <cuda.h>
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"

namespace vb {

int make_tmap_3d(CUtensorMap* m, int dtype, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
                 uint64_t batch_stride_elems, uint32_t box0, uint32_t box1);   // gemm.cu

constexpr int ATC_THREADS = 384;
constexpr uint32_t ATC_ROWS = 256;                       // smem rows reserved per operand and stage
constexpr uint32_t ATC_OP_BYTES = ATC_ROWS * 128;        // 32 KB
constexpr uint32_t ATC_STAGE_BYTES = 3 * ATC_OP_BYTES;   // Q, K, V
constexpr uint32_t ATC_SMEM = 2 * ATC_STAGE_BYTES + 256 + 1024;

struct AttnTcArgs {
    int B, H, S, n_qt, npad, total_heads;
    float scale_log2;
    __nv_bfloat16* o;
    long long ldo, tok_stride, batch_stride;
    float* lse;
    const uint8_t* kpm;
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnTcArgs args) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * ATC_STAGE_BYTES);
    uint64_t* kv_full = bars;          // [2]
    uint64_t* kv_empty = bars + 2;     // [2]
    uint64_t* s_full = bars + 4;       // [2] per slot
    uint64_t* p_full = bars + 6;       // [2]
    uint64_t* o_full = bars + 8;       // [2]
    uint64_t* slot_free = bars + 10;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

    const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = args.S, npad = args.npad, n_qt = args.n_qt;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);      // one arrival per softmax warp
            mbar_init(&o_full[i], 1);
            mbar_init(&slot_free[i], 4);
        }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_ptr_smem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t o_off = npad >> 1;   // O accumulator aliases the S columns right after the packed P

    if (warp_idx == 0) {
        // ===================================== TMA loader =====================================
        if (lane == 0) {
            int hc = 0;
            for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
                const int st = hc & 1, b = head / args.H, h = head - b * args.H;
                mbar_wait(&kv_empty[st], ((hc >> 1) & 1) ^ 1);
                uint8_t* sq = smem + st * ATC_STAGE_BYTES;
                mbar_arrive_expect_tx(&kv_full[st], n_qt * 128 * 128 + 2 * npad * 128);
                for (int qt = 0; qt < n_qt; ++qt) tma_load_3d(sq + qt * 128 * 128, &tmQ, &kv_full[st], h * 64, qt * 128, b);
                tma_load_3d(sq + ATC_OP_BYTES, &tmK, &kv_full[st], h * 64, 0, b);
                tma_load_3d(sq + 2 * ATC_OP_BYTES, &tmV, &kv_full[st], h * 64, 0, b);
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer ======================================
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(128, npad, 0, 0);
            const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
            constexpr uint64_t kdesc = umma_smem_desc_base(0, 1024);          // K-major SW128 (Q, K)
            const uint64_t vdesc = umma_smem_desc_base(npad * 128, 1024);     // MN-major SW128 (V): 8-key groups 1024 B apart
            int j = 0, hc = 0;
            bool have_prev = false;
            int p_slot = 0, p_sph = 0, p_stage = 0, p_last = 0;
            auto issue_o = [&]() {
                mbar_wait(&p_full[p_slot], p_sph);
                tcgen05_fence_after();
                const uint32_t sv = smem_u32(smem + p_stage * ATC_STAGE_BYTES + 2 * ATC_OP_BYTES);
                const uint32_t t_slot = tmem_base + p_slot * 256;
                for (int k = 0; k < npad / 16; ++k)
                    umma_bf16_ts(t_slot + o_off, t_slot + k * 8, umma_smem_desc(vdesc, sv + k * 2048), idesc_o, k > 0 ? 1u : 0u);
                umma_commit(&o_full[p_slot]);
                if (p_last) umma_commit(&kv_empty[p_stage]);
            };
            for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
                const int st = hc & 1;
                mbar_wait(&kv_full[st], (hc >> 1) & 1);
                tcgen05_fence_after();
                const uint32_t sq = smem_u32(smem + st * ATC_STAGE_BYTES);
                const uint32_t sk = sq + ATC_OP_BYTES;
                for (int qt = 0; qt < n_qt; ++qt, ++j) {
                    const int slot = j & 1, sph = (j >> 1) & 1;
                    mbar_wait(&slot_free[slot], sph ^ 1);
                    tcgen05_fence_after();
                    const uint32_t t_slot = tmem_base + slot * 256;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(t_slot, umma_smem_desc(kdesc, sq + qt * 128 * 128 + k * 32), umma_smem_desc(kdesc, sk + k * 32), idesc_s,
                                     k > 0 ? 1u : 0u);
                    umma_commit(&s_full[slot]);
                    if (have_prev) issue_o();
                    p_slot = slot; p_sph = sph; p_stage = st; p_last = (qt == n_qt - 1);
                    have_prev = true;
                }
            }
            if (have_prev) issue_o();
        }
    } else if (warp_idx >= 4) {
        // ===================================== softmax + epilogue ===============================
        const int grp = (warp_idx - 4) >> 2;            // this warpgroup serves TMEM slot `grp`
        const uint32_t quad = warp_idx & 3;
        const uint32_t t_lane = tmem_base + ((quad * 32) << 16) + grp * 256;
        const int row_in_tile = quad * 32 + lane;
        const float c = args.scale_log2;
        int j = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x) {
            const int b = head / args.H, h = head - b * args.H;
            for (int qt = 0; qt < n_qt; ++qt, ++j) {
                if ((j & 1) != grp) continue;
                const int ph = (j >> 1) & 1;
                const uint8_t* kpm = args.kpm ? args.kpm + (long long)b * S : nullptr;
                mbar_wait(&s_full[grp], ph);
                tcgen05_fence_after();
                // ---- pass 1: row maximum (raw scores) ----
                float mx = -INFINITY;
                for (int c0 = 0; c0 < npad; c0 += 32) {
                    uint32_t r[32];
                    if (c0 + 32 <= npad) {
                        tmem_ld_32x32b_x32(t_lane + c0, r);
                    } else {
                        uint32_t(&lo)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
                        tmem_ld_32x32b_x16(t_lane + c0, lo);
#pragma unroll
                        for (int i = 16; i < 32; ++i) r[i] = 0xff800000u;   // -inf
                    }
                    tmem_ld_wait();
                    if (c0 + 32 > S || kpm) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int key = c0 + i;
                            const bool dead = key >= S || (kpm && kpm[min(key, S - 1)] != 0);
                            if (dead) r[i] = 0xff800000u;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                }
                const float msafe = (mx == -INFINITY) ? 0.f : mx;
                const float nm = -msafe * c;
                // ---- pass 2: P = exp2(s*c - m*c) -> packed bf16 written over the S columns; row sum ----
                float l = 0.f;
                for (int c0 = 0; c0 < npad; c0 += 32) {
                    uint32_t r[32];
                    const bool full = c0 + 32 <= npad;
                    if (full) {
                        tmem_ld_32x32b_x32(t_lane + c0, r);
                    } else {
                        uint32_t(&lo)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
                        tmem_ld_32x32b_x16(t_lane + c0, lo);
#pragma unroll
                        for (int i = 16; i < 32; ++i) r[i] = 0xff800000u;
                    }
                    tmem_ld_wait();
                    if (c0 + 32 > S || kpm) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int key = c0 + i;
                            const bool dead = key >= S || (kpm && kpm[min(key, S - 1)] != 0);
                            if (dead) r[i] = 0xff800000u;
                        }
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = ex2f(fmaf(__uint_as_float(r[2 * i]), c, nm));
                        const float p1 = ex2f(fmaf(__uint_as_float(r[2 * i + 1]), c, nm));
                        l += p0 + p1;
                        pk[i] = pack_bf16x2(p0, p1);
                    }
                    {
                        uint32_t(&lo)[8] = *reinterpret_cast<uint32_t(*)[8]>(&pk[0]);
                        tmem_st_32x32b_x8(t_lane + (c0 >> 1), lo);
                        if (full) {
                            uint32_t(&hi)[8] = *reinterpret_cast<uint32_t(*)[8]>(&pk[8]);
                            tmem_st_32x32b_x8(t_lane + (c0 >> 1) + 8, hi);
                        }
                    }
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[grp]);
                // ---- epilogue: O / l -> bf16 rows, lse ----
                mbar_wait(&o_full[grp], ph);
                tcgen05_fence_after();
                const int q = qt * 128 + row_in_tile;
                const float inv = l > 0.f ? 1.f / l : 0.f;
                __nv_bfloat16* orow = args.o + ((long long)b * args.batch_stride + (long long)min(q, S - 1) * args.tok_stride) * args.ldo + h * 64;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_lane + o_off + half * 32, r);
                    tmem_ld_wait();
                    if (q < S) {
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) {
                            uint4 w;
                            w.x = pack_bf16x2(__uint_as_float(r[v4 * 8 + 0]) * inv, __uint_as_float(r[v4 * 8 + 1]) * inv);
                            w.y = pack_bf16x2(__uint_as_float(r[v4 * 8 + 2]) * inv, __uint_as_float(r[v4 * 8 + 3]) * inv);
                            w.z = pack_bf16x2(__uint_as_float(r[v4 * 8 + 4]) * inv, __uint_as_float(r[v4 * 8 + 5]) * inv);
                            w.w = pack_bf16x2(__uint_as_float(r[v4 * 8 + 6]) * inv, __uint_as_float(r[v4 * 8 + 7]) * inv);
                            *reinterpret_cast<uint4*>(orow + half * 32 + v4 * 8) = w;
                        }
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&slot_free[grp]);
                if (args.lse && q < S) args.lse[((long long)b * args.H + h) * S + q] = mx * c + log2f(l);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
#endif
}

// Returns VB_OK if launched, 1 if this shape is not handled here (caller falls back to the mma.sync kernels).
int attention_fwd_tc(const VbAttnDesc* d, cudaStream_t stream) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("VITB200_ATTN_TC");
        enabled = (e && e[0] == '0') ? 0 : 1;
    }
    if (!enabled || d->S > 256 || d->head_dim != 64 || d->tok_stride != 1) return 1;   // batch-first layouts only
    const int S = d->S, npad = (S + 15) / 16 * 16, n_qt = (S + 127) / 128;
    CUtensorMap tq, tk, tv;
    const uint64_t cols = (uint64_t)d->H * 64;
    int rc;
    // 3-D maps (column, token, batch): token pitch = tok_stride * ld, batch pitch = batch_stride * ld
    if ((rc = make_tmap_3d(&tq, VB_BF16, d->q, cols, S, d->B, d->tok_stride * d->ldq, d->batch_stride * d->ldq, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tk, VB_BF16, d->k, cols, S, d->B, d->tok_stride * d->ldk, d->batch_stride * d->ldk, 64, npad))) return rc;
    if ((rc = make_tmap_3d(&tv, VB_BF16, d->v, cols, S, d->B, d->tok_stride * d->ldv, d->batch_stride * d->ldv, 64, npad))) return rc;
    AttnTcArgs a{};
    a.B = d->B; a.H = d->H; a.S = S; a.n_qt = n_qt; a.npad = npad; a.total_heads = d->B * d->H;
    a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.o = (__nv_bfloat16*)d->o; a.ldo = d->ldo; a.tok_stride = d->tok_stride; a.batch_stride = d->batch_stride;
    a.lse = d->lse; a.kpm = d->key_padding_mask;
    static DeviceOnce configured;
    if (!configured.is_set()) {
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
        configured.set();
    }
    int grid = num_sms();
    if (grid > a.total_heads) grid = a.total_heads;
    attn_fwd_tc_kernel<<<grid, ATC_THREADS, ATC_SMEM, stream>>>(tq, tk, tv, a);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

}  // namespace vb

// Debug hook (not part of the documented ABI surface used by the engine): device buffer of >= 64*16 int64 that receives
// cycle stamps from CTA 0 of the tcgen05 attention kernels (attention_fwd_tc.cu, attention_bwd_tc.cu); NULL disables.
namespace vb { void attention_bwd_tc5_set_debug(long long* p); void attention_fwd_tc3_set_debug(long long* p); void attention_fwd_tc4_set_debug(long long* p); }
extern "C" VB_API int vb_debug_set_attn_timeline(void* device_buffer) {
    vb::attention_bwd_tc5_set_debug(reinterpret_cast<long long*>(device_buffer));
    vb::attention_fwd_tc3_set_debug(reinterpret_cast<long long*>(device_buffer));
    vb::attention_fwd_tc4_set_debug(reinterpret_cast<long long*>(device_buffer));
    return VB_OK;
}
