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
    static bool configured = false;
    if (!configured) {
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
        configured = true;
    }
    int grid = num_sms();
    if (grid > a.total_heads) grid = a.total_heads;
    attn_fwd_tc_kernel<<<grid, ATC_THREADS, ATC_SMEM, stream>>>(tq, tk, tv, a);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

// ==========================================================================================================
// Backward (S <= 256).  Per head, with Q, K, V, dO resident in shared memory (one TMA load each), four tile passes
// through ONE 512-column TMEM slot (scores at columns [0,256), their gradient at [256,512)):
//   A-type item (rows = 128 keys kt):  S^T = K_kt Q^T, dP^T = V_kt dO^T  ->  P^T = exp2(S^T c - lse[q]),
//        dS^T = P^T o (dP^T - delta[q])  (both re-packed to bf16 in place)  ->  dV_kt = P^T dO, dK_kt = dS^T Q / 8
//   B-type item (rows = 128 queries qt): S = Q_qt K^T, dP = dO_qt V^T -> dS = P o (dP - delta[row]) -> dQ_qt = dS K / 8
// All second-stage products read their A operand straight from TMEM; P / dS never touch shared or global memory.
// Element-wise work is done by 384 threads, three per tile row (each owns a third of the columns, in 16-column groups).
// delta = rowsum(dO o O) comes from attn_delta_kernel (attention.cu).
// ==========================================================================================================
constexpr int ATC_BWD_PARTS = 3;
constexpr int ATC_BWD_THREADS = 128 + ATC_BWD_PARTS * 128;
static long long* g_attn_dbg = nullptr;   // optional device buffer for in-kernel cycle stamps (tools/attn_timeline.py)

struct AttnBwdTcArgs {
    long long* dbg;
    int B, H, S, n_t, npad, ng, total_heads, stages;
    int gs[ATC_BWD_PARTS + 1];   // 16-column group ranges of the column parts
    uint32_t op_bytes, stage_bytes;
    float scale, scale_log2;
    const float* lse;
    const float* delta;
    __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv;
    long long lddq, lddk, lddv, batch_stride;
    const uint8_t* kpm;
};

__global__ void __launch_bounds__(ATC_BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const AttnBwdTcArgs args) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int S = args.S, npad = args.npad, n_t = args.n_t, ng = args.ng, stages = args.stages;
    const uint32_t op_bytes = args.op_bytes, stage_bytes = args.stage_bytes;
    // layout: [stage 0: Q K V dO][stage 1 ...][tail pad][lse/delta: stages x 2 x 256 floats][barriers]
    const uint32_t tail = (uint32_t)n_t * 128 * 128 > op_bytes ? (uint32_t)n_t * 128 * 128 - op_bytes : 0;
    float* stats = reinterpret_cast<float*>(smem + stages * stage_bytes + tail);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stats + stages * 512);
    uint64_t* kv_full = bars;         // [2] count 2: TMA thread (expect_tx) + stats loader
    uint64_t* kv_empty = bars + 2;    // [2]
    uint64_t* s_full = bars + 4;
    uint64_t* p_full = bars + 5;      // count 384
    uint64_t* o_full = bars + 6;
    uint64_t* tile_free = bars + 7;   // count 384
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

    const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 2); mbar_init(&kv_empty[i], 1); }
        mbar_init(s_full, 1); mbar_init(p_full, ATC_BWD_PARTS * 4); mbar_init(o_full, 1); mbar_init(tile_free, ATC_BWD_PARTS * 4);   // one arrival per warp
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_ptr_smem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int n_items = 2 * n_t;

    if (warp_idx == 0) {
        if (lane == 0) {   // ---------------- TMA loader ----------------
            int hc = 0;
            for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
                const int st = hc % stages, b = head / args.H, h = head - b * args.H;
                mbar_wait(&kv_empty[st], ((hc / stages) & 1) ^ 1);
                uint8_t* base = smem + st * stage_bytes;
                mbar_arrive_expect_tx(&kv_full[st], 4 * op_bytes);
                tma_load_3d(base, &tmQ, &kv_full[st], h * 64, 0, b);
                tma_load_3d(base + op_bytes, &tmK, &kv_full[st], h * 64, 0, b);
                tma_load_3d(base + 2 * op_bytes, &tmV, &kv_full[st], h * 64, 0, b);
                tma_load_3d(base + 3 * op_bytes, &tmdO, &kv_full[st], h * 64, 0, b);
            }
        }
    } else if (warp_idx == 3) {
        // ---------------- lse / delta loader (whole warp) ----------------
        int hc = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
            const int st = hc % stages;
            mbar_wait(&kv_empty[st], ((hc / stages) & 1) ^ 1);
            float* ls = stats + st * 512;
            float* ds = ls + 256;
            const float* gl = args.lse + (long long)head * S;
            const float* gd = args.delta + (long long)head * S;
            for (int i = lane; i < 256; i += 32) {
                ls[i] = i < S ? gl[i] : INFINITY;   // +inf => P = 0 for padded queries
                ds[i] = i < S ? gd[i] : 0.f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&kv_full[st]);
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {   // ---------------- MMA issuer ----------------
            const uint32_t idesc_s = umma_idesc_bf16(128, npad, 0, 0);
            const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
            constexpr uint64_t kdesc = umma_smem_desc_base(0, 1024);      // K-major SW128
            const uint64_t mdesc = umma_smem_desc_base(op_bytes, 1024);   // MN-major SW128 (64-wide atom, 8-row K groups)
            // packed bf16 group k of a row lives where its column part wrote it: part start (in fp32 columns) + 8 per group
            auto a_col = [&](int k) -> uint32_t {
                int p = 0;
#pragma unroll
                for (int i = 1; i < ATC_BWD_PARTS; ++i) p += (k >= args.gs[i]) ? 1 : 0;
                return (uint32_t)args.gs[p] * 16 + (uint32_t)(k - args.gs[p]) * 8;
            };
            int hc = 0, it = 0;
            for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
                const int st = hc % stages;
                mbar_wait(&kv_full[st], (hc / stages) & 1);
                tcgen05_fence_after();
                const uint32_t sQ = smem_u32(smem + st * stage_bytes), sK = sQ + op_bytes, sV = sK + op_bytes, sdO = sV + op_bytes;
                for (int item = 0; item < n_items; ++item, ++it) {
                    const bool typeA = item < n_t;
                    const int t = typeA ? item : item - n_t;
                    mbar_wait(tile_free, (it & 1) ^ 1);
                    tcgen05_fence_after();
                    if (args.dbg && blockIdx.x == 0 && it < 64) args.dbg[it * 16 + 0] = clock64();
                    // scores and their gradient: rows = keys (A) or queries (B)
                    const uint32_t a0 = (typeA ? sK : sQ) + t * 128 * 128, b0 = typeA ? sQ : sK;
                    const uint32_t a1 = (typeA ? sV : sdO) + t * 128 * 128, b1 = typeA ? sdO : sV;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base, umma_smem_desc(kdesc, a0 + k * 32), umma_smem_desc(kdesc, b0 + k * 32), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base + 256, umma_smem_desc(kdesc, a1 + k * 32), umma_smem_desc(kdesc, b1 + k * 32), idesc_s, k > 0 ? 1u : 0u);
                    umma_commit(s_full);
                    if (args.dbg && blockIdx.x == 0 && it < 64) args.dbg[it * 16 + 1] = clock64();
                    mbar_wait(p_full, it & 1);
                    tcgen05_fence_after();
                    if (args.dbg && blockIdx.x == 0 && it < 64) args.dbg[it * 16 + 2] = clock64();
                    if (typeA) {
                        for (int k = 0; k < ng; ++k)   // dV = P^T dO
                            umma_bf16_ts(tmem_base + 192, tmem_base + a_col(k), umma_smem_desc(mdesc, sdO + k * 2048), idesc_o, k > 0 ? 1u : 0u);
                        for (int k = 0; k < ng; ++k)   // dK = dS^T Q
                            umma_bf16_ts(tmem_base + 256 + 192, tmem_base + 256 + a_col(k), umma_smem_desc(mdesc, sQ + k * 2048), idesc_o, k > 0 ? 1u : 0u);
                    } else {
                        for (int k = 0; k < ng; ++k)   // dQ = dS K
                            umma_bf16_ts(tmem_base + 192, tmem_base + 256 + a_col(k), umma_smem_desc(mdesc, sK + k * 2048), idesc_o, k > 0 ? 1u : 0u);
                    }
                    umma_commit(o_full);
                    if (args.dbg && blockIdx.x == 0 && it < 64) args.dbg[it * 16 + 3] = clock64();
                }
                umma_commit(&kv_empty[st]);
            }
        }
    } else if (warp_idx >= 4) {
        // ---------------- element-wise + read-out: 384 threads, three per tile row ----------------
        const uint32_t quad = warp_idx & 3, part = (warp_idx - 4) >> 2;
        const uint32_t t_lane = tmem_base + ((quad * 32) << 16);
        const int row_in_tile = quad * 32 + lane;
        const int gb = args.gs[part], ge = args.gs[part + 1];          // my 16-column groups
        const uint32_t pk_base = (uint32_t)gb * 16;                    // my packed outputs go at pk_base + (g - gb) * 8
        const float c = args.scale_log2;
        int hc = 0, it = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
            const int st = hc % stages, b = head / args.H, h = head - b * args.H;
            const float* ls = stats + st * 512;
            const float* dls = ls + 256;
            const uint8_t* kpm = args.kpm ? args.kpm + (long long)b * S : nullptr;
            for (int item = 0; item < n_items; ++item, ++it) {
                const bool typeA = item < n_t;
                const int t = typeA ? item : item - n_t;
                const int row = t * 128 + row_in_tile;   // key (A) or query (B) index of this thread's row
                mbar_wait(s_full, it & 1);
                tcgen05_fence_after();
                const bool dbg_on = args.dbg && blockIdx.x == 0 && it < 64 && warp_idx == 4 && lane == 0;
                if (dbg_on) args.dbg[it * 16 + 4] = clock64();
                if (typeA) {
                    const bool dead_row = row >= S || (kpm && kpm[min(row, S - 1)] != 0);
                    const float kill = dead_row ? 0.f : 1.f;
                    for (int g = gb; g < ge; ++g) {
                        uint32_t sv[16], dv[16];
                        tmem_ld_32x32b_x16(t_lane + g * 16, sv);
                        tmem_ld_32x32b_x16(t_lane + 256 + g * 16, dv);
                        tmem_ld_wait();
                        uint32_t pp[8], pd[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 l2 = *reinterpret_cast<const float2*>(ls + g * 16 + 2 * i);
                            const float2 d2 = *reinterpret_cast<const float2*>(dls + g * 16 + 2 * i);
                            const float p0 = ex2f(fmaf(__uint_as_float(sv[2 * i]), c, -l2.x)) * kill;
                            const float p1 = ex2f(fmaf(__uint_as_float(sv[2 * i + 1]), c, -l2.y)) * kill;
                            pp[i] = pack_bf16x2(p0, p1);
                            pd[i] = pack_bf16x2(p0 * (__uint_as_float(dv[2 * i]) - d2.x), p1 * (__uint_as_float(dv[2 * i + 1]) - d2.y));
                        }
                        tmem_st_32x32b_x8(t_lane + pk_base + (g - gb) * 8, pp);
                        tmem_st_32x32b_x8(t_lane + 256 + pk_base + (g - gb) * 8, pd);
                    }
                } else {
                    const float nl = -ls[min(row, 255)], dl = dls[min(row, 255)];   // rows >= S: lse = +inf => P = 0
                    for (int g = gb; g < ge; ++g) {
                        uint32_t sv[16], dv[16];
                        tmem_ld_32x32b_x16(t_lane + g * 16, sv);
                        tmem_ld_32x32b_x16(t_lane + 256 + g * 16, dv);
                        tmem_ld_wait();
                        uint32_t pd[8];
                        const bool need_mask = (g * 16 + 16 > S) || kpm;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float p0 = ex2f(fmaf(__uint_as_float(sv[2 * i]), c, nl));
                            float p1 = ex2f(fmaf(__uint_as_float(sv[2 * i + 1]), c, nl));
                            if (need_mask) {
                                const int k0 = g * 16 + 2 * i;
                                if (k0 >= S || (kpm && kpm[min(k0, S - 1)] != 0)) p0 = 0.f;
                                if (k0 + 1 >= S || (kpm && kpm[min(k0 + 1, S - 1)] != 0)) p1 = 0.f;
                            }
                            pd[i] = pack_bf16x2(p0 * (__uint_as_float(dv[2 * i]) - dl), p1 * (__uint_as_float(dv[2 * i + 1]) - dl));
                        }
                        tmem_st_32x32b_x8(t_lane + 256 + pk_base + (g - gb) * 8, pd);
                    }
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (dbg_on) args.dbg[it * 16 + 5] = clock64();
                if (lane == 0) mbar_arrive(p_full);
                // ---- read-out: 32-column slices of the 128 x 64 results are dealt out to the column parts ----
                mbar_wait(o_full, it & 1);
                tcgen05_fence_after();
                if (dbg_on) args.dbg[it * 16 + 6] = clock64();
                const long long grow = (long long)b * args.batch_stride + min(row, S - 1);
                auto store32 = [&](uint32_t taddr, __nv_bfloat16* dst, float mul) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr, r);
                    tmem_ld_wait();
                    if (row < S) {
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) {
                            uint4 w;
                            w.x = pack_bf16x2(__uint_as_float(r[v4 * 8 + 0]) * mul, __uint_as_float(r[v4 * 8 + 1]) * mul);
                            w.y = pack_bf16x2(__uint_as_float(r[v4 * 8 + 2]) * mul, __uint_as_float(r[v4 * 8 + 3]) * mul);
                            w.z = pack_bf16x2(__uint_as_float(r[v4 * 8 + 4]) * mul, __uint_as_float(r[v4 * 8 + 5]) * mul);
                            w.w = pack_bf16x2(__uint_as_float(r[v4 * 8 + 6]) * mul, __uint_as_float(r[v4 * 8 + 7]) * mul);
                            *reinterpret_cast<uint4*>(dst + v4 * 8) = w;
                        }
                    }
                };
                if (typeA) {   // slices: dV lo, dV hi, dK lo -> parts 0, 1, 2; dK hi -> part 0
                    if (part == 0) store32(t_lane + 192, args.dv + grow * args.lddv + h * 64, 1.0f);
                    if (part == 1) store32(t_lane + 192 + 32, args.dv + grow * args.lddv + h * 64 + 32, 1.0f);
                    if (part == 2) store32(t_lane + 256 + 192, args.dk + grow * args.lddk + h * 64, args.scale);
                    if (part == 0) store32(t_lane + 256 + 192 + 32, args.dk + grow * args.lddk + h * 64 + 32, args.scale);
                } else {
                    if (part < 2) store32(t_lane + 192 + part * 32, args.dq + grow * args.lddq + h * 64 + part * 32, args.scale);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (dbg_on) args.dbg[it * 16 + 7] = clock64();
                if (lane == 0) mbar_arrive(tile_free);
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

// Returns VB_OK if launched, 1 if this shape is not handled here.  `delta` must already hold rowsum(dO o O).
int attention_bwd_tc(const VbAttnDesc* d, cudaStream_t stream) {
    if (d->S > 208 || d->head_dim != 64 || d->tok_stride != 1) return 1;   // packed operands must end below TMEM column 192
    const int S = d->S, npad = (S + 15) / 16 * 16, n_t = (S + 127) / 128;
    AttnBwdTcArgs a{};
    a.B = d->B; a.H = d->H; a.S = S; a.n_t = n_t; a.npad = npad; a.ng = npad / 16;
    {
        int acc = 0;
        for (int p = 0; p < ATC_BWD_PARTS; ++p) { a.gs[p] = acc; acc += a.ng / ATC_BWD_PARTS + (p < a.ng % ATC_BWD_PARTS ? 1 : 0); }
        a.gs[ATC_BWD_PARTS] = acc;
    }
    a.total_heads = d->B * d->H;
    a.op_bytes = (uint32_t)npad * 128;
    a.stage_bytes = 4 * a.op_bytes;
    const uint32_t tail = (uint32_t)n_t * 128 * 128 > a.op_bytes ? (uint32_t)n_t * 128 * 128 - a.op_bytes : 0;
    const uint32_t fixed = tail + 2 * 512 * sizeof(float) + 256 + 1024;
    a.stages = (232448 - fixed) / a.stage_bytes >= 2 ? 2 : 1;
    const uint32_t smem = a.stages * a.stage_bytes + fixed;
    if (smem > 232448) return 1;
    a.scale = 0.125f; a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.lse = d->lse; a.delta = d->delta;
    a.dq = (__nv_bfloat16*)d->dq; a.dk = (__nv_bfloat16*)d->dk; a.dv = (__nv_bfloat16*)d->dv;
    a.lddq = d->lddq; a.lddk = d->lddk; a.lddv = d->lddv; a.batch_stride = d->batch_stride;
    a.kpm = d->key_padding_mask;
    a.dbg = g_attn_dbg;
    CUtensorMap tq, tk, tv, tdo;
    const uint64_t cols = (uint64_t)d->H * 64;
    int rc;
    if ((rc = make_tmap_3d(&tq, VB_BF16, d->q, cols, S, d->B, d->ldq, d->batch_stride * d->ldq, 64, npad))) return rc;
    if ((rc = make_tmap_3d(&tk, VB_BF16, d->k, cols, S, d->B, d->ldk, d->batch_stride * d->ldk, 64, npad))) return rc;
    if ((rc = make_tmap_3d(&tv, VB_BF16, d->v, cols, S, d->B, d->ldv, d->batch_stride * d->ldv, 64, npad))) return rc;
    if ((rc = make_tmap_3d(&tdo, VB_BF16, d->dout, cols, S, d->B, d->lddo, d->batch_stride * d->lddo, 64, npad))) return rc;
    static uint32_t configured = 0;
    if (configured < smem) {
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    int grid = num_sms();
    if (grid > a.total_heads) grid = a.total_heads;
    attn_bwd_tc_kernel<<<grid, ATC_BWD_THREADS, smem, stream>>>(tq, tk, tv, tdo, a);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

}  // namespace vb

// Debug hook (not part of the documented ABI surface used by the engine): device buffer of >= 64*16 int64 that receives
// cycle stamps from CTA 0 of the tcgen05 attention backward; pass NULL to disable.
namespace vb { void attention_bwd_tc5_set_debug(long long* p); }
extern "C" VB_API int vb_debug_set_attn_timeline(void* device_buffer) {
    vb::g_attn_dbg = reinterpret_cast<long long*>(device_buffer);
    vb::attention_bwd_tc5_set_debug(reinterpret_cast<long long*>(device_buffer));
    return VB_OK;
}
