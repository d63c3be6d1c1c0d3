// tcgen05 / TMEM attention forward, one-block shapes with 193..208 tokens (ViT-B/L and DeiT at 224^2: S = 197 / 198), head_dim 64:
// the score row lives in REGISTERS.
//
// Same skeleton as attention_fwd_tc.cu (persistent CTA per SM over (batch, head); Q / K / V of a head by TMA in a 2-stage ring;
// S = Q K^T of tile j + 1 is computed into the second 208-column TMEM slot while tile j is in its softmax; O = P V with P read from
// TMEM; read-out warps normalise O and hand bf16 tiles to a TMA-store warp) — what differs is the softmax:
//   * ONE group of 16 warps, FOUR THREADS PER QUERY ROW, each owning 52 consecutive key columns.  A thread loads its 52 scores
//     from TMEM ONCE (x32 + x16 + x4), keeps them in registers, takes its partial row maximum, meets the three other parts of its
//     TMEM lane quadrant at a named barrier (maxima through shared memory) and exponentiates straight out of the registers.
//     The two-pass kernel read every score twice with seven latency-exposed 16-column loads per pass and had its two ping-pong
//     groups fight over the MUFU in their exp passes: 3.35 k cycles per tile against a 1.66 k MUFU floor.
//   * because every thread of a quadrant has its scores in registers before anyone stores, P can be packed CONTIGUOUSLY over the
//     slot's first 104 columns (part p writes packed columns [26 p, 26 p + 26)): the A operand of P V advances 8 columns per
//     16-key step.
// Replaces F.scaled_dot_product_attention reached from nn.MultiheadAttention (vanilla_vit.py:77, torch/nn/functional.py:6676-6688).
#include <cuda.h>
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"
#include "attn_tc_common.cuh"
#include "pdl.cuh"

namespace vb {

int make_tmap_3d(CUtensorMap* m, int dtype, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
                 uint64_t batch_stride_elems, uint32_t box0, uint32_t box1);   // gemm.cu

namespace fwd4 {
using namespace atc;

constexpr int kSmWarps = 16;                     // softmax warps: 4 TMEM lane quadrants x 4 column parts
constexpr int kRoWarps = 4;                      // read-out warps (one per quadrant)
constexpr int kThreads = 128 + (kSmWarps + kRoWarps) * 32;   // warps 0-3: TMA loader, MMA issuer, TMEM alloc + store warp, idle
constexpr uint32_t kN = 208;                     // padded keys / queries (13 groups of 16)
constexpr uint32_t kPartCols = 52;               // score columns per thread
constexpr uint32_t kOpBytes = kN * 128;          // Q, K or V of one head
constexpr uint32_t kStageBytes = 3 * kOpBytes;
constexpr uint32_t kOutOff = 2 * kStageBytes;    // 2 output tiles of 128 rows x 128 B
constexpr uint32_t kXmOff = kOutOff + 2 * 16384; // partial maxima  [4 tile slots][4 parts][128 rows]
constexpr uint32_t kXsOff = kXmOff + 4 * 4 * 128 * 4;
constexpr uint32_t kBarOff = kXsOff + 4 * 4 * 128 * 4;
constexpr uint32_t kSmemBytes = kBarOff + 256 + 1024;
static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
constexpr uint32_t kColO = 416;
// In-kernel cycle stamps (tools/attn_timeline_fwd.py) are compiled in only with -DVB_ATTN_DBG: even predicated off they cost ~5 us per
// launch in this kernel's softmax loop.
#ifdef VB_ATTN_DBG
#define VB_DBG_STAMP(slot) if (dbg_on) args.dbg[j * 16 + (slot)] = clock64()
#else
#define VB_DBG_STAMP(slot) do { } while (0)
#endif

struct Args {
    int B, H, S, n_qt, total_heads;
    float scale_log2;
    float* lse;
    long long* dbg;   // optional in-kernel cycle stamps (tools/attn_timeline_fwd.py)
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Args args) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
    uint64_t* kv_full = bars;          // [2]
    uint64_t* kv_empty = bars + 2;     // [2] tcgen05.commit after the head's last P V product
    uint64_t* s_full = bars + 4;       // [2] per TMEM slot
    uint64_t* p_full = bars + 6;       // [2] count kSmWarps: P of the slot is packed
    uint64_t* o_full = bars + 8;
    uint64_t* o_free = bars + 9;       // count kRoWarps: the O accumulator has been read out
    uint64_t* out_ready = bars + 10;   // [2] count kRoWarps: output tile staged
    uint64_t* out_free = bars + 12;    // [2] count 1: the TMA store has finished reading the tile
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

    const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = args.S, n_qt = args.n_qt;
    pdl_launch_dependents();

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&p_full[i], kSmWarps);
            mbar_init(&out_ready[i], kRoWarps); mbar_init(&out_free[i], 1);
        }
        mbar_init(o_full, 1); mbar_init(o_free, kRoWarps);
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_ptr_smem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();

    if (warp_idx == 0) {
        if (lane == 0) {   // ---------------- TMA loader ----------------
            int hc = 0;
            for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
                const int st = hc & 1, b = head / args.H, h = head - b * args.H;
                mbar_wait(&kv_empty[st], ((hc >> 1) & 1) ^ 1);
                uint8_t* sq = smem + st * kStageBytes;
                mbar_arrive_expect_tx(&kv_full[st], 3 * kOpBytes);
                tma_load_3d(sq, &tmQ, &kv_full[st], h * 64, 0, b);
                tma_load_3d(sq + kOpBytes, &tmK, &kv_full[st], h * 64, 0, b);
                tma_load_3d(sq + 2 * kOpBytes, &tmV, &kv_full[st], h * 64, 0, b);
            }
        }
    } else if (warp_idx == 2) {
        // ---------------- store warp ----------------
        int j = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x) {
            const int b = head / args.H, h = head - b * args.H;
            for (int qt = 0; qt < n_qt; ++qt, ++j) {
                const int buf = j & 1;
                mbar_wait(&out_ready[buf], (j >> 1) & 1);
                if (lane == 0) {
                    tma_store_3d(&tmO, smem + kOutOff + buf * 16384, h * 64, qt * 128, b);
                    tma_store_commit();
                    tma_store_wait_read<0>();
                    mbar_arrive(&out_free[buf]);
                }
                __syncwarp();
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    } else if (warp_idx == 1) {
        // ---------------- MMA issuer (uniform control flow; one elected lane issues) ----------------
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, kN, 0, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
        constexpr uint64_t kdesc = umma_smem_desc_base(0, 1024);            // K-major SW128 (Q, K)
        constexpr uint64_t vdesc = umma_smem_desc_base(kOpBytes, 1024);     // MN-major SW128 (V): 8-key groups 1024 B apart
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const int total_tiles = ((args.total_heads - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * n_qt;
        auto issue_s = [&](int jj) {
            const int hc = jj / n_qt, qt = jj - hc * n_qt, st = hc & 1;
            if (qt == 0) {
                mbar_wait(&kv_full[st], (hc >> 1) & 1);
                tcgen05_fence_after();
            }
            const uint32_t sq = smem_u32(smem + st * kStageBytes) + qt * 16384, sk = smem_u32(smem + st * kStageBytes + kOpBytes);
            const uint32_t d = tb + (jj & 1) * kN;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(d, umma_smem_desc(kdesc, sq + k * 32), umma_smem_desc(kdesc, sk + k * 32), idesc_s, k > 0 ? 1u : 0u);
                umma_commit(&s_full[jj & 1]);
            }
            __syncwarp();
        };
        // Program order S(0) S(1) PV(0) S(2) PV(1) S(3) ...: S(j + 2) re-uses the slot whose P was consumed by PV(j) (in-order tensor pipe).
        if (total_tiles > 0) issue_s(0);
        if (total_tiles > 1) issue_s(1);
        for (int j = 0; j < total_tiles; ++j) {
            const int hc = j / n_qt, qt = j - hc * n_qt, st = hc & 1, slot = j & 1, par = (j >> 1) & 1;
            const uint64_t bd = umma_smem_desc(vdesc, smem_u32(smem + st * kStageBytes + 2 * kOpBytes));
            const uint32_t a0 = tb + slot * kN;
            mbar_wait(&p_full[slot], par);
            mbar_wait(o_free, (j & 1) ^ 1);
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 13; ++k)   // packed P: 16 keys = 8 TMEM columns per step
                    umma_bf16_ts(tb + kColO, a0 + 8 * k, bd + (uint64_t)(k * 128), idesc_o, k > 0 ? 1u : 0u);
                umma_commit(o_full);
                if (qt == n_qt - 1) umma_commit(&kv_empty[st]);
            }
            __syncwarp();
            if (j + 2 < total_tiles) issue_s(j + 2);
        }
    } else if (warp_idx >= 4 && warp_idx < 4 + kSmWarps) {
        // ---------------- softmax: 52 scores per thread, held in registers ----------------
        const uint32_t quad = warp_idx & 3, part = (warp_idx - 4) >> 2;
        const uint32_t t_lane = tmem_base + ((quad * 32) << 16);
        const int row_in_tile = quad * 32 + lane;
        const float c = args.scale_log2;
        const f2 c2 = f2_pack(c, c);
        const uint32_t xm_u32 = smem_u32(smem + kXmOff), xs_u32 = smem_u32(smem + kXsOff);
        const int col0 = (int)(part * kPartCols);
        const int total_tiles = ((args.total_heads - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * n_qt;
        for (int j = 0; j < total_tiles; ++j) {
            const uint32_t slot = j & 1;
            const uint32_t t_s = t_lane + slot * kN;
            const uint32_t xoff = ((j & 3) * 4 * 128 + row_in_tile) * 4;     // [tile slot][part][row]
            [[maybe_unused]] const bool dbg_on = args.dbg && blockIdx.x == 0 && j < 64 && warp_idx == 4 && lane == 0;
            VB_DBG_STAMP(0);
            mbar_wait(&s_full[slot], (j >> 1) & 1);
            tcgen05_fence_after();
            VB_DBG_STAMP(1);
            uint32_t r[52];
            {
                uint32_t a[32], b[16], d[4];
                tmem_ld_32x32b_x32(t_s + col0, a);
                tmem_ld_32x32b_x16(t_s + col0 + 32, b);
                tmem_ld_32x32b_x4(t_s + col0 + 48, d);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = a[i];
#pragma unroll
                for (int i = 0; i < 16; ++i) r[32 + i] = b[i];
#pragma unroll
                for (int i = 0; i < 4; ++i) r[48 + i] = d[i];
            }
            if (part == 3) {   // keys >= S are dead (S >= 193: only the last part has any)
#pragma unroll
                for (int i = 36; i < 52; ++i)
                    if (col0 + i >= S) r[i] = 0xff800000u;
            }
            float m0 = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), m1 = fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3]));
#pragma unroll
            for (int i = 4; i < 52; i += 4) {          // FMNMX3: two running maxima, two new elements each per instruction
                m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
            }
            float mx = fmaxf(m0, m1);
            VB_DBG_STAMP(2);
            sts32(xm_u32 + xoff + part * 512, mx);
            named_bar_sync(1 + quad, 128);           // the four parts of this lane quadrant: maxima exchanged, every score in registers
            mx = fmaxf(fmaxf(lds32(xm_u32 + xoff), lds32(xm_u32 + xoff + 512)), fmaxf(lds32(xm_u32 + xoff + 1024), lds32(xm_u32 + xoff + 1536)));
            const float nm = -mx * c;
            const f2 nm2 = f2_pack(nm, nm);
            VB_DBG_STAMP(3);
            f2 l2 = f2_pack(0.f, 0.f);
            const int live_pairs = (part == 3) ? (S - col0 + 1) >> 1 : 26;   // pairs with at least one key < S (warp-uniform)
#pragma unroll
            for (int i = 0; i < 26; ++i) {
                if (i >= 18 && i >= live_pairs) {     // nothing but padded keys: P = 0 without touching the MUFU
                    r[i] = 0u;
                    continue;
                }
                // (Tried and rejected: 2^x of every 3rd / 5th pair as a degree-3 polynomial on the FMA pipes, FlashAttention-4 style.  The
                // exp pass itself shrank from 2.0 k to 1.65 k cycles per tile, but the extra issue slots starved the MMA-issue, read-out
                // and store warps that share the four schedulers: 77.8 us -> 88 us (1 in 5) / 108 us (1 in 3) per layer.)
                float x0, x1;
                f2_unpack(f2_fma(f2_pack_u(r[2 * i], r[2 * i + 1]), c2, nm2), x0, x1);
                const float p0 = ex2f(x0), p1 = ex2f(x1);
                l2 = f2_add(l2, f2_pack(p0, p1));
                r[i] = pack2(p0, p1);                 // in place: element i is written after elements 2 i, 2 i + 1 were read
            }
            {   // packed P of this part: columns [26 part, 26 part + 26) of the slot
                const uint32_t t_p = t_s + part * 26;
                uint32_t a[16], b[8];
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = r[i];
#pragma unroll
                for (int i = 0; i < 8; ++i) b[i] = r[16 + i];
                tmem_st_32x32b_x16(t_p, a);
                tmem_st_32x32b_x8(t_p + 16, b);
                tmem_st_32x32b_x2(t_p + 24, r[24], r[25]);
            }
            float l0, l1;
            f2_unpack(l2, l0, l1);
            sts32(xs_u32 + xoff + part * 512, l0 + l1);
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[slot]);
            VB_DBG_STAMP(4);
        }
    } else if (warp_idx >= 4 + kSmWarps) {
        // ---------------- read-out: O / l -> bf16 tile in shared memory (TMA store by warp 2), lse ----------------
        const uint32_t quad = warp_idx & 3;
        const uint32_t t_lane = tmem_base + ((quad * 32) << 16);
        const int row_in_tile = quad * 32 + lane;
        const float c = args.scale_log2;
        const uint32_t swz = (uint32_t)(row_in_tile & 7);
        const uint32_t xm_u32 = smem_u32(smem + kXmOff), xs_u32 = smem_u32(smem + kXsOff);
        const uint32_t out_row = smem_u32(smem + kOutOff) + row_in_tile * 128;
        int j = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x) {
            for (int qt = 0; qt < n_qt; ++qt, ++j) {
                const uint32_t xoff = ((j & 3) * 4 * 128 + row_in_tile) * 4;
                const int buf = j & 1;
                mbar_wait(&out_free[buf], ((j >> 1) & 1) ^ 1);
                mbar_wait(o_full, j & 1);
                tcgen05_fence_after();
                uint32_t r0[32], r1[32];
                tmem_ld_32x32b_x32(t_lane + kColO, r0);
                tmem_ld_32x32b_x32(t_lane + kColO + 32, r1);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_free);
                const float tot = (lds32(xs_u32 + xoff) + lds32(xs_u32 + xoff + 512)) + (lds32(xs_u32 + xoff + 1024) + lds32(xs_u32 + xoff + 1536));
                const float mx = fmaxf(fmaxf(lds32(xm_u32 + xoff), lds32(xm_u32 + xoff + 512)), fmaxf(lds32(xm_u32 + xoff + 1024), lds32(xm_u32 + xoff + 1536)));
                const float inv = tot > 0.f ? 1.f / tot : 0.f;
                const uint32_t dst = out_row + buf * 16384;
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    const uint32_t w0 = pack2(__uint_as_float(r0[v4 * 8 + 0]) * inv, __uint_as_float(r0[v4 * 8 + 1]) * inv);
                    const uint32_t w1 = pack2(__uint_as_float(r0[v4 * 8 + 2]) * inv, __uint_as_float(r0[v4 * 8 + 3]) * inv);
                    const uint32_t w2 = pack2(__uint_as_float(r0[v4 * 8 + 4]) * inv, __uint_as_float(r0[v4 * 8 + 5]) * inv);
                    const uint32_t w3 = pack2(__uint_as_float(r0[v4 * 8 + 6]) * inv, __uint_as_float(r0[v4 * 8 + 7]) * inv);
                    sts128(dst + (((uint32_t)v4 ^ swz) << 4), w0, w1, w2, w3);
                }
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    const uint32_t w0 = pack2(__uint_as_float(r1[v4 * 8 + 0]) * inv, __uint_as_float(r1[v4 * 8 + 1]) * inv);
                    const uint32_t w1 = pack2(__uint_as_float(r1[v4 * 8 + 2]) * inv, __uint_as_float(r1[v4 * 8 + 3]) * inv);
                    const uint32_t w2 = pack2(__uint_as_float(r1[v4 * 8 + 4]) * inv, __uint_as_float(r1[v4 * 8 + 5]) * inv);
                    const uint32_t w3 = pack2(__uint_as_float(r1[v4 * 8 + 6]) * inv, __uint_as_float(r1[v4 * 8 + 7]) * inv);
                    sts128(dst + (((uint32_t)(4 + v4) ^ swz) << 4), w0, w1, w2, w3);
                }
                const int q = qt * 128 + row_in_tile;
                if (args.lse && q < S) args.lse[(long long)head * S + q] = mx * c + log2f(tot);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&out_ready[buf]);
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

}  // namespace fwd4

static long long* g_dbg4 = nullptr;
void attention_fwd_tc4_set_debug(long long* p) { g_dbg4 = p; }

// Returns VB_OK if launched, 1 if this shape is not handled here (caller uses the two-pass kernel of attention_fwd_tc.cu).
int attention_fwd_tc4(const VbAttnDesc* d, cudaStream_t stream) {
    using namespace fwd4;
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("VITB200_ATTN_TC4");
        enabled = (e && e[0] == '0') ? 0 : 1;
    }
    const bool cross = d->S_kv > 0 && d->S_kv != d->S;
    if (!enabled || cross || d->S > (int)kN || d->S <= 192 || d->head_dim != 64 || d->tok_stride != 1 || d->key_padding_mask != nullptr ||
        d->dropout_p > 0.f)
        return 1;
    const int S = d->S;
    Args a{};
    a.B = d->B; a.H = d->H; a.S = S; a.n_qt = (S + 127) / 128; a.total_heads = d->B * d->H;
    a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.lse = d->lse;
    a.dbg = g_dbg4;
    CUtensorMap tq, tk, tv, to;
    const uint64_t cols = (uint64_t)d->H * 64;
    int rc;
    if ((rc = make_tmap_3d(&tq, VB_BF16, d->q, cols, S, d->B, d->ldq, d->batch_stride * d->ldq, 64, kN))) return rc;
    if ((rc = make_tmap_3d(&tk, VB_BF16, d->k, cols, S, d->B, d->ldk, d->batch_stride * d->ldk, 64, kN))) return rc;
    if ((rc = make_tmap_3d(&tv, VB_BF16, d->v, cols, S, d->B, d->ldv, d->batch_stride * d->ldv, 64, kN))) return rc;
    if ((rc = make_tmap_3d(&to, VB_BF16, d->o, cols, S, d->B, d->ldo, d->batch_stride * d->ldo, 64, 128))) return rc;
    int grid = num_sms();
    if (grid > a.total_heads) grid = a.total_heads;
    static DeviceOnce configured;
    VB_ONCE_PER_DEVICE(configured, VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)));
    VB_CUDA_CHECK(launch_pdl(attn_fwd_tc4_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, tq, tk, tv, to, a));
    return VB_OK;
}

}  // namespace vb
