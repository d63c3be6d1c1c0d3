// tcgen05 / TMEM attention forward for short sequences (S <= 208: every ViT / DeiT config), head_dim 64.
//
// One persistent CTA per SM walks (batch, head) pairs; Q, K, V of a head arrive by TMA in a 2-stage ring.  Per 128-query tile:
//   S = Q K^T        tcgen05.mma 128 x npad x 64 (SS) into one of two 208-column TMEM slots (tile j+1 is computed while the
//                    softmax of tile j runs)
//   softmax          TWO WARP GROUPS of 256 threads ping-pong the tiles (group j & 1 owns tile j and TMEM slot j & 1), so the
//                    exp pass of one tile (MUFU) overlaps the row-max pass and the hand-offs of the other.  Inside a group TWO
//                    THREADS PER QUERY ROW: 16-key group g belongs to column part g % 2.  Pass 1: partial row maxima,
//                    exchanged through shared memory between the two warps of a TMEM lane quadrant (named barrier).
//                    Pass 2: P = exp2(s c - m c) packed to bf16 over the group's own fp32 columns [16g, 16g+8) (hazard-free
//                    for any group-to-warp assignment), partial row sums.
//   O = P V          tcgen05.mma 128 x 64 x npad with the A operand read from TMEM, V as an MN-major smem operand; issued in
//                    two parts (keys < 128 as soon as their groups are packed, the rest after the whole pass)
//   read-out         O / l -> bf16 tile in shared memory -> TMA store by a dedicated warp (rows >= S are clipped), lse
// GEN instantiations (any number of queries / keys, key-padding masks, sequence-first strides, cross-attention: the DETR encoder and
// decoder, transformer.py:213-226, 145-147): a work item is (batch, head, 128-query tile, key block of KB <= 208 keys) instead of a
// whole head.  Masked / out-of-range keys get an additive -inf from a per-stage bias row in shared memory (filled by the otherwise
// idle warp 3).  With more than one key block per head the item writes a PARTIAL result — O normalised by the block's own row sum
// and the block's log-sum-exp — and attn_merge_kernel combines the blocks (log-sum-exp weights); with a single key block the item
// writes the final O / lse directly.  No running maximum, no accumulator rescaling.
// Replaces F.scaled_dot_product_attention reached from nn.MultiheadAttention (vanilla_vit.py:77,
// torch/nn/functional.py:6676-6688) and the explicit-softmax path of transformer.py:219 (torch/nn/functional.py:6630-6666).
#include <cuda.h>
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"
#include "attn_tc_common.cuh"
#include "dropout.cuh"
#include "pdl.cuh"

// In-kernel cycle stamps (tools/attn_timeline*.py) are compiled in only with -DVB_ATTN_DBG (VITB200_NVCC_DEFS=-DVB_ATTN_DBG python -m
// vitb200.build): even predicated off they cost issue slots in the element-wise loops.
#ifdef VB_ATTN_DBG
#define VB_DBG(cond, slot) do { if (cond) slot = clock64(); } while (0)
#else
#define VB_DBG(cond, slot) do { (void)(cond); } while (0)
#endif

namespace vb {

int make_tmap_3d(CUtensorMap* m, int dtype, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
                 uint64_t batch_stride_elems, uint32_t box0, uint32_t box1);   // gemm.cu

namespace fwd3 {
using namespace atc;

constexpr int kEwWarps = 16;                     // two softmax groups x (4 lane quadrants x 2 column parts)
constexpr int kRoWarps = 4;                      // read-out warps (one per TMEM lane quadrant): O / l -> bf16 tile, lse
constexpr int kThreads = 128 + (kEwWarps + kRoWarps) * 32;   // warps 0-3: TMA loader, MMA issuer, TMEM alloc + store warp, idle
constexpr uint32_t kMaxQ = 208;
constexpr uint32_t kGenMaxKeys = 176;           // GEN: keys per block (three stages of {Q tile, K block, V block} must fit)
// Shared-memory layout.  One-block kernels: 2 stages of {Q, K, V} of a whole head (208 rows each).  GEN: 3 stages of {128-query
// tile, K block, V block of <= 176 keys} — an item is a single tile there, so the ring has to be one stage deeper to keep the
// loads of item j + 2 in flight while item j is still in its P V product.
template <bool GEN>
struct Lay {
    static constexpr uint32_t kStages = GEN ? 3 : 2;
    static constexpr uint32_t kQ = GEN ? 128 * 128 : kMaxQ * 128;            // bytes of the Q slot
    static constexpr uint32_t kKV = GEN ? kGenMaxKeys * 128 : kMaxQ * 128;   // bytes of the K / V slots
    static constexpr uint32_t kStage = kQ + 2 * kKV;
    static constexpr uint32_t kOut = kStages * kStage;          // 2 output tiles of 128 rows x 128 B
    static constexpr uint32_t kXm = kOut + 2 * 16384;           // partial maxima  [4 tile slots][2 parts][128 rows]
    static constexpr uint32_t kXs = kXm + 4 * 2 * 128 * 4;
    static constexpr uint32_t kBias = kXs + 4 * 2 * 128 * 4;    // GEN: key-off bit masks [stage][16 words]: bit i of word g = key 16 g + i is masked / padded
    static constexpr uint32_t kBar = kBias + (GEN ? kStages * 16 * 4 : 0);
    static constexpr uint32_t kSmem = kBar + 256 + 1024;
    static_assert(kSmem <= 232448, "shared memory budget exceeded");
};
constexpr uint32_t kColO = 416;

struct Args {
    int B, H, S, nks, n_qt, total_heads;
    float scale_log2;
    float* lse;
    long long* dbg;   // optional in-kernel cycle stamps (tools/attn_timeline.py)
    uint32_t drop_thresh, drop_stream;   // attention dropout (DROP instantiations): keep iff hash >= thresh, P *= 1 / (1 - p)
    float drop_inv_keep;
    const uint32_t* drop_seed;
    // GEN: S = number of queries, Sk = number of keys, n_qb = 128-query tiles per head, n_kb = key blocks per head, KB = keys per block
    // (a multiple of 16, nks = KB / 16); total_heads = number of work items = B * H * n_qb * n_kb
    int Sk, n_qb, n_kb, KB;
    const uint8_t* kpm;      // [B, Sk] key-padding mask (1 = ignore) or null
    float* lse_part;         // [n_kb][B * H * S] block log-sum-exps when n_kb > 1
};

struct Item {
    int b, h, bh, qb, kb;
};
template <bool GEN>
__device__ __forceinline__ Item decode_item(const Args& a, int item) {
    Item it;
    if (GEN) {
        const int per_head = a.n_qb * a.n_kb;
        it.bh = item / per_head;
        const int r = item - it.bh * per_head;
        it.qb = r / a.n_kb;
        it.kb = r - it.qb * a.n_kb;
    } else {
        it.bh = item; it.qb = 0; it.kb = 0;
    }
    it.b = it.bh / a.H;
    it.h = it.bh - it.b * a.H;
    return it;
}

template <int NKS_T, bool DROP, bool GEN>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Args args) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
    using L = Lay<GEN>;
    constexpr int kNSt = (int)L::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
    uint64_t* kv_full = bars;          // [kNSt]
    uint64_t* kv_empty = bars + 3;     // [kNSt] tcgen05.commit after the head's last P V product
    uint64_t* s_full = bars + 6;       // [2] per TMEM slot
    uint64_t* p_full = bars + 8;       // [2 slots][2] count 8 (one softmax group): keys < 128 packed / all keys packed
    uint64_t* o_full = bars + 12;
    uint64_t* o_free = bars + 13;      // count kRoWarps: the O accumulator has been read out
    uint64_t* out_ready = bars + 14;   // [2] count kRoWarps: output tile staged
    uint64_t* out_free = bars + 16;    // [2] count 1: the TMA store has finished reading the tile
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);
    // stage / phase of the hc-th head (item) of this CTA
    auto stage_of = [](int hc) { return GEN ? hc % 3 : (hc & 1); };
    auto phase_of = [](int hc) { return GEN ? (hc / 3) & 1 : (hc >> 1) & 1; };

    const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = args.S, n_qt = args.n_qt;
    const int nks = NKS_T ? NKS_T : args.nks;
    const int npad = nks * 16;
    const int jA = nks < 8 ? nks : 8;
    pdl_launch_dependents();

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < kNSt; ++i) {
            mbar_init(&kv_full[i], GEN ? 2 : 1);   // GEN: TMA + the bias warp
            mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[2 * i], 8); mbar_init(&p_full[2 * i + 1], 8);
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
                const Item it = decode_item<GEN>(args, head);
                const int st = stage_of(hc), b = it.b, h = it.h;
                mbar_wait(&kv_empty[st], phase_of(hc) ^ 1);
                uint8_t* sq = smem + st * L::kStage;
                mbar_arrive_expect_tx(&kv_full[st], GEN ? (128 + 2 * npad) * 128 : 3 * npad * 128);
                tma_load_3d(sq, &tmQ, &kv_full[st], h * 64, it.qb * 128, b);
                tma_load_3d(sq + L::kQ, &tmK, &kv_full[st], h * 64, it.kb * args.KB, b);
                tma_load_3d(sq + L::kQ + L::kKV, &tmV, &kv_full[st], h * 64, it.kb * args.KB, b);
            }
        }
    } else if (GEN && warp_idx == 3) {
        // ---------------- key-bias warp (GEN): 0 for a key that takes part, -inf for a padded / masked / out-of-range one ----------------
        int hc = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
            const Item it = decode_item<GEN>(args, head);
            const int st = stage_of(hc);
            mbar_wait(&kv_empty[st], phase_of(hc) ^ 1);
            uint32_t* mw = reinterpret_cast<uint32_t*>(smem + L::kBias) + st * 16;   // one word per 16-key group: bit i = key i is off
            const int k0 = it.kb * args.KB;
            const uint8_t* mrow = args.kpm ? args.kpm + (long long)it.b * args.Sk : nullptr;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const int col = lane + 32 * r, key = k0 + col;
                bool off = !(col < npad && key < args.Sk);
                if (!off && mrow) off = mrow[key] != 0;
                const uint32_t bits = __ballot_sync(0xffffffffu, off);       // keys col0 .. col0 + 31 of this pass
                if (lane == 0 && 2 * r < 13) mw[2 * r] = bits & 0xFFFFu;
                if (lane == 0 && 2 * r + 1 < 13) mw[2 * r + 1] = bits >> 16;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&kv_full[st]);
        }
    } else if (warp_idx == 2) {
        // ---------------- store warp ----------------
        int j = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x) {
            const Item it = decode_item<GEN>(args, head);
            const int h = it.h;
            const int b = (GEN && args.n_kb > 1) ? it.kb * args.B + it.b : it.b;   // partial results: one batch slab per key block
            for (int qt = 0; qt < n_qt; ++qt, ++j) {
                const int buf = j & 1;
                mbar_wait(&out_ready[buf], (j >> 1) & 1);
                if (lane == 0) {
                    tma_store_3d(&tmO, smem + L::kOut + buf * 16384, h * 64, (GEN ? it.qb : qt) * 128, b);
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
        const uint32_t idesc_s = umma_idesc_bf16(128, npad, 0, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
        constexpr uint64_t kdesc = umma_smem_desc_base(0, 1024);            // K-major SW128 (Q, K)
        constexpr uint64_t vdesc = umma_smem_desc_base(L::kKV, 1024);       // MN-major SW128 (V): 8-key groups 1024 B apart
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const int total_tiles = ((args.total_heads - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * n_qt;
        // issues S of tile jj (head counter jj / n_qt, query tile jj % n_qt)
        auto issue_s = [&](int jj) {
            const int hc = jj / n_qt, qt = jj - hc * n_qt, st = stage_of(hc);
            if (qt == 0) {
                mbar_wait(&kv_full[st], phase_of(hc));
                tcgen05_fence_after();
            }
            const uint32_t sq = smem_u32(smem + st * L::kStage) + qt * 16384, sk = smem_u32(smem + st * L::kStage + L::kQ);
            const uint32_t d = tb + (jj & 1) * kMaxQ;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(d, umma_smem_desc(kdesc, sq + k * 32), umma_smem_desc(kdesc, sk + k * 32), idesc_s, k > 0 ? 1u : 0u);
                umma_commit(&s_full[jj & 1]);
            }
            __syncwarp();
        };
        // Program order S(0) S(1) PV(0) S(2) PV(1) S(3) ...: S(j+2) re-uses the slot whose P was consumed by PV(j) (in-order tensor pipe),
        // and it is what lets softmax group j & 1 start its next tile — so a group can never run a whole tile ahead of its slowest warp.
        if (total_tiles > 0) issue_s(0);
        if (total_tiles > 1) issue_s(1);
        for (int j = 0; j < total_tiles; ++j) {
            const int hc = j / n_qt, qt = j - hc * n_qt, st = stage_of(hc), slot = j & 1, par = (j >> 1) & 1;
            const uint64_t bd = umma_smem_desc(vdesc, smem_u32(smem + st * L::kStage + L::kQ + L::kKV));
            const uint32_t a0 = tb + slot * kMaxQ;
            mbar_wait(&p_full[2 * slot], par);
            mbar_wait(o_free, (j & 1) ^ 1);
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < jA) umma_bf16_ts(tb + kColO, a0 + 16 * k, bd + (uint64_t)(k * 128), idesc_o, k > 0 ? 1u : 0u);
            }
            __syncwarp();
            mbar_wait(&p_full[2 * slot + 1], par);
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 8; k < (NKS_T ? NKS_T : 13); ++k)
                    if (NKS_T || k < nks) umma_bf16_ts(tb + kColO, a0 + 16 * k, bd + (uint64_t)(k * 128), idesc_o, 1u);
                umma_commit(o_full);
                if (qt == n_qt - 1) umma_commit(&kv_empty[st]);
            }
            __syncwarp();
            if (j + 2 < total_tiles) issue_s(j + 2);
        }
    } else if (warp_idx >= 4 && warp_idx < 4 + kEwWarps) {
        // ---------------- softmax ----------------
        const uint32_t quad = warp_idx & 3, grp = (warp_idx - 4) >> 3, part = ((warp_idx - 4) >> 2) & 1;
        const uint32_t t_lane = tmem_base + ((quad * 32) << 16);
        const int row_in_tile = quad * 32 + lane;
        const float c = args.scale_log2;
        const uint32_t xm_u32 = smem_u32(smem + L::kXm), xs_u32 = smem_u32(smem + L::kXs);
        const uint32_t drop_key = DROP ? dropout_key(*args.drop_seed, args.drop_stream) : 0u;
        int j = 0, hc = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
            const Item it = decode_item<GEN>(args, head);
            const uint32_t bias_u32 = smem_u32(smem + L::kBias) + (uint32_t)stage_of(hc) * 16 * 4;
            for (int qt = 0; qt < n_qt; ++qt, ++j) {
                if ((uint32_t)(j & 1) != grp) continue;   // tile j belongs to softmax group j & 1
                // element (q, k) of the head -> base + k
                const uint32_t drop_base = GEN ? ((uint32_t)it.bh * (uint32_t)S + (uint32_t)(it.qb * 128 + row_in_tile)) * (uint32_t)args.Sk + (uint32_t)(it.kb * args.KB)
                                               : ((uint32_t)head * (uint32_t)S + (uint32_t)(qt * 128 + row_in_tile)) * (uint32_t)S;
                const uint32_t t_s = t_lane + (j & 1) * kMaxQ;
                const uint32_t xoff = ((j & 3) * 2 * 128 + row_in_tile) * 4;
                const bool dbg_on = args.dbg && blockIdx.x == 0 && j < 64 && (warp_idx == 4 || warp_idx == 12) && lane == 0;
                VB_DBG(dbg_on, args.dbg[j * 16 + 0]);
                mbar_wait(&s_full[j & 1], (j >> 1) & 1);
                tcgen05_fence_after();
                VB_DBG(dbg_on, args.dbg[j * 16 + 1]);
                // ---- pass 1: partial row maximum over my groups (raw scores; keys >= S are dead) ----
                float mx = -INFINITY;
                {
                    // two register sets: the TMEM load of the next group is in flight while the current one is reduced
                    uint32_t ra[16], rb[16];
                    auto reduce = [&](uint32_t (&r)[16], int g) {
                        if (GEN) {   // the mask word is the same for every row: warp-uniform branches, nothing to do for an all-on group
                            const uint32_t w = lds_u32(bias_u32 + g * 4);
                            if (w == 0xFFFFu) return;            // every key of the group is off
                            if (w != 0u) {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if ((w >> i) & 1u) r[i] = 0xff800000u;
                            }
                        } else if (g * 16 + 16 > S) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (g * 16 + i >= S) r[i] = 0xff800000u;
                        }
                        float m0 = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), m1 = fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3]));
#pragma unroll
                        for (int i = 4; i < 16; i += 2) {
                            m0 = fmaxf(m0, __uint_as_float(r[i]));
                            m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
                        }
                        mx = fmaxf(mx, fmaxf(m0, m1));
                    };
                    int g = part;
                    if (g < nks) tmem_ld_32x32b_x16(t_s + g * 16, ra);
                    for (; g < nks; g += 4) {
                        tmem_ld_wait();
                        if (g + 2 < nks) tmem_ld_32x32b_x16(t_s + (g + 2) * 16, rb);
                        reduce(ra, g);
                        if (g + 2 < nks) {
                            tmem_ld_wait();
                            if (g + 4 < nks) tmem_ld_32x32b_x16(t_s + (g + 4) * 16, ra);
                            reduce(rb, g + 2);
                        }
                    }
                }
                VB_DBG(dbg_on, args.dbg[j * 16 + 2]);
                sts32(xm_u32 + xoff + part * 512, mx);
                named_bar_sync(1 + grp * 4 + quad, 64);
                mx = fmaxf(lds32(xm_u32 + xoff), lds32(xm_u32 + xoff + 512));
                const float nm = (GEN && mx == -INFINITY) ? 0.f : -mx * c;   // every key of this block masked: P = 0, l = 0 (no NaN)
                VB_DBG(dbg_on, args.dbg[j * 16 + 3]);
                // ---- pass 2: P = exp2(s c - m c) packed over the group's own columns; partial row sum ----
                float l = 0.f;
                bool arrivedA = false;
                auto arriveA = [&]() {
                    tmem_st_wait();
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[2 * grp]);
                    arrivedA = true;
                };
                {
                    uint32_t ra[16], rb[16];
                    auto emit = [&](uint32_t (&r)[16], int g) {
                        // the volatile no-op pins the math below after the (volatile) prefetch of the next group
                        float nmg = nm;
                        asm volatile("" : "+f"(nmg));
                        if (GEN) {
                            const uint32_t w = lds_u32(bias_u32 + g * 4);
                            if (w == 0xFFFFu) {                  // every key of the group is off: P = 0 without touching the MUFU
                                const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                                tmem_st_32x32b_x8(t_s + g * 16, zero);
                                return;
                            }
                            if (w != 0u) {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if ((w >> i) & 1u) r[i] = 0xff800000u;
                            }
                        } else if (g * 16 + 16 > S) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (g * 16 + i >= S) r[i] = 0xff800000u;
                        }
                        uint32_t pk[8];
                        float l0 = 0.f, l1 = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float p0 = ex2f(fmaf(__uint_as_float(r[2 * i]), c, nmg));
                            float p1 = ex2f(fmaf(__uint_as_float(r[2 * i + 1]), c, nmg));
                            l0 += p0; l1 += p1;
                            if (DROP) {   // the row sum keeps the un-dropped probabilities; P V sees keep * P / (1 - p)
                                p0 = dropout_keep(drop_key, drop_base + g * 16 + 2 * i, args.drop_thresh) ? p0 * args.drop_inv_keep : 0.f;
                                p1 = dropout_keep(drop_key, drop_base + g * 16 + 2 * i + 1, args.drop_thresh) ? p1 * args.drop_inv_keep : 0.f;
                            }
                            pk[i] = pack2(p0, p1);
                        }
                        l += l0 + l1;
                        tmem_st_32x32b_x8(t_s + g * 16, pk);
                    };
                    int g = part;
                    if (g < nks) tmem_ld_32x32b_x16(t_s + g * 16, ra);
                    for (; g < nks; g += 4) {
                        if (g >= 8 && !arrivedA) arriveA();
                        tmem_ld_wait();
                        if (g + 2 < nks) tmem_ld_32x32b_x16(t_s + (g + 2) * 16, rb);
                        emit(ra, g);
                        if (g + 2 < nks) {
                            if (g + 2 >= 8 && !arrivedA) arriveA();
                            tmem_ld_wait();
                            if (g + 4 < nks) tmem_ld_32x32b_x16(t_s + (g + 4) * 16, ra);
                            emit(rb, g + 2);
                        }
                    }
                }
                if (!arrivedA) arriveA();
                sts32(xs_u32 + xoff + part * 512, l);
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[2 * grp + 1]);
                VB_DBG(dbg_on, args.dbg[j * 16 + 4]);
            }
        }
    } else if (warp_idx >= 4 + kEwWarps) {
        // ---------------- read-out: O / l -> bf16 tile in shared memory (TMA store by warp 2), lse ----------------
        const uint32_t quad = warp_idx & 3;
        const uint32_t t_lane = tmem_base + ((quad * 32) << 16);
        const int row_in_tile = quad * 32 + lane;
        const float c = args.scale_log2;
        const uint32_t swz = (uint32_t)(row_in_tile & 7);
        const uint32_t xm_u32 = smem_u32(smem + L::kXm), xs_u32 = smem_u32(smem + L::kXs);
        const uint32_t out_row = smem_u32(smem + L::kOut) + row_in_tile * 128;
        int j = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x) {
            const Item it = decode_item<GEN>(args, head);
            for (int qt = 0; qt < n_qt; ++qt, ++j) {
                const uint32_t xoff = ((j & 3) * 2 * 128 + row_in_tile) * 4;
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
                const float tot = lds32(xs_u32 + xoff) + lds32(xs_u32 + xoff + 512);
                const float mx = fmaxf(lds32(xm_u32 + xoff), lds32(xm_u32 + xoff + 512));
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
                const int q = (GEN ? it.qb : qt) * 128 + row_in_tile;
                if (GEN && args.n_kb > 1) {   // block log-sum-exp (-inf when every key of the block is masked)
                    if (q < S) args.lse_part[((long long)it.kb * args.B * args.H + it.bh) * S + q] = mx * c + log2f(tot);
                } else if (args.lse && q < S) {
                    args.lse[(long long)it.bh * S + q] = mx * c + log2f(tot);
                }
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

static long long* g_dbg = nullptr;
}  // namespace fwd3
void attention_fwd_tc3_set_debug(long long* p) { fwd3::g_dbg = p; }

// Returns VB_OK if launched, 1 if this shape is not handled here.
int attention_fwd_tc3(const VbAttnDesc* d, cudaStream_t stream) {
    using namespace fwd3;
    if (d->S > (int)kMaxQ || d->head_dim != 64 || d->tok_stride != 1 || d->key_padding_mask != nullptr) return 1;
    const int S = d->S, npad = (S + 15) / 16 * 16;
    Args a{};
    a.B = d->B; a.H = d->H; a.S = S; a.nks = npad / 16; a.n_qt = (S + 127) / 128; a.total_heads = d->B * d->H;
    a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.lse = d->lse;
    a.dbg = g_dbg;
    CUtensorMap tq, tk, tv, to;
    const uint64_t cols = (uint64_t)d->H * 64;
    int rc;
    if ((rc = make_tmap_3d(&tq, VB_BF16, d->q, cols, S, d->B, d->ldq, d->batch_stride * d->ldq, 64, npad))) return rc;
    if ((rc = make_tmap_3d(&tk, VB_BF16, d->k, cols, S, d->B, d->ldk, d->batch_stride * d->ldk, 64, npad))) return rc;
    if ((rc = make_tmap_3d(&tv, VB_BF16, d->v, cols, S, d->B, d->ldv, d->batch_stride * d->ldv, 64, npad))) return rc;
    if ((rc = make_tmap_3d(&to, VB_BF16, d->o, cols, S, d->B, d->ldo, d->batch_stride * d->ldo, 64, 128))) return rc;
    int grid = num_sms();
    if (grid > a.total_heads) grid = a.total_heads;
    const bool drop = d->dropout_p > 0.f;
    if (drop) {
        VB_REQUIRE(d->dropout_p < 1.f && d->dropout_seed != nullptr, "attention dropout: p must be < 1 and dropout_seed non-null");
        VB_REQUIRE((long long)a.total_heads * S * S < (1ll << 32), "attention dropout: more than 2^32 score elements");
        a.drop_thresh = dropout_threshold(d->dropout_p);
        a.drop_inv_keep = 1.0f / (1.0f - d->dropout_p);
        a.drop_seed = d->dropout_seed;
        a.drop_stream = d->dropout_stream;
    }
#define VB_FWD_LAUNCH(NKS, DR)                                                                                              \
    do {                                                                                                                    \
        static DeviceOnce configured;                                                                                     \
        if (!configured.is_set()) {                                                                                                  \
            VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc3_kernel<NKS, DR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<false>::kSmem)); \
            configured.set();                                                                                              \
        }                                                                                                                   \
        VB_CUDA_CHECK(launch_pdl(attn_fwd_tc3_kernel<NKS, DR, false>, dim3(grid), dim3(kThreads), Lay<false>::kSmem, stream, tq, tk, tv, to, a)); \
    } while (0)
    if (a.nks == 13) {
        if (drop) VB_FWD_LAUNCH(13, true); else VB_FWD_LAUNCH(13, false);
    } else {
        if (drop) VB_FWD_LAUNCH(0, true); else VB_FWD_LAUNCH(0, false);
    }
#undef VB_FWD_LAUNCH
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// General path: any S / S_kv, key-padding masks, sequence-first strides, cross-attention (GEN instantiations above)
// ---------------------------------------------------------------------------------------------------------------------------
namespace fwd3 {

// out[b, q, h, :] = sum_kb w_kb * part[kb][b][q][h][:],  w_kb = 2^(lse_kb - lse) with lse = log2 sum_kb 2^lse_kb: the exact
// recombination of per-key-block softmax results.  One warp per (token, head): 128 bytes per block, fp32 accumulation.
__global__ void __launch_bounds__(256) attn_merge_kernel(const __nv_bfloat16* __restrict__ part, const float* __restrict__ lse_part,
                                                         __nv_bfloat16* __restrict__ out, long long ldo, long long tok_stride,
                                                         long long batch_stride, float* __restrict__ lse, int B, int H, int S, int n_kb) {
    const int lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (b * S + q) * H + h
    if (unit >= (long long)B * S * H) return;
    const int h = (int)(unit % H);
    const long long tok = unit / H;
    const int q = (int)(tok % S), b = (int)(tok / S);
    const long long bh = (long long)b * H + h, slab = (long long)B * H * S;
    const long long D = (long long)H * 64;
    const float* lp = lse_part + bh * S + q;
    const __nv_bfloat16* pp = part + ((long long)b * S + q) * D + h * 64 + lane * 2;
    const long long pslab = (long long)B * S * D;
    // chunks of 8 key blocks: the 16 loads of a chunk are independent (one memory round trip per chunk, not per block); chunks are
    // combined with a running maximum
    float m = -INFINITY, tot = 0.f, a0 = 0.f, a1 = 0.f;
    for (int kb0 = 0; kb0 < n_kb; kb0 += 8) {
        float l[8];
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool ok = kb0 + i < n_kb;
            l[i] = ok ? __ldg(lp + (long long)(kb0 + i) * slab) : -INFINITY;
            v[i] = ok ? __ldg(reinterpret_cast<const uint32_t*>(pp + (long long)(kb0 + i) * pslab)) : 0u;
        }
        float mc = m;
#pragma unroll
        for (int i = 0; i < 8; ++i) mc = fmaxf(mc, l[i]);
        if (mc == -INFINITY) continue;                       // nothing but fully masked blocks so far
        const float resc = (m == -INFINITY) ? 0.f : exp2f(m - mc);
        tot *= resc; a0 *= resc; a1 *= resc;
        m = mc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float w = (l[i] == -INFINITY) ? 0.f : exp2f(l[i] - m);
            tot += w;
            a0 = fmaf(w, __uint_as_float(v[i] << 16), a0);
            a1 = fmaf(w, __uint_as_float(v[i] & 0xFFFF0000u), a1);
        }
    }
    const float inv = tot > 0.f ? 1.f / tot : 0.f;
    __nv_bfloat162 r = __floats2bfloat162_rn(a0 * inv, a1 * inv);
    *reinterpret_cast<__nv_bfloat162*>(out + ((long long)b * batch_stride + (long long)q * tok_stride) * ldo + h * 64 + lane * 2) = r;
    if (lse && lane == 0) lse[bh * S + q] = m + log2f(tot);
}

}  // namespace fwd3

// Key blocks per head: ceil(Sk / 176) blocks of equal size (rounded up to 16 keys), so that the last block is not mostly padding.
static void gen_key_blocks(int Sk, int* n_kb, int* KB) {
    const int nb = (Sk + (int)fwd3::kGenMaxKeys - 1) / (int)fwd3::kGenMaxKeys;
    int kb = ((Sk + nb - 1) / nb + 15) / 16 * 16;
    *n_kb = (Sk + kb - 1) / kb;
    *KB = kb;
}

size_t attention_fwd_gen_workspace(const VbAttnDesc* d) {
    const int Sk = d->S_kv > 0 ? d->S_kv : d->S;
    int n_kb, KB;
    gen_key_blocks(Sk, &n_kb, &KB);
    if (n_kb <= 1) return 0;
    const size_t opart = ((size_t)n_kb * d->B * d->S * d->H * 64 * 2 + 255) / 256 * 256;
    return opart + (size_t)n_kb * d->B * d->H * d->S * 4;
}

// Returns VB_OK if launched, < 0 on error.
int attention_fwd_gen(const VbAttnDesc* d, cudaStream_t stream) {
    using namespace fwd3;
    const int S = d->S, Sk = d->S_kv > 0 ? d->S_kv : d->S;
    Args a{};
    gen_key_blocks(Sk, &a.n_kb, &a.KB);
    a.B = d->B; a.H = d->H; a.S = S; a.Sk = Sk; a.nks = a.KB / 16; a.n_qt = 1;
    a.n_qb = (S + 127) / 128;
    const long long items = (long long)d->B * d->H * a.n_qb * a.n_kb;
    VB_REQUIRE(items < (1ll << 31), "attention: too many work items");
    a.total_heads = (int)items;
    a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.lse = d->lse;
    a.kpm = d->key_padding_mask;
    a.dbg = g_dbg;
    const uint64_t cols = (uint64_t)d->H * 64;
    __nv_bfloat16* opart = nullptr;
    if (a.n_kb > 1) {
        const size_t need = attention_fwd_gen_workspace(d);
        VB_REQUIRE(d->workspace != nullptr && (size_t)d->workspace_bytes >= need && (reinterpret_cast<uintptr_t>(d->workspace) & 255) == 0,
                   "attention_fwd: %d key blocks need a 256-byte aligned workspace of %zu bytes (vb_attention_workspace_bytes)", a.n_kb, need);
        opart = reinterpret_cast<__nv_bfloat16*>(d->workspace);
        const size_t off = ((size_t)a.n_kb * d->B * S * cols * 2 + 255) / 256 * 256;
        a.lse_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->workspace) + off);
    }
    CUtensorMap tq, tk, tv, to;
    int rc;
    if ((rc = make_tmap_3d(&tq, VB_BF16, d->q, cols, S, d->B, d->tok_stride * d->ldq, d->batch_stride * d->ldq, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tk, VB_BF16, d->k, cols, Sk, d->B, d->tok_stride * d->ldk, d->batch_stride * d->ldk, 64, a.KB))) return rc;
    if ((rc = make_tmap_3d(&tv, VB_BF16, d->v, cols, Sk, d->B, d->tok_stride * d->ldv, d->batch_stride * d->ldv, 64, a.KB))) return rc;
    if (a.n_kb > 1) rc = make_tmap_3d(&to, VB_BF16, opart, cols, S, (uint64_t)a.n_kb * d->B, cols, (uint64_t)S * cols, 64, 128);
    else rc = make_tmap_3d(&to, VB_BF16, d->o, cols, S, d->B, d->tok_stride * d->ldo, d->batch_stride * d->ldo, 64, 128);
    if (rc) return rc;
    int grid = num_sms();
    if (grid > a.total_heads) grid = a.total_heads;
    const bool drop = d->dropout_p > 0.f;
    if (drop) {
        VB_REQUIRE(d->dropout_p < 1.f && d->dropout_seed != nullptr, "attention dropout: p must be < 1 and dropout_seed non-null");
        VB_REQUIRE((long long)d->B * d->H * S * Sk < (1ll << 32), "attention dropout: more than 2^32 score elements");
        a.drop_thresh = dropout_threshold(d->dropout_p);
        a.drop_inv_keep = 1.0f / (1.0f - d->dropout_p);
        a.drop_seed = d->dropout_seed;
        a.drop_stream = d->dropout_stream;
    }
    static DeviceOnce configured;
    if (!configured.is_set()) {
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc3_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<true>::kSmem));
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc3_kernel<0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<true>::kSmem));
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc3_kernel<11, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<true>::kSmem));
        configured.set();
    }
    // 176-key blocks (every long sequence: S_kv > 176 splits into blocks of ~176) get the fully unrolled instantiation
    if (drop) VB_CUDA_CHECK(launch_pdl(attn_fwd_tc3_kernel<0, true, true>, dim3(grid), dim3(kThreads), Lay<true>::kSmem, stream, tq, tk, tv, to, a));
    else if (a.nks == 11) VB_CUDA_CHECK(launch_pdl(attn_fwd_tc3_kernel<11, false, true>, dim3(grid), dim3(kThreads), Lay<true>::kSmem, stream, tq, tk, tv, to, a));
    else VB_CUDA_CHECK(launch_pdl(attn_fwd_tc3_kernel<0, false, true>, dim3(grid), dim3(kThreads), Lay<true>::kSmem, stream, tq, tk, tv, to, a));
    if (a.n_kb > 1) {
        const long long units = (long long)d->B * S * d->H;
        attn_merge_kernel<<<(unsigned)((units + 7) / 8), 256, 0, stream>>>(opart, a.lse_part, reinterpret_cast<__nv_bfloat16*>(d->o), d->ldo,
                                                                          d->tok_stride, d->batch_stride, d->lse, d->B, d->H, S, a.n_kb);
        VB_CUDA_CHECK(cudaGetLastError());
    }
    return VB_OK;
}

}  // namespace vb
