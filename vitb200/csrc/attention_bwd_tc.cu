// tcgen05 / TMEM attention backward for short sequences (S <= 208: every ViT / DeiT config), head_dim 64.
//
// Five matrix products per head, P computed once (rows = keys):
//   per 128-key tile t:   S^T = K_t Q^T,  dP^T = V_t dO^T                  (SS MMAs, N = npad, fp32 in TMEM)
//     element-wise        P^T = exp2(S^T c - lse[q]),  dS^T = P^T o (dP^T - delta[q])
//                         P^T  -> packed bf16 written over the S^T columns (A operand of dV, read from TMEM)
//                         dS^T -> bf16 [key][query] staging tile in shared memory (A operand of dK and dQ)
//     dV_t  = P^T dO            (TS MMA)         dK_t = dS^T Q / 8   (SS MMA, staging read K-major)
//     dQ   += dS K_t / 8        (SS MMA, the same staging tile read MN-major; accumulated over the key tiles)
// One persistent CTA per SM walks (batch, head) pairs.  Q / dO (whole head) and K_t / V_t (per key tile) arrive by TMA in
// two 2-stage rings, so the loads of the next tile / head overlap the math of the current one.
//
// TMEM map (512 columns): [0,208) S^T, the packed P^T of 16-query group g overwrites columns [16g, 16g+8); [208,416) dP^T ->
// once the queries < 128 are consumed the dead columns hold the dV [208,272) and dK [272,336) accumulators, and after the
// whole element-wise pass the second-query-tile dQ accumulator [336,400); [448,512) is the dQ accumulator of query tile 0,
// which persists across the key tiles.  The dQ contribution of key tile 0 to query tile 1 is carried in registers (the
// TMEM budget is 32 columns short of keeping it resident).
// The fourth 64-query staging block (queries 192..207) aliases the V_t buffer, which is dead once dP^T is complete.
// Pipelining inside a tile: scores and the element-wise pass are split at query 128 (two commit / arrive points each),
// so the second-stage products of the first part overlap the element-wise work on the second.  Accumulators are read out
// into the dead K_t / V_t / Q buffers and leave by TMA store issued from a dedicated warp.
//
// GEN instantiations (any number of queries / keys, key-padding masks, sequence-first strides, cross-attention: the DETR encoder and
// decoder): a work item is (batch, head, 128-query block) looping over ALL key tiles; dQ of the block accumulates in TMEM across
// the key tiles as above, while dK_t / dV_t — partial sums over this query block only — are reduce-added (red.global.add.v4.f32)
// from registers into fp32 accumulators that attn_acc_to_bf16_kernel converts afterwards.  Masked keys get P = dS = 0.
// Replaces the autograd of F.scaled_dot_product_attention reached from nn.MultiheadAttention (vanilla_vit.py:77,
// torch/nn/functional.py:6676-6688).  delta = rowsum(dO o O) comes from attn_delta_kernel (attention.cu).
#include <cuda.h>
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"
#include "attn_tc_common.cuh"
#include "dropout.cuh"

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

namespace bwd5 {

constexpr int kEwWarps = 12;                           // element-wise / read-out warps (3 per TMEM lane quadrant)
constexpr int kThreads = 128 + kEwWarps * 32;          // warps 0-3: TMA, MMA, TMEM alloc, stats
constexpr uint32_t kMaxQ = 208;                        // padded queries / keys per head
constexpr uint32_t kBlk = 128 * 128;                   // one 64-wide staging block / K_t / V_t tile: 128 rows x 128 B
constexpr uint32_t kQBytes = kMaxQ * 128;              // Q or dO of one head
// Shared-memory layout.  GEN items have at most 128 queries: two staging blocks, 16 KB Q / dO slots, and two extra 16 KB buffers in
// which the fp32 dK_t halves wait for their TMA reduce-add (the fp32 dV_t halves park in the dead K_t / V_t buffers).
template <bool GEN>
struct Lay {
    static constexpr uint32_t kStaging = 0;                                   // dS^T blocks of 64 queries (block 3 aliases V_t)
    static constexpr uint32_t kQSlot = GEN ? kBlk : kQBytes;                  // Q or dO of one item
    static constexpr uint32_t kQdo = (GEN ? 2 : 3) * kBlk;                    // 2 stages x {Q, dO}
    static constexpr uint32_t kKv = kQdo + 4 * kQSlot;                        // 2 stages x {K_t, V_t}
    static constexpr uint32_t kDk = kKv + 4 * kBlk;                           // GEN: fp32 dK_t columns 0..31 / 32..63
    static constexpr uint32_t kStats = kDk + (GEN ? 2 * kBlk : 0);            // 2 stages x {-lse[208], -delta[208]} fp32
    static constexpr uint32_t kBar = kStats + 2 * 2 * kMaxQ * 4;
    static constexpr uint32_t kSmem = kBar + 256 + 1024;
    static_assert(kSmem <= 232448, "shared memory budget exceeded");
};

constexpr uint32_t kColST = 0, kColDPT = 208, kColDV = 208, kColDK = 272, kColDQ1 = 336, kColCarry = 416, kColDQ0 = 448;

struct Args {
    long long* dbg;
    int B, H, S, npad, nks, n_t, total_heads;
    float scale, scale_log2;
    const float* lse;
    const __nv_bfloat16* o;     // forward output [token, H*64] bf16: delta = rowsum(dO o O) is computed in-kernel by the stats warp
    long long ldo;
    long long batch_stride;
    uint32_t drop_thresh, drop_stream;   // attention dropout (DROP instantiations)
    float drop_inv_keep;
    const uint32_t* drop_seed;
    // GEN: S = number of queries, Sk = number of keys (n_t = key tiles), n_qb = 128-query blocks per head, total_heads = work items;
    // tok_stride / batch_stride address the rows of o; dk_acc / dv_acc are fp32 [B * Sk, H * 64] accumulators (zeroed by the host)
    int Sk, n_qb;
    long long tok_stride;
    const uint8_t* kpm;
    float* dk_acc;
    float* dv_acc;
    int late_release;   // A/B switch (VITB200_ATTN_BWD_LATE_RELEASE=1): hand TMEM back only after the whole read-out
    float* colsum;   // optional fp32 [3][H*64]: += column sums of dq and dv (in-projection bias gradient; the dk part is exactly 0)
};

using namespace atc;

struct Item {
    int b, h, bh, qb;
};
template <bool GEN>
__device__ __forceinline__ Item decode_item(const Args& a, int item) {
    Item it;
    if (GEN) {
        it.bh = item / a.n_qb;
        it.qb = item - it.bh * a.n_qb;
    } else {
        it.bh = item; it.qb = 0;
    }
    it.b = it.bh / a.H;
    it.h = it.bh - it.b * a.H;
    return it;
}

// 16 columns of one tile row, phase 1: P^T = exp2(S^T c - lse[q]) from the raw scores sv (one TMEM lane = one key), packed to bf16 (pp,
// the A operand of dV) and kept in fp32 (pf) for phase 2.  nls points at -lse[q] of the 16 queries; fp32x2-packed arithmetic.
// DROP (attention dropout): element (query q, key) of head-local index didx0 + j * S was kept iff its hash clears the threshold;
// dV sees keep * P / (1 - p); `keep` returns the 16 keep bits for phase 2.
template <bool DROP>
__device__ __forceinline__ void ew_phase1(const uint32_t (&sv)[16], uint32_t nls, f2 c2, f2 (&pf)[8], uint32_t (&pp)[8], uint32_t& keep,
                                          uint32_t dkey, uint32_t didx0, uint32_t S, uint32_t thresh, float inv_keep) {
    keep = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f2 l01, l23;
        lds_2xf2(nls + 16 * i, l01, l23);
        float x0, x1, x2, x3;
        f2_unpack(f2_fma(f2_pack_u(sv[4 * i + 0], sv[4 * i + 1]), c2, l01), x0, x1);
        f2_unpack(f2_fma(f2_pack_u(sv[4 * i + 2], sv[4 * i + 3]), c2, l23), x2, x3);
        const float p0 = ex2f(x0), p1 = ex2f(x1), p2 = ex2f(x2), p3 = ex2f(x3);
        pf[2 * i] = f2_pack(p0, p1);
        pf[2 * i + 1] = f2_pack(p2, p3);
        if (DROP) {
            const bool k0 = dropout_keep(dkey, didx0 + (4 * i + 0) * S, thresh), k1 = dropout_keep(dkey, didx0 + (4 * i + 1) * S, thresh);
            const bool k2 = dropout_keep(dkey, didx0 + (4 * i + 2) * S, thresh), k3 = dropout_keep(dkey, didx0 + (4 * i + 3) * S, thresh);
            keep |= ((k0 ? 1u : 0u) | (k1 ? 2u : 0u) | (k2 ? 4u : 0u) | (k3 ? 8u : 0u)) << (4 * i);
            pp[2 * i] = pack2(k0 ? p0 * inv_keep : 0.f, k1 ? p1 * inv_keep : 0.f);
            pp[2 * i + 1] = pack2(k2 ? p2 * inv_keep : 0.f, k3 ? p3 * inv_keep : 0.f);
        } else {
            pp[2 * i] = pack2(p0, p1);
            pp[2 * i + 1] = pack2(p2, p3);
        }
    }
}
// phase 2: dS^T = P^T o (dP^T - delta[q]) from the score gradients dv; dls points at -delta[q]; with dropout dP = keep * (dO V^T) / (1 - p)
// (delta = rowsum(dO o O) still holds).
template <bool DROP>
__device__ __forceinline__ void ew_phase2(const f2 (&pf)[8], const uint32_t (&dv)[16], uint32_t dls, uint32_t (&pd)[8], uint32_t keep, float inv_keep) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f2 d01, d23;
        lds_2xf2(dls + 16 * i, d01, d23);
        f2 g01 = f2_pack_u(dv[4 * i + 0], dv[4 * i + 1]), g23 = f2_pack_u(dv[4 * i + 2], dv[4 * i + 3]);
        if (DROP) {
            const uint32_t kb = keep >> (4 * i);
            g01 = f2_mul(g01, f2_pack((kb & 1u) ? inv_keep : 0.f, (kb & 2u) ? inv_keep : 0.f));
            g23 = f2_mul(g23, f2_pack((kb & 4u) ? inv_keep : 0.f, (kb & 8u) ? inv_keep : 0.f));
        }
        float s0, s1, s2, s3;
        f2_unpack(f2_mul(pf[2 * i], f2_add(g01, d01)), s0, s1);
        f2_unpack(f2_mul(pf[2 * i + 1], f2_add(g23, d23)), s2, s3);
        pd[2 * i] = pack2(s0, s1);
        pd[2 * i + 1] = pack2(s2, s3);
    }
}

// NKS_T > 0: number of 16-query steps known at compile time (fully unrolled MMA issue); NKS_T == 0: generic.
template <int NKS_T, bool DROP, bool GEN>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                    const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                    const __grid_constant__ CUtensorMap tmDV, const Args args) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    using L = Lay<GEN>;
    constexpr uint32_t kQdoOff = L::kQdo, kKvOff = L::kKv, kStagingOff = L::kStaging, kQSlot = L::kQSlot;
    float* stats = reinterpret_cast<float*>(smem + L::kStats);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
    uint64_t* qdo_full = bars;        // [2] TMA expect_tx (Q and dO of the head have landed)
    uint64_t* qdo_empty = bars + 2;   // [2] store warp: the head's dQ tiles (parked in the Q / dO stage) have left
    uint64_t* kv_full = bars + 4;     // [2]
    uint64_t* kv_empty = bars + 6;    // [2] store warp: the tile's dV / dK tiles (parked in the K_t / V_t stage) have left
    uint64_t* s_full = bars + 8;      // [2] scores of queries < 128 / >= 128 complete
    uint64_t* p_full = bars + 10;     // [2] count kEwWarps: element-wise pass over queries < 128 / all queries done
    uint64_t* o_full = bars + 12;
    uint64_t* tile_free = bars + 13;  // count kEwWarps: accumulators read out, TMEM free for the next tile's scores
    uint64_t* out_ready = bars + 14;  // count kEwWarps: output tiles staged in shared memory
    uint64_t* stats_full = bars + 15; // [2] stats warp: -lse and delta of the head are in shared memory
    uint64_t* dk_free = bars + 17;    // GEN: the TMA reduce-adds have finished reading the fp32 dK_t buffers
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

    const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = args.S, n_t = args.n_t;
    const int nks = NKS_T ? NKS_T : args.nks;
    const int npad = nks * 16;                                        // padded queries of an item
    const int npad_k = GEN ? (args.Sk + 15) / 16 * 16 : npad;       // padded keys of the head
    const int jA = nks < 8 ? nks : 8;          // 16-query groups of the first part (queries < 128)
    const bool has_q1 = npad > 128;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
        tma_prefetch_desc(&tmDQ); tma_prefetch_desc(&tmDK); tma_prefetch_desc(&tmDV);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); mbar_init(&stats_full[i], 1);
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], kEwWarps);
        }
        mbar_init(o_full, 1); mbar_init(tile_free, kEwWarps); mbar_init(out_ready, kEwWarps); mbar_init(dk_free, 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_ptr_smem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp_idx == 0) {
        if (lane == 0) {   // ---------------- TMA loader ----------------
            int hc = 0, ic = 0;
            for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
                const Item it = decode_item<GEN>(args, head);
                const int qs = hc & 1, b = it.b, h = it.h;
                mbar_wait(&qdo_empty[qs], ((hc >> 1) & 1) ^ 1);
                uint8_t* qb = smem + kQdoOff + qs * 2 * kQSlot;
                mbar_arrive_expect_tx(&qdo_full[qs], 2 * npad * 128);
                tma_load_3d(qb, &tmQ, &qdo_full[qs], h * 64, it.qb * 128, b);
                tma_load_3d(qb + kQSlot, &tmdO, &qdo_full[qs], h * 64, it.qb * 128, b);
                for (int t = 0; t < n_t; ++t, ++ic) {
                    const int ks = ic & 1;
                    mbar_wait(&kv_empty[ks], ((ic >> 1) & 1) ^ 1);
                    uint8_t* kb = smem + kKvOff + ks * 2 * kBlk;
                    mbar_arrive_expect_tx(&kv_full[ks], 2 * kBlk);
                    tma_load_3d(kb, &tmK, &kv_full[ks], h * 64, t * 128, b);
                    tma_load_3d(kb + kBlk, &tmV, &kv_full[ks], h * 64, t * 128, b);
                }
            }
        }
    } else if (warp_idx == 3) {
        // ---------------- statistics warp: -lse from global, delta[q] = sum_d dO[q,d] O[q,d] computed here ----------------
        // (replaces the separate attn_delta pre-kernel: 32 us per layer at batch 256).  dO of the head is already in shared memory
        // (TMA, 128-byte swizzle); O rows are read from global, 4 lanes x 32 bytes (LDG.256) per row, 8 rows per pass, fp32 accumulation.
        // The stage is filled one head ahead of its use, so this work is off the critical path.
        int hc = 0;
        const int rsub = lane >> 2, csub = lane & 3;   // 4 lanes x 32 bytes per row, 8 rows per pass, 26 passes
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
            const Item it = decode_item<GEN>(args, head);
            const int qs = hc & 1, b = it.b, h = it.h, q0 = it.qb * 128;
            const int nq = GEN ? min(128, S - q0) : S;                 // valid queries of this item
            mbar_wait(&qdo_empty[qs], ((hc >> 1) & 1) ^ 1);
            const bool sdbg = args.dbg && blockIdx.x == 0 && hc < 32 && lane == 0;   // stamps in the row of the head's first tile
            VB_DBG(sdbg, args.dbg[hc * 2 * 16 + 13]);
            float* nl = stats + qs * 2 * kMaxQ;
            float* dl = nl + kMaxQ;
            const float* gl = args.lse + (long long)it.bh * S + q0;
            const long long orow = GEN ? args.tok_stride * args.ldo : args.ldo;   // pitch between consecutive tokens of o
            const __nv_bfloat16* go = args.o + ((long long)b * args.batch_stride) * args.ldo + (long long)q0 * orow + h * 64 + csub * 16;
            constexpr int kBatch = 9;                           // 3 batches x 9 passes (the last pass of the third batch is empty)
            uint32_t ov[kBatch][8];
            auto load_batch = [&](int p0) {                     // independent 256-bit loads: one round trip per batch
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    const int row = (p0 + j) * 8 + rsub;
                    if (row < nq) {
                        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(ov[j][0]), "=r"(ov[j][1]), "=r"(ov[j][2]), "=r"(ov[j][3]), "=r"(ov[j][4]), "=r"(ov[j][5]),
                                       "=r"(ov[j][6]), "=r"(ov[j][7])
                                     : "l"(go + (long long)row * orow));
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) ov[j][i] = 0u;
                    }
                }
            };
            load_batch(0);                                      // O does not depend on the TMA: its latency hides behind the wait below
            {   // pull the NEXT head's O rows (and lse) into L2 now, so that its three load batches are L2 hits instead of HBM round trips
                const int nhead = head + (int)gridDim.x;
                if (nhead < args.total_heads) {
                    const Item nx = decode_item<GEN>(args, nhead);
                    const int nq0 = nx.qb * 128, nnq = GEN ? min(128, S - nq0) : S;
                    const __nv_bfloat16* no = args.o + ((long long)nx.b * args.batch_stride) * args.ldo + (long long)nq0 * orow + nx.h * 64;
#pragma unroll
                    for (int r = 0; r < 7; ++r) {
                        const int row = lane + 32 * r;
                        if (row < nnq) asm volatile("prefetch.global.L2 [%0];" ::"l"(no + (long long)row * orow));
                    }
                    if (lane * 32 < nnq) asm volatile("prefetch.global.L2 [%0];" ::"l"(args.lse + (long long)nx.bh * S + nq0 + lane * 32));
                }
            }
            float lv[7];
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const int i = lane + 32 * r;
                lv[r] = i < nq ? -__ldg(gl + i) : -INFINITY;   // -inf => P = 0 for padded queries
            }
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const int i = lane + 32 * r;
                if (i < (int)kMaxQ) nl[i] = lv[r];
            }
            mbar_wait(&qdo_full[qs], (hc >> 1) & 1);          // dO has landed
            VB_DBG(sdbg, args.dbg[hc * 2 * 16 + 14]);
            const uint8_t* sdo_p = smem + kQdoOff + qs * 2 * kQSlot + kQSlot;
            // acc += a.lo * b.lo + a.hi * b.hi with bf16 operands and fp32 accumulation: the mixed-precision FMA of sm_100
            // (FHFMA.BF16 with half-register selectors) needs no unpacking — one instruction per multiply-add
            auto dot2 = [](float acc, uint32_t a, uint32_t bb) {
                asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
                    "fma.rn.f32.bf16 %0, al, bl, %0;\n\tfma.rn.f32.bf16 %0, ah, bh, %0;\n\t}"
                    : "+f"(acc) : "r"(a), "r"(bb));
                return acc;
            };
            for (int p0 = 0; p0 < 27; p0 += kBatch) {
                if (p0 > 0) load_batch(p0);
                // the nine passes of a batch are independent: plain (schedulable) shared loads, nine accumulators, shuffles at the end
                float accs[kBatch];
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    const int row = min((p0 + j) * 8 + rsub, (int)kMaxQ - 1);
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {               // this lane's two 16-byte chunks of the swizzled dO row
                        const uint4 dd = *reinterpret_cast<const uint4*>(sdo_p + row * 128 + (((csub * 2 + c) ^ (row & 7)) << 4));
                        a0 = dot2(dot2(a0, dd.x, ov[j][4 * c]), dd.y, ov[j][4 * c + 1]);
                        a1 = dot2(dot2(a1, dd.z, ov[j][4 * c + 2]), dd.w, ov[j][4 * c + 3]);
                    }
                    accs[j] = a0 + a1;
                }
#pragma unroll
                for (int j = 0; j < kBatch; ++j) accs[j] += __shfl_xor_sync(0xffffffffu, accs[j], 1);
#pragma unroll
                for (int j = 0; j < kBatch; ++j) accs[j] += __shfl_xor_sync(0xffffffffu, accs[j], 2);
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    const int row = (p0 + j) * 8 + rsub;
                    if (csub == 0 && row < (int)kMaxQ) dl[row] = row < nq ? -accs[j] : 0.f;   // stored negated: dS = P o (dP + (-delta))
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&stats_full[qs]);
            VB_DBG(sdbg, args.dbg[hc * 2 * 16 + 15]);
        }
    } else if (warp_idx == 2) {
        // ---------------- store warp: output tiles (shared memory) -> global by TMA, then recycle the buffers ----------------
        int hc = 0, ic = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
            const Item it = decode_item<GEN>(args, head);
            const int qs = hc & 1, b = it.b, h = it.h;
            for (int t = 0; t < n_t; ++t, ++ic) {
                const int ks = ic & 1;
                mbar_wait(out_ready, ic & 1);
                const bool last = (t == n_t - 1);
                uint8_t* kb = smem + kKvOff + ks * 2 * kBlk;
                uint8_t* qb = smem + kQdoOff + qs * 2 * kQSlot;
                if (lane == 0) {
                    if (GEN) {   // partial sums over this item's query block: fp32 reduce-add into the accumulators (tmDV / tmDK are fp32 maps)
                        tma_reduce_add_3d(&tmDV, kb, h * 64, t * 128, b);
                        tma_reduce_add_3d(&tmDV, kb + kBlk, h * 64 + 32, t * 128, b);
                        tma_reduce_add_3d(&tmDK, smem + L::kDk, h * 64, t * 128, b);
                        tma_reduce_add_3d(&tmDK, smem + L::kDk + kBlk, h * 64 + 32, t * 128, b);
                    } else {
                        tma_store_3d(&tmDV, kb, h * 64, t * 128, b);
                        tma_store_3d(&tmDK, kb + kBlk, h * 64, t * 128, b);
                    }
                    if (last) {
                        tma_store_3d(&tmDQ, qb, h * 64, it.qb * 128, b);
                        if (has_q1) tma_store_3d(&tmDQ, qb + kBlk, h * 64, 128, b);
                    }
                    tma_store_commit();
                }
                __syncwarp();
                if (!GEN && args.colsum != nullptr) {
                    // In-projection bias gradient (torch/nn/functional.py:5835-5847): column sums of the staged bf16 dV / dQ tiles, taken
                    // HERE — by the otherwise idle store warp, while the TMA engine reads the same tiles — instead of by butterfly
                    // shuffles in the read-out of the element-wise warps (that was 60 us of a 240 us launch, on the critical path).
                    // The dK part is skipped: sum_k dS[q,k] = sum_k P (dP - delta) = 0 for every query, so the column sum of
                    // dK = dS^T Q (the key-bias gradient) is exactly zero — softmax ignores a constant key offset.
                    auto colsum_tile = [&](const uint8_t* tile, int valid_rows, float* dst) {
                        const uint32_t base = smem_u32(tile);
                        const int c = lane & 7, rg = lane >> 3;
                        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                        for (int i = 0; i < 32; ++i) {
                            const int r = 4 * i + rg;
                            if (r < valid_rows) {
                                uint32_t w[4];
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                                             : "r"(base + r * 128 + ((c ^ (r & 7)) << 4)));
#pragma unroll
                                for (int q = 0; q < 4; ++q)   // fp32 += bf16 (FHADD.BF16 with half-register selectors: no unpacking)
                                    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}"
                                        : "+f"(acc[2 * q]), "+f"(acc[2 * q + 1]) : "r"(w[q]));
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 8);
                            acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 16);
                        }
                        if (rg == 0) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) atomicAdd(dst + c * 8 + q, acc[q]);
                        }
                    };
                    const int hd = args.H * 64;
                    colsum_tile(kb, min(128, S - t * 128), args.colsum + 2 * hd + h * 64);
                    if (last) {
                        colsum_tile(qb, min(128, S), args.colsum + h * 64);
                        if (has_q1) colsum_tile(qb + kBlk, S - 128, args.colsum + h * 64);
                    }
                    __syncwarp();
                }
                if (lane == 0) {
                    tma_store_wait_read<0>();
                    mbar_arrive(&kv_empty[ks]);
                    if (GEN) mbar_arrive(dk_free);
                    if (last) mbar_arrive(&qdo_empty[qs]);
                }
                __syncwarp();
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    } else if (warp_idx == 1) {
        // ---------------- MMA issuer: the whole warp walks the loops (uniform control flow, descriptors in uniform
        // registers); one elected lane issues the tcgen05 instructions ----------------
        const int nA = jA * 16, nB = npad - nA;
        const uint32_t idesc_sA = umma_idesc_bf16(128, nA, 0, 0);
        const uint32_t idesc_sB = umma_idesc_bf16(128, nB > 0 ? nB : 16, 0, 0);
        constexpr uint32_t idesc_kn = umma_idesc_bf16(128, 64, 0, 1);   // A K-major (or TMEM), B MN-major
        constexpr uint32_t idesc_mn = umma_idesc_bf16(128, 64, 1, 1);   // A MN-major, B MN-major
        constexpr uint64_t kdesc = umma_smem_desc_base(0, 1024);        // K-major SW128
        constexpr uint64_t mdesc = umma_smem_desc_base(kBlk, 1024);     // MN-major SW128, 64-wide atoms one block apart
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t stg = smem_u32(smem + kStagingOff);
        int hc = 0, ic = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
            const int qs = hc & 1;
            mbar_wait(&qdo_full[qs], (hc >> 1) & 1);
            const uint32_t sQ = smem_u32(smem + kQdoOff + qs * 2 * kQSlot), sdO = sQ + kQSlot;
            for (int t = 0; t < n_t; ++t, ++ic) {
                const int ks = ic & 1;
                const uint32_t sK = smem_u32(smem + kKvOff + ks * 2 * kBlk), sV = sK + kBlk;
                mbar_wait(&kv_full[ks], (ic >> 1) & 1);
                tcgen05_fence_after();
                const bool dbg_on = args.dbg && blockIdx.x == 0 && ic < 64 && lane == 0;
                VB_DBG(dbg_on, args.dbg[ic * 16 + 0]);
                // S^T = K_t Q^T goes first and does not wait for the previous tile's read-out: its columns only hold the packed P^T that the
                // dV products in front of it in the (in-order) tensor pipe consume, so it runs while the accumulators are still being read.
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tb + kColST, umma_smem_desc(kdesc, sK + k * 32), umma_smem_desc(kdesc, sQ + k * 32), idesc_sA, k > 0 ? 1u : 0u);
                    if (nB > 0) {   // queries >= 128: rows 128.. of Q / dO, columns 128.. of the score regions
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss(tb + kColST + 128, umma_smem_desc(kdesc, sK + k * 32), umma_smem_desc(kdesc, sQ + kBlk + k * 32), idesc_sB, k > 0 ? 1u : 0u);
                    }
                }
                __syncwarp();
                // dP^T = V_t dO^T lands on the columns that hold the previous tile's dV / dK / dQ accumulators: after the read-out
                mbar_wait(tile_free, (ic & 1) ^ 1);
                tcgen05_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tb + kColDPT, umma_smem_desc(kdesc, sV + k * 32), umma_smem_desc(kdesc, sdO + k * 32), idesc_sA, k > 0 ? 1u : 0u);
                    umma_commit(&s_full[0]);
                    if (nB > 0) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss(tb + kColDPT + 128, umma_smem_desc(kdesc, sV + k * 32), umma_smem_desc(kdesc, sdO + kBlk + k * 32), idesc_sB, k > 0 ? 1u : 0u);
                        umma_commit(&s_full[1]);
                    }
                }
                __syncwarp();
                VB_DBG(dbg_on, args.dbg[ic * 16 + 1]);
                const uint64_t bd_do = umma_smem_desc(mdesc, sdO), bd_q = umma_smem_desc(mdesc, sQ), bd_k = umma_smem_desc(mdesc, sK);
                const uint64_t ad_q0 = umma_smem_desc(mdesc, stg);
                const uint64_t ad_q1 = umma_smem_desc(umma_smem_desc_base(sV - (stg + 2 * kBlk), 1024), stg + 2 * kBlk);
                const uint64_t ad_k = umma_smem_desc(kdesc, stg), ad_kv = umma_smem_desc(kdesc, sV);
                const int ksteps = (min(npad_k - t * 128, 128)) >> 4;
                // ---- first part: queries < 128 ----
                mbar_wait(&p_full[0], ic & 1);
                tcgen05_fence_after();
                VB_DBG(dbg_on, args.dbg[ic * 16 + 2]);
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)   // dV_t = P^T dO: A = packed P^T of group j at TMEM column 16 j
                        if (j < jA) umma_bf16_ts(tb + kColDV, tb + kColST + 16 * j, bd_do + (uint64_t)(j * 128), idesc_kn, j > 0 ? 1u : 0u);
#pragma unroll
                    for (int j = 0; j < 8; ++j)   // dK_t = dS^T Q: A = staging read K-major (64 queries per block)
                        if (j < jA)
                            umma_bf16_ss(tb + kColDK, ad_k + (uint64_t)((j >> 2) * (kBlk >> 4) + (j & 3) * 2), bd_q + (uint64_t)(j * 128), idesc_kn, j > 0 ? 1u : 0u);
#pragma unroll
                    for (int j = 0; j < 8; ++j)   // dQ(0..127) += dS K_t: A = staging blocks 0,1 read MN-major, B = K_t MN-major
                        if (j < ksteps)
                            umma_bf16_ss(tb + kColDQ0, ad_q0 + (uint64_t)(j * 128), bd_k + (uint64_t)(j * 128), idesc_mn, (t > 0 || j > 0) ? 1u : 0u);
                }
                __syncwarp();
                // ---- second part: queries >= 128 ----
                mbar_wait(&p_full[1], ic & 1);
                tcgen05_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 8; j < (NKS_T ? NKS_T : 13); ++j)
                        if (NKS_T || j < nks) umma_bf16_ts(tb + kColDV, tb + kColST + 16 * j, bd_do + (uint64_t)(j * 128), idesc_kn, 1u);
#pragma unroll
                    for (int j = 8; j < (NKS_T ? NKS_T : 13); ++j)   // block 3 of the staging (queries 192..) is the V_t buffer
                        if (NKS_T || j < nks)
                            umma_bf16_ss(tb + kColDK, (j < 12 ? ad_k + (uint64_t)((j >> 2) * (kBlk >> 4) + (j & 3) * 2) : ad_kv), bd_q + (uint64_t)(j * 128),
                                         idesc_kn, 1u);
                    if (has_q1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)   // dQ(128..) = dS K_t: staging blocks 2,3
                            if (j < ksteps)
                                umma_bf16_ss(tb + kColDQ1, ad_q1 + (uint64_t)(j * 128), bd_k + (uint64_t)(j * 128), idesc_mn, j > 0 ? 1u : 0u);
                    }
                    umma_commit(o_full);
                }
                __syncwarp();
                VB_DBG(dbg_on, args.dbg[ic * 16 + 3]);
            }
        }
    } else if (warp_idx >= 4) {
        // ---------------- element-wise + read-out: 384 threads, three per tile row; 16-query group g belongs to part g % 3 ----------------
        const uint32_t quad = warp_idx & 3, part = (warp_idx - 4) >> 2;
        const uint32_t t_lane = tmem_base + ((quad * 32) << 16);
        const int row_in_tile = quad * 32 + lane;
        const float c = args.scale_log2;
        const uint32_t swz = (uint32_t)(row_in_tile & 7);
        const uint32_t stg_row = smem_u32(smem + kStagingOff) + row_in_tile * 128;
        const uint32_t stats_u32 = smem_u32(stats);
        const uint32_t drop_key = DROP ? dropout_key(*args.drop_seed, args.drop_stream) : 0u;
        const f2 c2 = f2_pack(c, c);
        int hc = 0, ic = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
            const Item it = decode_item<GEN>(args, head);
            const int qs = hc & 1;
            const uint32_t nls = stats_u32 + qs * 2 * kMaxQ * 4;
            const uint32_t dls = nls + kMaxQ * 4;
            const uint32_t q_row = smem_u32(smem + kQdoOff + qs * 2 * kQSlot) + row_in_tile * 128;   // dQ tiles park in the Q / dO stage
            for (int t = 0; t < n_t; ++t, ++ic) {
                const int ks = ic & 1;
                const uint32_t kv_row = smem_u32(smem + kKvOff + ks * 2 * kBlk) + row_in_tile * 128;   // K_t row; V_t row = + kBlk
                // element (q, key) of the head -> base + q * (number of keys)
                const uint32_t kstride = GEN ? (uint32_t)args.Sk : (uint32_t)S;
                const uint32_t drop_base = ((uint32_t)it.bh * (uint32_t)S + (uint32_t)(it.qb * 128)) * kstride + (uint32_t)(t * 128 + row_in_tile);
                bool key_off = false;      // GEN: this row's key is padded or masked: P = dS = 0
                if (GEN) {
                    const int key = t * 128 + row_in_tile;
                    key_off = key >= args.Sk || (args.kpm != nullptr && args.kpm[(long long)it.b * args.Sk + key] != 0);
                }
                const bool dbg_on = args.dbg && blockIdx.x == 0 && ic < 64 && warp_idx == 4 && lane == 0;
                if (t == 0) mbar_wait(&stats_full[qs], (hc >> 1) & 1);   // -lse / delta of this head (written a head ahead)
                mbar_wait(&s_full[0], ic & 1);
                tcgen05_fence_after();
                VB_DBG(dbg_on, args.dbg[ic * 16 + 4]);
                bool arrivedA = false, waitedB = false;
                auto arriveA = [&]() {
                    tmem_st_wait();
                    fence_proxy_async_smem();   // staging writes must be visible to the tensor core's (async proxy) reads
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[0]);
                    arrivedA = true;
                };
                // Software-pipelined over this warp's 16-query groups (g, g + 3, g + 6, ...): the scores of the NEXT group are prefetched
                // from TMEM while the current group is exponentiated; the current group's score gradients are loaded at the top of its
                // own step, under the MUFU latency of the exponentials (a second prefetch buffer for them costs 16 registers this kernel
                // does not have: 512 threads x 128).
                auto wait_b = [&](int g) {
                    if (g >= 8 && !waitedB) {      // queries >= 128: their scores are committed separately
                        mbar_wait(&s_full[1], ic & 1);
                        tcgen05_fence_after();
                        waitedB = true;
                    }
                };
                // svc: scores of group g (already requested); svn: buffer for the scores of group g + 3
                auto step = [&](const uint32_t (&svc)[16], uint32_t (&svn)[16], int g) {
                    uint32_t dv[16], pp[8], pd[8], keep;
                    f2 pf[8];
                    tmem_ld_wait();                                            // svc has landed
                    tmem_ld_32x32b_x16(t_lane + kColDPT + g * 16, dv);
                    if (g + 3 < nks) {
                        wait_b(g + 3);
                        tmem_ld_32x32b_x16(t_lane + kColST + (g + 3) * 16, svn);
                    }
                    ew_phase1<DROP>(svc, nls + g * 64, c2, pf, pp, keep, drop_key, drop_base + (uint32_t)(g * 16) * kstride, kstride,
                                    args.drop_thresh, args.drop_inv_keep);
                    tmem_ld_wait();                                            // dv (and the prefetched scores) have landed
                    ew_phase2<DROP>(pf, dv, dls + g * 64, pd, keep, args.drop_inv_keep);
                    if (GEN && key_off) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) pp[i] = pd[i] = 0u;
                    }
                    tmem_st_32x32b_x8(t_lane + kColST + g * 16, pp);
                    const uint32_t dst = (g < 12) ? stg_row + (g >> 2) * kBlk : kv_row + kBlk;   // block 3 aliases V_t
                    const uint32_t ch = (uint32_t)(g & 3) * 2;
                    sts128(dst + ((ch ^ swz) << 4), pd[0], pd[1], pd[2], pd[3]);
                    sts128(dst + (((ch + 1) ^ swz) << 4), pd[4], pd[5], pd[6], pd[7]);
                    VB_DBG(dbg_on, args.dbg[ic * 16 + 8 + g / 3]);
                    if (g < 8 && g + 3 >= 8 && !arrivedA) arriveA();   // that was this warp's last group of the first part
                };
                {
                    uint32_t svA[16], svB[16];
                    int g = part;
                    if (g < nks) {
                        wait_b(g);
                        tmem_ld_32x32b_x16(t_lane + kColST + g * 16, svA);
                    }
                    for (; g < nks; g += 6) {
                        step(svA, svB, g);
                        if (g + 3 < nks) step(svB, svA, g + 3);
                    }
                }
                if (!arrivedA) arriveA();
                tmem_st_wait();
                fence_proxy_async_smem();
                tcgen05_fence_before();
                __syncwarp();
                VB_DBG(dbg_on, args.dbg[ic * 16 + 5]);
                if (lane == 0) mbar_arrive(&p_full[1]);
                // ---- read-out: accumulator slices -> bf16 tiles in the dead K_t / V_t (dV, dK) and Q / dO (dQ) buffers ----
                mbar_wait(o_full, ic & 1);
                tcgen05_fence_after();
                VB_DBG(dbg_on, args.dbg[ic * 16 + 6]);
                const bool last = (t == n_t - 1);
                auto stage32 = [&](const uint32_t (&r)[32], uint32_t row_addr, uint32_t hi, float mul) {
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
                        const uint32_t w0 = pack2(__uint_as_float(r[v4 * 8 + 0]) * mul, __uint_as_float(r[v4 * 8 + 1]) * mul);
                        const uint32_t w1 = pack2(__uint_as_float(r[v4 * 8 + 2]) * mul, __uint_as_float(r[v4 * 8 + 3]) * mul);
                        const uint32_t w2 = pack2(__uint_as_float(r[v4 * 8 + 4]) * mul, __uint_as_float(r[v4 * 8 + 5]) * mul);
                        const uint32_t w3 = pack2(__uint_as_float(r[v4 * 8 + 6]) * mul, __uint_as_float(r[v4 * 8 + 7]) * mul);
                        sts128(row_addr + (((hi * 4 + v4) ^ swz) << 4), w0, w1, w2, w3);
                    }
                };
                // GEN: a 32-column fp32 slice of dV_t / dK_t (partial over the item's query block) -> swizzled 128-byte rows for the TMA
                // reduce-add into the global accumulators
                auto stage32f = [&](const uint32_t (&r)[32], uint32_t row_addr, float mul) {
#pragma unroll
                    for (int v4 = 0; v4 < 8; ++v4)
                        sts128(row_addr + (((uint32_t)v4 ^ swz) << 4), __float_as_uint(__uint_as_float(r[4 * v4]) * mul),
                               __float_as_uint(__uint_as_float(r[4 * v4 + 1]) * mul), __float_as_uint(__uint_as_float(r[4 * v4 + 2]) * mul),
                               __float_as_uint(__uint_as_float(r[4 * v4 + 3]) * mul));
                };
                const uint32_t dk_row = smem_u32(smem + L::kDk) + row_in_tile * 128;   // GEN: dK_t columns 0..31; 32..63 at + kBlk
                // slices: part 0: dV lo, dK hi, [dQ0 lo]; part 1: dV hi, dQ1 lo, [dQ0 hi]; part 2: dK lo, dQ1 hi.
                // The TMEM loads come first and TMEM is handed back (tile_free) as soon as the last of them has landed, so the next tile's
                // score products run while the packing, the shared-memory stores and the column-sum butterflies are still in progress.
                auto release_tmem = [&]() {
                    if (args.late_release) return;
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tile_free);
                };
                const bool third = last && part < 2;          // this warp also reads a slice of the dQ(0..127) accumulator
                const bool second = part == 0 || has_q1;
                uint32_t ra[32], rb[32];
                tmem_ld_32x32b_x32(t_lane + (part == 0 ? kColDV : (part == 1 ? kColDV + 32 : kColDK)), ra);
                if (second) tmem_ld_32x32b_x32(t_lane + (part == 0 ? kColDK + 32 : kColDQ1 + (part - 1) * 32), rb);
                tmem_ld_wait();
                if (!third) release_tmem();
                if (GEN) {
                    if (part != 1) mbar_wait(dk_free, (ic & 1) ^ 1);   // the previous tile's dK_t has left its buffers
                    if (part == 0) stage32f(ra, kv_row, 1.0f);              // dV_t columns 0..31  -> dead K_t buffer
                    else if (part == 1) stage32f(ra, kv_row + kBlk, 1.0f);  // dV_t columns 32..63 -> dead V_t buffer
                    else stage32f(ra, dk_row, args.scale);                  // dK_t columns 0..31
                } else {
                    stage32(ra, part == 2 ? kv_row + kBlk : kv_row, part == 1 ? 1u : 0u, part == 2 ? args.scale : 1.0f);
                }
                if (third) {   // re-use ra for the dQ(0..127) slice; TMEM is free once it has landed
                    tmem_ld_32x32b_x32(t_lane + kColDQ0 + part * 32, ra);
                    tmem_ld_wait();
                    release_tmem();
                }
                if (part == 0) {
                    if (GEN) stage32f(rb, dk_row + kBlk, args.scale);          // dK_t columns 32..63
                    else stage32(rb, kv_row + kBlk, 1u, args.scale);
                } else if (has_q1) {
                    // The dQ contribution of key tile 0 to the queries >= 128 has to survive the next tile's score products, which
                    // overwrite its accumulator: it is parked as packed bf16 in the 32 TMEM columns nothing else uses (16 per part) —
                    // 32 registers less than carrying it in fp32, which is what pays for the double-buffered loads above.
                    const uint32_t t_carry = t_lane + kColCarry + (part - 1) * 16;
                    if (last) {
                        if (n_t > 1) {
                            uint32_t cr[16];
                            tmem_ld_32x32b_x16(t_carry, cr);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                rb[2 * i] = __float_as_uint(__uint_as_float(rb[2 * i]) + __uint_as_float(cr[i] << 16));
                                rb[2 * i + 1] = __float_as_uint(__uint_as_float(rb[2 * i + 1]) + __uint_as_float(cr[i] & 0xFFFF0000u));
                            }
                        }
                        stage32(rb, q_row + kBlk, part - 1, args.scale);
                    } else {
                        uint32_t cr[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) cr[i] = pack2(__uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1]));
                        tmem_st_32x32b_x16(t_carry, cr);
                        tmem_st_wait();
                    }
                }
                if (third) {
                    stage32(ra, q_row, part, args.scale);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (args.late_release) mbar_arrive(tile_free);
                    mbar_arrive(out_ready);   // output tiles staged: the store warp takes over
                }
                VB_DBG(dbg_on, args.dbg[ic * 16 + 7]);
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

}  // namespace bwd5

// Returns VB_OK if launched, 1 if this shape is not handled here.  delta = rowsum(dO o O) is computed inside (from d->o and d->dout).
int attention_bwd_tc5(const VbAttnDesc* d, cudaStream_t stream) {
    using namespace bwd5;
    if (d->S > (int)kMaxQ || d->head_dim != 64 || d->tok_stride != 1 || d->key_padding_mask != nullptr) return 1;
    VB_REQUIRE(d->o != nullptr && d->ldo % 16 == 0 && (reinterpret_cast<uintptr_t>(d->o) & 31) == 0,
               "attention_bwd: the forward output o (32-byte aligned, pitch % 16 == 0) is needed for delta = rowsum(dO o O)");
    const int S = d->S, npad = (S + 15) / 16 * 16;
    Args a{};
    a.B = d->B; a.H = d->H; a.S = S; a.npad = npad; a.nks = npad / 16;
    a.n_t = (S + 127) / 128;
    a.total_heads = d->B * d->H;
    a.scale = 0.125f; a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.lse = d->lse;
    a.o = reinterpret_cast<const __nv_bfloat16*>(d->o); a.ldo = d->ldo;
    a.batch_stride = d->batch_stride;
    a.tok_stride = 1;
    a.colsum = d->dqkv_colsum;
    { const char* e = getenv("VITB200_ATTN_BWD_LATE_RELEASE"); a.late_release = (e && e[0] == '1') ? 1 : 0; }
    a.dbg = g_dbg;
    CUtensorMap tq, tk, tv, tdo, tdq, tdk, tdv;
    const uint64_t cols = (uint64_t)d->H * 64;
    int rc;
    if ((rc = make_tmap_3d(&tq, VB_BF16, d->q, cols, S, d->B, d->ldq, d->batch_stride * d->ldq, 64, npad))) return rc;
    if ((rc = make_tmap_3d(&tk, VB_BF16, d->k, cols, S, d->B, d->ldk, d->batch_stride * d->ldk, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tv, VB_BF16, d->v, cols, S, d->B, d->ldv, d->batch_stride * d->ldv, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tdo, VB_BF16, d->dout, cols, S, d->B, d->lddo, d->batch_stride * d->lddo, 64, npad))) return rc;
    // outputs: 128-row tiles, rows >= S are clipped by the TMA store
    if ((rc = make_tmap_3d(&tdq, VB_BF16, d->dq, cols, S, d->B, d->lddq, d->batch_stride * d->lddq, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tdk, VB_BF16, d->dk, cols, S, d->B, d->lddk, d->batch_stride * d->lddk, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tdv, VB_BF16, d->dv, cols, S, d->B, d->lddv, d->batch_stride * d->lddv, 64, 128))) return rc;
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
#define VB_BWD_LAUNCH(NKS, DR)                                                                                              \
    do {                                                                                                                    \
        static DeviceOnce configured;                                                                                     \
        if (!configured.is_set()) {                                                                                                  \
            VB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc5_kernel<NKS, DR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<false>::kSmem)); \
            configured.set();                                                                                              \
        }                                                                                                                   \
        attn_bwd_tc5_kernel<NKS, DR, false><<<grid, kThreads, Lay<false>::kSmem, stream>>>(tq, tk, tv, tdo, tdq, tdk, tdv, a); \
    } while (0)
    if (a.nks == 13) {
        if (drop) VB_BWD_LAUNCH(13, true); else VB_BWD_LAUNCH(13, false);
    } else {
        if (drop) VB_BWD_LAUNCH(0, true); else VB_BWD_LAUNCH(0, false);
    }
#undef VB_BWD_LAUNCH
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// General path (GEN instantiations): any S / S_kv, key-padding masks, sequence-first strides, cross-attention
// ---------------------------------------------------------------------------------------------------------------------------
namespace bwd5 {
// out[b, s, :] (bf16, caller's strides) = acc[(b * S + s), :] (fp32, contiguous): dK / dV after the reduce-adds of every query block
__global__ void attn_acc_to_bf16_kernel(const float4* __restrict__ acc, __nv_bfloat16* __restrict__ out, long long ldo, long long tok_stride,
                                        long long batch_stride, int B, int S, int d4) {
    const long long total = (long long)B * S * d4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long tok = i / d4;
        const int c = (int)(i - tok * d4);
        const int s = (int)(tok % S), b = (int)(tok / S);
        const float4 v = __ldg(acc + i);
        uint2 w;
        w.x = pack2(v.x, v.y);
        w.y = pack2(v.z, v.w);
        *reinterpret_cast<uint2*>(out + ((long long)b * batch_stride + (long long)s * tok_stride) * ldo + c * 4) = w;
    }
}
}  // namespace bwd5

size_t attention_bwd_gen_workspace(const VbAttnDesc* d) {
    const int Sk = d->S_kv > 0 ? d->S_kv : d->S;
    return 2 * (size_t)d->B * Sk * d->H * 64 * sizeof(float);
}

int attention_bwd_gen(const VbAttnDesc* d, cudaStream_t stream) {
    using namespace bwd5;
    const int S = d->S, Sk = d->S_kv > 0 ? d->S_kv : d->S;
    const size_t need = attention_bwd_gen_workspace(d);
    VB_REQUIRE(d->workspace != nullptr && (size_t)d->workspace_bytes >= need && (reinterpret_cast<uintptr_t>(d->workspace) & 255) == 0,
               "attention_bwd: this shape needs a 256-byte aligned workspace of %zu bytes (vb_attention_workspace_bytes)", need);
    VB_REQUIRE(d->o != nullptr && (d->ldo * d->tok_stride) % 16 == 0 && d->ldo % 16 == 0 && (reinterpret_cast<uintptr_t>(d->o) & 31) == 0,
               "attention_bwd: the forward output o (32-byte aligned rows) is needed for delta = rowsum(dO o O)");
    Args a{};
    a.B = d->B; a.H = d->H; a.S = S; a.Sk = Sk; a.npad = 128; a.nks = 8;
    a.n_t = (Sk + 127) / 128;
    a.n_qb = (S + 127) / 128;
    const long long items = (long long)d->B * d->H * a.n_qb;
    VB_REQUIRE(items < (1ll << 31), "attention_bwd: too many work items");
    a.total_heads = (int)items;
    a.scale = 0.125f; a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.lse = d->lse;
    a.o = reinterpret_cast<const __nv_bfloat16*>(d->o); a.ldo = d->ldo;
    a.batch_stride = d->batch_stride; a.tok_stride = d->tok_stride;
    a.kpm = d->key_padding_mask;
    a.colsum = nullptr;
    a.late_release = 0;
    a.dbg = nullptr;
    const size_t half = (size_t)d->B * Sk * d->H * 64;
    a.dk_acc = reinterpret_cast<float*>(d->workspace);
    a.dv_acc = a.dk_acc + half;
    VB_CUDA_CHECK(cudaMemsetAsync(d->workspace, 0, need, stream));
    CUtensorMap tq, tk, tv, tdo, tdq;
    const uint64_t cols = (uint64_t)d->H * 64;
    int rc;
    if ((rc = make_tmap_3d(&tq, VB_BF16, d->q, cols, S, d->B, d->tok_stride * d->ldq, d->batch_stride * d->ldq, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tk, VB_BF16, d->k, cols, Sk, d->B, d->tok_stride * d->ldk, d->batch_stride * d->ldk, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tv, VB_BF16, d->v, cols, Sk, d->B, d->tok_stride * d->ldv, d->batch_stride * d->ldv, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tdo, VB_BF16, d->dout, cols, S, d->B, d->tok_stride * d->lddo, d->batch_stride * d->lddo, 64, 128))) return rc;
    if ((rc = make_tmap_3d(&tdq, VB_BF16, d->dq, cols, S, d->B, d->tok_stride * d->lddq, d->batch_stride * d->lddq, 64, 128))) return rc;
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
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc5_kernel<8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<true>::kSmem));
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc5_kernel<8, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<true>::kSmem));
        configured.set();
    }
    // here the dK / dV maps address the fp32 accumulators ([B, Sk, H * 64] contiguous; 32-column boxes for the TMA reduce-adds)
    CUtensorMap tdk, tdv;
    if ((rc = make_tmap_3d(&tdk, VB_F32, a.dk_acc, cols, Sk, d->B, cols, (uint64_t)Sk * cols, 32, 128))) return rc;
    if ((rc = make_tmap_3d(&tdv, VB_F32, a.dv_acc, cols, Sk, d->B, cols, (uint64_t)Sk * cols, 32, 128))) return rc;
    if (drop) attn_bwd_tc5_kernel<8, true, true><<<grid, kThreads, Lay<true>::kSmem, stream>>>(tq, tk, tv, tdo, tdq, tdk, tdv, a);
    else attn_bwd_tc5_kernel<8, false, true><<<grid, kThreads, Lay<true>::kSmem, stream>>>(tq, tk, tv, tdo, tdq, tdk, tdv, a);
    VB_CUDA_CHECK(cudaGetLastError());
    const int d4 = (int)(cols / 4);
    const long long n4 = (long long)d->B * Sk * d4;
    int cgrid = (int)((n4 + 255) / 256);
    if (cgrid > num_sms() * 8) cgrid = num_sms() * 8;
    attn_acc_to_bf16_kernel<<<cgrid, 256, 0, stream>>>(reinterpret_cast<const float4*>(a.dk_acc), reinterpret_cast<__nv_bfloat16*>(d->dk), d->lddk,
                                                       d->tok_stride, d->batch_stride, d->B, Sk, d4);
    attn_acc_to_bf16_kernel<<<cgrid, 256, 0, stream>>>(reinterpret_cast<const float4*>(a.dv_acc), reinterpret_cast<__nv_bfloat16*>(d->dv), d->lddv,
                                                       d->tok_stride, d->batch_stride, d->B, Sk, d4);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

void attention_bwd_tc5_set_debug(long long* p) { bwd5::g_dbg = p; }

}  // namespace vb
