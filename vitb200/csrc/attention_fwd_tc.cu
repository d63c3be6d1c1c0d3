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
// Replaces F.scaled_dot_product_attention reached from nn.MultiheadAttention (vanilla_vit.py:77,
// torch/nn/functional.py:6676-6688).  The mma.sync kernels in attention.cu remain for S > 208 and key-padding masks.
#include <cuda.h>
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"
#include "attn_tc_common.cuh"
#include "dropout.cuh"

namespace vb {

int make_tmap_3d(CUtensorMap* m, int dtype, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
                 uint64_t batch_stride_elems, uint32_t box0, uint32_t box1);   // gemm.cu

namespace fwd3 {
using namespace atc;

constexpr int kEwWarps = 16;                     // two softmax groups x (4 lane quadrants x 2 column parts)
constexpr int kRoWarps = 4;                      // read-out warps (one per TMEM lane quadrant): O / l -> bf16 tile, lse
constexpr int kThreads = 128 + (kEwWarps + kRoWarps) * 32;   // warps 0-3: TMA loader, MMA issuer, TMEM alloc + store warp, idle
constexpr uint32_t kMaxQ = 208;
constexpr uint32_t kOpBytes = kMaxQ * 128;      // Q, K or V of one head
constexpr uint32_t kStageBytes = 3 * kOpBytes;
constexpr uint32_t kOutOff = 2 * kStageBytes;   // 2 output tiles of 128 rows x 128 B
constexpr uint32_t kXmOff = kOutOff + 2 * 16384;   // partial maxima  [4 tile slots][2 parts][128 rows]
constexpr uint32_t kXsOff = kXmOff + 4 * 2 * 128 * 4;
constexpr uint32_t kBarOff = kXsOff + 4 * 2 * 128 * 4;
constexpr uint32_t kSmemBytes = kBarOff + 256 + 1024;
static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
constexpr uint32_t kColO = 416;

struct Args {
    int B, H, S, nks, n_qt, total_heads;
    float scale_log2;
    float* lse;
    long long* dbg;   // optional in-kernel cycle stamps (tools/attn_timeline.py)
    uint32_t drop_thresh, drop_stream;   // attention dropout (DROP instantiations): keep iff hash >= thresh, P *= 1 / (1 - p)
    float drop_inv_keep;
    const uint32_t* drop_seed;
};

template <int NKS_T, bool DROP>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Args args) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
    uint64_t* kv_full = bars;          // [2]
    uint64_t* kv_empty = bars + 2;     // [2] tcgen05.commit after the head's last P V product
    uint64_t* s_full = bars + 4;       // [2] per TMEM slot
    uint64_t* p_full = bars + 6;       // [2 slots][2] count 8 (one softmax group): keys < 128 packed / all keys packed
    uint64_t* o_full = bars + 10;
    uint64_t* o_free = bars + 11;      // count kRoWarps: the O accumulator has been read out
    uint64_t* out_ready = bars + 12;   // [2] count kRoWarps: output tile staged
    uint64_t* out_free = bars + 14;    // [2] count 1: the TMA store has finished reading the tile
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

    const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = args.S, n_qt = args.n_qt;
    const int nks = NKS_T ? NKS_T : args.nks;
    const int npad = nks * 16;
    const int jA = nks < 8 ? nks : 8;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); mbar_init(&s_full[i], 1);
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

    if (warp_idx == 0) {
        if (lane == 0) {   // ---------------- TMA loader ----------------
            int hc = 0;
            for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x, ++hc) {
                const int st = hc & 1, b = head / args.H, h = head - b * args.H;
                mbar_wait(&kv_empty[st], ((hc >> 1) & 1) ^ 1);
                uint8_t* sq = smem + st * kStageBytes;
                mbar_arrive_expect_tx(&kv_full[st], 3 * npad * 128);
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
        const uint32_t idesc_s = umma_idesc_bf16(128, npad, 0, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
        constexpr uint64_t kdesc = umma_smem_desc_base(0, 1024);            // K-major SW128 (Q, K)
        constexpr uint64_t vdesc = umma_smem_desc_base(kOpBytes, 1024);     // MN-major SW128 (V): 8-key groups 1024 B apart
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const int total_tiles = ((args.total_heads - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * n_qt;
        // issues S of tile jj (head counter jj / n_qt, query tile jj % n_qt)
        auto issue_s = [&](int jj) {
            const int hc = jj / n_qt, qt = jj - hc * n_qt, st = hc & 1;
            if (qt == 0) {
                mbar_wait(&kv_full[st], (hc >> 1) & 1);
                tcgen05_fence_after();
            }
            const uint32_t sq = smem_u32(smem + st * kStageBytes) + qt * 16384, sk = smem_u32(smem + st * kStageBytes + kOpBytes);
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
            const int hc = j / n_qt, qt = j - hc * n_qt, st = hc & 1, slot = j & 1, par = (j >> 1) & 1;
            const uint64_t bd = umma_smem_desc(vdesc, smem_u32(smem + st * kStageBytes + 2 * kOpBytes));
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
        const uint32_t xm_u32 = smem_u32(smem + kXmOff), xs_u32 = smem_u32(smem + kXsOff);
        const uint32_t drop_key = DROP ? dropout_key(*args.drop_seed, args.drop_stream) : 0u;
        int j = 0;
        for (int head = blockIdx.x; head < args.total_heads; head += gridDim.x) {
            for (int qt = 0; qt < n_qt; ++qt, ++j) {
                if ((uint32_t)(j & 1) != grp) continue;   // tile j belongs to softmax group j & 1
                const uint32_t drop_base = ((uint32_t)head * (uint32_t)S + (uint32_t)(qt * 128 + row_in_tile)) * (uint32_t)S;   // element (q, k) -> base + k
                const uint32_t t_s = t_lane + (j & 1) * kMaxQ;
                const uint32_t xoff = ((j & 3) * 2 * 128 + row_in_tile) * 4;
                const bool dbg_on = args.dbg && blockIdx.x == 0 && j < 64 && (warp_idx == 4 || warp_idx == 12) && lane == 0;
                if (dbg_on) args.dbg[j * 16 + 0] = clock64();
                mbar_wait(&s_full[j & 1], (j >> 1) & 1);
                tcgen05_fence_after();
                if (dbg_on) args.dbg[j * 16 + 1] = clock64();
                // ---- pass 1: partial row maximum over my groups (raw scores; keys >= S are dead) ----
                float mx = -INFINITY;
                {
                    // two register sets: the TMEM load of the next group is in flight while the current one is reduced
                    uint32_t ra[16], rb[16];
                    auto reduce = [&](uint32_t (&r)[16], int g) {
                        if (g * 16 + 16 > S) {
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
                if (dbg_on) args.dbg[j * 16 + 2] = clock64();
                sts32(xm_u32 + xoff + part * 512, mx);
                named_bar_sync(1 + grp * 4 + quad, 64);
                mx = fmaxf(lds32(xm_u32 + xoff), lds32(xm_u32 + xoff + 512));
                const float nm = -mx * c;
                if (dbg_on) args.dbg[j * 16 + 3] = clock64();
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
                        if (g * 16 + 16 > S) {
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
                if (dbg_on) args.dbg[j * 16 + 4] = clock64();
            }
        }
    } else if (warp_idx >= 4 + kEwWarps) {
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

static long long* g_dbg = nullptr;
}  // namespace fwd3
void attention_fwd_tc3_set_debug(long long* p) { fwd3::g_dbg = p; }

// Returns VB_OK if launched, 1 if this shape is not handled here (caller falls back to the mma.sync kernels).
int attention_fwd_tc3(const VbAttnDesc* d, cudaStream_t stream) {
    using namespace fwd3;
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("VITB200_ATTN_TC");
        enabled = (e && e[0] == '0') ? 0 : 1;
    }
    if (!enabled || d->S > (int)kMaxQ || d->head_dim != 64 || d->tok_stride != 1 || d->key_padding_mask != nullptr) return 1;
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
            VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc3_kernel<NKS, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)); \
            configured.set();                                                                                              \
        }                                                                                                                   \
        attn_fwd_tc3_kernel<NKS, DR><<<grid, kThreads, kSmemBytes, stream>>>(tq, tk, tv, to, a);                            \
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

}  // namespace vb
