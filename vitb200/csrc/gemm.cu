// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )     bf16 operands, fp32 accumulation in TMEM.
//
// One CTA per SM; with CG = 2 (default) the two CTAs of a cluster pair share 256x256 output tiles: each CTA loads its
// own 128 rows of A and HALF of the B tile, the leader issues tcgen05.mma.cta_group::2 (M = 256) that reads both
// CTAs' shared memory, halving the per-SM shared-memory traffic of the B operand.  Roles per CTA:
//   warp 0      TMA producer: fills a 4-6-stage ring of {A 128x64, B (256/CG)x64} bf16 tiles (128B swizzle)
//   warp 1      MMA issuer (leader CTA): one thread issues tcgen05.mma into one of two TMEM accumulators
//   warp 2      TMEM allocator (512 columns = 2 x 256 fp32 accumulator columns); with GemmArgs::a_colsum it then sums the
//               MN-major A tiles over k straight from the operand stages (the bias gradient of a weight-gradient GEMM)
//   warps 4-11  epilogue: tcgen05.ld -> registers -> fused math -> swizzled smem -> TMA store / reduce-add,
//               overlapping the MMA of the next tile (double-buffered accumulator)
// Operand majors are handled in the shared-memory descriptors (K-major or MN-major canonical SW128
// layouts), never by transposing in HBM: forward uses (K,K), dgrad (K,MN), wgrad (MN,MN).
//
// Replaces the cuBLASLt calls behind nn.Linear / F.linear on the reference path
// (vanilla_vit.py:33-42,77-79,212-213; torch/nn/functional.py:5835-5847,6690) and their autograd formulas.
#include <cuda.h>
#include <atomic>
#include <mutex>
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"
#include "pdl.cuh"

namespace vb {

constexpr uint32_t BM = 128, BN = 256, BK = 64, UK = 16;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr uint32_t B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr uint32_t EPI_BUF_BYTES = 32 * 128;     // 32 rows x 128 B
constexpr uint32_t kEpiWarps = 8;
constexpr uint32_t kFirstEpiWarp = 4;
constexpr uint32_t kThreads = (kFirstEpiWarp + kEpiWarps) * 32;  // 384
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSchedStages = 4;   // depth of the dynamic scheduler's unit ring

struct GemmArgs {
    int M, N, K, batches;
    int n_blocks, k_blocks, tiles_per_batch, num_tiles;
    int split_k, kb_per_split, num_units;
    int c_row_offset, aux_bcast, b_batched;
    int direct;
    int tail_cols;        // a last n-block with at most this many valid columns runs as a 256x128 MMA (128; 0 disables)
    float drelu_scale;    // VB_EPI_DRELU: kept elements are multiplied by this (1/(1-p) of the dropout after the activation)
    long long* dbg;       // optional: cycle accounting of cluster 0's MMA issuer (vb_debug_set_gemm_timeline)
    int* sched_counter;   // non-null: dynamic tile scheduling (units beyond the first per cluster are handed out by atomicAdd)
    const float* bias;
    float* a_colsum;      // non-null (AMAJ = 1, one batch): a_colsum[m] += sum over k of A[k, m], read from the operand stages (warp 2)
    void* C;
    void* C2;
    const void* AUX;
    long long ldc, ldc2, ldaux, bsc, bsc2, bsaux;
};

template <int EPI, int CG>
struct EpiTraits {
    static constexpr bool kHasAux = (EPI == VB_EPI_RESIDUAL || EPI == VB_EPI_DGELU || EPI == VB_EPI_DRELU);
    static constexpr bool kTwoOut = (EPI == VB_EPI_GELU);
    // staging buffers per epilogue warp: [aux-in] + out0 [+ out1 = second output of GELU].  Single-buffered on purpose: the MMA issuer
    // waits for a free accumulator < 1 % of the time but for operand stages ~20-35 % (tools/gemm_timeline.py), so shared memory
    // is worth more as one more pipeline stage than as a second store buffer.
    static constexpr uint32_t kBufsPerWarp = (kHasAux ? 1 : 0) + (kTwoOut ? 2 : 1);
    static constexpr uint32_t kBStageBytes = B_STAGE_BYTES / CG;
    static constexpr uint32_t kStageBytes = A_STAGE_BYTES + kBStageBytes;
    static constexpr uint32_t kEpiBytes = kEpiWarps * kBufsPerWarp * EPI_BUF_BYTES;
    static constexpr uint32_t kBudget = 232448 - 1024 /*align*/ - 512 /*barriers + scheduler ring*/;
    static constexpr uint32_t kStagesRaw = (kBudget - kEpiBytes) / kStageBytes;
    static constexpr uint32_t kStages = kStagesRaw > 7 ? 7 : kStagesRaw;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kEpiBytes + 512 + 1024;
};

// Exact-erf GELU via the Abramowitz-Stegun 7.1.26 rational approximation of erf (|abs err| < 1.5e-7, far below
// bf16 resolution): Phi(z) = 1 - q for z >= 0, q for z < 0, with q = 0.5 * poly(t) * exp(-z^2/2),
// t = 1 / (1 + 0.2316419 |z|).  One MUFU.RCP + one MUFU.EX2 + ~10 FMAs per element instead of erff()'s
// two divergent polynomial branches; the derivative shares the exponential.
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void phi_parts(float z, float& cdf, float& e) {
    const float t = fast_rcp(fmaf(0.2316419f, fabsf(z), 1.0f));
    e = fast_ex2((z * -0.72134752044448170f) * z);  // exp(-z^2 / 2)
    float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);   // coefficients pre-multiplied by 0.5
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, 0.5f * -0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    const float q = (poly * t) * e;                  // = 1 - Phi(|z|)
    // Phi(z) = 0.5 + sign(z) * (0.5 - q)
    cdf = 0.5f + __uint_as_float(__float_as_uint(0.5f - q) ^ (__float_as_uint(z) & 0x80000000u));
}
__device__ __forceinline__ float gelu_erf(float x) {
    float cdf, e;
    phi_parts(x, cdf, e);
    return x * cdf;
}
__device__ __forceinline__ float dgelu_erf(float x) {
    float cdf, e;
    phi_parts(x, cdf, e);
    return fmaf(x * 0.39894228040143268f, e, cdf);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// Packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2): one issue slot for two lanes of the epilogue math.  The
// epilogue warps are issue-bound (8 warps feed a 256x256 tile every 6144 cycles at K = 768), not FMA-pipe-bound, so
// halving the instruction count of the arithmetic is what moves the tensor pipe's duty cycle.
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_pack(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2_bcast(float x) { return f2_pack(x, x); }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
    f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
    f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
    f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// Two lanes of phi_parts(): gelu = x * Phi(x), dgelu = Phi(x) + x * pdf(x)   (same A&S 7.1.26 arithmetic as above)
__device__ __forceinline__ void gelu_pair(float x0, float x1, float& g0, float& g1, float& d0, float& d1) {
    const f2 x = f2_pack(x0, x1);
    const f2 ax = f2_pack(fabsf(x0), fabsf(x1));
    float den0, den1;
    f2_unpack(f2_fma(ax, f2_bcast(0.2316419f), f2_bcast(1.0f)), den0, den1);
    const f2 t = f2_pack(fast_rcp(den0), fast_rcp(den1));
    float a0, a1;
    f2_unpack(f2_mul(f2_mul(x, f2_bcast(-0.72134752044448170f)), x), a0, a1);
    const f2 e = f2_pack(fast_ex2(a0), fast_ex2(a1));   // exp(-x^2 / 2)
    f2 poly = f2_fma(f2_bcast(0.5f * 1.061405429f), t, f2_bcast(0.5f * -1.453152027f));
    poly = f2_fma(poly, t, f2_bcast(0.5f * 1.421413741f));
    poly = f2_fma(poly, t, f2_bcast(0.5f * -0.284496736f));
    poly = f2_fma(poly, t, f2_bcast(0.5f * 0.254829592f));
    const f2 q = f2_mul(f2_mul(poly, t), e);             // 1 - Phi(|x|)
    float h0, h1;
    f2_unpack(f2_fma(q, f2_bcast(-1.0f), f2_bcast(0.5f)), h0, h1);   // 0.5 - q
    h0 = __uint_as_float(__float_as_uint(h0) ^ (__float_as_uint(x0) & 0x80000000u));
    h1 = __uint_as_float(__float_as_uint(h1) ^ (__float_as_uint(x1) & 0x80000000u));
    const f2 cdf = f2_add(f2_pack(h0, h1), f2_bcast(0.5f));
    f2_unpack(f2_mul(x, cdf), g0, g1);
    f2_unpack(f2_fma(f2_mul(x, f2_bcast(0.39894228040143268f)), e, cdf), d0, d1);
}

struct UnitCoord {
    int m_blk, n_blk, batch, kb0, kb1;
};
// Units enumerate (k-split, batch, m-group, n-block) with n fastest; an m-group is CG consecutive 128-row blocks and
// CTA `rank` of the pair owns block m_group * CG + rank.
template <int CG>
__device__ __forceinline__ UnitCoord decode_unit(const GemmArgs& a, int u, int rank) {
    UnitCoord c;
    const int s = u / a.num_tiles;
    const int t = u - s * a.num_tiles;
    c.batch = t / a.tiles_per_batch;
    const int r = t - c.batch * a.tiles_per_batch;
    const int mg = r / a.n_blocks;
    c.m_blk = mg * CG + rank;
    c.n_blk = r - mg * a.n_blocks;
    c.kb0 = s * a.kb_per_split;
    c.kb1 = min(c.kb0 + a.kb_per_split, a.k_blocks);
    return c;
}

template <int AMAJ, int BMAJ, int EPI, int CDT, int CG>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
            const __grid_constant__ CUtensorMap tmAux, const GemmArgs args) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
    using T = EpiTraits<EPI, CG>;
    constexpr uint32_t kStages = T::kStages;
    constexpr uint32_t kBStage = T::kBStageBytes;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
    const int unit0 = blockIdx.x / CG, unit_stride = gridDim.x / CG;
    constexpr uint32_t CPC = (CDT == VB_BF16) ? 64 : 32;  // columns per 128-byte store chunk
    constexpr uint32_t kChunks = 128 / CPC;               // chunks per epilogue warp per tile

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * A_STAGE_BYTES;
    uint8_t* smem_epi = smem_b + kStages * kBStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + kEpiWarps * T::kBufsPerWarp * EPI_BUF_BYTES);
    uint64_t* full_bar = bars;                      // [kStages]
    uint64_t* empty_bar = bars + kStages;           // [kStages]
    uint64_t* tmem_full_bar = bars + 2 * kStages;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
    uint64_t* aux_bar = tmem_empty_bar + 2;         // [kEpiWarps]
    uint64_t* sched_full = aux_bar + kEpiWarps;          // [kSchedStages] scheduler ring: entry written
    uint64_t* sched_empty = sched_full + kSchedStages;   // [kSchedStages] (leader's copy) entry read by every consumer of the pair
    uint64_t* cs_done = sched_empty + kSchedStages;      // [kStages] a_colsum: the column-sum warp has finished reading the A stage
    uint64_t* cs_ready = cs_done + kStages;              // [kStages] a_colsum, non-leader CTA: the leader saw the stage's loads complete
    int* sched_unit = reinterpret_cast<int*>(cs_ready + kStages);   // [kSchedStages]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sched_unit + kSchedStages);

    const uint32_t warp_idx = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    pdl_launch_dependents();   // the next kernel's CTAs may take this SM's place as soon as this CTA exits (pdl.cuh)

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        if (T::kTwoOut) tma_prefetch_desc(&tmC2);
        if (T::kHasAux) tma_prefetch_desc(&tmAux);
    }
    if (warp_idx == 1 && lane == 0) {
        for (uint32_t i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], kEpiWarps * CG);   // (leader's copy) every epilogue warp of the pair arrives
        }
        for (uint32_t i = 0; i < kEpiWarps; ++i) mbar_init(&aux_bar[i], 1);
        const uint32_t cs_readers = (AMAJ == 1 && args.a_colsum != nullptr) ? 1 : 0;
        for (uint32_t i = 0; i < kSchedStages; ++i) {
            mbar_init(&sched_full[i], 1);
            // producer + epilogue warps (+ column-sum warp) of each CTA, MMA issuer of the leader
            mbar_init(&sched_empty[i], (1 + kEpiWarps + cs_readers) * CG + 1);
        }
        for (uint32_t i = 0; i < kStages; ++i) {
            mbar_init(&cs_done[i], 1);
            mbar_init(&cs_ready[i], 1);
        }
        fence_barrier_init();
    }
    if (warp_idx == 2) {
        if (CG == 2) tmem_alloc_2sm<kTmemCols>(tmem_ptr_smem);
        else tmem_alloc<kTmemCols>(tmem_ptr_smem);
    }
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all();   // peers' barriers are initialised before anyone signals them
    else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();                // everything above is independent of the kernels in front; from here on their results are needed

    // ---- tile scheduling ----------------------------------------------------------------------------------------------
    // Static: cluster c owns units c, c + #clusters, ...  Dynamic (args.sched_counter): the first unit is still the cluster index,
    // later ones come from a global atomic counter fetched by the scheduler warp of the leader CTA and broadcast to every role of
    // both CTAs through a small shared-memory ring — so a cluster that starts late (its SMs were held by a co-running kernel,
    // e.g. the NCCL all-reduce of the previous gradient bucket) simply takes fewer units instead of finishing a wave late.
    const bool dynamic = args.sched_counter != nullptr;
    const bool colsum_on = (AMAJ == 1) && args.a_colsum != nullptr;
    // a_colsum: k-block kb of a unit is summed by the unit whose n-block is kb mod n_blocks.  The roles keep a wrapping counter
    // (0 = "mine") instead of taking a modulo per k-block: that sat on the issue path of the TMA producer and the MMA issuer.
    auto cs_count0 = [&](const UnitCoord& c) -> int {
        if (!colsum_on) return 0;
        const int r = (c.kb0 - c.n_blk) % args.n_blocks;
        return r < 0 ? r + args.n_blocks : r;
    };
    const uint32_t sched_empty_leader = (CG == 2) ? mapa_u32(smem_u32(&sched_empty[0]), 0) : smem_u32(&sched_empty[0]);
    struct SchedReader {
        uint32_t stage = 0, phase = 0;
    };
    // next unit for this role (-1 = none); called by one lane (producer, MMA) or by the whole warp (epilogue: lane 0 releases)
    auto sched_next = [&](SchedReader& rd, int u, bool whole_warp) -> int {
        if (!dynamic) {
            const int n = u + unit_stride;
            return n < args.num_units ? n : -1;
        }
        mbar_wait(&sched_full[rd.stage], rd.phase);
        const int n = *reinterpret_cast<volatile int*>(&sched_unit[rd.stage]);
        if (whole_warp) __syncwarp();
        if (!whole_warp || lane == 0) {
            if (CG == 2 && rank != 0) mbar_arrive_cluster(sched_empty_leader + rd.stage * 8);
            else mbar_arrive(&sched_empty[rd.stage]);
        }
        if (++rd.stage == kSchedStages) { rd.stage = 0; rd.phase ^= 1; }
        return n;
    };

    if (warp_idx == 3) {
        // ===================================== dynamic tile scheduler (leader CTA) ==============
        if (dynamic && lane == 0 && rank == 0) {
            const int nclusters = unit_stride;
            const int total_fetches = (args.num_units > nclusters ? args.num_units - nclusters : 0) + nclusters;
            const uint32_t unit_peer = (CG == 2) ? mapa_u32(smem_u32(&sched_unit[0]), 1) : 0;
            const uint32_t full_peer = (CG == 2) ? mapa_u32(smem_u32(&sched_full[0]), 1) : 0;
            uint32_t stage = 0, phase = 0;
            while (true) {
                mbar_wait(&sched_empty[stage], phase ^ 1);
                const int v = atomicAdd(args.sched_counter, 1);
                if (v == total_fetches - 1) atomicExch(args.sched_counter, 0);   // the last fetch of the launch re-arms the counter
                int u = v + nclusters;
                if (u >= args.num_units) u = -1;
                sched_unit[stage] = u;
                if (CG == 2) {
                    st_shared_cluster_u32(unit_peer + stage * 4, (uint32_t)u);
                    mbar_arrive_cluster_release(full_peer + stage * 8);
                }
                mbar_arrive(&sched_full[stage]);
                if (u < 0) break;
                if (++stage == kSchedStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            uint32_t cs_mask = 0, cs_phase = 0;   // per-stage bits: awaiting the column-sum warp / phase of its barrier
            SchedReader rd;
            for (int u = unit0; u >= 0; u = sched_next(rd, u, false)) {
                const UnitCoord c = decode_unit<CG>(args, u, rank);
                const int bb = args.b_batched ? c.batch : 0;
                // this CTA's share of the B tile.  A last n-block with at most 128 valid columns (N = 384: DeiT-S) runs as a 256x128
                // MMA: each CTA of a pair then supplies 64 of them (the box still loads 128 rows; the extra ones are not read)
                const bool tail128 = args.N - c.n_blk * int(BN) <= args.tail_cols;
                const int n_row0 = c.n_blk * BN + rank * ((tail128 ? 128 : BN) / CG);
                int cs_cnt = cs_count0(c);
                for (int kb = c.kb0; kb < c.kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (AMAJ == 1) {
                        // a stage whose previous contents the column-sum warp reads is refilled only after it has said so
                        if (cs_mask & (1u << stage)) {
                            mbar_wait(&cs_done[stage], (cs_phase >> stage) & 1);
                            cs_phase ^= 1u << stage;
                            cs_mask &= ~(1u << stage);
                        }
                        if (colsum_on && cs_cnt == 0) cs_mask |= 1u << stage;
                        if (++cs_cnt == args.n_blocks) cs_cnt = 0;
                    }
                    uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
                    uint8_t* sb = smem_b + stage * kBStage;
                    if (CG == 1) {
                        mbar_arrive_expect_tx(&full_bar[stage], T::kStageBytes);
                        if (AMAJ == 0) {
                            tma_load_3d(sa, &tmA, &full_bar[stage], kb * BK, c.m_blk * BM, c.batch);
                        } else {
#pragma unroll
                            for (int j = 0; j < int(BM / 64); ++j)
                                tma_load_3d(sa + j * (BK * 128), &tmA, &full_bar[stage], c.m_blk * BM + j * 64, kb * BK, c.batch);
                        }
                        if (BMAJ == 0) {
                            tma_load_3d(sb, &tmB, &full_bar[stage], kb * BK, n_row0, bb);
                        } else {
#pragma unroll
                            for (int j = 0; j < int(BN / 64); ++j)
                                tma_load_3d(sb + j * (BK * 128), &tmB, &full_bar[stage], n_row0 + j * 64, kb * BK, bb);
                        }
                    } else {
                        // both CTAs' loads complete on the LEADER's full barrier; the leader expects both halves
                        const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * T::kStageBytes);
                        if (AMAJ == 0) {
                            tma_load_3d_2sm(sa, &tmA, fb, kb * BK, c.m_blk * BM, c.batch);
                        } else {
#pragma unroll
                            for (int j = 0; j < int(BM / 64); ++j)
                                tma_load_3d_2sm(sa + j * (BK * 128), &tmA, fb, c.m_blk * BM + j * 64, kb * BK, c.batch);
                        }
                        if (BMAJ == 0) {
                            tma_load_3d_2sm(sb, &tmB, fb, kb * BK, n_row0, bb);
                        } else {
#pragma unroll
                            for (int j = 0; j < int(BN / CG / 64); ++j)
                                tma_load_3d_2sm(sb + j * (BK * 128), &tmB, fb, n_row0 + j * 64, kb * BK, bb);
                        }
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer =======================================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc_full = umma_idesc_bf16(BM * CG, BN, AMAJ, BMAJ);
            constexpr uint32_t idesc_tail = umma_idesc_bf16(BM * CG, 128, AMAJ, BMAJ);   // n-block with <= 128 valid columns
            // K-major SW128: 8-row groups 1024 B apart (SBO); MN-major SW128: 64-wide MN atoms BK*128 B apart
            // (LBO) and 8-deep K groups 1024 B apart (SBO).
            constexpr uint64_t a_base = (AMAJ == 0) ? umma_smem_desc_base(0, 1024) : umma_smem_desc_base(BK * 128, 1024);
            constexpr uint64_t b_base = (BMAJ == 0) ? umma_smem_desc_base(0, 1024) : umma_smem_desc_base(BK * 128, 1024);
            constexpr uint32_t a_kstep = (AMAJ == 0) ? UK * 2 : UK * 128;
            constexpr uint32_t b_kstep = (BMAJ == 0) ? UK * 2 : UK * 128;
            uint32_t stage = 0, phase = 0, it = 0;
            SchedReader rd;
            const uint32_t cs_ready_peer = (CG == 2) ? mapa_u32(smem_u32(&cs_ready[0]), 1) : 0;
            const bool dbg_on = args.dbg != nullptr && blockIdx.x == 0;
            long long t_begin = 0, w_full = 0, w_tmem = 0, w_sched = 0, t0 = 0;
            if (dbg_on) t_begin = clock64();
            int u = unit0;
            while (u >= 0) {
                const UnitCoord c = decode_unit<CG>(args, u, 0);
                const uint32_t as = it & 1, ap = (it >> 1) & 1;
                if (dbg_on) t0 = clock64();
                mbar_wait(&tmem_empty_bar[as], ap ^ 1);
                if (dbg_on) w_tmem += clock64() - t0;
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                const uint32_t idesc = (args.N - c.n_blk * int(BN) <= args.tail_cols) ? idesc_tail : idesc_full;
                int cs_cnt = cs_count0(c);
                for (int kb = c.kb0; kb < c.kb1; ++kb) {
                    if (dbg_on) t0 = clock64();
                    mbar_wait(&full_bar[stage], phase);
                    if (dbg_on) w_full += clock64() - t0;
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
                    const uint32_t sb = smem_u32(smem_b + stage * kBStage);
#pragma unroll
                    for (uint32_t k = 0; k < BK / UK; ++k) {
                        const uint64_t ad = umma_smem_desc(a_base, sa + k * a_kstep);
                        const uint64_t bd = umma_smem_desc(b_base, sb + k * b_kstep);
                        if (CG == 2) umma_bf16_ss_2sm(d_tmem, ad, bd, idesc, (kb > c.kb0 || k > 0) ? 1u : 0u);
                        else umma_bf16_ss(d_tmem, ad, bd, idesc, (kb > c.kb0 || k > 0) ? 1u : 0u);
                    }
                    // frees the smem slot (in both CTAs of the pair) once these MMAs retire
                    if (CG == 2) umma_commit_2sm(&empty_bar[stage], 0x3);
                    else umma_commit(&empty_bar[stage]);
                    if constexpr (AMAJ == 1) {
                        if (colsum_on && cs_cnt == 0) {   // this stage is also read by the column-sum warps (after the issue: off its path)
                            mbar_arrive(&cs_ready[stage]);
                            if (CG == 2) mbar_arrive_cluster(cs_ready_peer + stage * 8);
                        }
                        if (++cs_cnt == args.n_blocks) cs_cnt = 0;
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue warps of both CTAs
                if (CG == 2) umma_commit_2sm(&tmem_full_bar[as], 0x3);
                else umma_commit(&tmem_full_bar[as]);
                if (dbg_on) t0 = clock64();
                u = sched_next(rd, u, false);
                if (dbg_on) w_sched += clock64() - t0;
                ++it;
            }
            if (dbg_on) {
                args.dbg[0] = clock64() - t_begin; args.dbg[1] = w_full; args.dbg[2] = w_tmem; args.dbg[3] = w_sched; args.dbg[4] = it;
            }
        }
    } else if (warp_idx == 2) {
        // ===================================== column sums of the A operand (a_colsum) ===========================================
        // A is MN-major: a stage holds BM/64 boxes of [BK k-rows][64 m] bf16 (128-byte rows, 16-byte chunks XOR-swizzled by k & 7).
        // Of the n_blocks units that load the same A tiles (same m-group and k-range, different n-block), unit j sums the k-blocks
        // with kb % n_blocks == j, so the extra shared-memory reads are spread evenly.  The stage is read WHILE the MMAs consume it
        // (reading it after they retire delayed the refill of every third stage and cost the weight-gradient GEMMs ~20 %): the MMA
        // issuer, who sees every TMA completion in order, relays it to this warp in both CTAs of a pair through cs_ready (the loads of
        // both CTAs complete on the leader's barrier only; a parity wait on full_bar by a warp that skips stages would be ambiguous).
        // The producer refills such a stage only after cs_done as well as the MMA's commit.
        // Lane (r = lane >> 3, ch = lane & 7) owns 8 columns (one 16-byte chunk) of rows r, r + 4, ...
        if constexpr (AMAJ == 1) {
            if (colsum_on) {
                static_assert(BM == 128 && BK == 64, "column-sum geometry");
                uint32_t stage = 0, ready_phase = 0;
                SchedReader rd;
                const uint32_t ch = lane & 7, r4 = lane >> 3;
                for (int u = unit0; u >= 0; u = sched_next(rd, u, true)) {
                    const UnitCoord c = decode_unit<CG>(args, u, rank);
                    float acc[2][8];
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
                    int cs_cnt = cs_count0(c);
                    for (int kb = c.kb0; kb < c.kb1; ++kb) {
                        const bool mine = cs_cnt == 0;
                        if (++cs_cnt == args.n_blocks) cs_cnt = 0;
                        if (mine) {
                            mbar_wait(&cs_ready[stage], (ready_phase >> stage) & 1);   // exactly one arrival per use (MMA issuer)
                            ready_phase ^= 1u << stage;
                            const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
#pragma unroll
                            for (uint32_t half_ = 0; half_ < 2; ++half_) {
                                uint32_t w[8][2][4];                      // 16 independent 16-byte loads in flight per lane
#pragma unroll
                                for (uint32_t it = 0; it < 8; ++it) {
                                    const uint32_t k = (half_ * 8 + it) * 4 + r4;
                                    const uint32_t off = k * 128 + ((ch ^ (k & 7)) << 4);
#pragma unroll
                                    for (int j = 0; j < 2; ++j)
                                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                                     : "=r"(w[it][j][0]), "=r"(w[it][j][1]), "=r"(w[it][j][2]), "=r"(w[it][j][3])
                                                     : "r"(sa + j * (BK * 128) + off));
                                }
#pragma unroll
                                for (uint32_t it = 0; it < 8; ++it)
#pragma unroll
                                    for (int j = 0; j < 2; ++j)
#pragma unroll
                                        for (int q = 0; q < 4; ++q)
                                            asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}"
                                                : "+f"(acc[j][2 * q]), "+f"(acc[j][2 * q + 1]) : "r"(w[it][j][q]));
                            }
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&cs_done[stage]);
                        }
                        if (++stage == kStages) stage = 0;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float v = acc[j][e];
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            v += __shfl_xor_sync(0xffffffffu, v, 16);
                            const int m = c.m_blk * int(BM) + j * 64 + int(ch) * 8 + e;
                            if (r4 == 0 && m < args.M) atomicAdd(args.a_colsum + m, v);
                        }
                }
            }
        }
    } else if (warp_idx >= kFirstEpiWarp) {
        // ===================================== epilogue ==========================================
        const uint32_t ew = warp_idx - kFirstEpiWarp;
        const uint32_t quad = warp_idx & 3;   // TMEM lane quadrant this warp may access
        const uint32_t half = ew >> 2;        // which 128-column half of the tile
        uint8_t* wbuf = smem_epi + ew * T::kBufsPerWarp * EPI_BUF_BYTES;
        uint8_t* buf_aux = wbuf;                                        // aux-in (only if kHasAux)
        uint8_t* buf_o0 = wbuf + (T::kHasAux ? EPI_BUF_BYTES : 0);      // output staging 0
        uint8_t* buf_o1 = buf_o0 + EPI_BUF_BYTES;                       // staging 1: GELU's second output, else double buffer
        uint32_t chunk_count = 0;                                       // running count of stored chunks (selects the out buffer)
        const uint32_t tmem_empty_remote = (CG == 2) ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : 0;
        const uint32_t swz = (lane & 7) << 4;
        const uint32_t row_off = lane * 128;
        uint32_t aux_phase = 0;

        // number of this warp's column chunks of unit u that lie inside N (chunks are ordered by column)
        auto valid_chunks = [&](int u) -> int {
            const UnitCoord c = decode_unit<CG>(args, u, rank);
            int n = 0;
#pragma unroll
            for (int ch = 0; ch < int(kChunks); ++ch)
                if (c.n_blk * BN + half * 128 + ch * int(CPC) < args.N) n = ch + 1;
            return n;
        };
        auto issue_aux = [&](int u, int ch) {
            const UnitCoord c = decode_unit<CG>(args, u, rank);
            const int col0 = c.n_blk * BN + half * 128 + ch * CPC;
            const int row0 = args.c_row_offset + c.m_blk * BM + quad * 32;
            mbar_arrive_expect_tx(&aux_bar[ew], EPI_BUF_BYTES);
            tma_load_3d(buf_aux, &tmAux, &aux_bar[ew], col0, row0, args.aux_bcast ? 0 : c.batch);
        };
        // The aux tile of the next chunk is prefetched one chunk ahead, across the unit boundary when the next unit is known
        // (aux_issued is warp-uniform; lane 0 issues).
        bool aux_issued = false;
        SchedReader rd;
        int cur = unit0;
        int nxt = sched_next(rd, cur, true);

        uint32_t it = 0;
        auto release_tmem = [&](uint32_t as) {
            if (CG == 2 && rank != 0) mbar_arrive_cluster(tmem_empty_remote + as * 8);
            else mbar_arrive(&tmem_empty_bar[as]);
        };
        for (; cur >= 0; cur = nxt, nxt = (cur >= 0 ? sched_next(rd, cur, true) : -1), ++it) {
            const int u = cur;
            const UnitCoord c = decode_unit<CG>(args, u, rank);
            const uint32_t as = it & 1, ap = (it >> 1) & 1;
            mbar_wait(&tmem_full_bar[as], ap);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((quad * 32) << 16) + as * BN + half * 128;
            const int row_in_batch = c.m_blk * BM + quad * 32 + lane;  // GEMM row of this thread
            const int row0 = args.c_row_offset + c.m_blk * BM + quad * 32;

            const int n_valid_chunks = (args.direct == 2) ? 0 : valid_chunks(u);   // direct == 2: profiling aid, epilogue skipped entirely
            if (n_valid_chunks == 0) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) release_tmem(as);
                continue;
            }

            for (int ch = 0; ch < n_valid_chunks; ++ch) {
                const int col0 = c.n_blk * BN + half * 128 + ch * CPC;
                if constexpr (T::kHasAux) {
                    if (!args.direct) {
                        if (!aux_issued && lane == 0) issue_aux(u, ch);   // no look-ahead was possible (first chunk, ragged N)
                        mbar_wait(&aux_bar[ew], aux_phase);
                        aux_phase ^= 1;
                    }
                }
                const bool write_c = (EPI != VB_EPI_GELU) || args.C != nullptr;
                // (Tried and rejected: global stores straight from registers instead of staging + TMA store — a lane owns a ROW, so one
                // store instruction touches 32 different cache lines and the LSU becomes the bottleneck at K = 768: 16-byte stores
                // 10.4 k vs 7.5 k cycles per tile (round 1); 256-bit STG.256 / LDG.256 for the aux operand (round 2): fc1 + GELU
                // 235 -> 265 us, fc2 dgrad x gelu' 205 -> 250 us, out-proj + residual 82 -> 103 us.  Also rejected in round 2: a
                // MUFU-free-reciprocal erf (degree-8 polynomial for erfcx, one MUFU per element instead of two): 235 us unchanged —
                // the two-output epilogue is bound by shared-memory staging traffic next to the operand reads, not by the math.
                // Late round 2, the fp32-residual epilogue at K = 768 (12.4-12.6 k cycles per tile, MMA floor 6.2 k): (a) in-place
                // double buffering of the aux tile — the TMA load of the next chunk's aux in flight during the whole chunk — 12.4 k;
                // (b) no staging at all: tcgen05.ld.16x256b fragments (four lanes own 32 contiguous bytes of a row), 8-byte LDG / STG
                // on whole sectors, L2 prefetch two chunks ahead — 14.3 k.  Same bytes, same time: that GEMM is bound by the L2->SM
                // throughput share of an SM (384 KB operands + 256 KB fp32 aux / out per CTA and tile))
                uint8_t* obuf = buf_o0;
#pragma unroll
                for (int g = 0; g < int(CPC / 32); ++g) {
                    const int colg = col0 + g * 32;
                    // ---- bias (issued first: the global-load latency hides behind the TMEM load) ----
                    float4 b4[8];
                    const bool bias_vec = args.bias != nullptr && colg + 32 <= args.N;
                    if (bias_vec) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(args.bias + colg) + j);
                    }
                    // ---- accumulator -> registers (32 fp32 columns of this thread's row) ----
                    uint32_t acc[32];
                    tmem_ld_32x32b_x32(t_addr + ch * CPC + g * 32, acc);
                    tmem_ld_wait();
                    if (ch == n_valid_chunks - 1 && g == int(CPC / 32) - 1) {
                        // all TMEM reads of this tile by this warp are done -> hand the accumulator back
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) release_tmem(as);
                    }
                    if (args.direct == 3) {   // profiling aid: TMEM loads only (keeps the loaded values alive, writes nothing)
                        uint32_t x_ = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) x_ ^= acc[j];
                        if (x_ == 0x12345678u && args.C2 != nullptr) reinterpret_cast<uint32_t*>(args.C2)[0] = x_;
                        continue;
                    }
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
                    if (bias_vec) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            f2_unpack(f2_add(f2_pack(v[4 * j], v[4 * j + 1]), f2_pack(b4[j].x, b4[j].y)), v[4 * j], v[4 * j + 1]);
                            f2_unpack(f2_add(f2_pack(v[4 * j + 2], v[4 * j + 3]), f2_pack(b4[j].z, b4[j].w)), v[4 * j + 2], v[4 * j + 3]);
                        }
                    } else if (args.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (colg + j < args.N) v[j] += __ldg(args.bias + colg + j);
                    }
                    // ---- aux operand (same dtype / geometry as C) ----
                    float x[32];
                    if constexpr (T::kHasAux) {
                        if (!args.direct) {
                            if constexpr (CDT == VB_F32) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    const uint4 w = *reinterpret_cast<const uint4*>(buf_aux + row_off + ((q << 4) ^ swz));
                                    x[q * 4 + 0] = __uint_as_float(w.x); x[q * 4 + 1] = __uint_as_float(w.y);
                                    x[q * 4 + 2] = __uint_as_float(w.z); x[q * 4 + 3] = __uint_as_float(w.w);
                                }
                            } else {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const uint4 w = *reinterpret_cast<const uint4*>(buf_aux + row_off + ((((g * 4 + q)) << 4) ^ swz));
                                    x[q * 8 + 0] = bf16_lo(w.x); x[q * 8 + 1] = bf16_hi(w.x);
                                    x[q * 8 + 2] = bf16_lo(w.y); x[q * 8 + 3] = bf16_hi(w.y);
                                    x[q * 8 + 4] = bf16_lo(w.z); x[q * 8 + 5] = bf16_hi(w.z);
                                    x[q * 8 + 6] = bf16_lo(w.w); x[q * 8 + 7] = bf16_hi(w.w);
                                }
                            }
                        } else {
                            const long long ab = args.aux_bcast ? 0 : (long long)c.batch * args.bsaux;
                            const long long off = ab + (long long)(args.c_row_offset + row_in_batch) * args.ldaux + colg;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const bool ok = row_in_batch < args.M && colg + j < args.N;
                                if constexpr (CDT == VB_F32) x[j] = ok ? reinterpret_cast<const float*>(args.AUX)[off + j] : 0.f;
                                else x[j] = ok ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(args.AUX)[off + j]) : 0.f;
                            }
                        }
                    }
                    // ---- fused math (fp32x2-packed where the operation allows it) ----
                    float gl[32];
                    if constexpr (EPI == VB_EPI_GELU) {   // C2 = gelu(x); C = gelu'(x) (what the backward epilogue multiplies by)
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float g0, g1, d0, d1;
                            gelu_pair(v[j], v[j + 1], g0, g1, d0, d1);
                            gl[j] = g0; gl[j + 1] = g1;
                            v[j] = d0; v[j + 1] = d1;
                        }
                    } else if constexpr (EPI == VB_EPI_RESIDUAL) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) f2_unpack(f2_add(f2_pack(v[j], v[j + 1]), f2_pack(x[j], x[j + 1])), v[j], v[j + 1]);
                    } else if constexpr (EPI == VB_EPI_DGELU) {   // AUX = gelu'(pre-activation), saved by the forward epilogue
#pragma unroll
                        for (int j = 0; j < 32; j += 2) f2_unpack(f2_mul(f2_pack(v[j], v[j + 1]), f2_pack(x[j], x[j + 1])), v[j], v[j + 1]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if constexpr (EPI == VB_EPI_RELU) v[j] = fmaxf(v[j], 0.f);
                            if constexpr (EPI == VB_EPI_DRELU) v[j] = x[j] > 0.f ? v[j] * args.drelu_scale : 0.f;
                        }
                    }
                    // ---- registers -> swizzled staging buffer (or straight to global on the bring-up path) ----
                    if (!args.direct) {
                        if (g == 0) {
                            // the staging buffer is single: the previous chunk's TMA store must have finished reading it
                            if (lane == 0) {
                                tma_store_wait_read<0>();
                            }
                            __syncwarp();
                        }
                        if constexpr (CDT == VB_F32) {
                            if (write_c) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    uint4 w;
                                    w.x = __float_as_uint(v[q * 4 + 0]); w.y = __float_as_uint(v[q * 4 + 1]);
                                    w.z = __float_as_uint(v[q * 4 + 2]); w.w = __float_as_uint(v[q * 4 + 3]);
                                    *reinterpret_cast<uint4*>(obuf + row_off + ((q << 4) ^ swz)) = w;
                                }
                            }
                        } else {
                            if (write_c) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    uint4 w;
                                    w.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]); w.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
                                    w.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]); w.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
                                    *reinterpret_cast<uint4*>(obuf + row_off + (((g * 4 + q) << 4) ^ swz)) = w;
                                }
                            }
                            if constexpr (T::kTwoOut) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    uint4 w;
                                    w.x = pack_bf16(gl[q * 8 + 0], gl[q * 8 + 1]); w.y = pack_bf16(gl[q * 8 + 2], gl[q * 8 + 3]);
                                    w.z = pack_bf16(gl[q * 8 + 4], gl[q * 8 + 5]); w.w = pack_bf16(gl[q * 8 + 6], gl[q * 8 + 7]);
                                    *reinterpret_cast<uint4*>(buf_o1 + row_off + (((g * 4 + q) << 4) ^ swz)) = w;
                                }
                            }
                        }
                    } else if (row_in_batch < args.M) {
                        const long long off = (long long)c.batch * args.bsc + (long long)(args.c_row_offset + row_in_batch) * args.ldc + colg;
                        const long long off2 = (long long)c.batch * args.bsc2 + (long long)(args.c_row_offset + row_in_batch) * args.ldc2 + colg;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (colg + j >= args.N) continue;
                            if constexpr (EPI == VB_EPI_ACCUM) {
                                atomicAdd(reinterpret_cast<float*>(args.C) + off + j, v[j]);
                            } else if (write_c) {
                                if constexpr (CDT == VB_F32) reinterpret_cast<float*>(args.C)[off + j] = v[j];
                                else reinterpret_cast<__nv_bfloat16*>(args.C)[off + j] = __float2bfloat16_rn(v[j]);
                            }
                            if constexpr (T::kTwoOut) reinterpret_cast<__nv_bfloat16*>(args.C2)[off2 + j] = __float2bfloat16_rn(gl[j]);
                        }
                    }
                }
                if (!args.direct) {
                    fence_proxy_async_smem();  // staging writes (and aux reads) ordered before the async proxy touches smem
                    __syncwarp();
                    if constexpr (T::kHasAux) {
                        // aux buffer consumed: prefetch the aux tile of the next chunk this warp will process
                        if (ch + 1 < n_valid_chunks) {
                            aux_issued = true;
                            if (lane == 0) issue_aux(u, ch + 1);
                        } else if (nxt >= 0 && valid_chunks(nxt) > 0) {
                            aux_issued = true;
                            if (lane == 0) issue_aux(nxt, 0);
                        } else {
                            aux_issued = false;
                        }
                    }
                    if (lane == 0) {
                        if constexpr (EPI == VB_EPI_ACCUM) {
                            tma_reduce_add_3d(&tmC, obuf, col0, row0, c.batch);
                        } else {
                            if (write_c) tma_store_3d(&tmC, obuf, col0, row0, c.batch);
                            if constexpr (T::kTwoOut) tma_store_3d(&tmC2, buf_o1, col0, row0, c.batch);
                        }
                        tma_store_commit();
                    }
                    ++chunk_count;
                }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all();   // the peer may still be reading our smem / signalling our barriers
    else __syncthreads();
    if (warp_idx == 2) {
        tcgen05_fence_after();
        if (CG == 2) tmem_dealloc_2sm<kTmemCols>(tmem_base);
        else tmem_dealloc<kTmemCols>(tmem_base);
    }
#endif
}

// ------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D map (inner, rows, batch) with a 128-byte-swizzled (box0 x box1 x 1) box.
int make_tmap_3d(CUtensorMap* m, int dtype, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
                 uint64_t batch_stride_elems, uint32_t box0, uint32_t box1) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(VB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const uint64_t es = (dtype == VB_BF16) ? 2 : 4;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(VB_ERR_ARG, "tensor base %p not 16-byte aligned", ptr);
    if ((ld_elems * es) % 16 != 0) return fail(VB_ERR_ARG, "row pitch %llu bytes not a multiple of 16", (unsigned long long)(ld_elems * es));
    if (d2 > 1 && (batch_stride_elems * es) % 16 != 0) return fail(VB_ERR_ARG, "batch stride not a multiple of 16 bytes");
    if (box0 * es > 128 || box1 > 256) return fail(VB_ERR_ARG, "bad TMA box %u x %u", box0, box1);
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {ld_elems * es, (d2 > 1 ? batch_stride_elems : d1 * ld_elems) * es};
    if (strides[1] == 0) strides[1] = strides[0];
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, dtype == VB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                    const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(VB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu, pitch %llu B, box %u x %u", (int)r,
                    (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                    (unsigned long long)strides[0], box0, box1);
    return VB_OK;
}

template <int AMAJ, int BMAJ, int EPI, int CDT, int CG>
static int launch(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const CUtensorMap& tC2,
                  const CUtensorMap& tX, const GemmArgs& args, int grid, cudaStream_t stream) {
    auto kern = gemm_kernel<AMAJ, BMAJ, EPI, CDT, CG>;
    constexpr uint32_t smem = EpiTraits<EPI, CG>::kSmemBytes;
    static_assert(smem <= 232448, "shared memory budget exceeded");
    static DeviceOnce configured;  // per instantiation; attribute is per-context, benign to repeat on races
    if (!configured.is_set()) {
        VB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.set();
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled_gemm() ? 2 : 1;
    VB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tA, tB, tC, tC2, tX, args));
    return VB_OK;
}

// Dynamic tile scheduling (see the kernel): VITB200_GEMM_DYNAMIC=0 selects the static round-robin schedule.
static int dynamic_scheduling() {
    static int dyn = -1;
    if (dyn < 0) {
        const char* e = getenv("VITB200_GEMM_DYNAMIC");
        dyn = (e && e[0] == '0') ? 0 : 1;
    }
    return dyn;
}
// Pool of self-re-arming unit counters (the last fetch of a launch resets its counter), handed out round-robin so that
// launches that may overlap never share one.  Allocated on first use (before any CUDA-graph capture: the first step is eager).
static int* next_sched_counter() {
    constexpr int kPool = 256;
    static int* pool[16] = {};
    static std::atomic<unsigned> next[16];   // forward and autograd threads may launch concurrently
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (pool[dev] == nullptr) {
        int* p = nullptr;
        if (cudaMalloc(&p, kPool * sizeof(int)) != cudaSuccess) return nullptr;
        if (cudaMemset(p, 0, kPool * sizeof(int)) != cudaSuccess) return nullptr;
        cudaDeviceSynchronize();
        pool[dev] = p;
    }
    return pool[dev] + (next[dev].fetch_add(1, std::memory_order_relaxed) % kPool);
}

static long long* g_gemm_dbg = nullptr;

static int default_cta_group() {
    static int cg = 0;
    if (cg == 0) {
        const char* e = getenv("VITB200_GEMM_CTA_GROUP");
        cg = (e && e[0] == '1') ? 1 : 2;
    }
    return cg;
}

}  // namespace vb

extern "C" int vb_gemm_bf16(const VbGemmDesc* d, void* stream_) {
    using namespace vb;
    if (!d) return fail(VB_ERR_ARG, "null descriptor");
    if (int rc = check_arch()) return rc;
    cudaStream_t stream = as_stream(stream_);
    VB_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0 && d->batches >= 1, "bad GEMM dims M=%d N=%d K=%d batches=%d", d->M, d->N, d->K, d->batches);
    VB_REQUIRE(d->A && d->B && d->C || (d->epilogue == VB_EPI_GELU && d->A && d->B && d->C2), "null operand pointer");
    VB_REQUIRE(d->a_major == 0 || d->a_major == 1, "bad a_major");
    VB_REQUIRE(d->b_major == 0 || d->b_major == 1, "bad b_major");
    VB_REQUIRE(d->c_dtype == VB_BF16 || d->c_dtype == VB_F32, "bad c_dtype");
    const int split = d->split_k > 1 ? d->split_k : 1;
    VB_REQUIRE(split == 1 || d->epilogue == VB_EPI_ACCUM, "split_k > 1 requires VB_EPI_ACCUM");
    VB_REQUIRE(d->epilogue != VB_EPI_ACCUM || d->c_dtype == VB_F32, "VB_EPI_ACCUM requires fp32 C");
    const bool has_aux = d->epilogue == VB_EPI_RESIDUAL || d->epilogue == VB_EPI_DGELU || d->epilogue == VB_EPI_DRELU;
    VB_REQUIRE(!has_aux || d->AUX, "epilogue %d needs AUX", d->epilogue);
    VB_REQUIRE(d->epilogue != VB_EPI_GELU || (d->C2 && d->c_dtype == VB_BF16), "VB_EPI_GELU needs bf16 C2");

    const int cg = default_cta_group();
    GemmArgs a{};
    a.M = d->M; a.N = d->N; a.K = d->K; a.batches = d->batches;
    const int m_blocks = ((d->M + BM - 1) / BM + cg - 1) / cg;   // m-groups of `cg` 128-row blocks
    a.n_blocks = (d->N + BN - 1) / BN;
    a.k_blocks = (d->K + BK - 1) / BK;
    a.tiles_per_batch = m_blocks * a.n_blocks;
    a.num_tiles = a.tiles_per_batch * d->batches;
    a.kb_per_split = (a.k_blocks + split - 1) / split;
    a.split_k = (a.k_blocks + a.kb_per_split - 1) / a.kb_per_split;  // drop empty splits
    a.num_units = a.num_tiles * a.split_k;
    a.c_row_offset = d->c_row_offset;
    a.aux_bcast = d->aux_batch_broadcast;
    a.b_batched = d->batch_stride_b != 0;
    a.direct = d->debug_direct_store;
    static const int tail_on = [] { const char* e = getenv("VITB200_GEMM_TAIL128"); return e ? atoi(e) : 1; }();
    a.tail_cols = tail_on ? 128 : 0;
    a.drelu_scale = d->drelu_scale > 0.f ? d->drelu_scale : 1.0f;
    a.sched_counter = nullptr;
    a.dbg = g_gemm_dbg;
    a.bias = d->bias;
    VB_REQUIRE(!d->a_colsum || (d->a_major == 1 && d->batches == 1), "a_colsum needs a_major = 1 and a single batch");
    a.a_colsum = d->a_colsum;
    a.C = d->C; a.C2 = d->C2; a.AUX = d->AUX;
    a.ldc = d->ldc; a.ldc2 = d->ldc2; a.ldaux = d->ldaux;
    a.bsc = d->batch_stride_c; a.bsc2 = d->batch_stride_c2; a.bsaux = d->batch_stride_aux;
    const uint64_t c_rows = d->c_rows > 0 ? d->c_rows : d->M + d->c_row_offset;

    CUtensorMap tA, tB, tC, tC2, tX;
    int rc;
    const uint64_t nb = d->batches;
    if (d->a_major == 0) rc = make_tmap_3d(&tA, VB_BF16, d->A, d->K, d->M, nb, d->lda, d->batch_stride_a, 64, BM);
    else rc = make_tmap_3d(&tA, VB_BF16, d->A, d->M, d->K, nb, d->lda, d->batch_stride_a, 64, BK);
    if (rc) return rc;
    const uint64_t nbb = a.b_batched ? nb : 1;
    if (d->b_major == 0) rc = make_tmap_3d(&tB, VB_BF16, d->B, d->K, d->N, nbb, d->ldb, d->batch_stride_b, 64, BN / cg);
    else rc = make_tmap_3d(&tB, VB_BF16, d->B, d->N, d->K, nbb, d->ldb, d->batch_stride_b, 64, BK);
    if (rc) return rc;
    const uint32_t cpc = d->c_dtype == VB_BF16 ? 64 : 32;
    const bool have_c = d->C != nullptr;
    if (have_c) {
        rc = make_tmap_3d(&tC, d->c_dtype, d->C, d->N, c_rows, nb, d->ldc, d->batch_stride_c, cpc, 32);
        if (rc) return rc;
    }
    if (d->epilogue == VB_EPI_GELU) {
        rc = make_tmap_3d(&tC2, VB_BF16, d->C2, d->N, c_rows, nb, d->ldc2, d->batch_stride_c2, cpc, 32);
        if (rc) return rc;
        if (!have_c) tC = tC2;
    } else {
        tC2 = tC;
    }
    if (has_aux) {
        rc = make_tmap_3d(&tX, d->c_dtype, d->AUX, d->N, c_rows, d->aux_batch_broadcast ? 1 : nb, d->ldaux, d->batch_stride_aux, cpc, 32);
        if (rc) return rc;
    } else {
        tX = tC;
    }

    int sms = num_sms();
    if (sms <= 0) return fail(VB_ERR_CUDA, "cannot determine SM count");
    int grid = d->max_ctas > 0 ? (d->max_ctas < sms ? d->max_ctas : sms) : sms;
    if (grid > a.num_units * cg) grid = a.num_units * cg;
    grid -= grid % cg;
    if (grid < cg) grid = cg;
    if (dynamic_scheduling() && a.num_units > grid / cg) a.sched_counter = next_sched_counter();   // more units than clusters

#define VB_LAUNCH(AM, BMJ, EP, CD)                                                              \
    do {                                                                                        \
        if (cg == 2) return launch<AM, BMJ, EP, CD, 2>(tA, tB, tC, tC2, tX, a, grid, stream);   \
        return launch<AM, BMJ, EP, CD, 1>(tA, tB, tC, tC2, tX, a, grid, stream);                \
    } while (0)
    const int am = d->a_major, bm = d->b_major, ep = d->epilogue, cd = d->c_dtype;
    if (am == 0 && bm == 0) {
        if (ep == VB_EPI_STORE && cd == VB_BF16) VB_LAUNCH(0, 0, VB_EPI_STORE, VB_BF16);
        if (ep == VB_EPI_STORE && cd == VB_F32) VB_LAUNCH(0, 0, VB_EPI_STORE, VB_F32);
        if (ep == VB_EPI_GELU && cd == VB_BF16) VB_LAUNCH(0, 0, VB_EPI_GELU, VB_BF16);
        if (ep == VB_EPI_RESIDUAL && cd == VB_F32) VB_LAUNCH(0, 0, VB_EPI_RESIDUAL, VB_F32);
        if (ep == VB_EPI_RELU && cd == VB_BF16) VB_LAUNCH(0, 0, VB_EPI_RELU, VB_BF16);
        if (ep == VB_EPI_ACCUM && cd == VB_F32) VB_LAUNCH(0, 0, VB_EPI_ACCUM, VB_F32);   // split-K forward for latency-bound (small M) shapes
    } else if (am == 0 && bm == 1) {
        if (ep == VB_EPI_STORE && cd == VB_BF16) VB_LAUNCH(0, 1, VB_EPI_STORE, VB_BF16);
        if (ep == VB_EPI_STORE && cd == VB_F32) VB_LAUNCH(0, 1, VB_EPI_STORE, VB_F32);
        if (ep == VB_EPI_DGELU && cd == VB_BF16) VB_LAUNCH(0, 1, VB_EPI_DGELU, VB_BF16);
        if (ep == VB_EPI_DRELU && cd == VB_BF16) VB_LAUNCH(0, 1, VB_EPI_DRELU, VB_BF16);
    } else if (am == 1 && bm == 0) {
        // A stored [K, M] (M contiguous), B stored [N, K]: the 1x1 input_proj convolution on an NCHW feature map (detr.py:125) —
        // forward (A = features [C_in, HW], B = weight), weight gradient (A = dY [tokens, hidden], B = features) and input gradient
        if (ep == VB_EPI_STORE && cd == VB_F32) VB_LAUNCH(1, 0, VB_EPI_STORE, VB_F32);
        if (ep == VB_EPI_ACCUM && cd == VB_F32) VB_LAUNCH(1, 0, VB_EPI_ACCUM, VB_F32);
    } else if (am == 1 && bm == 1) {
        if (ep == VB_EPI_ACCUM && cd == VB_F32) VB_LAUNCH(1, 1, VB_EPI_ACCUM, VB_F32);
        if (ep == VB_EPI_STORE && cd == VB_F32) VB_LAUNCH(1, 1, VB_EPI_STORE, VB_F32);
    }
#undef VB_LAUNCH
    return fail(VB_ERR_UNSUPPORTED, "no GEMM instantiation for a_major=%d b_major=%d epilogue=%d c_dtype=%d", am, bm, ep, cd);
}

// Debug hook: device buffer of >= 8 int64 receiving the cycle accounting of cluster 0's MMA issuer (total, waiting for operand
// stages, waiting for a free accumulator, waiting for the scheduler, tiles); NULL disables.  tools/gemm_timeline.py
extern "C" VB_API int vb_debug_set_gemm_timeline(void* device_buffer) {
    vb::g_gemm_dbg = reinterpret_cast<long long*>(device_buffer);
    return VB_OK;
}
