// Fused flash-style multi-head attention, head_dim = 64: softmax(Q K^T / 8) V with online softmax in registers
// and warp-shuffle row reductions; forward saves only the per-row log-sum-exp, backward recomputes P.
//
// Replaces F.scaled_dot_product_attention as reached from nn.MultiheadAttention at vanilla_vit.py:77
// (torch/nn/functional.py:6676-6688) and the explicit bmm/softmax/bmm path of the DETR encoder layer
// (transformer.py:219 -> torch/nn/functional.py:6630-6666), including its boolean key-padding mask.
//
// Q, K, V are read in place from the projection output (row = token, head h at column h*64), so no
// head-major copy is ever made; O is written token-major, ready to be the A operand of the out-proj GEMM.
// Token row index = b * batch_stride + s * tok_stride (batch-first ViT: (S,1); sequence-first DETR: (1,N)).
#include "common.h"
#include "dropout.cuh"
#include <cstdlib>
#include <cuda_bf16.h>

namespace vb {

constexpr int HD = 64;       // head dim
constexpr int TILE = 64;     // rows per smem tile (queries or keys)
constexpr int TILE_BYTES = TILE * HD * 2;

struct AttnParams {
    const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v;
    long long ldq, ldk, ldv;
    __nv_bfloat16* o; long long ldo;
    float* lse;                // [B, H, S], log2 domain: lse2 = max + log2(sum)
    const uint8_t* kpm;        // [B, S], 1 = key is padding; may be null
    int B, H, S;              // S = number of queries
    int Sk;                   // number of keys / values (= S for self-attention; cross-attention: transformer.py:145-147)
    long long tok_stride, batch_stride;
    float scale, scale_log2;   // 1/sqrt(hd), scale * log2(e)
    // backward
    const __nv_bfloat16* dout; long long lddo;
    float* delta;              // [B, H, S]
    __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv;
    long long lddq, lddk, lddv;
    // attention dropout (DROP instantiations of the 64x64-tile kernels): element ((b*H + h)*S + q)*S + key — the indexing of the
    // tcgen05 kernels — is kept iff its hash clears drop_thresh; kept probabilities are scaled by drop_inv_keep
    uint32_t drop_thresh, drop_stream;
    float drop_inv_keep;
    const uint32_t* drop_seed;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// smem tile: 64 rows x 128 B, 16-byte chunk index XOR-swizzled with (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// Loads rows [r0, r0+64) x 64 cols of head h (rows >= S zero-filled) with 128 threads.
__device__ __forceinline__ void load_tile(uint32_t sbase, const __nv_bfloat16* g, long long ld, long long tok_stride, long long row_base,
                                          int r0, int S, int h) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + i * 128;
        const int row = c >> 3, ch = c & 7;
        const int s = r0 + row;
        const bool valid = s < S;
        const __nv_bfloat16* src = g + (row_base + (long long)(valid ? s : 0) * tok_stride) * ld + h * HD + ch * 8;
        cp_async16(sbase + tile_off(row, ch), src, valid);
    }
}

// A fragments (16 rows x 64 k) of rows [row0, row0+16) of a tile: 4 k-tiles x 4 regs.
__device__ __forceinline__ void load_a_frags(uint32_t sbase, int row0, uint32_t (&f)[4][4]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) ldsm_x4(sbase + tile_off(row0 + (lane & 15), kt * 2 + (lane >> 4)), f[kt]);
}

// acc[8][4] (16 x 64) += A(16 x 64 k, register frags) * T^T where T is a smem tile [n=64][k=64] (non-transposed B operand).
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sT) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4(sT + tile_off(np * 16 + (lane & 7) + ((lane >> 4) << 3), kt * 2 + ((lane >> 3) & 1)), b);
            mma_bf16(acc[2 * np], a[kt], b[0], b[1]);
            mma_bf16(acc[2 * np + 1], a[kt], b[2], b[3]);
        }
    }
}

// acc[8][4] (16 x 64 n) += P(16 x 64 k, given as accumulator-layout floats p[8][4]) * T where T is a smem tile [k=64][n=64].
__device__ __forceinline__ void mma_p_tile(float (&acc)[8][4], const float (&p)[8][4], uint32_t sT) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
        uint32_t a[4];
        a[0] = pack2(p[2 * kq][0], p[2 * kq][1]);
        a[1] = pack2(p[2 * kq][2], p[2 * kq][3]);
        a[2] = pack2(p[2 * kq + 1][0], p[2 * kq + 1][1]);
        a[3] = pack2(p[2 * kq + 1][2], p[2 * kq + 1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4_t(sT + tile_off(kq * 16 + (lane & 15), np * 2 + (lane >> 4)), b);
            mma_bf16(acc[2 * np], a, b[0], b[1]);
            mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}

// Writes a warp's 16 x 64 accumulator tile (scaled) as bf16 to global rows via a swizzled smem staging tile.
__device__ __forceinline__ void store_rows(uint8_t* stage, int row0, const float (&acc)[8][4], float mul, __nv_bfloat16* g, long long ld,
                                           long long tok_stride, long long row_base, int s0, int S, int h) {
    const int lane = threadIdx.x & 31;
    const int r = lane >> 2, cq = (lane & 3) * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + cq;
        *reinterpret_cast<uint32_t*>(stage + tile_off(row0 + r, col >> 3) + (col & 7) * 2) = pack2(acc[nt][0] * mul, acc[nt][1] * mul);
        *reinterpret_cast<uint32_t*>(stage + tile_off(row0 + r + 8, col >> 3) + (col & 7) * 2) = pack2(acc[nt][2] * mul, acc[nt][3] * mul);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int id = lane + i * 32;
        const int row = id >> 3, ch = id & 7;
        const int s = s0 + row0 + row;
        if (s < S) {
            const uint4 w = *reinterpret_cast<const uint4*>(stage + tile_off(row0 + row, ch));
            *reinterpret_cast<uint4*>(g + (row_base + (long long)s * tok_stride) * ld + h * HD + ch * 8) = w;
        }
    }
}

// ----------------------------------------------------------------------------------------------------------
// Forward: grid (ceil(S/64), H, B), 128 threads; warp w owns query rows [q0 + 16w, q0 + 16w + 16).
// ----------------------------------------------------------------------------------------------------------
template <bool DROP>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;          // 2 buffers
    uint8_t* sV = smem + 3 * TILE_BYTES;      // 2 buffers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
    const long long row_base = (long long)b * p.batch_stride;
    const int nkv = (p.Sk + TILE - 1) / TILE;

    load_tile(smem_addr(sQ), p.q, p.ldq, p.tok_stride, row_base, q0, p.S, h);
    load_tile(smem_addr(sK), p.k, p.ldk, p.tok_stride, row_base, 0, p.Sk, h);
    load_tile(smem_addr(sV), p.v, p.ldv, p.tok_stride, row_base, 0, p.Sk, h);
    cp_async_commit();

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const uint32_t drop_key = DROP ? dropout_key(*p.drop_seed, p.drop_stream) : 0u;
    const uint32_t drop_row0 = (((uint32_t)b * (uint32_t)p.H + (uint32_t)h) * (uint32_t)p.S + (uint32_t)(q0 + warp * 16 + (lane >> 2))) * (uint32_t)p.Sk;
    const uint32_t drop_row1 = drop_row0 + 8u * (uint32_t)p.Sk;

    for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j + 1 < nkv) {
            load_tile(smem_addr(sK + (buf ^ 1) * TILE_BYTES), p.k, p.ldk, p.tok_stride, row_base, (j + 1) * TILE, p.Sk, h);
            load_tile(smem_addr(sV + (buf ^ 1) * TILE_BYTES), p.v, p.ldv, p.tok_stride, row_base, (j + 1) * TILE, p.Sk, h);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (j == 0) load_a_frags(smem_addr(sQ), warp * 16, qf);

        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        mma_a_tileT(s, qf, smem_addr(sK + buf * TILE_BYTES));

        // scale to log2 domain, mask invalid / padded keys
        const int kbase = j * TILE + (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = kbase + nt * 8 + e;
                bool dead = key >= p.Sk;
                if (!dead && p.kpm) dead = p.kpm[(long long)b * p.Sk + key] != 0;
                s[nt][e] = dead ? -INFINITY : s[nt][e] * p.scale_log2;
                s[nt][e + 2] = dead ? -INFINITY : s[nt][e + 2] * p.scale_log2;
            }
        }
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float ms0 = (mx0 == -INFINITY) ? 0.f : mx0, ms1 = (mx1 == -INFINITY) ? 0.f : mx1;
        const float a0 = exp2f(m0 - ms0), a1 = exp2f(m1 - ms1);
        m0 = mx0; m1 = mx1;
        l0 *= a0; l1 *= a1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - ms0); s[nt][1] = exp2f(s[nt][1] - ms0);
            s[nt][2] = exp2f(s[nt][2] - ms1); s[nt][3] = exp2f(s[nt][3] - ms1);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
            o[nt][0] *= a0; o[nt][1] *= a0; o[nt][2] *= a1; o[nt][3] *= a1;
        }
        if (DROP) {   // the row sum keeps the un-dropped P (softmax first, dropout on the probabilities: functional.py:6650-6652)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t key = (uint32_t)(kbase + nt * 8 + e);
                    s[nt][e] = dropout_keep(drop_key, drop_row0 + key, p.drop_thresh) ? s[nt][e] * p.drop_inv_keep : 0.f;
                    s[nt][e + 2] = dropout_keep(drop_key, drop_row1 + key, p.drop_thresh) ? s[nt][e + 2] * p.drop_inv_keep : 0.f;
                }
            }
        }
        mma_p_tile(o, s, smem_addr(sV + buf * TILE_BYTES));
        __syncthreads();  // everyone done with this K/V buffer before it is refilled
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= inv0; o[nt][1] *= inv0; o[nt][2] *= inv1; o[nt][3] *= inv1; }
    if (p.lse && (lane & 3) == 0) {
        const int r0 = q0 + warp * 16 + (lane >> 2);
        float* lse = p.lse + ((long long)b * p.H + h) * p.S;
        if (r0 < p.S) lse[r0] = m0 + log2f(l0);
        if (r0 + 8 < p.S) lse[r0 + 8] = m1 + log2f(l1);
    }
    store_rows(sQ, warp * 16, o, 1.0f, p.o, p.ldo, p.tok_stride, row_base, q0, p.S, h);
}

// ----------------------------------------------------------------------------------------------------------
// delta[b,h,s] = sum_d dO[b,s,h,d] * O[b,s,h,d].  Eight lanes per (token, head): each loads 16 bytes of O and dO
// (a warp covers 4 consecutive heads = 512 contiguous bytes per tensor), 3 shuffle steps; 4 units in flight per thread.
// HBM-bound: 2 x 128 B read + 4 B written per (token, head).
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const AttnParams p) {
    const long long total = (long long)p.B * p.S * p.H;
    const int sub = threadIdx.x & 7;
    const long long units_per_iter = (long long)gridDim.x * (blockDim.x >> 3);
    long long u0 = (long long)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
    for (; u0 < total; u0 += 4 * units_per_iter) {
        uint4 ov[4], dv[4];
        long long didx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long u = u0 + i * units_per_iter;
            const bool ok = u < total;
            const long long uu = ok ? u : 0;
            const int h = (int)(uu % p.H);
            const long long bs = uu / p.H;
            const int s = (int)(bs % p.S);
            const int b = (int)(bs / p.S);
            const long long row = (long long)b * p.batch_stride + (long long)s * p.tok_stride;
            ov[i] = __ldg(reinterpret_cast<const uint4*>(p.o + row * p.ldo + h * HD + sub * 8));
            dv[i] = __ldg(reinterpret_cast<const uint4*>(p.dout + row * p.lddo + h * HD + sub * 8));
            didx[i] = ok ? ((long long)b * p.H + h) * p.S + s : -1;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t oo[4] = {ov[i].x, ov[i].y, ov[i].z, ov[i].w}, dd[4] = {dv[i].x, dv[i].y, dv[i].z, dv[i].w};
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                acc += __uint_as_float(oo[j] << 16) * __uint_as_float(dd[j] << 16) + __uint_as_float(oo[j] & 0xFFFF0000u) * __uint_as_float(dd[j] & 0xFFFF0000u);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            if (sub == 0 && didx[i] >= 0) p.delta[didx[i]] = acc;
        }
    }
}

// ----------------------------------------------------------------------------------------------------------
// Backward dK, dV: grid (ceil(S/64), H, B); warp w owns keys [k0 + 16w, +16); loops over query blocks.
//   S^T = K Q^T, P^T = exp2(S^T*c - lse[q]), dV += P^T dO, dP^T = V dO^T, dS^T = P^T o (dP^T - delta[q]), dK += dS^T Q
// ----------------------------------------------------------------------------------------------------------
template <bool DROP>
__global__ void __launch_bounds__(128) attn_bwd_dkdv_kernel(const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sK = smem;
    uint8_t* sV = smem + TILE_BYTES;
    uint8_t* sQ = smem + 2 * TILE_BYTES;    // 2 buffers
    uint8_t* sdO = smem + 4 * TILE_BYTES;   // 2 buffers
    float* sLse = reinterpret_cast<float*>(smem + 6 * TILE_BYTES);  // [2][64]
    float* sDelta = sLse + 2 * TILE;                                // [2][64]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
    const long long row_base = (long long)b * p.batch_stride;
    const int nq = (p.S + TILE - 1) / TILE;
    const float* lse = p.lse + ((long long)b * p.H + h) * p.S;
    const float* delta = p.delta + ((long long)b * p.H + h) * p.S;

    auto load_q_block = [&](int j, int buf) {
        load_tile(smem_addr(sQ + buf * TILE_BYTES), p.q, p.ldq, p.tok_stride, row_base, j * TILE, p.S, h);
        load_tile(smem_addr(sdO + buf * TILE_BYTES), p.dout, p.lddo, p.tok_stride, row_base, j * TILE, p.S, h);
        if (threadIdx.x < TILE) {
            const int s = j * TILE + threadIdx.x;
            sLse[buf * TILE + threadIdx.x] = s < p.S ? lse[s] : 0.f;
            sDelta[buf * TILE + threadIdx.x] = s < p.S ? delta[s] : 0.f;
        }
    };
    load_tile(smem_addr(sK), p.k, p.ldk, p.tok_stride, row_base, k0, p.Sk, h);
    load_tile(smem_addr(sV), p.v, p.ldv, p.tok_stride, row_base, k0, p.Sk, h);
    load_q_block(0, 0);
    cp_async_commit();

    uint32_t kf[4][4], vf[4][4];
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
    // key validity of this thread's two rows
    const int key0 = k0 + warp * 16 + (lane >> 2), key1 = key0 + 8;
    bool dead0 = key0 >= p.Sk, dead1 = key1 >= p.Sk;
    if (p.kpm) {
        if (!dead0) dead0 = p.kpm[(long long)b * p.Sk + key0] != 0;
        if (!dead1) dead1 = p.kpm[(long long)b * p.Sk + key1] != 0;
    }

    const uint32_t drop_key = DROP ? dropout_key(*p.drop_seed, p.drop_stream) : 0u;
    const uint32_t drop_head = ((uint32_t)b * (uint32_t)p.H + (uint32_t)h) * (uint32_t)p.S;

    for (int j = 0; j < nq; ++j) {
        const int buf = j & 1;
        if (j + 1 < nq) {
            load_q_block(j + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (j == 0) {
            load_a_frags(smem_addr(sK), warp * 16, kf);
            load_a_frags(smem_addr(sV), warp * 16, vf);
        }
        const uint32_t sQb = smem_addr(sQ + buf * TILE_BYTES), sdOb = smem_addr(sdO + buf * TILE_BYTES);
        float st[8][4], dpt[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
        mma_a_tileT(st, kf, sQb);     // S^T[key, q]
        mma_a_tileT(dpt, vf, sdOb);   // dP^T[key, q]
        const int qb = (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int qi = nt * 8 + qb + e;
                const bool qdead = j * TILE + qi >= p.S;
                const float l = sLse[buf * TILE + qi], dl = sDelta[buf * TILE + qi];
                const float p0 = (qdead || dead0) ? 0.f : exp2f(st[nt][e] * p.scale_log2 - l);
                const float p1 = (qdead || dead1) ? 0.f : exp2f(st[nt][e + 2] * p.scale_log2 - l);
                float kp0 = 1.f, kp1 = 1.f;
                if (DROP) {   // dV sees keep * P / (1 - p); dS = P o (keep * dP / (1 - p) - delta)
                    const uint32_t qrow = (drop_head + (uint32_t)(j * TILE + qi)) * (uint32_t)p.Sk;
                    kp0 = dropout_keep(drop_key, qrow + (uint32_t)key0, p.drop_thresh) ? p.drop_inv_keep : 0.f;
                    kp1 = dropout_keep(drop_key, qrow + (uint32_t)key1, p.drop_thresh) ? p.drop_inv_keep : 0.f;
                }
                st[nt][e] = p0 * kp0; st[nt][e + 2] = p1 * kp1;
                dpt[nt][e] = p0 * (dpt[nt][e] * kp0 - dl);
                dpt[nt][e + 2] = p1 * (dpt[nt][e + 2] * kp1 - dl);
            }
        }
        mma_p_tile(dv, st, sdOb);   // dV += P^T dO
        mma_p_tile(dk, dpt, sQb);   // dK += dS^T Q
        __syncthreads();
    }
    store_rows(sK, warp * 16, dk, p.scale, p.dk, p.lddk, p.tok_stride, row_base, k0, p.Sk, h);
    store_rows(sV, warp * 16, dv, 1.0f, p.dv, p.lddv, p.tok_stride, row_base, k0, p.Sk, h);
}

// ----------------------------------------------------------------------------------------------------------
// Backward dQ: grid (ceil(S/64), H, B); warp w owns queries [q0 + 16w, +16); loops over key blocks.
//   S = Q K^T, P = exp2(S*c - lse[row]), dP = dO V^T, dS = P o (dP - delta[row]), dQ += dS K
// ----------------------------------------------------------------------------------------------------------
template <bool DROP>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sdO = smem + TILE_BYTES;
    uint8_t* sK = smem + 2 * TILE_BYTES;   // 2 buffers
    uint8_t* sV = smem + 4 * TILE_BYTES;   // 2 buffers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
    const long long row_base = (long long)b * p.batch_stride;
    const int nkv = (p.Sk + TILE - 1) / TILE;

    load_tile(smem_addr(sQ), p.q, p.ldq, p.tok_stride, row_base, q0, p.S, h);
    load_tile(smem_addr(sdO), p.dout, p.lddo, p.tok_stride, row_base, q0, p.S, h);
    load_tile(smem_addr(sK), p.k, p.ldk, p.tok_stride, row_base, 0, p.Sk, h);
    load_tile(smem_addr(sV), p.v, p.ldv, p.tok_stride, row_base, 0, p.Sk, h);
    cp_async_commit();

    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    const float* lse = p.lse + ((long long)b * p.H + h) * p.S;
    const float* delta = p.delta + ((long long)b * p.H + h) * p.S;
    const float lse0 = r0 < p.S ? lse[r0] : 0.f, lse1 = r1 < p.S ? lse[r1] : 0.f;
    const float dl0 = r0 < p.S ? delta[r0] : 0.f, dl1 = r1 < p.S ? delta[r1] : 0.f;

    const uint32_t drop_key = DROP ? dropout_key(*p.drop_seed, p.drop_stream) : 0u;
    const uint32_t drop_row0 = (((uint32_t)b * (uint32_t)p.H + (uint32_t)h) * (uint32_t)p.S + (uint32_t)r0) * (uint32_t)p.Sk;
    const uint32_t drop_row1 = drop_row0 + 8u * (uint32_t)p.Sk;
    uint32_t qf[4][4], dof[4][4];
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

    for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j + 1 < nkv) {
            load_tile(smem_addr(sK + (buf ^ 1) * TILE_BYTES), p.k, p.ldk, p.tok_stride, row_base, (j + 1) * TILE, p.Sk, h);
            load_tile(smem_addr(sV + (buf ^ 1) * TILE_BYTES), p.v, p.ldv, p.tok_stride, row_base, (j + 1) * TILE, p.Sk, h);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (j == 0) {
            load_a_frags(smem_addr(sQ), warp * 16, qf);
            load_a_frags(smem_addr(sdO), warp * 16, dof);
        }
        const uint32_t sKb = smem_addr(sK + buf * TILE_BYTES), sVb = smem_addr(sV + buf * TILE_BYTES);
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
        mma_a_tileT(s, qf, sKb);     // S[q, key]
        mma_a_tileT(dp, dof, sVb);   // dP[q, key]
        const int kbase = j * TILE + (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = kbase + nt * 8 + e;
                bool dead = key >= p.Sk;
                if (!dead && p.kpm) dead = p.kpm[(long long)b * p.Sk + key] != 0;
                const float p0 = dead ? 0.f : exp2f(s[nt][e] * p.scale_log2 - lse0);
                const float p1 = dead ? 0.f : exp2f(s[nt][e + 2] * p.scale_log2 - lse1);
                float kp0 = 1.f, kp1 = 1.f;
                if (DROP) {
                    kp0 = dropout_keep(drop_key, drop_row0 + (uint32_t)key, p.drop_thresh) ? p.drop_inv_keep : 0.f;
                    kp1 = dropout_keep(drop_key, drop_row1 + (uint32_t)key, p.drop_thresh) ? p.drop_inv_keep : 0.f;
                }
                dp[nt][e] = p0 * (dp[nt][e] * kp0 - dl0);
                dp[nt][e + 2] = p1 * (dp[nt][e + 2] * kp1 - dl1);
            }
        }
        mma_p_tile(dq, dp, sKb);     // dQ += dS K
        __syncthreads();
    }
    store_rows(sQ, warp * 16, dq, p.scale, p.dq, p.lddq, p.tok_stride, row_base, q0, p.S, h);
}


// ==========================================================================================================
// Short-sequence kernels (S <= 256: every ViT/DeiT config): one CTA per (batch, head), the whole head's Q, K, V
// (and dO) resident in shared memory, exact 16-row / 8-column tile counts (no 64-granular padding waste), two CTAs
// per SM so one head's loads overlap the other's math.  Warp w owns 16-row tiles w and w + NW.
// ==========================================================================================================
__device__ __forceinline__ void load_rows(uint32_t sbase, const __nv_bfloat16* g, long long ld, long long tok_stride, long long row_base,
                                          int rows_pad, int S, int h, int nthreads) {
    for (int c = threadIdx.x; c < rows_pad * 8; c += nthreads) {
        const int row = c >> 3, ch = c & 7;
        const bool valid = row < S;
        const __nv_bfloat16* src = g + (row_base + (long long)(valid ? row : 0) * tok_stride) * ld + h * HD + ch * 8;
        cp_async16(sbase + tile_off(row, ch), src, valid);
    }
}

// 4x4 transpose inside each lane quad: in[j] = this lane's packed column pair of n-tile j; out[i] = column pair i of n-tile (lane & 3).
__device__ __forceinline__ void quad_transpose(const uint32_t (&in)[4], uint32_t (&out)[4]) {
    // two butterfly rounds (xor 1, xor 2) with selects only: no divergent branches around the shuffles
    const bool odd = (threadIdx.x & 1) != 0, hi = (threadIdx.x & 2) != 0;
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? in[0] : in[1], 1);
    const uint32_t r1 = __shfl_xor_sync(0xffffffffu, odd ? in[2] : in[3], 1);
    const uint32_t a0 = odd ? r0 : in[0], a1 = odd ? in[1] : r0, a2 = odd ? r1 : in[2], a3 = odd ? in[3] : r1;
    const uint32_t u0 = __shfl_xor_sync(0xffffffffu, hi ? a0 : a2, 2);
    const uint32_t u1 = __shfl_xor_sync(0xffffffffu, hi ? a1 : a3, 2);
    out[0] = hi ? u0 : a0; out[1] = hi ? u1 : a1; out[2] = hi ? a2 : u0; out[3] = hi ? a3 : u1;
}

// Stores a warp's 16 x 64 accumulator tile (rows row0.., scaled) as bf16 with 16-byte row-contiguous stores.
__device__ __forceinline__ void store_acc_tile(const float (&acc)[8][4], float mul, __nv_bfloat16* g, long long ld, long long tok_stride,
                                               long long row_base, int row0, int S, int h) {
    const int lane = threadIdx.x & 31;
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int s = row0 + r + half * 8;
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
            uint32_t in[4], out[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) in[j] = pack2(acc[grp * 4 + j][half * 2] * mul, acc[grp * 4 + j][half * 2 + 1] * mul);
            quad_transpose(in, out);
            if (s < S) {
                uint4 w = make_uint4(out[0], out[1], out[2], out[3]);
                *reinterpret_cast<uint4*>(g + (row_base + (long long)s * tok_stride) * ld + h * HD + (grp * 4 + c) * 8) = w;
            }
        }
    }
}

// Per-lane shared-memory offsets of the ldmatrix patterns, computed once (all row bases used below are multiples
// of 16, so the XOR swizzle term only depends on the lane and the 16-byte chunk index):
//   a[i]: A-operand / transposed-B pattern, row (lane & 15), chunk 2*i + (lane >> 4)
//   b[i]: non-transposed-B pattern, row (lane & 7) + 8 * (lane >> 4), chunk 2*i + ((lane >> 3) & 1)
struct LaneOff {
    uint32_t a[4], b[4];
};
__device__ __forceinline__ LaneOff make_lane_off() {
    const int lane = threadIdx.x & 31;
    LaneOff o;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o.a[i] = (lane & 15) * 128 + (((i * 2 + (lane >> 4)) ^ (lane & 7)) << 4);
        o.b[i] = ((lane & 7) + ((lane >> 4) << 3)) * 128 + (((i * 2 + ((lane >> 3) & 1)) ^ (lane & 7)) << 4);
    }
    return o;
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// acc[NT][4] (16 x 8*NT) += A(16 rows starting at smem address sA0, k = 64) * T^T, T = rows of a smem matrix [n][64]
// starting at address sT0.  A fragments are re-read from shared memory per k-step (keeps them out of the register budget).
template <int NT, bool FULL>
__device__ __forceinline__ void mma_rows_tileT_impl(float (&acc)[NT][4], uint32_t sA0, uint32_t sT0, int np_count, const LaneOff& off) {
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
        uint32_t a[4];
        ldsm_x4(sA0 + off.a[kt], a);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
            if (FULL || np < np_count) {
                uint32_t b[4];
                ldsm_x4(sT0 + np * 2048 + off.b[kt], b);
                mma_bf16(acc[2 * np], a, b[0], b[1]);
                mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
            }
        }
    }
}
// Full blocks take a branch-free path (no per-tile predicates => no convergence barriers around ldmatrix/mma).
template <int NT>
__device__ __forceinline__ void mma_rows_tileT(float (&acc)[NT][4], uint32_t sA0, uint32_t sT0, int np_count, const LaneOff& off) {
    if (np_count == NT / 2) mma_rows_tileT_impl<NT, true>(acc, sA0, sT0, np_count, off);
    else mma_rows_tileT_impl<NT, false>(acc, sA0, sT0, np_count, off);
}

// acc[8][4] (16 x 64) += P(16 x 16*KQ, accumulator-layout floats p[2*KQ][4]) * T, T = rows of a smem matrix [k][64] from sT0.
template <int KQ, bool FULL>
__device__ __forceinline__ void mma_p_rows_impl(float (&acc)[8][4], const float (&p)[2 * KQ][4], uint32_t sT0, int kq_count, const LaneOff& off) {
#pragma unroll
    for (int kq = 0; kq < KQ; ++kq) {
        if (FULL || kq < kq_count) {
            uint32_t a[4];
            a[0] = pack2(p[2 * kq][0], p[2 * kq][1]);
            a[1] = pack2(p[2 * kq][2], p[2 * kq][3]);
            a[2] = pack2(p[2 * kq + 1][0], p[2 * kq + 1][1]);
            a[3] = pack2(p[2 * kq + 1][2], p[2 * kq + 1][3]);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b[4];
                ldsm_x4_t(sT0 + kq * 2048 + off.a[np], b);
                mma_bf16(acc[2 * np], a, b[0], b[1]);
                mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
            }
        }
    }
}
template <int KQ>
__device__ __forceinline__ void mma_p_rows(float (&acc)[8][4], const float (&p)[2 * KQ][4], uint32_t sT0, int kq_count, const LaneOff& off) {
    if (kq_count == KQ) mma_p_rows_impl<KQ, true>(acc, p, sT0, kq_count, off);
    else mma_p_rows_impl<KQ, false>(acc, p, sT0, kq_count, off);
}

// One block of NP*16 keys of the online-softmax forward for a warp's 16 query rows (compile-time tile counts, so
// the tail block of a sequence costs only its own tiles).  MASK: apply the key >= S / key-padding mask.
template <int NP, bool MASK>
__device__ __forceinline__ void fwd_block(float (&o)[8][4], float& m0, float& m1, float& l0, float& l1, uint32_t sA0, uint32_t sK0,
                                          uint32_t sV0, int k0, int S, int b, const uint8_t* kpm, float c, const LaneOff& off) {
    const int lane = threadIdx.x & 31;
    float s[2 * NP][4];
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_rows_tileT_impl<2 * NP, true>(s, sA0, sK0, NP, off);
    if (MASK) {
        const int kbase = k0 + (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < 2 * NP; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = kbase + nt * 8 + e;
                bool dead = key >= S;
                if (!dead && kpm) dead = kpm[(long long)b * S + key] != 0;
                if (dead) { s[nt][e] = -INFINITY; s[nt][e + 2] = -INFINITY; }
            }
        }
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 2 * NP; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float ms0 = (mx0 == -INFINITY) ? 0.f : mx0, ms1 = (mx1 == -INFINITY) ? 0.f : mx1;
    const float a0 = ex2((m0 - ms0) * c), a1 = ex2((m1 - ms1) * c);
    const float n0 = -ms0 * c, n1 = -ms1 * c;
    m0 = mx0; m1 = mx1;
    l0 *= a0; l1 *= a1;
#pragma unroll
    for (int nt = 0; nt < 2 * NP; ++nt) {
        s[nt][0] = ex2(fmaf(s[nt][0], c, n0)); s[nt][1] = ex2(fmaf(s[nt][1], c, n0));
        s[nt][2] = ex2(fmaf(s[nt][2], c, n1)); s[nt][3] = ex2(fmaf(s[nt][3], c, n1));
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= a0; o[nt][1] *= a0; o[nt][2] *= a1; o[nt][3] *= a1; }
    mma_p_rows_impl<NP, true>(o, s, sV0, NP, off);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 2) attn_fwd_short_kernel(const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = p.S;
    const int n_mt = (S + 15) >> 4, rows_pad = n_mt * 16;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + rows_pad * 128;
    uint8_t* sV = sK + rows_pad * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.x, b = blockIdx.y;
    const long long row_base = (long long)b * p.batch_stride;
    load_rows(smem_addr(sQ), p.q, p.ldq, p.tok_stride, row_base, rows_pad, S, h, NW * 32);
    load_rows(smem_addr(sK), p.k, p.ldk, p.tok_stride, row_base, rows_pad, S, h, NW * 32);
    load_rows(smem_addr(sV), p.v, p.ldv, p.tok_stride, row_base, rows_pad, S, h, NW * 32);
    cp_async_commit();
    const LaneOff off = make_lane_off();
    cp_async_wait<0>();
    __syncthreads();
    const uint32_t aQ = smem_addr(sQ), aK = smem_addr(sK), aV = smem_addr(sV);
    const int n_nt = (S + 7) >> 3;  // valid 8-key tiles
    const float c = p.scale_log2;

    for (int mt = warp; mt < n_mt; mt += NW) {
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // running max in raw-score units
        const int n_full = p.kpm ? 0 : (S >> 6);                     // 64-key blocks that need no masking at all
        int k0 = 0;
        for (; k0 < n_full * 64; k0 += 64)
            fwd_block<4, false>(o, m0, m1, l0, l1, aQ + mt * 2048, aK + k0 * 128, aV + k0 * 128, k0, S, b, p.kpm, c, off);
        for (; k0 < S; k0 += 64) {                                   // tail (or key-padding-masked) blocks
            const int pairs = (min(8, n_nt - (k0 >> 3)) + 1) >> 1;
            if (pairs == 4) fwd_block<4, true>(o, m0, m1, l0, l1, aQ + mt * 2048, aK + k0 * 128, aV + k0 * 128, k0, S, b, p.kpm, c, off);
            else if (pairs == 3) fwd_block<3, true>(o, m0, m1, l0, l1, aQ + mt * 2048, aK + k0 * 128, aV + k0 * 128, k0, S, b, p.kpm, c, off);
            else if (pairs == 2) fwd_block<2, true>(o, m0, m1, l0, l1, aQ + mt * 2048, aK + k0 * 128, aV + k0 * 128, k0, S, b, p.kpm, c, off);
            else fwd_block<1, true>(o, m0, m1, l0, l1, aQ + mt * 2048, aK + k0 * 128, aV + k0 * 128, k0, S, b, p.kpm, c, off);
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= inv0; o[nt][1] *= inv0; o[nt][2] *= inv1; o[nt][3] *= inv1; }
        if (p.lse && (lane & 3) == 0) {
            const int r0 = mt * 16 + (lane >> 2);
            float* lse = p.lse + ((long long)b * p.H + h) * S;
            if (r0 < S) lse[r0] = m0 * c + log2f(l0);
            if (r0 + 8 < S) lse[r0 + 8] = m1 * c + log2f(l1);
        }
        store_acc_tile(o, 1.0f, p.o, p.ldo, p.tok_stride, row_base, mt * 16, S, h);
    }
}

// Phase-A block: 16 keys (this warp) x NP*16 queries.  lse = +inf beyond S makes P vanish there; kill0/1 zero dead keys.
template <int NP>
__device__ __forceinline__ void bwd_a_block(float (&dk)[8][4], float (&dv)[8][4], uint32_t sK0, uint32_t sV0, uint32_t sQ0, uint32_t sdO0,
                                            const float* lse_q, const float* delta_q, float kill0, float kill1, float c, const LaneOff& off) {
    const int lane = threadIdx.x & 31;
    float st[2 * NP][4], dpt[2 * NP][4];
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
    mma_rows_tileT_impl<2 * NP, true>(st, sK0, sQ0, NP, off);     // S^T[key, q]
    mma_rows_tileT_impl<2 * NP, true>(dpt, sV0, sdO0, NP, off);   // dP^T[key, q]
    const float* lq = lse_q + (lane & 3) * 2;
    const float* dq_ = delta_q + (lane & 3) * 2;
#pragma unroll
    for (int nt = 0; nt < 2 * NP; ++nt) {
        const float2 l2 = *reinterpret_cast<const float2*>(lq + nt * 8);
        const float2 d2 = *reinterpret_cast<const float2*>(dq_ + nt * 8);
        const float p00 = ex2(fmaf(st[nt][0], c, -l2.x)) * kill0, p01 = ex2(fmaf(st[nt][1], c, -l2.y)) * kill0;
        const float p10 = ex2(fmaf(st[nt][2], c, -l2.x)) * kill1, p11 = ex2(fmaf(st[nt][3], c, -l2.y)) * kill1;
        st[nt][0] = p00; st[nt][1] = p01; st[nt][2] = p10; st[nt][3] = p11;
        dpt[nt][0] = p00 * (dpt[nt][0] - d2.x); dpt[nt][1] = p01 * (dpt[nt][1] - d2.y);
        dpt[nt][2] = p10 * (dpt[nt][2] - d2.x); dpt[nt][3] = p11 * (dpt[nt][3] - d2.y);
    }
    mma_p_rows_impl<NP, true>(dv, st, sdO0, NP, off);   // dV += P^T dO
    mma_p_rows_impl<NP, true>(dk, dpt, sQ0, NP, off);   // dK += dS^T Q
}

// Phase-B block: 16 queries (this warp) x NP*16 keys.
template <int NP, bool MASK>
__device__ __forceinline__ void bwd_b_block(float (&dq)[8][4], uint32_t sQ0, uint32_t sdO0, uint32_t sK0, uint32_t sV0, int k0, int S, int b,
                                            const uint8_t* kpm, float nl0, float nl1, float dl0, float dl1, float c, const LaneOff& off) {
    const int lane = threadIdx.x & 31;
    float s[2 * NP][4], dp[2 * NP][4];
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
    mma_rows_tileT_impl<2 * NP, true>(s, sQ0, sK0, NP, off);      // S[q, key]
    mma_rows_tileT_impl<2 * NP, true>(dp, sdO0, sV0, NP, off);    // dP[q, key]
    if (MASK) {
        const int kbase = k0 + (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < 2 * NP; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = kbase + nt * 8 + e;
                bool dead = key >= S;
                if (!dead && kpm) dead = kpm[(long long)b * S + key] != 0;
                if (dead) { s[nt][e] = -INFINITY; s[nt][e + 2] = -INFINITY; }
            }
        }
    }
#pragma unroll
    for (int nt = 0; nt < 2 * NP; ++nt) {
        const float p00 = ex2(fmaf(s[nt][0], c, nl0)), p01 = ex2(fmaf(s[nt][1], c, nl0));
        const float p10 = ex2(fmaf(s[nt][2], c, nl1)), p11 = ex2(fmaf(s[nt][3], c, nl1));
        dp[nt][0] = p00 * (dp[nt][0] - dl0); dp[nt][1] = p01 * (dp[nt][1] - dl0);
        dp[nt][2] = p10 * (dp[nt][2] - dl1); dp[nt][3] = p11 * (dp[nt][3] - dl1);
    }
    mma_p_rows_impl<NP, true>(dq, dp, sK0, NP, off);   // dQ += dS K
}

// Backward, whole head per CTA.  Phase A: warp owns 16 keys -> dK, dV (loops over 32-query blocks).
// Phase B: warp owns 16 queries -> dQ (loops over 32-key blocks).  delta = rowsum(dO o O) is computed in-kernel.
template <int NW>
__global__ void __launch_bounds__(NW * 32, 2) attn_bwd_short_kernel(const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = p.S;
    const int n_mt = (S + 15) >> 4, rows_pad = n_mt * 16;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + rows_pad * 128;
    uint8_t* sV = sK + rows_pad * 128;
    uint8_t* sdO = sV + rows_pad * 128;
    float* sLse = reinterpret_cast<float*>(sdO + rows_pad * 128);   // [rows_pad + 32], entries >= S hold +inf (=> P = 0)
    float* sDelta = sLse + rows_pad + 32;                           // [rows_pad + 32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.x, b = blockIdx.y;
    const long long row_base = (long long)b * p.batch_stride;
    load_rows(smem_addr(sQ), p.q, p.ldq, p.tok_stride, row_base, rows_pad, S, h, NW * 32);
    load_rows(smem_addr(sK), p.k, p.ldk, p.tok_stride, row_base, rows_pad, S, h, NW * 32);
    load_rows(smem_addr(sV), p.v, p.ldv, p.tok_stride, row_base, rows_pad, S, h, NW * 32);
    load_rows(smem_addr(sdO), p.dout, p.lddo, p.tok_stride, row_base, rows_pad, S, h, NW * 32);
    cp_async_commit();
    for (int r = threadIdx.x; r < rows_pad + 32; r += NW * 32) sLse[r] = r < S ? p.lse[((long long)b * p.H + h) * S + r] : INFINITY;
    const LaneOff off = make_lane_off();
    cp_async_wait<0>();
    __syncthreads();
    // delta[r] = sum_d dO[r,d] * O[r,d]  (dO from smem, O from global)
    for (int r = threadIdx.x; r < rows_pad + 32; r += NW * 32) {
        float acc = 0.f;
        if (r < S) {
            const __nv_bfloat16* orow = p.o + (row_base + (long long)r * p.tok_stride) * p.ldo + h * HD;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const uint4 ov = *reinterpret_cast<const uint4*>(orow + ch * 8);
                const uint4 dv = *reinterpret_cast<const uint4*>(sdO + tile_off(r, ch));
                const uint32_t oo[4] = {ov.x, ov.y, ov.z, ov.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc += __uint_as_float(oo[j] << 16) * __uint_as_float(dd[j] << 16) +
                           __uint_as_float(oo[j] & 0xFFFF0000u) * __uint_as_float(dd[j] & 0xFFFF0000u);
            }
        }
        sDelta[r] = acc;
    }
    __syncthreads();
    const uint32_t aQ = smem_addr(sQ), aK = smem_addr(sK), aV = smem_addr(sV), adO = smem_addr(sdO);
    const int n_nt = (S + 7) >> 3;
    const float c = p.scale_log2;

    // ---------------- phase A: dK, dV ----------------
    for (int mt = warp; mt < n_mt; mt += NW) {
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
        const int key0 = mt * 16 + (lane >> 2), key1 = key0 + 8;
        bool dead0 = key0 >= S, dead1 = key1 >= S;
        if (p.kpm) {
            if (!dead0) dead0 = p.kpm[(long long)b * S + key0] != 0;
            if (!dead1) dead1 = p.kpm[(long long)b * S + key1] != 0;
        }
        const float kill0 = dead0 ? 0.f : 1.f, kill1 = dead1 ? 0.f : 1.f;
        for (int q0 = 0; q0 < S; q0 += 32) {
            const int pairs = (min(4, n_nt - (q0 >> 3)) + 1) >> 1;
            if (pairs == 2) bwd_a_block<2>(dk, dv, aK + mt * 2048, aV + mt * 2048, aQ + q0 * 128, adO + q0 * 128, sLse + q0, sDelta + q0, kill0, kill1, c, off);
            else bwd_a_block<1>(dk, dv, aK + mt * 2048, aV + mt * 2048, aQ + q0 * 128, adO + q0 * 128, sLse + q0, sDelta + q0, kill0, kill1, c, off);
        }
        store_acc_tile(dk, p.scale, p.dk, p.lddk, p.tok_stride, row_base, mt * 16, S, h);
        store_acc_tile(dv, 1.0f, p.dv, p.lddv, p.tok_stride, row_base, mt * 16, S, h);
    }
    // ---------------- phase B: dQ ----------------
    for (int mt = warp; mt < n_mt; mt += NW) {
        float dq[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
        const int r0 = mt * 16 + (lane >> 2), r1 = r0 + 8;
        // rows >= S carry lse = +inf => P = 0 => dQ = 0 (and are never stored)
        const float nl0 = -sLse[r0], nl1 = -sLse[r1], dl0 = sDelta[r0], dl1 = sDelta[r1];
        const int n_full = p.kpm ? 0 : (S >> 5);
        int k0 = 0;
        for (; k0 < n_full * 32; k0 += 32)
            bwd_b_block<2, false>(dq, aQ + mt * 2048, adO + mt * 2048, aK + k0 * 128, aV + k0 * 128, k0, S, b, p.kpm, nl0, nl1, dl0, dl1, c, off);
        for (; k0 < S; k0 += 32) {
            const int pairs = (min(4, n_nt - (k0 >> 3)) + 1) >> 1;
            if (pairs == 2) bwd_b_block<2, true>(dq, aQ + mt * 2048, adO + mt * 2048, aK + k0 * 128, aV + k0 * 128, k0, S, b, p.kpm, nl0, nl1, dl0, dl1, c, off);
            else bwd_b_block<1, true>(dq, aQ + mt * 2048, adO + mt * 2048, aK + k0 * 128, aV + k0 * 128, k0, S, b, p.kpm, nl0, nl1, dl0, dl1, c, off);
        }
        store_acc_tile(dq, p.scale, p.dq, p.lddq, p.tok_stride, row_base, mt * 16, S, h);
    }
}

template <int NW>
static int launch_short_fwd(const AttnParams& p, cudaStream_t st) {
    const int rows_pad = ((p.S + 15) / 16) * 16;
    const int smem = 3 * rows_pad * 128;
    auto kern = attn_fwd_short_kernel<NW>;
    static DeviceOnce configured;   // opt in once per device for the largest shape this path serves (S <= 256)
    VB_ONCE_PER_DEVICE(configured, VB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 256 * 128)));
    kern<<<dim3(p.H, p.B), NW * 32, smem, st>>>(p);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}
template <int NW>
static int launch_short_bwd(const AttnParams& p, cudaStream_t st) {
    const int rows_pad = ((p.S + 15) / 16) * 16;
    const int smem = 4 * rows_pad * 128 + 2 * (rows_pad + 32) * (int)sizeof(float);
    auto kern = attn_bwd_short_kernel<NW>;
    static DeviceOnce configured;
    VB_ONCE_PER_DEVICE(configured, VB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                       4 * 256 * 128 + 2 * (256 + 32) * (int)sizeof(float))));
    kern<<<dim3(p.H, p.B), NW * 32, smem, st>>>(p);
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

// one thread block covers 32 (token, head) units per iteration, 4 iterations in flight
static unsigned delta_grid(long long units) {
    long long g = (units + 127) / 128;
    const long long cap = (long long)num_sms() * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

int attention_fwd_tc(const VbAttnDesc* d, cudaStream_t stream);   // attention_tc.cu (older tcgen05 path: S <= 256, key-padding masks)
int attention_fwd_tc3(const VbAttnDesc* d, cudaStream_t stream);  // attention_fwd_tc.cu (S <= 208, no mask: two-pass softmax, two threads per row)
int attention_fwd_tc4(const VbAttnDesc* d, cudaStream_t stream);  // attention_fwd_tc4.cu (193 <= S <= 208: score row in registers, four threads per row)
int attention_bwd_tc5(const VbAttnDesc* d, cudaStream_t stream);   // attention_bwd_tc.cu (five-product tcgen05 path, S <= 208)
int attention_fwd_gen(const VbAttnDesc* d, cudaStream_t stream);   // attention_fwd_tc.cu: any S / S_kv, masks, strides (tcgen05)
int attention_bwd_gen(const VbAttnDesc* d, cudaStream_t stream);   // attention_bwd_tc.cu
size_t attention_fwd_gen_workspace(const VbAttnDesc* d);
size_t attention_bwd_gen_workspace(const VbAttnDesc* d);

// The one-block tcgen05 kernels serve batch-first self-attention over <= 208 tokens without a mask (every ViT / DeiT config); everything
// else (DETR: S = 1050 / 4200, key-padding masks, sequence-first strides, 100 queries against the S-token memory) goes to the general
// tcgen05 kernels.  VITB200_ATTN_GEN=0 selects the mma.sync kernels of this file instead (A/B comparisons).
static bool one_block_shape(const VbAttnDesc* d) {
    const bool cross = d->S_kv > 0 && d->S_kv != d->S;
    return !cross && d->S <= 208 && d->tok_stride == 1 && d->key_padding_mask == nullptr;
}
static bool gen_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("VITB200_ATTN_GEN");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

static int check_common(const VbAttnDesc* d) {
    VB_REQUIRE(d != nullptr, "attention: null descriptor");
    VB_REQUIRE(d->head_dim == 64, "attention: head_dim %d unsupported (every reference config has 64)", d->head_dim);
    VB_REQUIRE(d->B > 0 && d->H > 0 && d->S > 0, "attention: bad dims B=%d H=%d S=%d", d->B, d->H, d->S);
    VB_REQUIRE(d->q && d->k && d->v && d->o, "attention: null tensor");
    VB_REQUIRE(d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && d->ldo % 8 == 0, "attention: row pitches must be multiples of 8");
    VB_REQUIRE(((uintptr_t)d->q & 15) == 0 && ((uintptr_t)d->k & 15) == 0 && ((uintptr_t)d->v & 15) == 0 && ((uintptr_t)d->o & 15) == 0,
               "attention: tensors must be 16-byte aligned");
    return VB_OK;
}

static AttnParams to_params(const VbAttnDesc* d) {
    AttnParams p{};
    p.q = (const __nv_bfloat16*)d->q; p.k = (const __nv_bfloat16*)d->k; p.v = (const __nv_bfloat16*)d->v;
    p.ldq = d->ldq; p.ldk = d->ldk; p.ldv = d->ldv;
    p.o = (__nv_bfloat16*)d->o; p.ldo = d->ldo;
    p.lse = d->lse; p.kpm = d->key_padding_mask;
    p.B = d->B; p.H = d->H; p.S = d->S;
    p.Sk = d->S_kv > 0 ? d->S_kv : d->S;
    p.tok_stride = d->tok_stride; p.batch_stride = d->batch_stride;
    p.scale = 0.125f;
    p.scale_log2 = 0.125f * 1.4426950408889634f;
    p.dout = (const __nv_bfloat16*)d->dout; p.lddo = d->lddo; p.delta = d->delta;
    p.dq = (__nv_bfloat16*)d->dq; p.dk = (__nv_bfloat16*)d->dk; p.dv = (__nv_bfloat16*)d->dv;
    p.lddq = d->lddq; p.lddk = d->lddk; p.lddv = d->lddv;
    if (d->dropout_p > 0.f) {
        p.drop_thresh = dropout_threshold(d->dropout_p);
        p.drop_inv_keep = 1.0f / (1.0f - d->dropout_p);
        p.drop_seed = d->dropout_seed;
        p.drop_stream = d->dropout_stream;
    }
    return p;
}

}  // namespace vb

// 64x64-tile mma.sync kernels: any S, key-padding masks, both layouts, optional attention dropout
static int launch_generic_fwd(const VbAttnDesc* d, const vb::AttnParams& p, cudaStream_t st) {
    using namespace vb;
    const int smem = 5 * TILE_BYTES;
    static DeviceOnce configured;
    if (!configured.is_set()) {
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.set();
    }
    dim3 grid((d->S + TILE - 1) / TILE, d->H, d->B);
    if (d->dropout_p > 0.f) {
        VB_REQUIRE(d->dropout_p < 1.f && d->dropout_seed, "attention: dropout needs 0 <= p < 1 and a device seed");
        attn_fwd_kernel<true><<<grid, 128, smem, st>>>(p);
    } else {
        attn_fwd_kernel<false><<<grid, 128, smem, st>>>(p);
    }
    VB_CUDA_CHECK(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_attention_fwd(const VbAttnDesc* d, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    if (int rc = check_common(d)) return rc;
    const AttnParams p = to_params(d);
    if (!one_block_shape(d) && gen_enabled()) return attention_fwd_gen(d, as_stream(stream));
    if (p.Sk != p.S) return launch_generic_fwd(d, p, as_stream(stream));   // cross-attention: query and key counts differ
    if (d->S <= 256) {
        int tc = attention_fwd_tc4(d, as_stream(stream));
        if (tc <= 0) return tc;
        tc = attention_fwd_tc3(d, as_stream(stream));
        if (tc <= 0) return tc;
        if (d->dropout_p > 0.f) return launch_generic_fwd(d, p, as_stream(stream));   // masks / sequence-first layouts with dropout
        tc = attention_fwd_tc(d, as_stream(stream));
        if (tc <= 0) return tc;   // launched (0) or failed with an error (< 0); 1 = shape not handled there
        const int n_mt = (d->S + 15) / 16;
        cudaStream_t st = as_stream(stream);
        if (n_mt <= 6) return launch_short_fwd<3>(p, st);
        if (n_mt <= 10) return launch_short_fwd<5>(p, st);
        if (n_mt <= 14) return launch_short_fwd<7>(p, st);
        return launch_short_fwd<8>(p, st);
    }
    return launch_generic_fwd(d, p, as_stream(stream));
}

extern "C" int vb_colsum_bf16(const void* x, int64_t ld, int32_t rows, int32_t cols, float* out_accum, void* stream);   // elementwise.cu

// paths that do not sum their output tiles in-kernel: dqkv_colsum[which] += column sums of dq / dk / dv
static int bwd_colsums(const VbAttnDesc* d, void* stream) {
    if (d->dqkv_colsum == nullptr) return VB_OK;
    const int cols = d->H * 64;
    const long long rows = (long long)d->B * d->S, rows_kv = (long long)d->B * (d->S_kv > 0 ? d->S_kv : d->S);
    // token rows are contiguous for both layouts the engine uses (batch-first: b*S + s; sequence-first: s*N + b)
    if (int rc = vb_colsum_bf16(d->dq, d->lddq, (int)rows, cols, d->dqkv_colsum, stream)) return rc;
    if (int rc = vb_colsum_bf16(d->dk, d->lddk, (int)rows_kv, cols, d->dqkv_colsum + cols, stream)) return rc;
    return vb_colsum_bf16(d->dv, d->lddv, (int)rows_kv, cols, d->dqkv_colsum + 2 * cols, stream);
}

extern "C" int64_t vb_attention_workspace_bytes(const VbAttnDesc* d, int32_t backward) {
    using namespace vb;
    if (d == nullptr || one_block_shape(d) || !gen_enabled()) return 0;
    return (int64_t)(backward ? attention_bwd_gen_workspace(d) : attention_fwd_gen_workspace(d));
}

extern "C" int vb_attention_bwd(const VbAttnDesc* d, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    if (int rc = check_common(d)) return rc;
    VB_REQUIRE(d->dout && d->dq && d->dk && d->dv && d->lse && d->delta, "attention_bwd: null tensor");
    VB_REQUIRE(d->lddo % 8 == 0 && d->lddq % 8 == 0 && d->lddk % 8 == 0 && d->lddv % 8 == 0, "attention_bwd: row pitches must be multiples of 8");
    const AttnParams p = to_params(d);
    cudaStream_t st = as_stream(stream);
    // S <= 208 without a key-padding mask (every ViT / DeiT config): five-product tcgen05 backward (attention_bwd_tc.cu).
    // VITB200_ATTN_TC_BWD=0 selects the mma.sync kernels below (A/B comparisons).
    const char* tc_env = getenv("VITB200_ATTN_TC_BWD");
    const bool cross = p.Sk != p.S;
    if (!one_block_shape(d) && gen_enabled()) {
        if (int rc = attention_bwd_gen(d, st)) return rc;
        return bwd_colsums(d, stream);
    }
    if (!cross && d->S <= 208 && d->tok_stride == 1 && d->key_padding_mask == nullptr && !(tc_env && tc_env[0] == '0')) {
        const int tc = attention_bwd_tc5(d, st);   // computes delta = rowsum(dO o O) itself
        if (tc <= 0) return tc;
    }
    const bool drop = d->dropout_p > 0.f;
    if (drop) VB_REQUIRE(d->dropout_p < 1.f && d->dropout_seed, "attention_bwd: dropout needs 0 <= p < 1 and a device seed");
    if (d->S <= 256 && !drop && !cross) {
        const int n_mt = (d->S + 15) / 16;
        int rc;
        if (n_mt <= 6) rc = launch_short_bwd<3>(p, st);
        else if (n_mt <= 10) rc = launch_short_bwd<5>(p, st);
        else if (n_mt <= 14) rc = launch_short_bwd<7>(p, st);
        else rc = launch_short_bwd<8>(p, st);
        return rc ? rc : bwd_colsums(d, stream);
    }
    const int smem = 6 * TILE_BYTES + 4 * TILE * (int)sizeof(float);
    static DeviceOnce configured;
    if (!configured.is_set()) {
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * TILE_BYTES));
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        VB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * TILE_BYTES));
        configured.set();
    }
    const long long nwarps = (long long)d->B * d->S * d->H;
    attn_delta_kernel<<<delta_grid(nwarps), 256, 0, st>>>(p);
    VB_CUDA_CHECK(cudaGetLastError());
    dim3 grid((d->S + TILE - 1) / TILE, d->H, d->B), grid_kv((p.Sk + TILE - 1) / TILE, d->H, d->B);
    if (drop) {
        attn_bwd_dkdv_kernel<true><<<grid_kv, 128, smem, st>>>(p);
        VB_CUDA_CHECK(cudaGetLastError());
        attn_bwd_dq_kernel<true><<<grid, 128, 6 * TILE_BYTES, st>>>(p);
    } else {
        attn_bwd_dkdv_kernel<false><<<grid_kv, 128, smem, st>>>(p);
        VB_CUDA_CHECK(cudaGetLastError());
        attn_bwd_dq_kernel<false><<<grid, 128, 6 * TILE_BYTES, st>>>(p);
    }
    VB_CUDA_CHECK(cudaGetLastError());
    return bwd_colsums(d, stream);
}
