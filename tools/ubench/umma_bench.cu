// Microbenchmarks of tcgen05.mma issue patterns, TMEM load/store throughput and a few ALU rates on sm_100a
// (design input for the attention kernels and the GEMM epilogues).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vitb200/csrc -o tools/ubench/umma_bench tools/ubench/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vb;

// TS: A from TMEM; N: MMA N; AMN/BMN: operand majors; NACC: independent accumulators cycled through; 64 MMAs per rep
template <int TS, int N, int AMN, int BMN, int NACC>
__global__ void __launch_bounds__(128, 1) mma_kernel(long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tb = tptr;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, AMN, BMN);
        constexpr uint64_t kd = umma_smem_desc_base(0, 1024);
        constexpr uint64_t md = umma_smem_desc_base(32768, 1024);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 65536);
        constexpr uint32_t acc_stride = (N + 31) / 32 * 32;
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
#pragma unroll
            for (int o = 0; o < 4; ++o) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int kk = i & 3;
                    const uint64_t ad = umma_smem_desc(AMN ? md : kd, sa + (AMN ? kk * 2048 : kk * 32));
                    const uint64_t bd = umma_smem_desc(BMN ? md : kd, sb + (BMN ? kk * 2048 : kk * 32));
                    const uint32_t d = tb + 256 + (i % NACC) * acc_stride;
                    if (TS) umma_bf16_ts(d, tb + (i * 8), bd, idesc, (o > 0 || i >= NACC) ? 1u : 0u);
                    else umma_bf16_ss(d, ad, bd, idesc, (o > 0 || i >= NACC) ? 1u : 0u);
                }
            }
            long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, rep & 1);
            long long t2 = clock64();
            out[rep * 2] = t1 - t0;
            out[rep * 2 + 1] = t2 - t0;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

// TMEM load/store throughput: blockDim/32 warps, each touches 256 columns of its lane quadrant `iters` times
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, long long* out, float* sink) {
    __shared__ uint32_t tptr;
    if (threadIdx.x < 32) tmem_alloc<512>(&tptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t t = tptr + (((warp & 3) * 32) << 16);
    uint32_t acc[4] = {0, 0, 0, 0};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t r[32];
            if (MODE == 0) {          // one x32 load, wait, consume
                tmem_ld_32x32b_x32(t + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i & 3] ^= r[i];
            } else if (MODE == 1) {   // two x32 loads in flight
                uint32_t r2[32];
                tmem_ld_32x32b_x32(t + c0, r);
                tmem_ld_32x32b_x32(t + c0 + 256, r2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i & 3] ^= r[i] + r2[i];
            } else if (MODE == 2) {   // x16 loads
                uint32_t(&lo)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
                uint32_t(&hi)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[16]);
                tmem_ld_32x32b_x16(t + c0, lo);
                tmem_ld_wait();
                tmem_ld_32x32b_x16(t + c0 + 16, hi);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i & 3] ^= r[i];
            } else {                  // stores
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = it + i;
#pragma unroll
                for (int q = 0; q < 4; ++q) tmem_st_32x32b_x8(t + c0 + q * 8, w);
                tmem_st_wait();
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = __uint_as_float(acc[0] ^ acc[1] ^ acc[2] ^ acc[3]);
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tptr);
}

__device__ __forceinline__ float ex2_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t fma2_(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2_(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// MODE 0 ex2, 1 ffma, 2 fma.f32x2, 3 rcp, 4 cvt.rn.bf16x2, 5 mul.f32x2, 6 ex2+2ffma mix (softmax-like)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) alu_kernel(int iters, long long* out, float* sink, float seed) {
    float a[8];
    uint64_t p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = threadIdx.x * 1e-3f + i * seed;
        float2 v = make_float2(a[i], a[i] + 0.5f);
        p[i] = *reinterpret_cast<uint64_t*>(&v);
    }
    float2 m2f = make_float2(0.999f * seed, 1.001f), c2f = make_float2(1e-3f, -1e-3f * seed);
    const uint64_t m2 = *reinterpret_cast<uint64_t*>(&m2f), c2 = *reinterpret_cast<uint64_t*>(&c2f);
    const float mf = 0.999f * seed, cf = 1e-3f * seed;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = ex2_(a[i]);
                else if (MODE == 1) a[i] = fmaf(a[i], mf, cf);
                else if (MODE == 2) p[i] = fma2_(p[i], m2, c2);
                else if (MODE == 3) a[i] = rcp_(a[i]);
                else if (MODE == 4) {
                    __nv_bfloat162 v = __floats2bfloat162_rn(a[i], a[(i + 1) & 7]);
                    a[i] = __uint_as_float(*reinterpret_cast<uint32_t*>(&v));
                } else if (MODE == 5) p[i] = mul2_(p[i], m2);
                else if (MODE == 6) a[i] = ex2_(fmaf(a[i], mf, cf)) * mf;
                else {   // 1 ex2 + 7 FMA-pipe ops per element (the attention-backward mix)
                    float x = ex2_(fmaf(a[i], mf, cf));
                    float y = fmaf(x, mf, cf); y = fmaf(y, mf, x); y = fmaf(y, cf, x); y = fmaf(y, mf, cf); y = fmaf(y, x, cf);
                    a[i] = y * mf;
                }
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&p[i]); s += a[i] + v.x + v.y; }
    sink[threadIdx.x] = s;
}

template <int TS, int N, int AMN, int BMN, int NACC>
void run_mma(const char* name, long long* out) {
    auto k = mma_kernel<TS, N, AMN, BMN, NACC>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<1, 128, 200 * 1024>>>(out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    printf("%-28s 64 MMAs: issue %6lld clk, complete %6lld clk -> %.1f clk/MMA (floor %.0f)\n", name, out[4], out[5], double(out[5]) / 64,
           128.0 * N / 256);
}
template <int MODE>
void run_tmem(const char* name, long long* out, float* sink) {
    for (int nw : {4, 8, 12, 16}) {
        const int iters = 64;
        tmem_kernel<MODE><<<1, nw * 32>>>(iters, out, sink);
        cudaDeviceSynchronize();
        const double bytes = double(nw) * 32 * 256 * 4 * iters * (MODE == 1 ? 2 : 1);
        printf("tmem %-22s warps %2d: %7lld clk, %.1f B/clk/SM, %.0f clk per 32x32 block per warp\n", name, nw, out[0], bytes / out[0],
               double(out[0]) / (iters * 8 * (MODE == 1 ? 2 : 1)));
    }
}
template <int MODE>
void run_alu(const char* name, long long* out, float* sink) {
    for (int nw : {4, 8, 16, 32}) {
        const int iters = 256;
        alu_kernel<MODE><<<1, nw * 32>>>(iters, out, sink, 1.0f);
        cudaDeviceSynchronize();
        const double ops = double(nw) * 32 * 32 * iters;
        printf("alu %-12s warps %2d: %7lld clk, %.1f thread-instr/clk/SM\n", name, nw, out[0], ops / out[0]);
    }
}

int main(int argc, char** argv) {
    long long* out; float* sink;
    const bool alu_only = argc > 1;
    cudaMallocManaged(&out, 64 * sizeof(long long));
    cudaMalloc(&sink, 4096 * sizeof(float));
    if (!alu_only) {
    run_mma<0, 256, 0, 0, 1>("SS N=256 K/K 1acc", out);
    run_mma<0, 208, 0, 0, 1>("SS N=208 K/K 1acc", out);
    run_mma<0, 128, 0, 0, 1>("SS N=128 K/K 1acc", out);
    run_mma<0, 64, 0, 0, 1>("SS N=64 K/K 1acc", out);
    run_mma<0, 64, 0, 0, 2>("SS N=64 K/K 2acc", out);
    run_mma<0, 64, 0, 0, 4>("SS N=64 K/K 4acc", out);
    run_mma<0, 64, 0, 1, 1>("SS N=64 K/MN 1acc", out);
    run_mma<0, 64, 1, 1, 1>("SS N=64 MN/MN 1acc", out);
    run_mma<0, 64, 1, 1, 4>("SS N=64 MN/MN 4acc", out);
    run_mma<0, 64, 1, 0, 1>("SS N=64 MN/K 1acc", out);
    run_mma<1, 64, 0, 1, 1>("TS N=64 -/MN 1acc", out);
    run_mma<1, 64, 0, 1, 2>("TS N=64 -/MN 2acc", out);
    run_mma<1, 64, 0, 1, 4>("TS N=64 -/MN 4acc", out);
    run_mma<1, 64, 0, 0, 1>("TS N=64 -/K 1acc", out);
    run_mma<1, 128, 0, 1, 1>("TS N=128 -/MN 1acc", out);
    run_mma<1, 208, 0, 0, 1>("TS N=208 -/K 1acc", out);
    run_mma<1, 256, 0, 0, 1>("TS N=256 -/K 1acc", out);
    run_mma<0, 112, 0, 0, 2>("SS N=112 K/K 2acc", out);
    run_mma<0, 96, 0, 0, 2>("SS N=96 K/K 2acc", out);
    run_tmem<0>("ld x32+wait", out, sink);
    run_tmem<1>("2 ld x32 then wait", out, sink);
    run_tmem<2>("ld x16+wait", out, sink);
    run_tmem<3>("4 st x8 then wait", out, sink);
    }
    run_alu<0>("ex2", out, sink);
    run_alu<1>("ffma", out, sink);
    run_alu<2>("fma.f32x2", out, sink);
    run_alu<3>("rcp", out, sink);
    run_alu<4>("cvt.bf16x2", out, sink);
    run_alu<5>("mul.f32x2", out, sink);
    run_alu<6>("ex2+ffma+fmul", out, sink);
    run_alu<7>("ex2+7fma", out, sink);
    printf("done\n");
    return 0;
}
