// Counter-based dropout masks shared by the element-wise dropout kernels and the attention kernels.
//
// keep(idx) = fmix32(fmix32(idx) + key) >> 8 >= round(p * 2^24), key = fmix32(seed ^ stream_id * 0x9E3779B9): two rounds of the
// murmur3 finaliser per element with the key added between them (a single keyed round fmix32(idx ^ key) would make every site's and
// every step's mask an XOR-permutation of ONE fixed table), no state, so forward and backward regenerate the same mask from
// (seed, stream, element index).  `seed` lives in device
// memory (the training step is replayed from a CUDA graph; the host bumps the counter between replays), `stream_id` names the
// dropout site (layer, position).  Replaces the Philox streams behind nn.Dropout / attention dropout at vanilla_vit.py:38,42,
// 67-68,94 — same Bernoulli(1-p) keep / scale-by-1/(1-p) semantics, different (documented) random stream.
#pragma once
#include <cstdint>
#ifndef __CUDACC__   // host-only builds (tests/test_dropout_cpu.py compiles this header with g++)
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace vb {

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ uint32_t dropout_key(uint32_t seed, uint32_t stream_id) { return fmix32(seed ^ (stream_id * 0x9E3779B9u)); }
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
    const float t = p * 16777216.0f;
    return t <= 0.f ? 0u : (t >= 16777216.0f ? 16777216u : (uint32_t)(t + 0.5f));
}
__host__ __device__ __forceinline__ bool dropout_keep(uint32_t key, uint32_t idx, uint32_t thresh24) {
    return (fmix32(fmix32(idx) + key) >> 8) >= thresh24;
}

}  // namespace vb
