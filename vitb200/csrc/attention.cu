// Fused multi-head attention, head_dim = 64: the C-ABI entry points and the dispatch between the tcgen05 kernels.
//
//   one-block shapes (batch-first self-attention over <= 208 tokens, no mask: every ViT / DeiT config)
//       forward   193 <= S <= 208, no dropout : attention_fwd_tc4.cu  (score row in registers, one TMEM pass)
//                 otherwise                   : attention_fwd_tc.cu   (two-pass softmax; attention dropout)
//       backward                              : attention_bwd_tc.cu   (five products, P computed once, delta in-kernel)
//   everything else (S > 208, key-padding masks, sequence-first strides, cross-attention: the DETR encoder / decoder)
//       the GEN instantiations of the same two kernels (key blocks + log-sum-exp merge; fp32 dK / dV reduce-add), which need the
//       caller-allocated VbAttnDesc.workspace (vb_attention_workspace_bytes)
// There is no mma.sync path any more (round 1 shipped 64x64-tile HMMA kernels for the general shapes; they are gone).
//
// Replaces F.scaled_dot_product_attention as reached from nn.MultiheadAttention at vanilla_vit.py:77
// (torch/nn/functional.py:6676-6688) and the explicit bmm / softmax / bmm path of the DETR layers
// (transformer.py:219, 145-147 -> torch/nn/functional.py:6630-6666), including the boolean key-padding mask.
// Q, K, V are read in place from the projection output (row = token, head h at column h*64), so no head-major copy is ever made;
// O is written token-major, ready to be the A operand of the out-proj GEMM.
// Token row index = b * batch_stride + s * tok_stride (batch-first ViT: (S,1); sequence-first DETR: (1,N)).
#include "common.h"
#include <cstdlib>

namespace vb {

int attention_fwd_tc3(const VbAttnDesc* d, cudaStream_t stream);   // attention_fwd_tc.cu
int attention_fwd_tc4(const VbAttnDesc* d, cudaStream_t stream);   // attention_fwd_tc4.cu
int attention_bwd_tc5(const VbAttnDesc* d, cudaStream_t stream);   // attention_bwd_tc.cu
int attention_fwd_gen(const VbAttnDesc* d, cudaStream_t stream);   // attention_fwd_tc.cu (GEN)
int attention_bwd_gen(const VbAttnDesc* d, cudaStream_t stream);   // attention_bwd_tc.cu (GEN)
size_t attention_fwd_gen_workspace(const VbAttnDesc* d);
size_t attention_bwd_gen_workspace(const VbAttnDesc* d);
void attention_bwd_tc5_set_debug(long long* p);
void attention_fwd_tc3_set_debug(long long* p);
void attention_fwd_tc4_set_debug(long long* p);

static int check_common(const VbAttnDesc* d) {
    VB_REQUIRE(d != nullptr, "attention: null descriptor");
    VB_REQUIRE(d->head_dim == 64, "attention: head_dim %d unsupported (every reference config has 64)", d->head_dim);
    VB_REQUIRE(d->B > 0 && d->H > 0 && d->S > 0 && d->S_kv >= 0, "attention: bad dims B=%d H=%d S=%d S_kv=%d", d->B, d->H, d->S, d->S_kv);
    VB_REQUIRE(d->q && d->k && d->v && d->o, "attention: null tensor");
    VB_REQUIRE(d->tok_stride >= 1 && d->batch_stride >= 1, "attention: bad token / batch strides");
    VB_REQUIRE(d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && d->ldo % 8 == 0, "attention: row pitches must be multiples of 8");
    VB_REQUIRE(((uintptr_t)d->q & 15) == 0 && ((uintptr_t)d->k & 15) == 0 && ((uintptr_t)d->v & 15) == 0 && ((uintptr_t)d->o & 15) == 0,
               "attention: tensors must be 16-byte aligned");
    VB_REQUIRE(d->dropout_p >= 0.f && d->dropout_p < 1.f && (d->dropout_p == 0.f || d->dropout_seed != nullptr),
               "attention: dropout needs 0 <= p < 1 and a device seed");
    return VB_OK;
}

static bool one_block_shape(const VbAttnDesc* d) {
    const bool cross = d->S_kv > 0 && d->S_kv != d->S;
    return !cross && d->S <= 208 && d->tok_stride == 1 && d->key_padding_mask == nullptr;
}

}  // namespace vb

extern "C" int vb_colsum_bf16(const void* x, int64_t ld, int32_t rows, int32_t cols, float* out_accum, void* stream);   // elementwise.cu

extern "C" int64_t vb_attention_workspace_bytes(const VbAttnDesc* d, int32_t backward) {
    using namespace vb;
    if (d == nullptr || one_block_shape(d)) return 0;
    return (int64_t)(backward ? attention_bwd_gen_workspace(d) : attention_fwd_gen_workspace(d));
}

extern "C" int vb_attention_fwd(const VbAttnDesc* d, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    if (int rc = check_common(d)) return rc;
    cudaStream_t st = as_stream(stream);
    if (!one_block_shape(d)) return attention_fwd_gen(d, st);
    int tc = attention_fwd_tc4(d, st);
    if (tc <= 0) return tc;      // launched (0) or failed with an error (< 0); 1 = shape not handled there
    tc = attention_fwd_tc3(d, st);
    if (tc <= 0) return tc;
    return fail(VB_ERR_UNSUPPORTED, "attention_fwd: no kernel for B=%d H=%d S=%d", d->B, d->H, d->S);
}

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

extern "C" int vb_attention_bwd(const VbAttnDesc* d, void* stream) {
    using namespace vb;
    if (int rc = check_arch()) return rc;
    if (int rc = check_common(d)) return rc;
    VB_REQUIRE(d->dout && d->dq && d->dk && d->dv && d->lse, "attention_bwd: null tensor");
    VB_REQUIRE(d->lddo % 8 == 0 && d->lddq % 8 == 0 && d->lddk % 8 == 0 && d->lddv % 8 == 0, "attention_bwd: row pitches must be multiples of 8");
    cudaStream_t st = as_stream(stream);
    if (!one_block_shape(d)) {
        if (int rc = attention_bwd_gen(d, st)) return rc;
        return bwd_colsums(d, stream);
    }
    const int tc = attention_bwd_tc5(d, st);   // computes delta = rowsum(dO o O) itself; sums dq / dv columns in its store warp
    if (tc <= 0) return tc;
    return fail(VB_ERR_UNSUPPORTED, "attention_bwd: no kernel for B=%d H=%d S=%d", d->B, d->H, d->S);
}

extern "C" VB_API int vb_debug_set_attn_timeline(void* device_buffer) {
    vb::attention_bwd_tc5_set_debug(reinterpret_cast<long long*>(device_buffer));
    vb::attention_fwd_tc3_set_debug(reinterpret_cast<long long*>(device_buffer));
    vb::attention_fwd_tc4_set_debug(reinterpret_cast<long long*>(device_buffer));
    return VB_OK;
}
