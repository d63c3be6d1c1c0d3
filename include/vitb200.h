/*
 * vitb200.h — C ABI of libvitb200.so: hand-written sm_100a kernels for the ViT / DeiT / DETR-encoder
 * forward+backward hot path of neeresh/vision-transformers.
 *
 * The reference has no FFI of its own: its boundary is the Python nn.Module contract, and
 * every arithmetic step is a PyTorch library call. Each entry point below therefore names the reference
 * call site (file:line under /root/reference, or torch/... for the PyTorch function the reference routes to)
 * whose arithmetic it replaces. The Python mirror of the module contract lives in vitb200/{vit,deit,detr}.py.
 *
 * Conventions
 *   - All pointers are caller-allocated DEVICE pointers (PyTorch's caching allocator owns every byte).
 *     The library never frees or retains them past the call.
 *   - Every entry point takes an explicit cudaStream_t (as void*), launches asynchronously and never
 *     synchronises the device.
 *   - Return value: 0 on success, a negative VB_ERR_* code otherwise; vb_last_error() gives the text.
 *     Nothing throws across the boundary and nothing calls exit().
 *   - There is no CPU fallback and no multi-backend dispatch: on a device that is not sm_100 the compute
 *     entry points return VB_ERR_ARCH.
 *   - Matrices are row-major; "ld*" are row pitches in ELEMENTS. bf16 = 16-bit brain float.
 */
#ifndef VITB200_H
#define VITB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VB_API __attribute__((visibility("default")))
#else
#define VB_API
#endif

#define VB_OK 0
#define VB_ERR_ARG (-1)     /* bad argument (alignment, shape, null pointer) */
#define VB_ERR_CUDA (-2)    /* CUDA runtime / driver error */
#define VB_ERR_ARCH (-3)    /* device is not sm_100 */
#define VB_ERR_UNSUPPORTED (-4)

/* ---- library ---------------------------------------------------------------------------------------- */
VB_API int vb_version(void);                 /* ABI version, currently 1 */
VB_API const char* vb_last_error(void);      /* thread-local text of the last failure */
VB_API int vb_device_check(int device);      /* VB_OK iff `device` is an sm_100 part and the kernels can load */
VB_API int vb_sm_count(int device);          /* number of SMs, or a negative error */

/* ---- GEMM: C[M,N] = epilogue( A[M,K] * B[N,K]^T ), bf16 operands, fp32 accumulation in TMEM ----------
 * Replaces every nn.Linear / F.linear on the path and its autograd dgrad/wgrad:
 *   QKV in-proj   torch/nn/functional.py:5835-5847 (called from vanilla_vit.py:77)
 *   out-proj      torch/nn/functional.py:6690 + residual vanilla_vit.py:78-79
 *   mlp.0 + GELU  vanilla_vit.py:33-38,50        mlp.3 + residual vanilla_vit.py:41-42,83
 *   head          vanilla_vit.py:212-213         conv_proj as GEMM vanilla_vit.py:129,196-198
 *   DETR linear1/linear2/in-proj  transformer.py:195-199,218-224
 * Operand majors: 0 = the contraction index K is contiguous (matrix stored [rows, K]);
 *                 1 = the row index is contiguous (matrix stored [K, rows]).
 *   forward  Y = X W^T           : a_major 0, b_major 0
 *   dgrad    dX = dY W           : a_major 0 (dY [M,N]), b_major 1 (W stored [N_w, K_w], contraction over N_w)
 *   wgrad    dW += dY^T X        : a_major 1 (dY stored [tokens, N_out]), b_major 1 (X stored [tokens, K_in])
 */
enum {
    VB_EPI_STORE = 0,    /* C = acc + bias                                                */
    VB_EPI_GELU = 1,     /* x = acc + bias; C2 = gelu_erf(x); C (optional) = gelu_erf'(x)  */
    VB_EPI_RESIDUAL = 2, /* C = acc + bias + AUX                                          */
    VB_EPI_RELU = 3,     /* C = max(acc + bias, 0)                                        */
    VB_EPI_DGELU = 4,    /* C = acc * AUX                   (AUX = gelu'(x) saved by VB_EPI_GELU) */
    VB_EPI_DRELU = 5,    /* C = AUX > 0 ? acc * drelu_scale : 0   (AUX = saved relu output) */
    VB_EPI_ACCUM = 6     /* C += acc  (fp32 C, TMA reduce-add; used by wgrad and split-K) */
};
enum { VB_BF16 = 0, VB_F32 = 1 };

typedef struct VbGemmDesc {
    int32_t M, N, K;          /* per-batch GEMM dims */
    int32_t batches;          /* >= 1; batch b uses A + b*batch_stride_a etc. */
    int32_t a_major, b_major; /* 0 = K contiguous, 1 = row index contiguous */
    int32_t epilogue;         /* VB_EPI_* */
    int32_t c_dtype;          /* VB_BF16 or VB_F32; AUX and C2 use the same dtype */
    int32_t split_k;          /* >= 1; > 1 requires VB_EPI_ACCUM */
    int32_t c_row_offset;     /* added to the row index when writing C/C2 and reading AUX (token offset) */
    int32_t c_rows;           /* rows per batch of C/C2/AUX (>= M + c_row_offset); 0 means M */
    int32_t aux_batch_broadcast; /* 1: AUX has a single batch that every batch reads (position embedding) */
    const void* A; int64_t lda; int64_t batch_stride_a;   /* bf16 */
    const void* B; int64_t ldb; int64_t batch_stride_b;   /* bf16; batch_stride_b = 0 shares B */
    void* C; int64_t ldc; int64_t batch_stride_c;
    void* C2; int64_t ldc2; int64_t batch_stride_c2;       /* VB_EPI_GELU second output or NULL */
    const void* AUX; int64_t ldaux; int64_t batch_stride_aux;
    const float* bias;        /* [N] fp32 or NULL */
    int32_t max_ctas;         /* 0 = one persistent CTA per SM */
    int32_t debug_direct_store; /* 1 = bypass the TMA-store epilogue (slow, for bring-up tests) */
    float drelu_scale;        /* VB_EPI_DRELU only; 0 means 1.  1/(1-p) when AUX is the DROPPED relu output keep*relu(x)/(1-p)
                               * (transformer.py:223: linear2(dropout(activation(linear1(src))))) */
    int32_t reserved0;
    float* a_colsum;          /* optional, a_major = 1 and batches = 1 only: a_colsum[m] += sum_k A[k, m] (fp32, atomicAdd), summed from the
                               * operand tiles while they sit in shared memory for the MMA.  In a weight-gradient GEMM dW = dY^T X the A
                               * operand is dY [tokens, out]: this is the bias gradient of the same nn.Linear (vanilla_vit.py:34,40;
                               * torch/nn/functional.py:5835 in_proj_bias) with no second pass over dY */
} VbGemmDesc;

VB_API int vb_gemm_bf16(const VbGemmDesc* desc, void* stream);

/* ---- LayerNorm over the last dimension of an fp32 [rows, dim] stream -------------------------------------
 * Replaces nn.LayerNorm(eps=1e-6) vanilla_vit.py:66,70,100 (ATen native_layer_norm) and the eps=1e-5 norms of
 * transformer.py:201-202.  dim must be a multiple of 128 and <= 1024.  Row pitches in elements.
 * fwd: y = (x - mean) * rstd * gamma + beta, written as bf16 (y_bf16) and/or fp32 (y_f32); mean/rstd optional;
 *      y2_bf16 (optional) = bf16(y + add) with add fp32 [rows, dim] (DETR q = k = src + pos, transformer.py:218).
 * bwd: dx = dres + LN'(dy + dy_add); dy_add (optional) is bf16; dgamma/dbeta/dx_colsum are ACCUMULATED
 *      (atomicAdd) and may be NULL; dx_colsum[c] += sum_rows dx[row, c] (bias gradient of the linear layer
 *      feeding this residual stream). */
VB_API int vb_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y_bf16,
                            int64_t ldy_bf16, float* y_f32, int64_t ldy_f32, float* mean, float* rstd, int32_t rows,
                            int32_t dim, float eps, const float* add, int64_t ldadd, void* y2_bf16, int64_t ldy2,
                            void* stream);
VB_API int vb_layernorm_bwd(const void* dy, int32_t dy_dtype, int64_t lddy, const float* x, int64_t ldx,
                            const float* mean, const float* rstd, const float* gamma, const float* dres, int64_t lddres,
                            float* dx, int64_t lddx, void* dx_bf16, int64_t lddx_bf16, float* dgamma, float* dbeta,
                            float* dx_colsum, int32_t rows, int32_t dim, const void* dy_add_bf16, int64_t lddy_add,
                            void* stream);

/* ---- Fused multi-head attention, head_dim 64 ---------------------------------------------------------------
 * Replaces F.scaled_dot_product_attention reached from nn.MultiheadAttention at vanilla_vit.py:77
 * (torch/nn/functional.py:6676-6688) and the explicit softmax path of transformer.py:219
 * (torch/nn/functional.py:6630-6666) with its boolean key_padding_mask (1 = ignore key).
 * q/k/v/o (and dq/dk/dv/dout) are bf16 matrices whose row is a token and whose columns [h*64, h*64+64) hold
 * head h; the row of token s of batch element b is b*batch_stride + s*tok_stride (ViT batch-first: S, 1;
 * DETR sequence-first: 1, N).  lse is [B,H,S] fp32 in the log2 domain; delta is [B,H,S] fp32 scratch.
 * The softmax scale is 1/sqrt(64), as in the reference. */
typedef struct VbAttnDesc {
    int32_t B, H, S, head_dim;
    int64_t tok_stride, batch_stride;
    const void* q; const void* k; const void* v;
    int64_t ldq, ldk, ldv;
    void* o; int64_t ldo;
    float* lse;
    const uint8_t* key_padding_mask; /* [B,S] or NULL */
    const void* dout; int64_t lddo;  /* backward only from here */
    float* delta;
    void* dq; void* dk; void* dv;
    int64_t lddq, lddk, lddv;
    /* attention dropout (nn.MultiheadAttention(dropout=p), vanilla_vit.py:67): P is multiplied by keep / (1 - p) before P V;
     * the mask is regenerated in backward from (*dropout_seed, dropout_stream, element).  p = 0 disables.  Only the tcgen05
     * kernels (S <= 208, batch-first, no key-padding mask) and the 64x64-tile mma.sync kernels (everything else) implement it. */
    float dropout_p;
    uint32_t dropout_stream;
    const uint32_t* dropout_seed; /* device pointer */
    /* backward only, optional: fp32 [3][H*64]; the column sums of dq | dk | dv over all tokens are ACCUMULATED into it (the bias
     * gradient of the packed in-projection, torch/nn/functional.py:5835-5847).  The tcgen05 backward sums its accumulator slices
     * in registers (warp butterfly + atomicAdd) and leaves the dk part untouched: sum_k dS[q,k] = 0 makes the key-bias gradient
     * exactly zero (softmax ignores a constant key offset).  The other paths run the column-sum kernel on dq, dk, dv. */
    float* dqkv_colsum;
    /* number of keys / values when it differs from the number of queries S (cross-attention of the DETR decoder layer,
     * transformer.py:145-147: 100 queries against the S-token memory); 0 = self-attention.  key_padding_mask is then [B,S_kv], lse and
     * delta stay [B,H,S].  Served by the 64x64-tile mma.sync kernels. */
    int32_t S_kv;
    int32_t reserved0;
    /* scratch for the shapes the one-block kernels do not cover (S > 208, key-padding masks, sequence-first strides, cross-
     * attention): forward = per-key-block partial outputs + log-sum-exps, backward = fp32 dK / dV accumulators.  Caller-allocated
     * device memory, 256-byte aligned, at least vb_attention_workspace_bytes() bytes; may be NULL when that returns 0. */
    void* workspace;
    int64_t workspace_bytes;
} VbAttnDesc;
VB_API int vb_attention_fwd(const VbAttnDesc* desc, void* stream);
VB_API int vb_attention_bwd(const VbAttnDesc* desc, void* stream);
/* bytes of VbAttnDesc.workspace the forward (backward = 0) or backward (backward = 1) call of this descriptor needs; 0 = none */
VB_API int64_t vb_attention_workspace_bytes(const VbAttnDesc* desc, int32_t backward);
/* profiling aid: device buffer (>= 1024 int64) receiving in-kernel cycle stamps of the tcgen05 backward; NULL disables */
VB_API int vb_debug_set_attn_timeline(void* device_buffer);
/* same for the GEMM: >= 8 int64 of cluster 0's MMA issuer (total cycles, waiting for operand stages, for a free accumulator, for
 * the tile scheduler, tiles issued); NULL disables */
VB_API int vb_debug_set_gemm_timeline(void* device_buffer);

/* ---- Helper kernels around the encoder (HBM-bound) -----------------------------------------------------------
 * vb_cast_f32_to_bf16: flat fp32 -> bf16 cast (master parameters -> tensor-core operands), n % 4 == 0.
 * vb_patchify: images [B,C,H,W] fp32 -> [B, (H/p)(W/p), C*p*p] bf16 with k = c*p*p + i*p + j, the operand of
 *              conv_proj-as-GEMM (nn.Conv2d(k=s=p) + reshape/permute, vanilla_vit.py:129,196-198).
 * vb_token_rows: x[b,t,:] = token_t + pos[t] (pos may be NULL: CPVT / CPE-ViT add it later) for the n_prefix (1 = cls, 2 = cls+dist) leading rows of the
 *              fp32 stream [B,S,D] (torch.cat + pos add, vanilla_vit.py:202-203,104).
 * vb_colsum_bf16: out[c] += sum_r x[r,c] (bias gradients of nn.Linear, autograd of vanilla_vit.py:34,41).
 * vb_embed_bwd: from dx [B,S,D] fp32: dpos += sum_b dx[b]; dtok_t += sum_b dx[b,t]; dbias += sum_{b,s>=n_prefix} dx[b,s];
 *              dx_patches_bf16 [B*(S-n_prefix), D] = compact bf16 copy (A operand of the conv_proj wgrad GEMM).
 *              possum_scratch is [S,D] fp32 scratch. */
VB_API int vb_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* out_bf16 = bf16(a + b), b may be NULL (DETR: bf16 operand copies of src and src + pos, transformer.py:218) */
VB_API int vb_add_cast_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream);
/* out[r,:] = x[r,:] + pos[r % period,:], fp32 contiguous [rows, D] (Encoder.forward on caller-supplied tokens: input + pos_embedding,
 * vanilla_vit.py:104; period = sequence length) */
VB_API int vb_add_rows_bcast(const float* x, const float* pos, float* out, int64_t rows, int32_t period, int32_t D, void* stream);
/* out = a + b_bf16 (+ c_bf16); accum += b_bf16 if accum != NULL (DETR backward: d_src and d_pos assembly) */
VB_API int vb_add3(const float* a, const void* b_bf16, const void* c_bf16, float* out, float* accum, int64_t n, void* stream);
/* dst_bf16[r, c] = bf16(src[r, c]), row-pitched (pitches in elements, lddst even): the NCHW backbone feature map as the bf16 operand of
 * input_proj, the 1x1 convolution in front of the DETR transformer (detr.py:125), run as a GEMM with a_major = 1 */
VB_API int vb_cast_rows_bf16(const float* src, int64_t ldsrc, void* dst_bf16, int64_t lddst, int64_t rows, int32_t cols, void* stream);
/* AbsolutePositionalEncoding (detr.py:33-63) in the encoder's sequence-first layout: pos[(y*w + x), n, :] = [col_embed[x] | row_embed[y]]
 * (fp32, [h*w, N, 2*pf]); bwd ACCUMULATES the embedding-table gradients from dpos of the same layout */
VB_API int vb_pos_embed_2d_fwd(const float* row_embed, const float* col_embed, float* pos, int32_t h, int32_t w, int32_t N, int32_t pf,
                               void* stream);
VB_API int vb_pos_embed_2d_bwd(const float* dpos, float* drow_accum, float* dcol_accum, int32_t h, int32_t w, int32_t N, int32_t pf,
                               void* stream);
VB_API int vb_patchify(const float* images, void* patches_bf16, int32_t B, int32_t C, int32_t H, int32_t W, int32_t patch,
                       void* stream);
/* inverse of vb_patchify: d images [B,C,H,W] fp32 from the bf16 patch-matrix gradient (input gradient of conv_proj, autograd of
 * vanilla_vit.py:196) */
VB_API int vb_unpatchify(const void* dpatches_bf16, float* dimages, int32_t B, int32_t C, int32_t H, int32_t W, int32_t patch,
                         void* stream);
VB_API int vb_token_rows(float* x, const float* tok0, const float* tok1, const float* pos, int32_t B, int32_t S, int32_t D,
                         int32_t n_prefix, void* stream);
VB_API int vb_colsum_bf16(const void* x, int64_t ld, int32_t rows, int32_t cols, float* out_accum, void* stream);
/* y_accum[n] += sum_k x[k] * W[k, n] (fp32, W row-major [K, N] with pitch ldw); x_accum[k] += x[k] if x_accum != NULL.  Used for the value-
 * bias gradient of the packed in-projection: colsum(dV) = colsum(dO) = colsum(d) W_out_proj (torch/nn/functional.py:5835-5847, 6690). */
VB_API int vb_vecmat_accum(const float* x, const float* W, int64_t ldw, float* y_accum, float* x_accum, int32_t K, int32_t N, void* stream);
VB_API int vb_embed_bwd(const float* dx, float* possum_scratch, void* dx_patches_bf16, float* dpos, float* dtok0, float* dtok1,
                        float* dbias, int32_t B, int32_t S, int32_t D, int32_t n_prefix, void* stream);

/* ---- Dropout ------------------------------------------------------------------------------
 * Hidden-state dropout of the ViT encoder: nn.Dropout at vanilla_vit.py:38 (mlp.2), :42 (mlp.4), :68/:78
 * (EncoderBlock.dropout) and :94/:104 (Encoder.dropout).  keep(i) = hash(*seed_dev, stream_id, i) >= p with i the
 * row-major element index (row * cols + col); kept elements are scaled by 1 / (1 - p).  The same (seed, stream_id,
 * rows x cols) regenerates the mask in backward.  The random stream differs from PyTorch's Philox stream (documented).
 * vb_dropout_f32: dst = keep ? src / (1 - p) : 0  (+ aux), written as fp32 (dst_f32) and/or bf16 (dst_bf16); in place allowed.
 * vb_dropout_bf16_pair: in place on x1 and (optional) x2 with one mask (GELU output and its saved derivative).
 * vb_dropout_mask_u8: the keep mask itself (tests / oracle). */
VB_API int vb_dropout_f32(const float* src, int64_t ldsrc, const float* aux, int64_t ldaux, float* dst_f32, int64_t lddst,
                          void* dst_bf16, int64_t lddstb, int32_t rows, int32_t cols, float p, const uint32_t* seed_dev,
                          uint32_t stream_id, void* stream);
VB_API int vb_dropout_bf16_pair(void* x1, void* x2, int64_t ld, int32_t rows, int32_t cols, float p, const uint32_t* seed_dev,
                                uint32_t stream_id, void* stream);
VB_API int vb_dropout_mask_u8(uint8_t* out, int64_t n, float p, const uint32_t* seed_dev, uint32_t stream_id, void* stream);

/* ---- Conditional positional encoding / PEG: depthwise 3x3 convolution over the G x G patch-token grid of a
 * token-major fp32 stream [B, S = n_prefix + G*G, D]; the prefix (class) rows pass through.  Replaces ConditionalPositionalEncoding
 * (cpe_vit.py:16-30 == cpvt.py:16-30: nn.Conv2d(D, D, 3, padding=1, groups=D) between two layout transposes).  w is the conv weight
 * [D, 1, 3, 3] as stored by nn.Conv2d, bias [D].
 *   fwd:        out = [x | conv(x) + bias] (+ pos[s,:] broadcast over the batch, NULL to skip) (+ x - sub, NULL to skip)
 *   bwd_data:   dx = [dy | conv^T(dy)] (optional); sum_f32 / sum_bf16 = dy + dx (optional; CPVT block tail, cpvt.py:93-96)
 *   bwd_weight: dw += sum dy * shifted x, db += sum dy over the patch rows (fp32 accumulation into the gradient buffer) */
VB_API int vb_dwconv3x3_fwd(const float* x, const float* w, const float* bias, const float* pos, const float* sub, float* out,
                            int32_t B, int32_t S, int32_t D, int32_t n_prefix, int32_t G, void* stream);
VB_API int vb_dwconv3x3_bwd_data(const float* dy, const float* w, float* dx, float* sum_f32, void* sum_bf16, int32_t B, int32_t S,
                                 int32_t D, int32_t n_prefix, int32_t G, void* stream);
VB_API int vb_dwconv3x3_bwd_weight(const float* dy, const float* x, float* dw_accum, float* db_accum, int32_t B, int32_t S, int32_t D,
                                   int32_t n_prefix, int32_t G, void* stream);

/* ---- Training-step tail -------------------------------------------------------------------
 * vb_cross_entropy: mean-reduction softmax cross-entropy (nn.CrossEntropyLoss at vanilla_vit.py:220,237) forward
 *   and gradient in one pass: *loss_accum += weight * sum_b (lse(z_b) - z_b[y_b]); dz = grad_scale * weight *
 *   (softmax(z) - onehot(y)) written as bf16 and/or fp32 (pass weight = 1/B for the mean).  labels are int64.
 *   correct_accum (optional) counts argmax(z) == y.
 * vb_adam_step: torch.optim.Adam semantics (vanilla_vit.py:221: lr 1e-4, betas .9/.999, eps 1e-8) over flat
 *   fp32 buffers, n % 4 == 0; gradients are multiplied by grad_scale first; params_bf16 (optional) receives the
 *   refreshed bf16 shadow used by the GEMMs.  If step_counter_dev != NULL the step number lives on the device
 *   (incremented by the call), so that a whole training step can be replayed from one CUDA graph. */
VB_API int vb_cross_entropy(const float* logits, int64_t ld, const int64_t* labels, int32_t B, int32_t C, float* loss_accum,
                            float weight, void* dlogits_bf16, int64_t lddz, float* dlogits_f32, int64_t lddzf,
                            float grad_scale, int32_t* correct_accum, void* stream);
/* vb_distill_loss: DistillationLoss.forward (utils/distillation_loss.py:30-75) on the two student logits and the teacher's logits,
 *   forward + both gradients in one pass: *loss_accum += (1-alpha)*CE(z, y) + alpha*D with D = CE(z_kd, argmax teacher) for
 *   kind = 2 ('hard', :70-71; first maximal index as torch.argmax) or D = tau^2/(B*C) * KL(softmax(teacher/tau) || softmax(z_kd/tau))
 *   for kind = 1 ('soft', :55-65, the legacy numel() normalisation).  Gradients (times grad_scale) go to bf16 and/or fp32 buffers;
 *   correct_accum counts argmax(z) == y (the accuracy deit.py:72-74 tracks).  'none' is vb_cross_entropy. */
VB_API int vb_distill_loss(const float* logits, int64_t ld, const float* logits_kd, int64_t ldkd, const float* teacher_logits,
                           int64_t ldt, const int64_t* labels, int32_t B, int32_t C, int32_t kind, float alpha, float tau,
                           float* loss_accum, void* dlogits_bf16, int64_t lddz, void* dlogits_kd_bf16, int64_t lddzkd,
                           float* dlogits_f32, int64_t lddzf, float* dlogits_kd_f32, int64_t lddzkdf, float grad_scale,
                           int32_t* correct_accum, void* stream);
VB_API int vb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16, int64_t n,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                        int32_t* step_counter_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITB200_H */
