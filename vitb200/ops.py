"""Tensor-level wrappers over the C ABI: torch tensors in, raw device pointers + strides out.

PyTorch is used for device memory and streams only; every arithmetic step happens in libvitb200.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import VbGemmDesc

# kernels launched by libvitb200.so since import (bench.py reports the per-step delta as `gpu_launches`)
LAUNCHES = {"n": 0}
_KERNELS_PER_CALL = {"vb_gemm_bf16": 1, "vb_layernorm_fwd": 1, "vb_layernorm_bwd": 1, "vb_attention_fwd": 1, "vb_attention_bwd": 1,
                     "vb_cast_f32_to_bf16": 1, "vb_patchify": 1, "vb_token_rows": 1, "vb_colsum_bf16": 1, "vb_embed_bwd": 2,
                     "vb_cross_entropy": 1, "vb_adam_step": 2, "vb_add_cast_bf16": 1, "vb_add3": 1, "vb_add_rows_bcast": 1, "vb_dropout_f32": 1, "vb_dropout_bf16_pair": 1,
                     "vb_dropout_mask_u8": 1, "vb_cast_rows_bf16": 1, "vb_pos_embed_2d_fwd": 1, "vb_pos_embed_2d_bwd": 1, "vb_vecmat_accum": 1}

EPI_STORE, EPI_GELU, EPI_RESIDUAL, EPI_RELU, EPI_DGELU, EPI_DRELU, EPI_ACCUM = range(7)
BF16, F32 = 0, 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _mat(t, name):
    """Returns (ptr, ld, batch_stride, batches, rows, cols) of a 2-D or 3-D row-major tensor."""
    if t.dim() == 2:
        assert t.stride(1) == 1, f"{name}: innermost stride must be 1"
        return t.data_ptr(), t.stride(0), 0, 1, t.shape[0], t.shape[1]
    assert t.dim() == 3 and t.stride(2) == 1, f"{name}: expected [batch, rows, cols] with unit inner stride"
    return t.data_ptr(), t.stride(1), t.stride(0), t.shape[0], t.shape[1], t.shape[2]


def gemm(A, B, C, *, a_major=0, b_major=0, epilogue=EPI_STORE, bias=None, aux=None, C2=None, split_k=1,
         c_row_offset=0, aux_broadcast=False, max_ctas=0, direct=False, drelu_scale=1.0, a_colsum=None):
    """C = epilogue(A @ B^T). ``A`` is [M,K] (a_major 0) or stored [K,M] (a_major 1); ``B`` is [N,K] or stored [K,N].

    3-D tensors add a leading batch dimension (B may stay 2-D to be shared). ``C``/``C2``/``aux`` are [M(+off),N].
    ``a_colsum`` (fp32 [M], a_major 1 only): += the sum of A over the contraction index, taken from the operand tiles in shared memory
    (the bias gradient of a weight-gradient GEMM).
    """
    lib = _lib.load()
    d = VbGemmDesc()
    pa, lda, bsa, nba, ra, ca = _mat(A, "A")
    pb, ldb, bsb, nbb, rb, cb = _mat(B, "B")
    M, K = (ra, ca) if a_major == 0 else (ca, ra)
    N, Kb = (rb, cb) if b_major == 0 else (cb, rb)
    assert K == Kb, f"contraction mismatch: A gives K={K}, B gives K={Kb}"
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    out = C if C is not None else C2
    pc, ldc, bsc, nbc, rc, cc = _mat(out, "C")
    assert cc == N and nbc == nba, (cc, N, nbc, nba)
    d.M, d.N, d.K, d.batches = M, N, K, nba
    d.a_major, d.b_major, d.epilogue, d.c_dtype = a_major, b_major, epilogue, _dt(out)
    d.split_k, d.c_row_offset, d.c_rows = split_k, c_row_offset, rc
    d.aux_batch_broadcast = 1 if aux_broadcast else 0
    d.A, d.lda, d.batch_stride_a = pa, lda, bsa
    d.B, d.ldb, d.batch_stride_b = pb, ldb, (bsb if nbb > 1 else 0)
    if C is not None:
        d.C, d.ldc, d.batch_stride_c = pc, ldc, bsc
    if C2 is not None:
        p2, ld2, bs2, _, _, _ = _mat(C2, "C2")
        d.C2, d.ldc2, d.batch_stride_c2 = p2, ld2, bs2
    if aux is not None:
        assert aux.dtype == out.dtype, "AUX must have the dtype of C"
        px, ldx, bsx, _, _, _ = _mat(aux, "aux")
        d.AUX, d.ldaux, d.batch_stride_aux = px, ldx, bsx
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
        d.bias = bias.data_ptr()
    d.max_ctas = max_ctas
    d.debug_direct_store = int(direct)
    d.drelu_scale = float(drelu_scale)
    if a_colsum is not None:
        assert a_colsum.dtype == torch.float32 and a_colsum.is_contiguous() and a_colsum.numel() == M and a_major == 1 and nba == 1
        d.a_colsum = a_colsum.data_ptr()
    _lib.check(lib.vb_gemm_bf16(ctypes.byref(d), _stream()), "vb_gemm_bf16")
    return out


def _p(t):
    return None if t is None else t.data_ptr()


def layernorm_fwd(x, gamma, beta, eps, *, y_bf16=None, y_f32=None, mean=None, rstd=None, add=None, y2_bf16=None):
    """x: fp32 [rows, D] view (unit inner stride). Writes y_bf16 and/or y_f32 (same shape) and optional mean/rstd."""
    lib = _lib.load()
    rows, D = x.shape
    assert x.dtype == torch.float32 and x.stride(1) == 1
    rc = lib.vb_layernorm_fwd(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(),
                              _p(y_bf16), y_bf16.stride(0) if y_bf16 is not None else 0,
                              _p(y_f32), y_f32.stride(0) if y_f32 is not None else 0,
                              _p(mean), _p(rstd), rows, D, float(eps), _p(add), add.stride(0) if add is not None else 0,
                              _p(y2_bf16), y2_bf16.stride(0) if y2_bf16 is not None else 0, _stream())
    _lib.check(rc, "vb_layernorm_fwd")


def layernorm_bwd(dy, x, mean, rstd, gamma, *, dres=None, dx=None, dx_bf16=None, dgamma=None, dbeta=None, dx_colsum=None,
                  dy_add=None):
    lib = _lib.load()
    rows, D = x.shape
    rc = lib.vb_layernorm_bwd(dy.data_ptr(), _dt(dy), dy.stride(0), x.data_ptr(), x.stride(0), mean.data_ptr(), rstd.data_ptr(),
                              gamma.data_ptr(), _p(dres), dres.stride(0) if dres is not None else 0,
                              _p(dx), dx.stride(0) if dx is not None else 0,
                              _p(dx_bf16), dx_bf16.stride(0) if dx_bf16 is not None else 0,
                              _p(dgamma), _p(dbeta), _p(dx_colsum), rows, D, _p(dy_add),
                              dy_add.stride(0) if dy_add is not None else 0, _stream())
    _lib.check(rc, "vb_layernorm_bwd")


def _attn_desc(q, k, v, o, lse, B, H, S, tok_stride, batch_stride, kpm, S_kv=0):
    d = _lib.VbAttnDesc()
    d.B, d.H, d.S, d.head_dim = B, H, S, 64
    d.S_kv = S_kv if S_kv != S else 0
    d.tok_stride, d.batch_stride = tok_stride, batch_stride
    d.q, d.k, d.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    d.ldq, d.ldk, d.ldv = q.stride(0), k.stride(0), v.stride(0)
    d.o, d.ldo = o.data_ptr(), o.stride(0)
    d.lse = _p(lse)
    d.key_padding_mask = _p(kpm)
    return d


_ATTN_WS = {}   # (device, backward) -> scratch tensor, grown on demand (VbAttnDesc.workspace: caller-allocated, never retained by the library)


def _attn_workspace(lib, d, device, backward):
    need = lib.vb_attention_workspace_bytes(ctypes.byref(d), 1 if backward else 0)
    if need <= 0:
        return
    key = (device, backward)
    ws = _ATTN_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _ATTN_WS[key] = torch.empty(int(need * 1.25) + 256, device=device, dtype=torch.uint8)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()


def _attn_dropout(d, dropout):
    """dropout: None or (p, seed_dev int32/uint32 tensor, stream_id)."""
    if dropout is not None and dropout[0] > 0:
        d.dropout_p, d.dropout_seed, d.dropout_stream = float(dropout[0]), dropout[1].data_ptr(), int(dropout[2])


def attention_fwd(q, k, v, o, lse, *, B, H, S, tok_stride, batch_stride, key_padding_mask=None, dropout=None, S_kv=0):
    """q/k/v/o: bf16 2-D views [tokens, H*64] (any row pitch); lse: fp32 [B,H,S] or None.  S_kv: number of keys when != S (cross-attention)."""
    lib = _lib.load()
    d = _attn_desc(q, k, v, o, lse, B, H, S, tok_stride, batch_stride, key_padding_mask, S_kv)
    _attn_dropout(d, dropout)
    _attn_workspace(lib, d, q.device, False)
    _lib.check(lib.vb_attention_fwd(ctypes.byref(d), _stream()), "vb_attention_fwd")


def attention_bwd(q, k, v, o, lse, dout, dq, dk, dv, delta, *, B, H, S, tok_stride, batch_stride, key_padding_mask=None, dropout=None,
                  dqkv_colsum=None, S_kv=0):
    """dqkv_colsum: optional fp32 [3*H*64] accumulator of the column sums of dq|dk|dv (bias gradient of the packed in-projection)."""
    lib = _lib.load()
    d = _attn_desc(q, k, v, o, lse, B, H, S, tok_stride, batch_stride, key_padding_mask, S_kv)
    _attn_dropout(d, dropout)
    if dqkv_colsum is not None:
        assert dqkv_colsum.dtype == torch.float32 and dqkv_colsum.is_contiguous() and dqkv_colsum.numel() == 3 * H * 64
        d.dqkv_colsum = dqkv_colsum.data_ptr()
    d.dout, d.lddo = dout.data_ptr(), dout.stride(0)
    d.delta = delta.data_ptr()
    d.dq, d.dk, d.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    d.lddq, d.lddk, d.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
    _attn_workspace(lib, d, q.device, True)
    _lib.check(lib.vb_attention_bwd(ctypes.byref(d), _stream()), "vb_attention_bwd")


def dropout_f32(src, p, seed_dev, stream_id, *, aux=None, dst=None, dst_bf16=None):
    """dst / dst_bf16 = keep ? src / (1 - p) : 0 (+ aux); src, aux, dst fp32 [rows, cols] views, dst_bf16 bf16."""
    lib = _lib.load()
    rows, cols = src.shape
    assert src.dtype == torch.float32 and src.stride(1) == 1
    _lib.check(lib.vb_dropout_f32(src.data_ptr(), src.stride(0), _p(aux), aux.stride(0) if aux is not None else 0, _p(dst),
                                  dst.stride(0) if dst is not None else 0, _p(dst_bf16), dst_bf16.stride(0) if dst_bf16 is not None else 0,
                                  rows, cols, float(p), seed_dev.data_ptr(), int(stream_id), _stream()), "vb_dropout_f32")


def dropout_bf16_pair(x1, x2, p, seed_dev, stream_id):
    lib = _lib.load()
    rows, cols = x1.shape
    assert x1.dtype == torch.bfloat16 and x1.stride(1) == 1 and (x2 is None or (x2.shape == x1.shape and x2.stride(0) == x1.stride(0)))
    _lib.check(lib.vb_dropout_bf16_pair(x1.data_ptr(), _p(x2), x1.stride(0), rows, cols, float(p), seed_dev.data_ptr(), int(stream_id),
                                        _stream()), "vb_dropout_bf16_pair")


def dropout_mask(n, p, seed_dev, stream_id, device):
    """The keep mask (uint8, n elements) of dropout site `stream_id` (lets tests replay the kernels' masks)."""
    lib = _lib.load()
    out = torch.empty(n, device=device, dtype=torch.uint8)
    _lib.check(lib.vb_dropout_mask_u8(out.data_ptr(), n, float(p), seed_dev.data_ptr(), int(stream_id), _stream()), "vb_dropout_mask_u8")
    return out


def cast_bf16(src, dst):
    lib = _lib.load()
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    _lib.check(lib.vb_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "vb_cast_f32_to_bf16")


def cast_rows_bf16(src, dst):
    """dst (bf16 [rows, cols] view, any even pitch) = bf16(src) (fp32 [rows, cols] view)."""
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.shape == dst.shape and src.stride(1) == 1 and dst.stride(1) == 1
    _lib.check(_lib.load().vb_cast_rows_bf16(src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), src.shape[0], src.shape[1], _stream()),
               "vb_cast_rows_bf16")


def pos_embed_2d_fwd(row_embed, col_embed, pos, h, w, N):
    """pos: fp32 [h*w, N, 2*pf] contiguous."""
    pf = row_embed.shape[1]
    assert pos.is_contiguous() and pos.shape == (h * w, N, 2 * pf) and row_embed.is_contiguous() and col_embed.is_contiguous()
    _lib.check(_lib.load().vb_pos_embed_2d_fwd(row_embed.data_ptr(), col_embed.data_ptr(), pos.data_ptr(), h, w, N, pf, _stream()), "vb_pos_embed_2d_fwd")


def pos_embed_2d_bwd(dpos, drow, dcol, h, w, N):
    pf = drow.shape[1]
    assert dpos.is_contiguous() and dpos.dtype == torch.float32 and dpos.shape == (h * w, N, 2 * pf)
    _lib.check(_lib.load().vb_pos_embed_2d_bwd(dpos.data_ptr(), drow.data_ptr(), dcol.data_ptr(), h, w, N, pf, _stream()), "vb_pos_embed_2d_bwd")


def patchify(images, out, patch):
    lib = _lib.load()
    B, C, H, W = images.shape
    assert images.dtype == torch.float32 and images.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
    _lib.check(lib.vb_patchify(images.data_ptr(), out.data_ptr(), B, C, H, W, patch, _stream()), "vb_patchify")


def unpatchify(dpatches, dimages, patch):
    """dimages [B,3,H,W] fp32 <- dpatches [B, P, 3*p*p] bf16 (contiguous): the inverse of patchify."""
    lib = _lib.load()
    B, C, H, W = dimages.shape
    assert dimages.dtype == torch.float32 and dimages.is_contiguous() and dpatches.dtype == torch.bfloat16 and dpatches.is_contiguous()
    _lib.check(lib.vb_unpatchify(dpatches.data_ptr(), dimages.data_ptr(), B, C, H, W, patch, _stream()), "vb_unpatchify")


def token_rows(x, tok0, tok1, pos, n_prefix):
    lib = _lib.load()
    B, S, D = x.shape
    assert x.is_contiguous() and (pos is None or pos.is_contiguous())
    _lib.check(lib.vb_token_rows(x.data_ptr(), tok0.data_ptr(), _p(tok1), _p(pos), B, S, D, n_prefix, _stream()), "vb_token_rows")


def colsum_bf16(x, out_accum):
    lib = _lib.load()
    rows, cols = x.shape
    assert x.dtype == torch.bfloat16 and x.stride(1) == 1 and out_accum.dtype == torch.float32
    _lib.check(lib.vb_colsum_bf16(x.data_ptr(), x.stride(0), rows, cols, out_accum.data_ptr(), _stream()), "vb_colsum_bf16")


def vecmat_accum(x, W, y_accum, x_accum=None):
    """y_accum += x @ W (fp32 vector [K] times fp32 matrix [K, N] view); x_accum += x (optional)."""
    K, N = W.shape
    assert x.dtype == torch.float32 and W.dtype == torch.float32 and W.stride(1) == 1 and x.numel() == K and y_accum.numel() == N
    _lib.check(_lib.load().vb_vecmat_accum(x.data_ptr(), W.data_ptr(), W.stride(0), y_accum.data_ptr(), _p(x_accum), K, N, _stream()),
               "vb_vecmat_accum")


def embed_bwd(dx, possum, dx_patches, dpos, dtok0, dtok1, dbias, n_prefix):
    lib = _lib.load()
    B, S, D = dx.shape
    assert dx.is_contiguous() and dx.dtype == torch.float32
    _lib.check(lib.vb_embed_bwd(dx.data_ptr(), possum.data_ptr(), _p(dx_patches), _p(dpos), _p(dtok0), _p(dtok1), _p(dbias),
                                B, S, D, n_prefix, _stream()), "vb_embed_bwd")


def dwconv_fwd(x, w, bias, out, *, n_prefix, pos=None, sub=None):
    """out = [prefix rows of x | depthwise 3x3 conv over the patch-token grid + bias] (+ pos broadcast) (+ x - sub); fp32 [B, S, D]."""
    lib = _lib.load()
    B, S, D = x.shape
    G = int(round((S - n_prefix) ** 0.5))
    for t in (x, out, sub):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.shape == x.shape)
    assert w.is_contiguous() and w.numel() == D * 9 and (pos is None or (pos.is_contiguous() and pos.numel() == S * D))
    _lib.check(lib.vb_dwconv3x3_fwd(x.data_ptr(), w.data_ptr(), _p(bias), _p(pos), _p(sub), out.data_ptr(), B, S, D, n_prefix, G, _stream()),
               "vb_dwconv3x3_fwd")


def dwconv_bwd_data(dy, w, *, n_prefix, dx=None, sum_f32=None, sum_bf16=None):
    """dx = [dy | conv^T(dy)]; sum_* = dy + dx."""
    lib = _lib.load()
    B, S, D = dy.shape
    G = int(round((S - n_prefix) ** 0.5))
    assert dy.dtype == torch.float32 and dy.is_contiguous()
    for t in (dx, sum_f32, sum_bf16):
        assert t is None or (t.is_contiguous() and t.numel() == dy.numel())
    _lib.check(lib.vb_dwconv3x3_bwd_data(dy.data_ptr(), w.data_ptr(), _p(dx), _p(sum_f32), _p(sum_bf16), B, S, D, n_prefix, G, _stream()),
               "vb_dwconv3x3_bwd_data")


def dwconv_bwd_weight(dy, x, dw, db, *, n_prefix):
    lib = _lib.load()
    B, S, D = dy.shape
    G = int(round((S - n_prefix) ** 0.5))
    assert dy.dtype == torch.float32 and dy.is_contiguous() and x.dtype == torch.float32 and x.is_contiguous() and x.shape == dy.shape
    assert dw.dtype == torch.float32 and dw.numel() == D * 9
    _lib.check(lib.vb_dwconv3x3_bwd_weight(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), _p(db), B, S, D, n_prefix, G, _stream()),
               "vb_dwconv3x3_bwd_weight")


def cross_entropy(logits, labels, loss_accum, *, weight, dlogits_bf16=None, dlogits_f32=None, grad_scale=1.0, correct_accum=None):
    """logits: fp32 [B, C] view; labels int64 [B]; loss_accum: fp32 scalar tensor that is ADDED to."""
    lib = _lib.load()
    B, C = logits.shape
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and labels.dtype == torch.int64 and labels.is_contiguous()
    rc = lib.vb_cross_entropy(logits.data_ptr(), logits.stride(0), labels.data_ptr(), B, C, loss_accum.data_ptr(), float(weight),
                              _p(dlogits_bf16), dlogits_bf16.stride(0) if dlogits_bf16 is not None else 0,
                              _p(dlogits_f32), dlogits_f32.stride(0) if dlogits_f32 is not None else 0,
                              float(grad_scale), _p(correct_accum), _stream())
    _lib.check(rc, "vb_cross_entropy")


def distill_loss(logits, logits_kd, teacher_logits, labels, loss_accum, *, kind, alpha, tau, dlogits_bf16=None, dlogits_kd_bf16=None,
                 dlogits_f32=None, dlogits_kd_f32=None, grad_scale=1.0, correct_accum=None):
    """DistillationLoss.forward (utils/distillation_loss.py:30-75) + both logits gradients; kind: 'soft' | 'hard'."""
    lib = _lib.load()
    B, C = logits.shape
    for t in (logits, logits_kd, teacher_logits):
        assert t.dtype == torch.float32 and t.stride(1) == 1 and t.shape == (B, C)
    assert labels.dtype == torch.int64 and labels.is_contiguous()
    st = lambda t: t.stride(0) if t is not None else 0
    rc = lib.vb_distill_loss(logits.data_ptr(), logits.stride(0), logits_kd.data_ptr(), logits_kd.stride(0), teacher_logits.data_ptr(),
                             teacher_logits.stride(0), labels.data_ptr(), B, C, {"soft": 1, "hard": 2}[kind], float(alpha), float(tau),
                             loss_accum.data_ptr(), _p(dlogits_bf16), st(dlogits_bf16), _p(dlogits_kd_bf16), st(dlogits_kd_bf16),
                             _p(dlogits_f32), st(dlogits_f32), _p(dlogits_kd_f32), st(dlogits_kd_f32), float(grad_scale),
                             _p(correct_accum), _stream())
    _lib.check(rc, "vb_distill_loss")


def adam_step(params, grads, exp_avg, exp_avg_sq, params_bf16, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0,
              step_counter=None):
    """step_counter: optional int32 device tensor holding the step number (incremented by the call; CUDA-graph friendly)."""
    lib = _lib.load()
    n = params.numel()
    rc = lib.vb_adam_step(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), _p(params_bf16), n,
                          float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale),
                          _p(step_counter), _stream())
    _lib.check(rc, "vb_adam_step")


def add_cast_bf16(a, b, out):
    """out_bf16 = bf16(a + b); a, b fp32 contiguous (b may be None)."""
    lib = _lib.load()
    assert a.is_contiguous() and out.is_contiguous() and (b is None or b.is_contiguous())
    _lib.check(lib.vb_add_cast_bf16(a.data_ptr(), _p(b), out.data_ptr(), a.numel(), _stream()), "vb_add_cast_bf16")


def add_rows_bcast(x, pos, out):
    """out[r] = x[r] + pos[r % period]; x/out fp32 contiguous [rows, D], pos fp32 contiguous [period, D]."""
    lib = _lib.load()
    rows, D = x.shape
    assert x.is_contiguous() and pos.is_contiguous() and out.is_contiguous() and x.dtype == torch.float32
    _lib.check(lib.vb_add_rows_bcast(x.data_ptr(), pos.data_ptr(), out.data_ptr(), rows, pos.shape[0], D, _stream()), "vb_add_rows_bcast")


def add3(a, b_bf16, c_bf16, out, accum=None):
    """out = a + b (+ c); accum += b (optional). a/out/accum fp32, b/c bf16, all contiguous."""
    lib = _lib.load()
    _lib.check(lib.vb_add3(a.data_ptr(), b_bf16.data_ptr(), _p(c_bf16), out.data_ptr(), _p(accum), a.numel(), _stream()), "vb_add3")
