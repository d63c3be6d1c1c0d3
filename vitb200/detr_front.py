"""What sits between the DETR backbone and the transformer encoder in the reference (SURVEY.md §8 f3 remainder):

* ``InputProjection``  — ``Detr.input_proj = nn.Conv2d(backbone.in_channels, hidden_dim, kernel_size=1)`` (detr.py:125).  A 1x1
  convolution over an NCHW feature map is a GEMM whose A operand is stored [C_in, H*W] per image (the contraction index C_in is NOT
  contiguous): it runs on the tcgen05 GEMM with ``a_major = 1`` straight from the NCHW layout — no NHWC copy — and writes the
  result directly in the encoder's sequence-first layout [H*W, N, hidden] (the ``src.flatten(2).permute(2, 0, 1)`` of
  transformer.py:49-50 is fused into the store geometry).  The module still *returns* [N, hidden, H, W] — as a zero-copy permuted view
  of that buffer — so the reference's ``Transformer.forward`` view arithmetic lands back on the contiguous tensor.
* ``AbsolutePositionalEncoding`` — the learned row / column embedding of detr.py:33-63, produced by one kernel in the same layout.
* ``NestedTensor`` / ``nested_tensor_from_tensor_list`` — utils/coco/util/misc.py:284-332 (batching images of different sizes with a
  padding mask) and ``mask_at`` for the mask at feature-map resolution (nearest-neighbour, as DETR's backbone does).
* ``Transformer`` — transformer.py:25-63: the constructor is the reference's; its ``forward`` is the one the code spells out with the
  three typos resolved (``memory.permte`` :63, ``hs.transpose(1, 1)`` :63, and the decoder layer's attribute name, see detr.py here).
"""
from typing import List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .detr import TransformerDecoder, TransformerDecoderLayer, TransformerEncoder, TransformerEncoderLayer


class NestedTensor(object):
    """utils/coco/util/misc.py:284-304."""

    def __init__(self, tensors, mask: Optional[torch.Tensor]):
        self.tensors = tensors
        self.mask = mask

    def to(self, device):
        return NestedTensor(self.tensors.to(device), self.mask.to(device) if self.mask is not None else None)

    def decompose(self):
        return self.tensors, self.mask

    def __repr__(self):
        return str(self.tensors)


def nested_tensor_from_tensor_list(tensor_list: List[torch.Tensor]):
    """utils/coco/util/misc.py:307-332: zero-pad [C, H_i, W_i] images to the largest H / W; mask is True on padding."""
    if tensor_list[0].ndim != 3:
        raise ValueError("not supported")
    max_size = [max(img.shape[d] for img in tensor_list) for d in range(3)]
    b, (c, h, w) = len(tensor_list), max_size
    tensor = torch.zeros((b, c, h, w), dtype=tensor_list[0].dtype, device=tensor_list[0].device)
    mask = torch.ones((b, h, w), dtype=torch.bool, device=tensor_list[0].device)
    for img, pad_img, m in zip(tensor_list, tensor, mask):
        pad_img[: img.shape[0], : img.shape[1], : img.shape[2]].copy_(img)
        m[: img.shape[1], : img.shape[2]] = False
    return NestedTensor(tensor, mask)


def mask_at(mask, size):
    """The padding mask at feature-map resolution (nearest neighbour), as the DETR backbone derives it for its NestedTensor outputs."""
    return F.interpolate(mask[None].float(), size=tuple(size)).to(torch.bool)[0]


class _InputProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        N, Cin, h, w = x.shape
        hidden, HW = weight.shape[0], h * w
        dev = x.device
        hw_pad = (HW + 7) // 8 * 8                       # TMA row pitches are multiples of 16 bytes
        xb = torch.empty(N, Cin, hw_pad, device=dev, dtype=torch.bfloat16)
        if hw_pad != HW:
            xb[:, :, HW:].zero_()
        ops.cast_rows_bf16(x.reshape(N * Cin, HW), xb.view(N * Cin, hw_pad)[:, :HW])
        wb = torch.empty(hidden, Cin, device=dev, dtype=torch.bfloat16)
        ops.cast_bf16(weight.reshape(-1), wb.view(-1))
        out = torch.empty(HW, N, hidden, device=dev, dtype=torch.float32)
        # per image: out[:, n, :] [HW, hidden] = x[n]^T [HW, C_in] . W^T  — A stored [C_in, HW] (a_major = 1), rows of C interleaved over n
        ops.gemm(xb[:, :, :HW], wb, out.permute(1, 0, 2), a_major=1, b_major=0, bias=bias)
        ctx.save_for_backward(xb, wb)
        ctx.shape = (N, Cin, h, w, hidden, hw_pad)
        return out.view(h, w, N, hidden).permute(2, 3, 0, 1)       # [N, hidden, h, w], a view of the sequence-first buffer

    @staticmethod
    def backward(ctx, grad_out):
        xb, wb = ctx.saved_tensors
        N, Cin, h, w, hidden, hw_pad = ctx.shape
        HW, dev = h * w, grad_out.device
        g = grad_out.permute(2, 3, 0, 1)                            # [h, w, N, hidden]: contiguous when it comes from the encoder
        if not g.is_contiguous() or g.dtype != torch.float32:
            g = g.contiguous().float()
        gb = torch.empty(HW, N, hidden, device=dev, dtype=torch.bfloat16)
        ops.cast_bf16(g.view(-1), gb.view(-1))
        dW = db = dx = None
        if ctx.needs_input_grad[1]:
            dW = torch.zeros(hidden, Cin, device=dev, dtype=torch.float32)
            for n in range(N):     # dW += dY_n^T X_n: A = dY_n stored [tokens, hidden] (a_major 1), B = X_n stored [C_in, tokens] (b_major 0)
                ops.gemm(gb[:, n, :], xb[n, :, :HW], dW, a_major=1, b_major=0, epilogue=ops.EPI_ACCUM)
            dW = dW.view(hidden, Cin, 1, 1)
        if ctx.needs_input_grad[2]:
            db = torch.zeros(hidden, device=dev, dtype=torch.float32)
            ops.colsum_bf16(gb.view(HW * N, hidden), db)
        if ctx.needs_input_grad[0]:
            p4 = (HW + 3) // 4 * 4
            dxp = torch.empty(N, Cin, p4, device=dev, dtype=torch.float32)
            for n in range(N):     # dX_n [C_in, HW] = W^T dY_n^T: A = W stored [hidden, C_in] (a_major 1), B = dY_n [HW, hidden] (b_major 0)
                ops.gemm(wb, gb[:, n, :], dxp[n, :, :HW], a_major=1, b_major=0)
            dx = dxp[:, :, :HW].reshape(N, Cin, h, w)
        return dx, dW, db


class InputProjection(nn.Conv2d):
    """``nn.Conv2d(in_channels, hidden_dim, kernel_size=1)`` (detr.py:125) — same parameters, same state_dict keys."""

    def __init__(self, in_channels, out_channels, kernel_size=1, **kw):
        if kernel_size not in (1, (1, 1)) or kw:
            raise NotImplementedError("vitb200.InputProjection is the 1x1 projection of detr.py:125")
        super().__init__(in_channels, out_channels, kernel_size=1)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("vitb200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        if self.in_channels % 8 != 0 or self.out_channels % 8 != 0:
            raise NotImplementedError("vitb200.InputProjection needs channel counts that are multiples of 8")
        x = x.contiguous().float()
        return _InputProjFn.apply(x, self.weight, self.bias)


class _PosEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, row_w, col_w, h, w, N):
        pf = row_w.shape[1]
        pos = torch.empty(h * w, N, 2 * pf, device=row_w.device, dtype=torch.float32)
        ops.pos_embed_2d_fwd(row_w.contiguous(), col_w.contiguous(), pos, h, w, N)
        ctx.geom = (h, w, N, row_w.shape, col_w.shape)
        return pos.view(h, w, N, 2 * pf).permute(2, 3, 0, 1)       # [N, 2 pf, h, w]

    @staticmethod
    def backward(ctx, grad):
        h, w, N, rs, cs = ctx.geom
        g = grad.permute(2, 3, 0, 1)
        if not g.is_contiguous() or g.dtype != torch.float32:
            g = g.contiguous().float()
        drow = torch.zeros(rs, device=grad.device, dtype=torch.float32)
        dcol = torch.zeros(cs, device=grad.device, dtype=torch.float32)
        ops.pos_embed_2d_bwd(g.view(h * w, N, -1), drow, dcol, h, w, N)
        return drow, dcol, None, None, None


class AbsolutePositionalEncoding(nn.Module):
    """detr.py:33-63: learned row / column embeddings (50 positions each), concatenated per pixel."""

    def __init__(self, positional_features=256):
        super().__init__()
        self.row_embed = nn.Embedding(50, positional_features)
        self.col_embed = nn.Embedding(50, positional_features)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.uniform_(self.row_embed.weight)
        nn.init.uniform_(self.col_embed.weight)

    def forward(self, x):
        t = x.tensors if hasattr(x, "tensors") else x
        h, w = t.shape[-2:]
        if h > self.row_embed.num_embeddings or w > self.col_embed.num_embeddings:
            raise IndexError("feature map larger than the embedding tables (50 x 50: detr.py:44-45)")
        if not self.row_embed.weight.is_cuda:
            raise RuntimeError("vitb200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        return _PosEmbedFn.apply(self.row_embed.weight, self.col_embed.weight, h, w, t.shape[0])


class Transformer(nn.Module):
    """transformer.py:25-63."""

    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048, dropout=0.1,
                 activation="relu", normalize_before=False, return_intermediate_dec=False):
        super().__init__()
        encoder_layer = TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        encoder_norm = nn.LayerNorm(d_model) if normalize_before else None
        self.encoder = TransformerEncoder(encoder_layer, num_encoder_layers, encoder_norm)
        decoder_layer = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        decoder_norm = nn.LayerNorm(d_model)
        self.decoder = TransformerDecoder(decoder_layer, num_decoder_layers, decoder_norm, return_intermediate=return_intermediate_dec)
        self._reset_parameters()
        self.d_model = d_model
        self.nhead = nhead

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, src, mask, query_embed, pos_embed):
        bs, c, h, w = src.shape
        src = src.flatten(2).permute(2, 0, 1)                     # :49-50 (a view of the sequence-first buffer after InputProjection)
        pos_embed = pos_embed.flatten(2).permute(2, 0, 1)         # :51
        query_embed = query_embed.unsqueeze(1).repeat(1, bs, 1)   # :52
        mask = mask.flatten(1)                                    # :53
        tgt = torch.zeros_like(query_embed)
        memory = self.encoder(src, src_key_padding_mask=mask, pos=pos_embed)
        hs = self.decoder(tgt, memory, memory_key_padding_mask=mask, pos=pos_embed, query_pos=query_embed)
        return hs.transpose(1, 2), memory.permute(1, 2, 0).view(bs, c, h, w)      # :63 with its two typos resolved
