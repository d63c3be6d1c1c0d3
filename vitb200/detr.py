"""Drop-in DETR transformer encoder (models/object_detection/transformer.py: TransformerEncoderLayer :192-247,
TransformerEncoder :98-115) on the hand-written sm_100a kernels.

Same constructors, ``forward(src, src_mask=None, src_key_padding_mask=None, pos=None)`` signature, sequence-first
[S, N, C] tensors and state_dict keys (``layers.{i}.self_attn.in_proj_weight`` ...).  The post-norm path
(``normalize_before=False``, the reference default: transformer.py:27-28) and the pre-norm variant (``forward_pre``,
transformer.py:228-241) are implemented; ``q = k = src + pos``,
``v = src`` (transformer.py:218-219), boolean key-padding mask, ReLU (or GELU) feed-forward, LayerNorm eps 1e-5.
The layer objects are parameter containers: the stack is executed by the enclosing ``TransformerEncoder``.
"""
import copy

import torch
from torch import nn

from . import ops
from .engine import FlatParams, WorkspaceLease, dropout_start_seed, getstate_without_engine
from .vit import _EncoderFn  # noqa: F401  (same single-node autograd pattern)

DETR_ROLES = ("norm2_w", "norm2_b", "lin2_w", "lin2_b", "lin1_w", "lin1_b", "norm1_w", "norm1_b", "out_w", "out_b", "in_w", "in_b")


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, dropout, activation, normalize_before):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        if activation not in ("relu", "gelu"):
            raise RuntimeError(F"activation should be relu/gelu, not {activation}.")
        self.activation_name = activation
        self.normalize_before = normalize_before
        self.d_model, self.nhead, self.dim_feedforward, self.dropout_p = d_model, nhead, dim_feedforward, dropout

    def roles(self):
        return {"norm2_w": self.norm2.weight, "norm2_b": self.norm2.bias, "lin2_w": self.linear2.weight, "lin2_b": self.linear2.bias,
                "lin1_w": self.linear1.weight, "lin1_b": self.linear1.bias, "norm1_w": self.norm1.weight, "norm1_b": self.norm1.bias,
                "out_w": self.self_attn.out_proj.weight, "out_b": self.self_attn.out_proj.bias,
                "in_w": self.self_attn.in_proj_weight, "in_b": self.self_attn.in_proj_bias}

    def forward(self, src, src_mask=None, src_key_padding_mask=None, pos=None):
        # a single layer is a one-layer stack
        enc = self.__dict__.get("_solo")
        if enc is None:
            enc = TransformerEncoder.__new__(TransformerEncoder)
            nn.Module.__init__(enc)
            enc.layers = nn.ModuleList([self])
            enc.num_layers, enc.norm = 1, None
            enc.__dict__["_engine"] = None
            self.__dict__["_solo"] = enc
        return enc(src, mask=src_mask, src_key_padding_mask=src_key_padding_mask, pos=pos)

    __getstate__ = getstate_without_engine

    def __deepcopy__(self, memo):
        solo = self.__dict__.pop("_solo", None)
        try:
            new = self.__class__.__new__(self.__class__)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                new.__dict__[k] = copy.deepcopy(v, memo)
        finally:
            if solo is not None:
                self.__dict__["_solo"] = solo
        return new


class DetrEngine(FlatParams):
    def __init__(self, layers, norm, d_model, nhead, dim_feedforward, activation, eps=1e-5, pre_norm=False):
        if d_model % 64 != 0 or d_model // nhead != 64 or d_model > 1024:
            raise NotImplementedError(f"vitb200 kernels need head_dim == 64 and d_model <= 1024 (got d_model {d_model}, {nhead} heads)")
        self.D, self.H, self.F, self.L, self.eps = d_model, nhead, dim_feedforward, len(layers), eps
        self.act = activation
        self.pre_norm = bool(pre_norm)   # TransformerEncoderLayer.forward_pre (transformer.py:228-241) instead of forward_post (:213-226)
        self.has_norm = norm is not None
        self.p_drop = 0.0            # the layer's single rate (transformer.py:195,198,204-205); set per call by TransformerEncoder.forward
        self._drop_counter = None
        self._order = []
        seg = []
        if self.has_norm:
            self._order += [(("g", "norm_w"), norm.weight), (("g", "norm_b"), norm.bias)]
            seg.append(2)
        for li in range(self.L - 1, -1, -1):
            r = layers[li].roles()
            self._order += [((li, k), r[k]) for k in DETR_ROLES]
            seg.append(len(DETR_ROLES))
        self._layout(seg)

    @staticmethod
    def drop_site(layer, site):
        """stream id of a dropout site: 0 = dropout1 (attention output, transformer.py:220/236), 1 = dropout (after the activation,
        :223/239), 2 = dropout2 (FFN output, :224/240), 3 = attention weights (nn.MultiheadAttention(dropout=), :195)."""
        return layer * 8 + site

    def _begin_dropout(self, ws):
        """New masks for this forward: bump the device-side counter and snapshot it into the workspace (graph-capturable)."""
        dev = self.flat.device
        if self._drop_counter is None or self._drop_counter.device != dev:
            self._drop_counter = torch.full((1,), dropout_start_seed(), device=dev, dtype=torch.int32)
        self._drop_counter.add_(1)
        if ws.get("drop_seed") is None:
            ws["drop_seed"] = torch.zeros(1, device=dev, dtype=torch.int32)
            ws["tmp32"] = torch.empty(ws["M"], self.D, device=dev, dtype=torch.float32)
        ws["drop_seed"].copy_(self._drop_counter)

    def _ffn_fwd(self, ws, buf, li, pd):
        """act(linear1(x1_bf)) (+ dropout, transformer.py:223/239) -> the bf16 A operand of linear2."""
        if self.act == "relu":
            ops.gemm(buf["x1_bf"], self.w((li, "lin1_w")), buf["a"], epilogue=ops.EPI_RELU, bias=self.f((li, "lin1_b")))
            if pd > 0:   # a = keep * relu(x) / (1 - p): its sign pattern is the backward mask (VB_EPI_DRELU with drelu_scale)
                ops.dropout_bf16_pair(buf["a"], None, pd, ws["drop_seed"], self.drop_site(li, 1))
            return buf["a"]
        if "g" not in buf:
            buf["g"] = torch.empty_like(buf["a"])
        ops.gemm(buf["x1_bf"], self.w((li, "lin1_w")), buf["a"], C2=buf["g"], epilogue=ops.EPI_GELU, bias=self.f((li, "lin1_b")))
        if pd > 0:       # the same mask scales gelu(x) and the saved gelu'(x)
            ops.dropout_bf16_pair(buf["g"], buf["a"], pd, ws["drop_seed"], self.drop_site(li, 1))
        return buf["g"]

    def _linear_residual(self, ws, A, wkey, bkey, li, site, residual, out, pd):
        """out = residual + dropout(A W^T + b): one GEMM with the residual epilogue for p = 0, else an fp32 GEMM output followed by the
        fused keep * x / (1 - p) + residual kernel."""
        if pd > 0:
            ops.gemm(A, self.w((li, wkey)), ws["tmp32"], bias=self.f((li, bkey)))
            ops.dropout_f32(ws["tmp32"], pd, ws["drop_seed"], self.drop_site(li, site), aux=residual, dst=out)
        else:
            ops.gemm(A, self.w((li, wkey)), out, epilogue=ops.EPI_RESIDUAL, bias=self.f((li, bkey)), aux=residual)

    def _masked_operand(self, ws, d32, d_bf, li, site, bias_key):
        """backward of a dropped linear output: the bf16 GEMM operand is keep * d / (1 - p) (mask regenerated from the forward's seed) and
        its column sum is that layer's bias gradient."""
        ops.dropout_f32(d32, ws["p_drop"], ws["drop_seed"], self.drop_site(li, site), dst_bf16=d_bf)
        ops.colsum_bf16(d_bf, self.gview(bias_key))

    def workspace(self, S, N, training):
        key = (S, N, training)
        ws = self._free_workspace(key)
        if ws is not None:
            return ws
        dev, D, Fd, M = self.flat.device, self.D, self.F, S * N
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *shape, dtype=bf: torch.empty(*shape, device=dev, dtype=dtype)
        ws = {"S": S, "N": N, "M": M, "layer": []}
        for _ in range(self.L if training else 1):
            ws["layer"].append({"x_bf": e(M, D), "qk_bf": e(M, D), "qkb": e(M, 2 * D), "vb": e(M, D), "o": e(M, D),
                                "lse": e(N, self.H, S, dtype=f32), "x1pre": e(M, D, dtype=f32), "x1": e(M, D, dtype=f32),
                                "x1_bf": e(M, D), "a": e(M, Fd), "x2pre": e(M, D, dtype=f32), "mean1": e(M, dtype=f32),
                                "rstd1": e(M, dtype=f32), "mean2": e(M, dtype=f32), "rstd2": e(M, dtype=f32)})
        ws["x"] = [e(M, D, dtype=f32) for _ in range(2)]
        ws["meanf"], ws["rstdf"] = e(M, dtype=f32), e(M, dtype=f32)
        ws["y"] = e(M, D, dtype=f32)
        if training:
            ws["dA"], ws["dB"] = e(M, D, dtype=f32), e(M, D, dtype=f32)
            ws["dA_bf"], ws["dB_bf"] = e(M, D), e(M, D)
            ws["dh"], ws["dh2"] = e(M, D), e(M, D)
            ws["da"] = e(M, Fd)
            ws["dqk"], ws["dv"] = e(M, 2 * D), e(M, D)
            ws["delta"] = e(N, self.H, S, dtype=f32)
            ws["dpos"] = e(M, D, dtype=f32)
        self._ws[key].append(ws)
        return ws

    def forward(self, src, pos, kpm, training):
        self.ensure_bound()
        for t, name in ((src, "src"), (pos, "pos"), (kpm, "src_key_padding_mask")):
            self.check_input(t, name)
        S, N, D = src.shape
        if D != self.D or (pos is not None and pos.shape != src.shape) or (kpm is not None and tuple(kpm.shape) != (N, S)):
            raise RuntimeError(f"vitb200: expected src [S, N, {self.D}], pos of the same shape and src_key_padding_mask [N, S]; got "
                               f"{tuple(src.shape)}, {None if pos is None else tuple(pos.shape)}, {None if kpm is None else tuple(kpm.shape)}")
        ws = self.workspace(S, N, training)
        self.refresh_bf16()
        M, H = ws["M"], self.H
        src2 = src.contiguous().float().view(M, D)
        pos2 = pos.contiguous().float().view(M, D) if pos is not None else None
        ws["pos"], ws["kpm"] = pos2, kpm
        pd = ws["p_drop"] = float(self.p_drop)   # 0 unless the module is in train() mode
        if pd > 0:
            self._begin_dropout(ws)
        x = src2
        if self.pre_norm:
            return self._forward_pre(ws, x, pos2, kpm, training, S, N)
        for li in range(self.L):
            buf = ws["layer"][li if training else 0]
            if li == 0:
                ops.add_cast_bf16(x, None, buf["x_bf"])
                ops.add_cast_bf16(x, pos2, buf["qk_bf"])
            in_w, in_b = self.w((li, "in_w")), self.f((li, "in_b"))
            ops.gemm(buf["qk_bf"], in_w[:2 * D], buf["qkb"], bias=in_b[:2 * D])
            ops.gemm(buf["x_bf"], in_w[2 * D:], buf["vb"], bias=in_b[2 * D:])
            ops.attention_fwd(buf["qkb"][:, :D], buf["qkb"][:, D:], buf["vb"], buf["o"], buf["lse"] if training else None, B=N, H=H, S=S,
                              tok_stride=N, batch_stride=1, key_padding_mask=kpm,
                              dropout=(pd, ws["drop_seed"], self.drop_site(li, 3)) if pd > 0 else None)
            self._linear_residual(ws, buf["o"], "out_w", "out_b", li, 0, x, buf["x1pre"], pd)
            ops.layernorm_fwd(buf["x1pre"], self.f((li, "norm1_w")), self.f((li, "norm1_b")), self.eps, y_bf16=buf["x1_bf"], y_f32=buf["x1"],
                              mean=buf["mean1"], rstd=buf["rstd1"])
            act_out = self._ffn_fwd(ws, buf, li, pd)
            self._linear_residual(ws, act_out, "lin2_w", "lin2_b", li, 2, buf["x1"], buf["x2pre"], pd)
            x_next = ws["x"][li & 1]
            last = li == self.L - 1
            nbuf = None if last else ws["layer"][(li + 1) if training else 0]
            # the next layer's operands come out of this LayerNorm directly: bf16(x) for v, bf16(x + pos) for q = k
            if last:
                ops.layernorm_fwd(buf["x2pre"], self.f((li, "norm2_w")), self.f((li, "norm2_b")), self.eps, y_f32=x_next,
                                  mean=buf["mean2"], rstd=buf["rstd2"])
            elif pos2 is not None:
                ops.layernorm_fwd(buf["x2pre"], self.f((li, "norm2_w")), self.f((li, "norm2_b")), self.eps, y_f32=x_next,
                                  y_bf16=nbuf["x_bf"], mean=buf["mean2"], rstd=buf["rstd2"], add=pos2, y2_bf16=nbuf["qk_bf"])
            else:
                ops.layernorm_fwd(buf["x2pre"], self.f((li, "norm2_w")), self.f((li, "norm2_b")), self.eps, y_f32=x_next,
                                  y_bf16=nbuf["x_bf"], mean=buf["mean2"], rstd=buf["rstd2"])
                ops.add_cast_bf16(x_next, None, nbuf["qk_bf"])
            x = x_next
        if self.has_norm:
            ops.layernorm_fwd(x, self.f(("g", "norm_w")), self.f(("g", "norm_b")), self.eps, y_f32=ws["y"], mean=ws["meanf"], rstd=ws["rstdf"])
            ws["x_last"] = x
            return ws["y"].view(S, N, D), ws
        return x.view(S, N, D), ws

    # ------------------------------------------------------------------------------------------------------------------
    # Pre-norm layers (normalize_before=True): src2 = norm1(src); q = k = src2 + pos; src += attn(q, k, src2);
    # src2 = norm2(src); src += linear2(act(linear1(src2)))   — transformer.py:228-241; final encoder norm :112-113.
    # Same kernels as the post-norm path; the residual stream stays fp32 and is never normalised in place.
    # ------------------------------------------------------------------------------------------------------------------
    def _forward_pre(self, ws, x, pos2, kpm, training, S, N):
        M, D, H = ws["M"], self.D, self.H
        pd = ws["p_drop"]
        if "xin" not in ws:
            ws["xin"] = [torch.empty(M, D, device=x.device, dtype=torch.float32) for _ in range(self.L + 1 if training else 2)]
        xs = ws["xin"]
        xs[0].copy_(x)
        for li in range(self.L):
            buf = ws["layer"][li if training else 0]
            x_in = xs[li] if training else xs[li & 1]
            x_out = xs[li + 1] if training else xs[(li + 1) & 1]
            if pos2 is not None:   # bf16(norm1(x)) feeds v, bf16(norm1(x) + pos) feeds q = k
                ops.layernorm_fwd(x_in, self.f((li, "norm1_w")), self.f((li, "norm1_b")), self.eps, y_bf16=buf["x_bf"],
                                  mean=buf["mean1"], rstd=buf["rstd1"], add=pos2, y2_bf16=buf["qk_bf"])
                qk_in = buf["qk_bf"]
            else:
                ops.layernorm_fwd(x_in, self.f((li, "norm1_w")), self.f((li, "norm1_b")), self.eps, y_bf16=buf["x_bf"],
                                  mean=buf["mean1"], rstd=buf["rstd1"])
                qk_in = buf["x_bf"]
            in_w, in_b = self.w((li, "in_w")), self.f((li, "in_b"))
            ops.gemm(qk_in, in_w[:2 * D], buf["qkb"], bias=in_b[:2 * D])
            ops.gemm(buf["x_bf"], in_w[2 * D:], buf["vb"], bias=in_b[2 * D:])
            ops.attention_fwd(buf["qkb"][:, :D], buf["qkb"][:, D:], buf["vb"], buf["o"], buf["lse"] if training else None, B=N, H=H, S=S,
                              tok_stride=N, batch_stride=1, key_padding_mask=kpm,
                              dropout=(pd, ws["drop_seed"], self.drop_site(li, 3)) if pd > 0 else None)
            self._linear_residual(ws, buf["o"], "out_w", "out_b", li, 0, x_in, buf["x1"], pd)
            ops.layernorm_fwd(buf["x1"], self.f((li, "norm2_w")), self.f((li, "norm2_b")), self.eps, y_bf16=buf["x1_bf"],
                              mean=buf["mean2"], rstd=buf["rstd2"])
            act_out = self._ffn_fwd(ws, buf, li, pd)
            self._linear_residual(ws, act_out, "lin2_w", "lin2_b", li, 2, buf["x1"], x_out, pd)
        x = xs[self.L] if training else xs[self.L & 1]
        if self.has_norm:
            ops.layernorm_fwd(x, self.f(("g", "norm_w")), self.f(("g", "norm_b")), self.eps, y_f32=ws["y"], mean=ws["meanf"], rstd=ws["rstdf"])
            ws["x_last"] = x
            return ws["y"].view(S, N, D), ws
        return x.view(S, N, D), ws

    def _backward_pre(self, ws, grad_out):
        S, N, M, D, H, L = ws["S"], ws["N"], ws["M"], self.D, self.H, self.L
        d, d_bf, dh, dh2 = ws["dA"], ws["dA_bf"], ws["dh"], ws["dh2"]   # d: fp32 gradient of the residual stream, updated in place
        g = grad_out.contiguous().float().view(M, D)
        pos2, kpm = ws["pos"], ws["kpm"]
        dpos = None
        if pos2 is not None:
            dpos = ws["dpos"]
            dpos.zero_()
        seg = 0
        pd = ws["p_drop"]
        drop3 = lambda li: (pd, ws["drop_seed"], self.drop_site(li, 3)) if pd > 0 else None
        drelu = dict(drelu_scale=1.0 / (1.0 - pd)) if (pd > 0 and self.act == "relu") else {}
        last_b2 = self.gview((L - 1, "lin2_b"))
        if self.has_norm:
            ops.layernorm_bwd(g, ws["x_last"], ws["meanf"], ws["rstdf"], self.f(("g", "norm_w")), dx=d, dx_bf16=d_bf if pd == 0 else None,
                              dgamma=self.gview(("g", "norm_w")), dbeta=self.gview(("g", "norm_b")), dx_colsum=last_b2 if pd == 0 else None)
            self._seg_done(seg)
            seg += 1
        else:
            d.copy_(g)
            if pd == 0:
                ops.cast_bf16(d.view(-1), d_bf.view(-1))
                ops.colsum_bf16(d_bf, last_b2)
        if pd > 0:
            self._masked_operand(ws, d, d_bf, L - 1, 2, (L - 1, "lin2_b"))
        for li in range(L - 1, -1, -1):
            buf = ws["layer"][li]
            x_in = ws["xin"][li]
            # ---- FFN: x_out = x1 + linear2(act(linear1(norm2(x1)))) ----
            act_out = buf["a"] if self.act == "relu" else buf["g"]
            self._wgrad(d_bf, act_out, (li, "lin2_w"))
            ops.gemm(d_bf, self.w((li, "lin2_w")), ws["da"], b_major=1, epilogue=ops.EPI_DRELU if self.act == "relu" else ops.EPI_DGELU,
                     aux=buf["a"], **drelu)
            self._wgrad(ws["da"], buf["x1_bf"], (li, "lin1_w"), bias_key=(li, "lin1_b"))
            ops.gemm(ws["da"], self.w((li, "lin1_w")), dh, b_major=1)
            ops.layernorm_bwd(dh, buf["x1"], buf["mean2"], buf["rstd2"], self.f((li, "norm2_w")), dres=d, dx=d, dx_bf16=d_bf if pd == 0 else None,
                              dgamma=self.gview((li, "norm2_w")), dbeta=self.gview((li, "norm2_b")),
                              dx_colsum=self.gview((li, "out_b")) if pd == 0 else None)
            if pd > 0:
                self._masked_operand(ws, d, d_bf, li, 0, (li, "out_b"))
            # ---- attention: x1 = x_in + out_proj(attn(q = k = norm1(x_in) + pos, v = norm1(x_in))) ----
            self._wgrad(d_bf, buf["o"], (li, "out_w"))
            ops.gemm(d_bf, self.w((li, "out_w")), dh, b_major=1)      # dO
            dqk, dv = ws["dqk"], ws["dv"]
            ops.attention_bwd(buf["qkb"][:, :D], buf["qkb"][:, D:], buf["vb"], buf["o"], buf["lse"], dh, dqk[:, :D], dqk[:, D:], dv,
                              ws["delta"], B=N, H=H, S=S, tok_stride=N, batch_stride=1, key_padding_mask=kpm, dropout=drop3(li))
            in_w = self.w((li, "in_w"))
            gb = self.gview((li, "in_b"))
            self._wgrad(dqk, buf["qk_bf"] if pos2 is not None else buf["x_bf"], (li, "in_w"), rows=(0, 2 * D), bias_grad=gb[:2 * D])
            self._wgrad(dv, buf["x_bf"], (li, "in_w"), rows=(2 * D, 3 * D), bias_grad=gb[2 * D:])
            ops.gemm(dqk, in_w[:2 * D], dh, b_major=1)                 # d(norm1(x) + pos) through q and k
            ops.gemm(dv, in_w[2 * D:], dh2, b_major=1)                 # d norm1(x) through v
            if dpos is not None:
                ops.add3(dpos, dh, None, dpos, None)                   # d pos += dh
            prev_b2 = self.gview((li - 1, "lin2_b")) if li > 0 else None
            ops.layernorm_bwd(dh, x_in, buf["mean1"], buf["rstd1"], self.f((li, "norm1_w")), dres=d, dx=d,
                              dx_bf16=d_bf if (li > 0 and pd == 0) else None, dgamma=self.gview((li, "norm1_w")),
                              dbeta=self.gview((li, "norm1_b")), dx_colsum=prev_b2 if pd == 0 else None, dy_add=dh2)
            if pd > 0 and li > 0:
                self._masked_operand(ws, d, d_bf, li - 1, 2, (li - 1, "lin2_b"))
            self._seg_done(seg)
            seg += 1
        return d.view(S, N, D), (dpos.view(S, N, D) if dpos is not None else None)

    def backward(self, ws, grad_out):
        self.prepare_grads()
        if self.pre_norm:
            return self._backward_pre(ws, grad_out)
        S, N, M, D, H = ws["S"], ws["N"], ws["M"], self.D, self.H
        dA, dB, dA_bf, dB_bf, dh, dh2 = ws["dA"], ws["dB"], ws["dA_bf"], ws["dB_bf"], ws["dh"], ws["dh2"]
        g = grad_out.contiguous().float().view(M, D)
        pos2, kpm = ws["pos"], ws["kpm"]
        dpos = None
        if pos2 is not None:
            dpos = ws["dpos"]
            dpos.zero_()
        seg = 0
        pd = ws["p_drop"]
        drop3 = lambda li: (pd, ws["drop_seed"], self.drop_site(li, 3)) if pd > 0 else None
        drelu = dict(drelu_scale=1.0 / (1.0 - pd)) if (pd > 0 and self.act == "relu") else {}
        if self.has_norm:
            ops.layernorm_bwd(g, ws["x_last"], ws["meanf"], ws["rstdf"], self.f(("g", "norm_w")), dx=dA, dgamma=self.gview(("g", "norm_w")),
                              dbeta=self.gview(("g", "norm_b")))
            self._seg_done(seg)
            seg += 1
            g = dA
        for li in range(self.L - 1, -1, -1):
            buf = ws["layer"][li]
            # norm2: gradient w.r.t. x2pre = x1 + linear2(act(linear1(x1)))
            ops.layernorm_bwd(g, buf["x2pre"], buf["mean2"], buf["rstd2"], self.f((li, "norm2_w")), dx=dB, dx_bf16=dB_bf if pd == 0 else None,
                              dgamma=self.gview((li, "norm2_w")), dbeta=self.gview((li, "norm2_b")),
                              dx_colsum=self.gview((li, "lin2_b")) if pd == 0 else None)
            if pd > 0:   # dB stays the un-masked gradient of x2pre (the residual branch); the linear2 branch sees keep * dB / (1 - p)
                self._masked_operand(ws, dB, dB_bf, li, 2, (li, "lin2_b"))
            act_out = buf["a"] if self.act == "relu" else buf["g"]
            self._wgrad(dB_bf, act_out, (li, "lin2_w"))
            if self.act == "relu":
                ops.gemm(dB_bf, self.w((li, "lin2_w")), ws["da"], b_major=1, epilogue=ops.EPI_DRELU, aux=buf["a"], **drelu)
            else:
                ops.gemm(dB_bf, self.w((li, "lin2_w")), ws["da"], b_major=1, epilogue=ops.EPI_DGELU, aux=buf["a"])
            self._wgrad(ws["da"], buf["x1_bf"], (li, "lin1_w"), bias_key=(li, "lin1_b"))
            ops.gemm(ws["da"], self.w((li, "lin1_w")), dh, b_major=1)
            # norm1: its output x1 received dB (residual) + dh (through the FFN); input is x1pre = src + out_proj(attn)
            ops.layernorm_bwd(dB, buf["x1pre"], buf["mean1"], buf["rstd1"], self.f((li, "norm1_w")), dx=dA, dx_bf16=dA_bf if pd == 0 else None,
                              dgamma=self.gview((li, "norm1_w")), dbeta=self.gview((li, "norm1_b")),
                              dx_colsum=self.gview((li, "out_b")) if pd == 0 else None, dy_add=dh)
            if pd > 0:
                self._masked_operand(ws, dA, dA_bf, li, 0, (li, "out_b"))
            self._wgrad(dA_bf, buf["o"], (li, "out_w"))
            ops.gemm(dA_bf, self.w((li, "out_w")), dh, b_major=1)      # dO
            dqk, dv = ws["dqk"], ws["dv"]
            ops.attention_bwd(buf["qkb"][:, :D], buf["qkb"][:, D:], buf["vb"], buf["o"], buf["lse"], dh, dqk[:, :D], dqk[:, D:], dv,
                              ws["delta"], B=N, H=H, S=S, tok_stride=N, batch_stride=1, key_padding_mask=kpm, dropout=drop3(li))
            in_w = self.w((li, "in_w"))
            gb = self.gview((li, "in_b"))
            self._wgrad(dqk, buf["qk_bf"], (li, "in_w"), rows=(0, 2 * D), bias_grad=gb[:2 * D])
            self._wgrad(dv, buf["x_bf"], (li, "in_w"), rows=(2 * D, 3 * D), bias_grad=gb[2 * D:])
            ops.gemm(dqk, in_w[:2 * D], dh, b_major=1)                 # d(src + pos) through q and k
            ops.gemm(dv, in_w[2 * D:], dh2, b_major=1)                 # d src through v
            # d src = dA (residual around attention) + dh + dh2; d pos += dh
            ops.add3(dA, dh, dh2, dB, dpos)
            g = dB
            # ping-pong: next iteration's norm2 backward reads g (= dB) and writes dB again -> swap roles
            dA, dB = dB, dA
            dA_bf, dB_bf = dB_bf, dA_bf
            self._seg_done(seg)
            seg += 1
        return g.view(S, N, D), (dpos.view(S, N, D) if dpos is not None else None)


class _DetrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, kpm, src, pos, *params):
        engine.ensure_bound()
        S, N = src.shape[0], src.shape[1]

        def fresh():
            engine.bf16_fresh = False     # a captured forward must contain the fp32 -> bf16 parameter cast
        out, ws = engine.graphed(("enc_fwd", float(engine.p_drop)), S * N, [src, pos, kpm],
                                 lambda s_, p_, k_: engine.forward(s_, p_, k_, training=True), before_capture=fresh,
                                 busy=lambda res: res[1].get("leased"))
        ctx.engine, ctx.ws, ctx.n_params, ctx.has_pos = engine, ws, len(params), pos is not None
        ctx.lease = WorkspaceLease(ws)
        return out.clone()

    @staticmethod
    def backward(ctx, grad_out):
        eng, ws = ctx.engine, ctx.ws
        eng.prepare_grads()               # p.grad bookkeeping stays outside a captured backward
        dsrc, dpos = eng.graphed(("enc_bwd", id(ws), ws["p_drop"], ws["pos"] is not None, ws["kpm"] is not None), ws["M"],
                                 [grad_out], lambda g_: eng.backward(ws, g_))
        dsrc = dsrc.clone()
        dpos = dpos.clone() if (ctx.has_pos and ctx.needs_input_grad[3]) else None
        ctx.lease.release()
        return (None, None, dsrc, dpos) + (None,) * ctx.n_params


class TransformerEncoder(nn.Module):
    __getstate__ = getstate_without_engine

    def __init__(self, encoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm
        self.__dict__["_engine"] = None

    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            l0 = self.layers[0]
            eng = DetrEngine(list(self.layers), self.norm, l0.d_model, l0.nhead, l0.dim_feedforward, l0.activation_name,
                             eps=float(l0.norm1.eps), pre_norm=l0.normalize_before)
            self.__dict__["_engine"] = eng
        return eng

    def forward(self, src, mask=None, src_key_padding_mask=None, pos=None):
        l0 = self.layers[0]
        if mask is not None:
            raise NotImplementedError("vitb200: src_mask (attn_mask) is not supported; DETR passes None (transformer.py:59)")
        eng = self._get_engine()
        eng.p_drop = float(l0.dropout_p) if self.training else 0.0   # one rate for the four sites of a layer (transformer.py:195-205)
        kpm = None
        if src_key_padding_mask is not None:
            kpm = src_key_padding_mask.to(device=src.device, dtype=torch.uint8).contiguous()
        params = [p for _, p in eng._order]
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or src.requires_grad)
        if needs_grad:
            return _DetrFn.apply(eng, kpm, src, pos, *params)
        out, _ = eng.forward(src, pos, kpm, training=False)
        return out.clone()


# ======================================================================================================================
# DETR transformer DECODER: TransformerDecoderLayer / TransformerDecoder, transformer.py:66-95, 118-189.
#
# As written the reference layer cannot run: __init__ registers the cross-attention as ``multi_head_attn`` (transformer.py:122) while
# forward_post / forward_pre call ``self.multihead_attn`` (:148, :172).  Decision: the state_dict key is the one
# __init__ creates (``multi_head_attn.*``) and the forward is the one the code spells out with that attribute resolved — the parity fixtures
# come from the live reference run with exactly that alias added (tools/make_golden.py).  Post-norm layers (the reference default,
# transformer.py:27-28); sequence-first tensors; the decoder's own norm and ``return_intermediate`` as at :78-95.
# ======================================================================================================================
DEC_ROLES = ("norm3_w", "norm3_b", "lin2_w", "lin2_b", "lin1_w", "lin1_b", "norm2_w", "norm2_b", "cout_w", "cout_b", "cin_w", "cin_b",
             "norm1_w", "norm1_b", "out_w", "out_b", "in_w", "in_b")


class TransformerDecoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multi_head_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        if activation not in ("relu", "gelu"):
            raise RuntimeError(F"activation should be relu/gelu, not {activation}.")
        self.activation_name = activation
        self.normalize_before = normalize_before
        self.d_model, self.nhead, self.dim_feedforward, self.dropout_p = d_model, nhead, dim_feedforward, dropout

    @property
    def multihead_attn(self):          # the name the reference's forward uses (transformer.py:148,172)
        return self.multi_head_attn

    def roles(self):
        return {"norm1_w": self.norm1.weight, "norm1_b": self.norm1.bias, "norm2_w": self.norm2.weight, "norm2_b": self.norm2.bias,
                "norm3_w": self.norm3.weight, "norm3_b": self.norm3.bias, "lin1_w": self.linear1.weight, "lin1_b": self.linear1.bias,
                "lin2_w": self.linear2.weight, "lin2_b": self.linear2.bias,
                "in_w": self.self_attn.in_proj_weight, "in_b": self.self_attn.in_proj_bias,
                "out_w": self.self_attn.out_proj.weight, "out_b": self.self_attn.out_proj.bias,
                "cin_w": self.multi_head_attn.in_proj_weight, "cin_b": self.multi_head_attn.in_proj_bias,
                "cout_w": self.multi_head_attn.out_proj.weight, "cout_b": self.multi_head_attn.out_proj.bias}

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None, pos=None,
                query_pos=None):
        dec = self.__dict__.get("_solo")
        if dec is None:
            dec = TransformerDecoder.__new__(TransformerDecoder)
            nn.Module.__init__(dec)
            dec.layers = nn.ModuleList([self])
            dec.num_layers, dec.norm, dec.return_intermediate = 1, None, False
            dec.__dict__["_engine"] = None
            self.__dict__["_solo"] = dec
        return dec(tgt, memory, tgt_mask, memory_mask, tgt_key_padding_mask, memory_key_padding_mask, pos, query_pos)[0]

    __deepcopy__ = TransformerEncoderLayer.__deepcopy__
    __getstate__ = getstate_without_engine


class DetrDecoderEngine(DetrEngine):
    """Kernel sequencing for L post-norm decoder layers: self-attention over the Q object queries, cross-attention of the queries
    against the S-token encoder memory (vb_attention with S_kv), feed-forward; three LayerNorms; six dropout sites per layer
    (0 = dropout1, 1 = dropout, 2 = dropout3, 3 = self-attention weights, 4 = cross-attention weights, 5 = dropout2)."""

    def __init__(self, layers, norm, d_model, nhead, dim_feedforward, activation, eps=1e-5, return_intermediate=False, pre_norm=False):
        if d_model % 64 != 0 or d_model // nhead != 64 or d_model > 1024:
            raise NotImplementedError(f"vitb200 kernels need head_dim == 64 and d_model <= 1024 (got d_model {d_model}, {nhead} heads)")
        self.D, self.H, self.F, self.L, self.eps = d_model, nhead, dim_feedforward, len(layers), eps
        self.act, self.pre_norm, self.has_norm = activation, bool(pre_norm), norm is not None
        self.return_intermediate = bool(return_intermediate) and self.has_norm
        self.p_drop, self._drop_counter = 0.0, None
        self._order, seg = [], []
        if self.has_norm:
            self._order += [(("g", "norm_w"), norm.weight), (("g", "norm_b"), norm.bias)]
            seg.append(2)
        for li in range(self.L - 1, -1, -1):
            r = layers[li].roles()
            self._order += [((li, k), r[k]) for k in DEC_ROLES]
            seg.append(len(DEC_ROLES))
        self._layout(seg)

    def workspace(self, Q, S, N, training):
        key = (Q, S, N, training)
        ws = self._free_workspace(key)
        if ws is not None:
            return ws
        dev, D, Fd, H = self.flat.device, self.D, self.F, self.H
        Mq, Ms = Q * N, S * N
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *shape, dtype=bf: torch.empty(*shape, device=dev, dtype=dtype)
        ws = {"Q": Q, "S": S, "N": N, "M": Mq, "Ms": Ms, "layer": []}
        for _ in range(self.L):    # intermediates may be returned, so every layer keeps its own buffers in both modes
            ws["layer"].append({
                "x_bf": e(Mq, D), "qk_bf": e(Mq, D), "sqk": e(Mq, 2 * D), "sv": e(Mq, D), "so": e(Mq, D), "slse": e(N, H, Q, dtype=f32),
                "u1": e(Mq, D, dtype=f32), "t1": e(Mq, D, dtype=f32), "t1_bf": e(Mq, D), "t1q_bf": e(Mq, D),
                "mean1": e(Mq, dtype=f32), "rstd1": e(Mq, dtype=f32),
                "cq": e(Mq, D), "ck": e(Ms, D), "cv": e(Ms, D), "co": e(Mq, D), "clse": e(N, H, Q, dtype=f32),
                "u2": e(Mq, D, dtype=f32), "t2": e(Mq, D, dtype=f32), "x1_bf": e(Mq, D), "mean2": e(Mq, dtype=f32), "rstd2": e(Mq, dtype=f32),
                "a": e(Mq, Fd), "u3": e(Mq, D, dtype=f32), "mean3": e(Mq, dtype=f32), "rstd3": e(Mq, dtype=f32),
                "t3": e(Mq, D, dtype=f32)})
        ws["m_bf"], ws["mk_bf"] = e(Ms, D), e(Ms, D)
        n_out = self.L if self.return_intermediate else 1
        ws["y"] = e(n_out, Mq, D, dtype=f32)
        ws["meanf"], ws["rstdf"] = e(n_out, Mq, dtype=f32), e(n_out, Mq, dtype=f32)
        if training:
            for k in ("dT", "dU", "dU2", "dqpos"):
                ws[k] = e(Mq, D, dtype=f32)
            ws["dU_bf"], ws["dh"], ws["dh2"] = e(Mq, D), e(Mq, D), e(Mq, D)
            ws["da"] = e(Mq, Fd)
            ws["dsqk"], ws["dsv"], ws["dcq"] = e(Mq, 2 * D), e(Mq, D), e(Mq, D)
            ws["dck"], ws["dcv"], ws["dmk"], ws["dmv"] = e(Ms, D), e(Ms, D), e(Ms, D), e(Ms, D)
            ws["delta"] = e(N, H, Q, dtype=f32)
            ws["dmem"], ws["dpos"] = e(Ms, D, dtype=f32), e(Ms, D, dtype=f32)
        self._ws[key].append(ws)
        return ws

    def forward(self, tgt, memory, mem_kpm, pos, query_pos, training):
        self.ensure_bound()
        for t, name in ((tgt, "tgt"), (memory, "memory"), (pos, "pos"), (query_pos, "query_pos"), (mem_kpm, "memory_key_padding_mask")):
            self.check_input(t, name)
        Q, N, D = tgt.shape
        S = memory.shape[0]
        if D != self.D or tuple(memory.shape[1:]) != (N, D) or (pos is not None and pos.shape != memory.shape) \
                or (query_pos is not None and query_pos.shape != tgt.shape) or (mem_kpm is not None and tuple(mem_kpm.shape) != (N, S)):
            raise RuntimeError(f"vitb200: expected tgt/query_pos [Q, N, {self.D}], memory/pos [S, N, {self.D}], memory_key_padding_mask "
                               f"[N, S]; got tgt {tuple(tgt.shape)}, memory {tuple(memory.shape)}")
        ws = self.workspace(Q, S, N, training)
        self.refresh_bf16()
        Mq, Ms, H = ws["M"], ws["Ms"], self.H
        t0 = tgt.contiguous().float().view(Mq, D)
        mem = memory.contiguous().float().view(Ms, D)
        pos2 = pos.contiguous().float().view(Ms, D) if pos is not None else None
        qpos = query_pos.contiguous().float().view(Mq, D) if query_pos is not None else None
        ws["has_pos"], ws["has_qpos"], ws["kpm"] = pos2 is not None, qpos is not None, mem_kpm
        pd = ws["p_drop"] = float(self.p_drop)
        if pd > 0:
            self._begin_dropout(ws)
        site = self.drop_site
        seed = ws.get("drop_seed")
        ops.add_cast_bf16(mem, None, ws["m_bf"])                     # value operand of every layer's cross-attention
        ops.add_cast_bf16(mem, pos2, ws["mk_bf"])                    # key operand: memory + pos (transformer.py:146)
        if self.pre_norm:
            x = self._forward_pre_layers(ws, t0, qpos, mem_kpm, training, Q, S, N)
            return self._finish_forward(ws, x, Q, N, D)
        b0 = ws["layer"][0]
        ops.add_cast_bf16(t0, None, b0["x_bf"])
        ops.add_cast_bf16(t0, qpos, b0["qk_bf"])
        x = t0
        for li in range(self.L):
            buf = ws["layer"][li]
            # ---- self-attention over the queries: q = k = tgt + query_pos, v = tgt (transformer.py:142-143) ----
            in_w, in_b = self.w((li, "in_w")), self.f((li, "in_b"))
            ops.gemm(buf["qk_bf"], in_w[:2 * D], buf["sqk"], bias=in_b[:2 * D])
            ops.gemm(buf["x_bf"], in_w[2 * D:], buf["sv"], bias=in_b[2 * D:])
            ops.attention_fwd(buf["sqk"][:, :D], buf["sqk"][:, D:], buf["sv"], buf["so"], buf["slse"] if training else None, B=N, H=H, S=Q,
                              tok_stride=N, batch_stride=1, dropout=(pd, seed, site(li, 3)) if pd > 0 else None)
            self._linear_residual(ws, buf["so"], "out_w", "out_b", li, 0, x, buf["u1"], pd)                       # :144
            ops.layernorm_fwd(buf["u1"], self.f((li, "norm1_w")), self.f((li, "norm1_b")), self.eps, y_f32=buf["t1"], y_bf16=buf["t1_bf"],
                              mean=buf["mean1"], rstd=buf["rstd1"], add=qpos, y2_bf16=buf["t1q_bf"] if qpos is not None else None)
            # ---- cross-attention: query = tgt + query_pos, key = memory + pos, value = memory (:145-147) ----
            cw, cb = self.w((li, "cin_w")), self.f((li, "cin_b"))
            ops.gemm(buf["t1q_bf"] if qpos is not None else buf["t1_bf"], cw[:D], buf["cq"], bias=cb[:D])
            ops.gemm(ws["mk_bf"], cw[D:2 * D], buf["ck"], bias=cb[D:2 * D])
            ops.gemm(ws["m_bf"], cw[2 * D:], buf["cv"], bias=cb[2 * D:])
            ops.attention_fwd(buf["cq"], buf["ck"], buf["cv"], buf["co"], buf["clse"] if training else None, B=N, H=H, S=Q, S_kv=S,
                              tok_stride=N, batch_stride=1, key_padding_mask=mem_kpm,
                              dropout=(pd, seed, site(li, 4)) if pd > 0 else None)
            self._linear_residual(ws, buf["co"], "cout_w", "cout_b", li, 5, buf["t1"], buf["u2"], pd)             # :150
            ops.layernorm_fwd(buf["u2"], self.f((li, "norm2_w")), self.f((li, "norm2_b")), self.eps, y_f32=buf["t2"], y_bf16=buf["x1_bf"],
                              mean=buf["mean2"], rstd=buf["rstd2"])
            # ---- feed-forward (:152-154) ----
            act_out = self._ffn_fwd(ws, buf, li, pd)
            self._linear_residual(ws, act_out, "lin2_w", "lin2_b", li, 2, buf["t2"], buf["u3"], pd)
            nbuf = ws["layer"][li + 1] if li + 1 < self.L else None
            ops.layernorm_fwd(buf["u3"], self.f((li, "norm3_w")), self.f((li, "norm3_b")), self.eps, y_f32=buf["t3"],
                              y_bf16=nbuf["x_bf"] if nbuf else None, mean=buf["mean3"], rstd=buf["rstd3"],
                              add=qpos if nbuf else None, y2_bf16=nbuf["qk_bf"] if (nbuf and qpos is not None) else None)
            if nbuf and qpos is None:
                ops.add_cast_bf16(buf["t3"], None, nbuf["qk_bf"])
            x = buf["t3"]
        return self._finish_forward(ws, x, Q, N, D)

    def _finish_forward(self, ws, x, Q, N, D):
        if not self.has_norm:
            return x.view(1, Q, N, D), ws                            # output.unsqueeze(0) (:95)
        outs = range(self.L) if self.return_intermediate else [self.L - 1]
        for j, li in enumerate(outs):                                # self.norm(output) per returned layer (:83-88)
            ops.layernorm_fwd(ws["layer"][li]["t3"], self.f(("g", "norm_w")), self.f(("g", "norm_b")), self.eps, y_f32=ws["y"][j],
                              mean=ws["meanf"][j], rstd=ws["rstdf"][j])
        return ws["y"].view(-1, Q, N, D), ws

    # ------------------------------------------------------------------------------------------------------------------
    # Pre-norm layers (normalize_before=True, TransformerDecoderLayer.forward_pre, transformer.py:158-178): tgt2 = norm1(tgt);
    # q = k = tgt2 + query_pos; tgt += dropout1(self_attn(q, k, tgt2)); tgt2 = norm2(tgt); tgt += dropout2(cross(tgt2 + query_pos,
    # memory + pos, memory)); tgt2 = norm3(tgt); tgt += dropout3(ffn(tgt2)).  The fp32 stream is t0 -> t1 -> t2 -> t3.
    # ------------------------------------------------------------------------------------------------------------------
    def _forward_pre_layers(self, ws, t0, qpos, mem_kpm, training, Q, S, N):
        D, H = self.D, self.H
        pd, seed, site = ws["p_drop"], ws.get("drop_seed"), self.drop_site
        if "t0" not in ws:
            ws["t0"] = torch.empty_like(ws["layer"][0]["t3"])
        ws["t0"].copy_(t0)
        x = ws["t0"]
        for li in range(self.L):
            buf = ws["layer"][li]
            ops.layernorm_fwd(x, self.f((li, "norm1_w")), self.f((li, "norm1_b")), self.eps, y_bf16=buf["x_bf"], mean=buf["mean1"],
                              rstd=buf["rstd1"], add=qpos, y2_bf16=buf["qk_bf"] if qpos is not None else None)
            qk_in = buf["qk_bf"] if qpos is not None else buf["x_bf"]
            in_w, in_b = self.w((li, "in_w")), self.f((li, "in_b"))
            ops.gemm(qk_in, in_w[:2 * D], buf["sqk"], bias=in_b[:2 * D])
            ops.gemm(buf["x_bf"], in_w[2 * D:], buf["sv"], bias=in_b[2 * D:])
            ops.attention_fwd(buf["sqk"][:, :D], buf["sqk"][:, D:], buf["sv"], buf["so"], buf["slse"] if training else None, B=N, H=H, S=Q,
                              tok_stride=N, batch_stride=1, dropout=(pd, seed, site(li, 3)) if pd > 0 else None)
            self._linear_residual(ws, buf["so"], "out_w", "out_b", li, 0, x, buf["t1"], pd)
            ops.layernorm_fwd(buf["t1"], self.f((li, "norm2_w")), self.f((li, "norm2_b")), self.eps, y_bf16=buf["t1_bf"], mean=buf["mean2"],
                              rstd=buf["rstd2"], add=qpos, y2_bf16=buf["t1q_bf"] if qpos is not None else None)
            cw, cb = self.w((li, "cin_w")), self.f((li, "cin_b"))
            ops.gemm(buf["t1q_bf"] if qpos is not None else buf["t1_bf"], cw[:D], buf["cq"], bias=cb[:D])
            ops.gemm(ws["mk_bf"], cw[D:2 * D], buf["ck"], bias=cb[D:2 * D])
            ops.gemm(ws["m_bf"], cw[2 * D:], buf["cv"], bias=cb[2 * D:])
            ops.attention_fwd(buf["cq"], buf["ck"], buf["cv"], buf["co"], buf["clse"] if training else None, B=N, H=H, S=Q, S_kv=S,
                              tok_stride=N, batch_stride=1, key_padding_mask=mem_kpm, dropout=(pd, seed, site(li, 4)) if pd > 0 else None)
            self._linear_residual(ws, buf["co"], "cout_w", "cout_b", li, 5, buf["t1"], buf["t2"], pd)
            ops.layernorm_fwd(buf["t2"], self.f((li, "norm3_w")), self.f((li, "norm3_b")), self.eps, y_bf16=buf["x1_bf"], mean=buf["mean3"],
                              rstd=buf["rstd3"])
            act_out = self._ffn_fwd(ws, buf, li, pd)
            self._linear_residual(ws, act_out, "lin2_w", "lin2_b", li, 2, buf["t2"], buf["t3"], pd)
            x = buf["t3"]
        return x

    def _backward_pre_layers(self, ws, g_all):
        Q, S, N, Mq, D, H, L = ws["Q"], ws["S"], ws["N"], ws["M"], self.D, self.H, self.L
        d, d_bf, dh, dh2 = ws["dT"], ws["dU_bf"], ws["dh"], ws["dh2"]     # d: fp32 gradient of the stream, updated in place
        dmem, dpos, dqpos = ws["dmem"], ws["dpos"], ws["dqpos"]
        pd, kpm, seed, site = ws["p_drop"], ws["kpm"], ws.get("drop_seed"), self.drop_site
        drelu = dict(drelu_scale=1.0 / (1.0 - pd)) if (pd > 0 and self.act == "relu") else {}
        fuse = pd == 0
        have_next = False
        for li in range(L - 1, -1, -1):
            buf = ws["layer"][li]
            x_in = ws["layer"][li - 1]["t3"] if li > 0 else ws["t0"]
            gets_norm = self.has_norm and (self.return_intermediate or li == L - 1)
            lin2_b = self.gview((li, "lin2_b"))
            if gets_norm:      # the decoder norm's backward also emits this layer's bf16 operand and linear2 bias gradient
                j = li if self.return_intermediate else 0
                ops.layernorm_bwd(g_all[j], buf["t3"], ws["meanf"][j], ws["rstdf"][j], self.f(("g", "norm_w")), dres=d if have_next else None,
                                  dx=d, dx_bf16=d_bf if fuse else None, dgamma=self.gview(("g", "norm_w")), dbeta=self.gview(("g", "norm_b")),
                                  dx_colsum=lin2_b if fuse else None)
                if not self.return_intermediate:
                    self._seg_done(0)
            elif not have_next:
                d.copy_(g_all[0])
                if fuse:
                    ops.cast_bf16(d.view(-1), d_bf.view(-1))
                    ops.colsum_bf16(d_bf, lin2_b)
            if pd > 0:
                self._masked_operand(ws, d, d_bf, li, 2, (li, "lin2_b"))
            # ---- feed-forward ----
            act_out = buf["a"] if self.act == "relu" else buf["g"]
            self._wgrad(d_bf, act_out, (li, "lin2_w"))
            ops.gemm(d_bf, self.w((li, "lin2_w")), ws["da"], b_major=1, epilogue=ops.EPI_DRELU if self.act == "relu" else ops.EPI_DGELU,
                     aux=buf["a"], **drelu)
            self._wgrad(ws["da"], buf["x1_bf"], (li, "lin1_w"), bias_key=(li, "lin1_b"))
            ops.gemm(ws["da"], self.w((li, "lin1_w")), dh, b_major=1)
            ops.layernorm_bwd(dh, buf["t2"], buf["mean3"], buf["rstd3"], self.f((li, "norm3_w")), dres=d, dx=d, dx_bf16=d_bf if fuse else None,
                              dgamma=self.gview((li, "norm3_w")), dbeta=self.gview((li, "norm3_b")),
                              dx_colsum=self.gview((li, "cout_b")) if fuse else None)
            if pd > 0:
                self._masked_operand(ws, d, d_bf, li, 5, (li, "cout_b"))
            # ---- cross-attention ----
            self._wgrad(d_bf, buf["co"], (li, "cout_w"))
            ops.gemm(d_bf, self.w((li, "cout_w")), dh, b_major=1)
            ops.attention_bwd(buf["cq"], buf["ck"], buf["cv"], buf["co"], buf["clse"], dh, ws["dcq"], ws["dck"], ws["dcv"], ws["delta"],
                              B=N, H=H, S=Q, S_kv=S, tok_stride=N, batch_stride=1, key_padding_mask=kpm,
                              dropout=(pd, seed, site(li, 4)) if pd > 0 else None)
            cw, gcb = self.w((li, "cin_w")), self.gview((li, "cin_b"))
            self._wgrad(ws["dcq"], buf["t1q_bf"] if ws["has_qpos"] else buf["t1_bf"], (li, "cin_w"), rows=(0, D), bias_grad=gcb[:D])
            self._wgrad(ws["dck"], ws["mk_bf"], (li, "cin_w"), rows=(D, 2 * D), bias_grad=gcb[D:2 * D])
            self._wgrad(ws["dcv"], ws["m_bf"], (li, "cin_w"), rows=(2 * D, 3 * D), bias_grad=gcb[2 * D:])
            ops.gemm(ws["dcq"], cw[:D], dh, b_major=1)
            ops.gemm(ws["dck"], cw[D:2 * D], ws["dmk"], b_major=1)
            ops.gemm(ws["dcv"], cw[2 * D:], ws["dmv"], b_major=1)
            ops.add3(dmem, ws["dmk"], ws["dmv"], dmem, dpos)
            ops.add3(dqpos, dh, None, dqpos, None)
            ops.layernorm_bwd(dh, buf["t1"], buf["mean2"], buf["rstd2"], self.f((li, "norm2_w")), dres=d, dx=d, dx_bf16=d_bf if fuse else None,
                              dgamma=self.gview((li, "norm2_w")), dbeta=self.gview((li, "norm2_b")),
                              dx_colsum=self.gview((li, "out_b")) if fuse else None)
            if pd > 0:
                self._masked_operand(ws, d, d_bf, li, 0, (li, "out_b"))
            # ---- self-attention ----
            self._wgrad(d_bf, buf["so"], (li, "out_w"))
            ops.gemm(d_bf, self.w((li, "out_w")), dh, b_major=1)
            dsqk, dsv = ws["dsqk"], ws["dsv"]
            ops.attention_bwd(buf["sqk"][:, :D], buf["sqk"][:, D:], buf["sv"], buf["so"], buf["slse"], dh, dsqk[:, :D], dsqk[:, D:], dsv,
                              ws["delta"], B=N, H=H, S=Q, tok_stride=N, batch_stride=1, dropout=(pd, seed, site(li, 3)) if pd > 0 else None)
            in_w, gb = self.w((li, "in_w")), self.gview((li, "in_b"))
            self._wgrad(dsqk, buf["qk_bf"] if ws["has_qpos"] else buf["x_bf"], (li, "in_w"), rows=(0, 2 * D), bias_grad=gb[:2 * D])
            self._wgrad(dsv, buf["x_bf"], (li, "in_w"), rows=(2 * D, 3 * D), bias_grad=gb[2 * D:])
            ops.gemm(dsqk, in_w[:2 * D], dh, b_major=1)       # d(norm1(t0) + query_pos)
            ops.gemm(dsv, in_w[2 * D:], dh2, b_major=1)       # d norm1(t0) through the values
            ops.add3(dqpos, dh, None, dqpos, None)
            # the previous layer's operand / linear2 bias gradient come from here unless its own decoder-norm backward emits them
            prev_fused = fuse and li > 0 and not (self.has_norm and self.return_intermediate)
            ops.layernorm_bwd(dh, x_in, buf["mean1"], buf["rstd1"], self.f((li, "norm1_w")), dres=d, dx=d, dx_bf16=d_bf if prev_fused else None,
                              dgamma=self.gview((li, "norm1_w")), dbeta=self.gview((li, "norm1_b")),
                              dx_colsum=self.gview((li - 1, "lin2_b")) if prev_fused else None, dy_add=dh2)
            have_next = True
            if not self.return_intermediate:
                self._seg_done((L - 1 - li) + (1 if self.has_norm else 0))
        if self.return_intermediate:
            for i in range(L + 1):
                self._seg_done(i)
        return d

    def backward(self, ws, grad_out):
        self.prepare_grads()
        Q, S, N, Mq, Ms, D, H, L = ws["Q"], ws["S"], ws["N"], ws["M"], ws["Ms"], self.D, self.H, self.L
        g_all = grad_out.contiguous().float().view(-1, Mq, D)
        dT, dU, dU2, dU_bf, dh, dh2 = ws["dT"], ws["dU"], ws["dU2"], ws["dU_bf"], ws["dh"], ws["dh2"]
        dmem, dpos, dqpos = ws["dmem"], ws["dpos"], ws["dqpos"]
        dmem.zero_()
        dpos.zero_()
        dqpos.zero_()
        pd, kpm = ws["p_drop"], ws["kpm"]
        seed = ws.get("drop_seed")
        site = self.drop_site
        drelu = dict(drelu_scale=1.0 / (1.0 - pd)) if (pd > 0 and self.act == "relu") else {}
        fuse = pd == 0
        seg = 0
        if self.pre_norm:
            d = self._backward_pre_layers(ws, g_all)
            return (d.view(Q, N, D), dmem.view(S, N, D), dpos.view(S, N, D) if ws["has_pos"] else None,
                    dqpos.view(Q, N, D) if ws["has_qpos"] else None)
        have_next = False     # dT holds the gradient that layer li + 1 sends to this layer's output
        for li in range(L - 1, -1, -1):
            buf = ws["layer"][li]
            # ---- gradient of this layer's output t3: from the next layer and / or through the decoder norm ----
            if self.has_norm and (self.return_intermediate or li == L - 1):
                j = li if self.return_intermediate else 0
                ops.layernorm_bwd(g_all[j], buf["t3"], ws["meanf"][j], ws["rstdf"][j], self.f(("g", "norm_w")), dres=dT if have_next else None,
                                  dx=dT, dgamma=self.gview(("g", "norm_w")), dbeta=self.gview(("g", "norm_b")))
                if not self.return_intermediate:
                    self._seg_done(0)
            elif not self.has_norm and li == L - 1:
                dT.copy_(g_all[0])
            # ---- feed-forward: t3 = norm3(u3), u3 = t2 + dropout3(linear2(dropout(act(linear1(t2))))) ----
            ops.layernorm_bwd(dT, buf["u3"], buf["mean3"], buf["rstd3"], self.f((li, "norm3_w")), dx=dU, dx_bf16=dU_bf if fuse else None,
                              dgamma=self.gview((li, "norm3_w")), dbeta=self.gview((li, "norm3_b")),
                              dx_colsum=self.gview((li, "lin2_b")) if fuse else None)
            if pd > 0:
                self._masked_operand(ws, dU, dU_bf, li, 2, (li, "lin2_b"))
            act_out = buf["a"] if self.act == "relu" else buf["g"]
            self._wgrad(dU_bf, act_out, (li, "lin2_w"))
            ops.gemm(dU_bf, self.w((li, "lin2_w")), ws["da"], b_major=1, epilogue=ops.EPI_DRELU if self.act == "relu" else ops.EPI_DGELU,
                     aux=buf["a"], **drelu)
            self._wgrad(ws["da"], buf["x1_bf"], (li, "lin1_w"), bias_key=(li, "lin1_b"))
            ops.gemm(ws["da"], self.w((li, "lin1_w")), dh, b_major=1)
            # ---- cross-attention: t2 = norm2(u2), u2 = t1 + dropout2(out_proj(attn(t1 + qpos, memory + pos, memory))) ----
            ops.layernorm_bwd(dU, buf["u2"], buf["mean2"], buf["rstd2"], self.f((li, "norm2_w")), dx=dU2, dx_bf16=dU_bf if fuse else None,
                              dgamma=self.gview((li, "norm2_w")), dbeta=self.gview((li, "norm2_b")),
                              dx_colsum=self.gview((li, "cout_b")) if fuse else None, dy_add=dh)
            if pd > 0:
                self._masked_operand(ws, dU2, dU_bf, li, 5, (li, "cout_b"))
            self._wgrad(dU_bf, buf["co"], (li, "cout_w"))
            ops.gemm(dU_bf, self.w((li, "cout_w")), dh, b_major=1)                                   # dO of the cross-attention
            ops.attention_bwd(buf["cq"], buf["ck"], buf["cv"], buf["co"], buf["clse"], dh, ws["dcq"], ws["dck"], ws["dcv"], ws["delta"],
                              B=N, H=H, S=Q, S_kv=S, tok_stride=N, batch_stride=1, key_padding_mask=kpm,
                              dropout=(pd, seed, site(li, 4)) if pd > 0 else None)
            cw, gcb = self.w((li, "cin_w")), self.gview((li, "cin_b"))
            self._wgrad(ws["dcq"], buf["t1q_bf"] if ws["has_qpos"] else buf["t1_bf"], (li, "cin_w"), rows=(0, D), bias_grad=gcb[:D])
            self._wgrad(ws["dck"], ws["mk_bf"], (li, "cin_w"), rows=(D, 2 * D), bias_grad=gcb[D:2 * D])
            self._wgrad(ws["dcv"], ws["m_bf"], (li, "cin_w"), rows=(2 * D, 3 * D), bias_grad=gcb[2 * D:])
            ops.gemm(ws["dcq"], cw[:D], dh, b_major=1)                                               # d(t1 + query_pos)
            ops.gemm(ws["dck"], cw[D:2 * D], ws["dmk"], b_major=1)                                   # d(memory + pos)
            ops.gemm(ws["dcv"], cw[2 * D:], ws["dmv"], b_major=1)                                    # d memory through the values
            ops.add3(dmem, ws["dmk"], ws["dmv"], dmem, dpos)                                         # d memory += ..; d pos += dmk
            ops.add3(dqpos, dh, None, dqpos, None)
            # ---- self-attention: t1 = norm1(u1), u1 = t0 + dropout1(out_proj(attn(t0 + qpos, t0 + qpos, t0))) ----
            ops.layernorm_bwd(dU2, buf["u1"], buf["mean1"], buf["rstd1"], self.f((li, "norm1_w")), dx=dU, dx_bf16=dU_bf if fuse else None,
                              dgamma=self.gview((li, "norm1_w")), dbeta=self.gview((li, "norm1_b")),
                              dx_colsum=self.gview((li, "out_b")) if fuse else None, dy_add=dh)
            if pd > 0:
                self._masked_operand(ws, dU, dU_bf, li, 0, (li, "out_b"))
            self._wgrad(dU_bf, buf["so"], (li, "out_w"))
            ops.gemm(dU_bf, self.w((li, "out_w")), dh, b_major=1)                                    # dO of the self-attention
            dsqk, dsv = ws["dsqk"], ws["dsv"]
            ops.attention_bwd(buf["sqk"][:, :D], buf["sqk"][:, D:], buf["sv"], buf["so"], buf["slse"], dh, dsqk[:, :D], dsqk[:, D:], dsv,
                              ws["delta"], B=N, H=H, S=Q, tok_stride=N, batch_stride=1, dropout=(pd, seed, site(li, 3)) if pd > 0 else None)
            in_w, gb = self.w((li, "in_w")), self.gview((li, "in_b"))
            self._wgrad(dsqk, buf["qk_bf"], (li, "in_w"), rows=(0, 2 * D), bias_grad=gb[:2 * D])
            self._wgrad(dsv, buf["x_bf"], (li, "in_w"), rows=(2 * D, 3 * D), bias_grad=gb[2 * D:])
            ops.gemm(dsqk, in_w[:2 * D], dh, b_major=1)                                              # d(t0 + query_pos) through q and k
            ops.gemm(dsv, in_w[2 * D:], dh2, b_major=1)                                              # d t0 through v
            ops.add3(dU, dh, dh2, dT, dqpos)                                                         # d t0 = dU + dh + dh2; d qpos += dh
            have_next = True
            if not self.return_intermediate:
                self._seg_done(seg + (1 if self.has_norm else 0))
            seg += 1
        if self.return_intermediate:     # the shared norm's gradient is complete only now: release the segments in order
            for i in range(L + 1):
                self._seg_done(i)
        return (dT.view(Q, N, D), dmem.view(S, N, D), dpos.view(S, N, D) if ws["has_pos"] else None,
                dqpos.view(Q, N, D) if ws["has_qpos"] else None)


class _DetrDecFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, kpm, tgt, memory, pos, query_pos, *params):
        engine.ensure_bound()

        def fresh():
            engine.bf16_fresh = False
        out, ws = engine.graphed(("dec_fwd", float(engine.p_drop)), (tgt.shape[0] + memory.shape[0]) * tgt.shape[1],
                                 [tgt, memory, kpm, pos, query_pos],
                                 lambda t_, m_, k_, p_, q_: engine.forward(t_, m_, k_, p_, q_, training=True), before_capture=fresh,
                                 busy=lambda res: res[1].get("leased"))
        ctx.engine, ctx.ws, ctx.n_params = engine, ws, len(params)
        ctx.lease = WorkspaceLease(ws)
        return out.clone()

    @staticmethod
    def backward(ctx, grad_out):
        eng, ws = ctx.engine, ctx.ws
        eng.prepare_grads()
        dt, dm, dp, dq = eng.graphed(("dec_bwd", id(ws), ws["p_drop"], ws["has_pos"], ws["has_qpos"], ws["kpm"] is not None),
                                     ws["M"] + ws["Ms"], [grad_out], lambda g_: eng.backward(ws, g_))
        need = ctx.needs_input_grad
        res = (None, None, dt.clone() if need[2] else None, dm.clone() if need[3] else None,
               dp.clone() if (dp is not None and need[4]) else None, dq.clone() if (dq is not None and need[5]) else None)
        ctx.lease.release()
        return res + (None,) * ctx.n_params


class TransformerDecoder(nn.Module):
    __getstate__ = getstate_without_engine

    def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=False):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(decoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm
        self.return_intermediate = return_intermediate
        self.__dict__["_engine"] = None

    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            l0 = self.layers[0]
            if self.return_intermediate and self.norm is None:
                raise TypeError("return_intermediate needs the decoder norm (transformer.py:85 calls self.norm unconditionally)")
            eng = DetrDecoderEngine(list(self.layers), self.norm, l0.d_model, l0.nhead, l0.dim_feedforward, l0.activation_name,
                                    eps=float(l0.norm1.eps), return_intermediate=self.return_intermediate, pre_norm=l0.normalize_before)
            self.__dict__["_engine"] = eng
        return eng

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None, pos=None,
                query_pos=None):
        if tgt_mask is not None or memory_mask is not None or tgt_key_padding_mask is not None:
            raise NotImplementedError("vitb200: tgt_mask / memory_mask / tgt_key_padding_mask are not supported; DETR passes None for all "
                                      "three (transformer.py:60)")
        eng = self._get_engine()
        eng.p_drop = float(self.layers[0].dropout_p) if self.training else 0.0
        kpm = None
        if memory_key_padding_mask is not None:
            kpm = memory_key_padding_mask.to(device=tgt.device, dtype=torch.uint8).contiguous()
        params = [p for _, p in eng._order]
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or tgt.requires_grad or memory.requires_grad
                                                  or (pos is not None and pos.requires_grad)
                                                  or (query_pos is not None and query_pos.requires_grad))
        if needs_grad:
            return _DetrDecFn.apply(eng, kpm, tgt, memory, pos, query_pos, *params)
        out, _ = eng.forward(tgt, memory, kpm, pos, query_pos, training=False)
        return out.clone()
