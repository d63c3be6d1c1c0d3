"""Drop-in ViT classifier with the reference's constructor, forward() signature and state_dict keys, executed by
the hand-written sm_100a kernels of libvitb200.so.

Mirrors models/image_classification/vanilla_vit.py: MLP :22-44, MLPBlock :47-56, EncoderBlock :59-83,
Encoder :86-106, ViT :109-215.  The sub-modules are constructed in the reference's order (conv_proj, class_token,
pos_embedding, per block: ln_1, nn.MultiheadAttention, ln_2, mlp Linear/Linear, final ln, head) so that the same
``torch.manual_seed`` gives a bit-identical initial ``state_dict``; they are parameter containers only — the
arithmetic of ``forward`` never goes through them.
"""
import math
from collections import OrderedDict
from functools import partial
from typing import Callable, List, Optional

import torch
from torch import nn

from .engine import VitEngine, WorkspaceLease, getstate_without_engine


class MLP(torch.nn.Sequential):
    """Parameter container with the layer indices of the reference MLP (Linear 0, act 1, Dropout 2, Linear 3, Dropout 4)."""

    def __init__(self, in_channels: int, hidden_channels: List[int], norm_layer: Optional[Callable[..., torch.nn.Module]] = None,
                 activation_layer: Optional[Callable[..., torch.nn.Module]] = torch.nn.ReLU, inplace: Optional[bool] = None,
                 bias: bool = True, dropout: float = 0.0):
        params = {} if inplace is None else {"inplace": inplace}
        layers = []
        in_dim = in_channels
        for hidden_dim in hidden_channels[:-1]:
            layers.append(torch.nn.Linear(in_dim, hidden_dim, bias=bias))
            if norm_layer is not None:
                layers.append(norm_layer(hidden_dim))
            layers.append(activation_layer(**params))
            layers.append(torch.nn.Dropout(dropout, **params))
            in_dim = hidden_dim
        layers.append(torch.nn.Linear(in_dim, hidden_channels[-1], bias=bias))
        layers.append(torch.nn.Dropout(dropout, **params))
        super().__init__(*layers)


class MLPBlock(MLP):
    def __init__(self, in_dim: int, mlp_dim: int, dropout: float):
        super().__init__(in_dim, [mlp_dim, in_dim], activation_layer=nn.GELU, inplace=None, dropout=dropout)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.normal_(m.bias, std=1e-6)


def _norm_eps(norm_module):
    return float(getattr(norm_module, "eps", 1e-6))


def _deepcopy_without_engine(module, memo):
    """copy.deepcopy of a module that caches an engine (flat buffers, workspaces, CUDA graphs): the copy starts without one."""
    import copy
    eng = module.__dict__.pop("_engine", None)
    try:
        new = module.__class__.__new__(module.__class__)
        memo[id(module)] = new
        for k, v in module.__dict__.items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        new.__dict__["_engine"] = None
    finally:
        module.__dict__["_engine"] = eng
    return new


class EncoderBlock(nn.Module):
    """Parameter container for one pre-norm block (vanilla_vit.py:59-71); executed by the enclosing Encoder/ViT."""

    def __init__(self, num_heads: int, hidden_dim: int, mlp_dim: int, dropout: float, attention_dropout: float,
                 norm_layer: Callable[..., torch.nn.Module] = partial(nn.LayerNorm, eps=1e-6)):
        super().__init__()
        self.num_heads = num_heads
        self.ln_1 = norm_layer(hidden_dim)
        self.self_attention = nn.MultiheadAttention(hidden_dim, num_heads, dropout=attention_dropout, batch_first=True)
        self.dropout = nn.Dropout(dropout)
        self.ln_2 = norm_layer(hidden_dim)
        self.mlp = MLPBlock(hidden_dim, mlp_dim, dropout)

    def roles(self):
        return {"ln1_w": self.ln_1.weight, "ln1_b": self.ln_1.bias,
                "qkv_w": self.self_attention.in_proj_weight, "qkv_b": self.self_attention.in_proj_bias,
                "proj_w": self.self_attention.out_proj.weight, "proj_b": self.self_attention.out_proj.bias,
                "ln2_w": self.ln_2.weight, "ln2_b": self.ln_2.bias,
                "fc1_w": self.mlp[0].weight, "fc1_b": self.mlp[0].bias, "fc2_w": self.mlp[3].weight, "fc2_b": self.mlp[3].bias}

    # Stand-alone use (vanilla_vit.py:73-83: x = dropout(MHA(ln_1(input))) + input; return x + mlp(ln_2(x))): a one-layer block-mode
    # engine over this module's own parameters, one per sequence length.  Inside an Encoder / ViT the parent's engine runs the block
    # (this forward is then never called); using both alternately works — each engine re-binds the parameters when it finds them moved.
    def _get_engine(self, seq_length):
        engines = self.__dict__.get("_engine")
        if engines is None:
            engines = self.__dict__["_engine"] = {}
        eng = engines.get(seq_length)
        if eng is None:
            eng = VitEngine(image_size=4, patch_size=4, hidden_dim=self.ln_1.weight.shape[0], num_heads=self.num_heads,
                            mlp_dim=self.mlp[0].out_features, num_layers=1, num_classes=1, n_prefix=0, eps=_norm_eps(self.ln_1),
                            globals_={}, layers=[self.roles()], seq_length=seq_length, block_mode=True)
            eng.tokens_want = "block"
            engines[seq_length] = eng
        return eng

    def forward(self, input: torch.Tensor):
        torch._assert(input.dim() == 3, f"Expected (batch_size, seq_length, hidden_dim) got {input.shape}")
        torch._assert(input.shape[2] == self.ln_1.weight.shape[0], f"Expected hidden_dim {self.ln_1.weight.shape[0]} got {input.shape}")
        if not input.is_cuda:
            raise RuntimeError("vitb200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        eng = self._get_engine(int(input.shape[1]))
        eng.p_drop, eng.p_attn = (float(self.dropout.p), float(self.self_attention.dropout)) if self.training else (0.0, 0.0)
        params = [p for _, p in eng._order]
        x = input.contiguous().float()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            return _TokensFn.apply(eng, x, *params)
        outs, _ = eng.forward(x, training=(eng.p_drop > 0 or eng.p_attn > 0), want="block")
        return outs[0].clone()

    __getstate__ = getstate_without_engine

    def __deepcopy__(self, memo):
        return _deepcopy_without_engine(self, memo)


class Encoder(nn.Module):
    def __init__(self, seq_length: int, num_layers: int, num_heads: int, hidden_dim: int, mlp_dim: int, dropout: float,
                 attention_dropout: float, norm_layer: Callable[..., torch.nn.Module] = partial(nn.LayerNorm, eps=1e-6)):
        super().__init__()
        self.pos_embedding = nn.Parameter(torch.empty(1, seq_length, hidden_dim).normal_(std=0.02))
        self.dropout = nn.Dropout(dropout)
        layers: "OrderedDict[str, nn.Module]" = OrderedDict()
        for i in range(num_layers):
            layers[f"encoder_layer_{i}"] = EncoderBlock(num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, norm_layer)
        self.layers = nn.Sequential(layers)
        self.ln = norm_layer(hidden_dim)
        self.__dict__["_engine"] = None

    # Stand-alone use (the reference's Encoder is a public class: vanilla_vit.py:88-106; T2T_ViT carries an identical copy,
    # t2t_vit.py:89-110): input [B, S, D] -> input + pos_embedding -> dropout -> L blocks -> final LayerNorm, on the same kernels
    # through a tokens-mode engine over this module's own parameters.  Inside a ViT the parent's engine runs the stack instead.
    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            blocks = list(self.layers)
            b0 = blocks[0]
            D = self.pos_embedding.shape[-1]
            eng = VitEngine(image_size=4, patch_size=4, hidden_dim=D, num_heads=b0.num_heads, mlp_dim=b0.mlp[0].out_features,
                            num_layers=len(blocks), num_classes=1, n_prefix=0, eps=_norm_eps(self.ln),
                            globals_={"pos": self.pos_embedding, "lnf_w": self.ln.weight, "lnf_b": self.ln.bias},
                            layers=[b.roles() for b in blocks], seq_length=self.pos_embedding.shape[1])
            self.__dict__["_engine"] = eng
        return eng

    def forward(self, input: torch.Tensor):
        torch._assert(input.dim() == 3, f"Expected (batch_size, seq_length, hidden_dim) got {input.shape}")
        eng = self._get_engine()
        torch._assert(input.shape[1] == eng.S and input.shape[2] == eng.D, f"Expected (batch_size, {eng.S}, {eng.D}) got {input.shape}")
        b0 = self.layers[0]
        eng.p_drop, eng.p_attn = (float(self.dropout.p), float(b0.self_attention.dropout)) if self.training else (0.0, 0.0)
        params = [p for _, p in eng._order]
        x = input.contiguous().float()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            return _TokensFn.apply(eng, x, *params)
        outs, _ = eng.forward(x, training=(eng.p_drop > 0 or eng.p_attn > 0), want="features")
        return outs[0].clone()

    __getstate__ = getstate_without_engine

    def __deepcopy__(self, memo):
        return _deepcopy_without_engine(self, memo)


class _EncoderFn(torch.autograd.Function):
    """One autograd node for the whole model: forward runs the kernel sequence and keeps the workspace, backward
    writes parameter gradients straight into the flat gradient buffer that every p.grad is a view of."""

    @staticmethod
    def forward(ctx, engine, want, images, *params):
        if images.dtype != torch.float32 or not images.is_contiguous():
            images = images.contiguous().float()
        outs, ws = engine.forward_train(images, want=want)
        ctx.engine, ctx.ws, ctx.want, ctx.n_params = engine, ws, want, len(params)
        ctx.lease = WorkspaceLease(ws)       # the saved activations live in ws until the backward has run
        ctx.set_materialize_grads(False)
        res = tuple(o.clone() for o in outs)
        return res if len(res) > 1 else res[0]

    @staticmethod
    def backward(ctx, *grads):
        dimg = ctx.engine.backward_train(ctx.ws, list(grads), want=ctx.want, input_grad=ctx.needs_input_grad[2])
        dimg = dimg.clone() if dimg is not None else None
        ctx.lease.release()
        return (None, None, dimg) + (None,) * ctx.n_params


class _TokensFn(torch.autograd.Function):
    """Stand-alone Encoder: like _EncoderFn, but the input tokens receive a gradient too."""

    @staticmethod
    def forward(ctx, engine, tokens, *params):
        want = getattr(engine, "tokens_want", "features")   # "block": a bare EncoderBlock (no position embedding, no final norm)
        outs, ws = engine.forward(tokens, training=True, want=want)
        ctx.engine, ctx.ws, ctx.n_params, ctx.want = engine, ws, len(params), want
        ctx.lease = WorkspaceLease(ws)
        return outs[0].clone()

    @staticmethod
    def backward(ctx, grad):
        d = ctx.engine.backward(ctx.ws, [grad], want=ctx.want).clone()
        ctx.lease.release()
        return (None, d) + (None,) * ctx.n_params


def run_engine(engine, images, want, params, module_training, dropout_ps):
    needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or images.requires_grad)
    # dropout is active in train() mode only, as in nn.Dropout / nn.MultiheadAttention (vanilla_vit.py:38,42,67-68,94)
    engine.p_drop, engine.p_attn = (float(dropout_ps[0]), float(dropout_ps[1])) if module_training else (0.0, 0.0)
    if needs_grad:
        return _EncoderFn.apply(engine, want, images, *params)
    if engine.p_drop > 0 or engine.p_attn > 0:   # train() under no_grad: dropout still applies; use the training workspaces
        outs, _ = engine.forward(images, training=True, want=want)
    else:
        outs = engine.forward_inference(images, want=want)
    res = tuple(o.clone() for o in outs)
    return res if len(res) > 1 else res[0]


class ViT(nn.Module):
    """Same constructor and public attributes as the reference ViT (vanilla_vit.py:109-151)."""

    def __init__(self, image_size, patch_size, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, num_classes,
                 norm_layer: Callable[..., torch.nn.Module] = partial(nn.LayerNorm, eps=1e-6), *args, **kwargs):
        super().__init__(*args, **kwargs)
        if not hasattr(self, "device"):  # BaseTransformer.__init__ sets it (base.py:16-21) when used through dropin.install
            self.device = "cuda" if torch.cuda.is_available() else "cpu"
        torch._assert(image_size % patch_size == 0, "Input shape indivisible by patch size!")
        self.image_size = image_size
        self.patch_size = patch_size
        self.hidden_dim = hidden_dim
        self.mlp_dim = mlp_dim
        self.attention_dropout = attention_dropout
        self.dropout = dropout
        self.num_classes = num_classes
        self.norm_layer = norm_layer
        self.num_patches = (image_size // patch_size) ** 2
        self.num_layers = num_layers
        self.num_heads = num_heads

        self.conv_proj = nn.Conv2d(in_channels=3, out_channels=hidden_dim, kernel_size=patch_size, stride=patch_size)
        seq_length = (image_size // patch_size) ** 2
        self.class_token = nn.Parameter(torch.zeros(1, 1, hidden_dim))
        seq_length += 1
        self.encoder = Encoder(seq_length=seq_length, num_layers=num_layers, num_heads=num_heads, hidden_dim=hidden_dim,
                               mlp_dim=mlp_dim, dropout=dropout, attention_dropout=attention_dropout, norm_layer=norm_layer)
        heads_layers: "OrderedDict[str, nn.Module]" = OrderedDict()
        heads_layers["head"] = nn.Linear(hidden_dim, num_classes)
        self.heads = nn.Sequential(heads_layers)

        fan_in = self.conv_proj.in_channels * self.conv_proj.kernel_size[0] * self.conv_proj.kernel_size[1]
        nn.init.trunc_normal_(self.conv_proj.weight, std=math.sqrt(1 / fan_in))
        if self.conv_proj.bias is not None:
            nn.init.zeros_(self.conv_proj.bias)
        nn.init.zeros_(self.heads.head.weight)
        nn.init.zeros_(self.heads.head.bias)
        self.__dict__["_engine"] = None

    # -- engine plumbing ------------------------------------------------------------------------------------------
    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            blocks = list(self.encoder.layers)
            globals_ = {"cls": self.class_token, "pos": self.encoder.pos_embedding, "conv_w": self.conv_proj.weight,
                        "conv_b": self.conv_proj.bias, "lnf_w": self.encoder.ln.weight, "lnf_b": self.encoder.ln.bias,
                        "head_w": self.heads.head.weight, "head_b": self.heads.head.bias}
            eng = VitEngine(image_size=self.image_size, patch_size=self.patch_size, hidden_dim=self.hidden_dim,
                            num_heads=self.num_heads, mlp_dim=self.mlp_dim, num_layers=self.num_layers,
                            num_classes=self.num_classes, n_prefix=1, eps=_norm_eps(self.encoder.ln), globals_=globals_,
                            layers=[b.roles() for b in blocks])
            self.__dict__["_engine"] = eng
        return eng

    __getstate__ = getstate_without_engine

    def __deepcopy__(self, memo):
        import copy
        eng = self.__dict__.pop("_engine", None)
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                new.__dict__[k] = copy.deepcopy(v, memo)
            new.__dict__["_engine"] = None
        finally:
            self.__dict__["_engine"] = eng
        return new

    def _check_images(self, images):
        n, c, h, w = images.shape
        torch._assert(h == self.image_size, f"Wrong image height! Expected {self.image_size} but got {h}!")
        torch._assert(w == self.image_size, f"Wrong image width! Expected {self.image_size} but got {w}!")
        torch._assert(c == 3, f"Expected 3 input channels but got {c}!")

    def _run(self, images, want):
        self._check_images(images)
        eng = self._get_engine()
        params = [p for _, p in eng._order]
        return run_engine(eng, images, want, params, self.training, (self.dropout, self.attention_dropout))

    def forward_features(self, images: torch.Tensor):
        """[B,3,H,W] -> [B,S,D] after the final LayerNorm (vanilla_vit.py:186-207)."""
        return self._run(images, "features")

    def forward(self, images: torch.Tensor):
        """[B,3,H,W] -> logits [B,num_classes] (vanilla_vit.py:209-215)."""
        return self._run(images, "logits")
