"""Drop-in CPE-ViT / CPVT / CPVT-GAP classifiers: the ViT encoder with conditional positional encodings —
a depthwise 3x3 convolution over the patch-token grid — on the same hand-written sm_100a kernels plus ``csrc/dwconv.cu``.

Mirrors models/image_classification/cpe_vit.py (``CPEViT``: CPE after the class-token cat, then the ordinary Encoder with its learned
``pos_embedding``: :16-30, :100-115, :143, :197-198), cpvt.py (``CPVT``: no learned position embedding, a PEG at the end of EVERY
encoder block as written at :93-96 — ``x = x + y; x = peg(x); return x + y``) and cpvt_gap.py (``CPVTGAP``: the same forward; its
``gap`` pooling module is constructed but never called, cpvt_gap.py:149,208-214).  Constructors, ``forward`` / ``forward_features``
signatures, public attributes and ``state_dict`` keys are the reference's; sub-modules are created in the reference's order, so the
same ``torch.manual_seed`` gives a bit-identical initial ``state_dict``.  The modules are parameter containers: the arithmetic runs
in ``VitEngine``.
"""
import math
from collections import OrderedDict
from functools import partial
from typing import Callable

import torch
from torch import nn

from .engine import VitEngine, getstate_without_engine
from .vit import EncoderBlock as _VitBlock
from .vit import _norm_eps, run_engine


class ConditionalPositionalEncoding(nn.Module):
    """Parameter container for nn.Conv2d(d_model, d_model, 3, padding=1, groups=d_model) (cpe_vit.py:16-19)."""

    def __init__(self, d_model, kernel_size=3):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("vitb200: the PEG kernel is a 3x3 depthwise convolution (the reference's only use)")
        self.conv = nn.Conv2d(d_model, d_model, kernel_size=kernel_size, padding=kernel_size // 2, groups=d_model)

    def forward(self, images):
        raise RuntimeError("vitb200.ConditionalPositionalEncoding is executed by its parent model (fused path)")


class EncoderBlock(_VitBlock):
    """cpvt.py:67-97: the ViT block followed by a PEG."""

    def __init__(self, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, norm_layer=partial(nn.LayerNorm, eps=1e-6)):
        super().__init__(num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, norm_layer)
        self.peg = ConditionalPositionalEncoding(hidden_dim)

    def roles(self):
        r = super().roles()
        r.update({"peg_w": self.peg.conv.weight, "peg_b": self.peg.conv.bias})
        return r


class Encoder(nn.Module):
    """cpvt.py:99-113 (no learned position embedding) / cpe_vit.py:95-115 (with one)."""

    def __init__(self, seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout,
                 norm_layer: Callable[..., nn.Module] = partial(nn.LayerNorm, eps=1e-6), *, learned_pos, peg_blocks):
        super().__init__()
        if learned_pos:
            self.pos_embedding = nn.Parameter(torch.empty(1, seq_length, hidden_dim).normal_(std=0.02))
        self.dropout = nn.Dropout(dropout)
        block = EncoderBlock if peg_blocks else _VitBlock
        layers: "OrderedDict[str, nn.Module]" = OrderedDict()
        for i in range(num_layers):
            layers[f"encoder_layer_{i}"] = block(num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, norm_layer)
        self.layers = nn.Sequential(layers)
        self.ln = norm_layer(hidden_dim)

    def forward(self, input):
        raise RuntimeError("vitb200 CPVT/CPE-ViT Encoder is executed by its parent model (fused path)")


class _CpeBase(nn.Module):
    _LEARNED_POS = False
    _PEG_BLOCKS = False

    def __init__(self, image_size, patch_size, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, num_classes,
                 norm_layer: Callable[..., nn.Module] = partial(nn.LayerNorm, eps=1e-6), *args, **kwargs):
        super().__init__(*args, **kwargs)
        torch._assert(image_size % patch_size == 0, "Input shape indivisible by patch size!")
        g = image_size // patch_size
        self.image_size, self.patch_size, self.hidden_dim, self.mlp_dim = image_size, patch_size, hidden_dim, mlp_dim
        self.attention_dropout, self.dropout, self.num_classes, self.norm_layer = attention_dropout, dropout, num_classes, norm_layer
        self.num_patches, self.num_layers, self.num_heads = g * g, num_layers, num_heads
        self.conv_proj = nn.Conv2d(in_channels=3, out_channels=hidden_dim, kernel_size=patch_size, stride=patch_size)
        seq_length = g * g + 1
        self.class_token = nn.Parameter(torch.zeros(1, 1, hidden_dim))
        self.pos_embedding = ConditionalPositionalEncoding(hidden_dim)   # CPE (cpe_vit.py:143, cpvt.py:144)
        self.encoder = Encoder(seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout, norm_layer,
                               learned_pos=self._LEARNED_POS, peg_blocks=self._PEG_BLOCKS)
        self._extra_modules()
        heads_layers: "OrderedDict[str, nn.Module]" = OrderedDict()
        heads_layers["head"] = nn.Linear(hidden_dim, num_classes)
        self.heads = nn.Sequential(heads_layers)
        fan_in = self.conv_proj.in_channels * self.conv_proj.kernel_size[0] * self.conv_proj.kernel_size[1]
        nn.init.trunc_normal_(self.conv_proj.weight, std=math.sqrt(1 / fan_in))
        if self.conv_proj.bias is not None:
            nn.init.zeros_(self.conv_proj.bias)
        nn.init.zeros_(self.heads.head.weight)
        nn.init.zeros_(self.heads.head.bias)
        self.device = "cuda"                                             # cpe_vit.py:164, cpvt.py:165
        self.__dict__["_engine"] = None

    def _extra_modules(self):
        pass

    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            blocks = list(self.encoder.layers)
            g = {"cls": self.class_token, "conv_w": self.conv_proj.weight, "conv_b": self.conv_proj.bias,
                 "cpe_w": self.pos_embedding.conv.weight, "cpe_b": self.pos_embedding.conv.bias,
                 "lnf_w": self.encoder.ln.weight, "lnf_b": self.encoder.ln.bias, "head_w": self.heads.head.weight,
                 "head_b": self.heads.head.bias}
            if self._LEARNED_POS:
                g["pos"] = self.encoder.pos_embedding
            eng = VitEngine(image_size=self.image_size, patch_size=self.patch_size, hidden_dim=self.hidden_dim, num_heads=self.num_heads,
                            mlp_dim=self.mlp_dim, num_layers=self.num_layers, num_classes=self.num_classes, n_prefix=1,
                            eps=_norm_eps(self.encoder.ln), globals_=g, layers=[b.roles() for b in blocks])
            self.__dict__["_engine"] = eng
        return eng

    __getstate__ = getstate_without_engine

    def __deepcopy__(self, memo):
        import copy
        eng = self.__dict__.pop("_engine", None)
        try:
            new = self.__class__.__new__(self.__class__)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                new.__dict__[k] = copy.deepcopy(v, memo)
            new.__dict__["_engine"] = None
        finally:
            self.__dict__["_engine"] = eng
        return new

    def _run(self, images, want):
        n, c, h, w = images.shape
        torch._assert(h == self.image_size, f"Wrong image height! Expected {self.image_size} but got {h}!")
        torch._assert(w == self.image_size, f"Wrong image width! Expected {self.image_size} but got {w}!")
        eng = self._get_engine()
        return run_engine(eng, images, want, [p for _, p in eng._order], self.training, (self.dropout, self.attention_dropout))

    def forward_features(self, images: torch.Tensor):
        return self._run(images, "features")

    def forward(self, images: torch.Tensor):
        return self._run(images, "logits")


class CPEViT(_CpeBase):
    """models/image_classification/cpe_vit.py:117-212."""
    _LEARNED_POS = True
    _PEG_BLOCKS = False


class CPVT(_CpeBase):
    """models/image_classification/cpvt.py:118-213."""
    _LEARNED_POS = False
    _PEG_BLOCKS = True


class CPVTGAP(CPVT):
    """models/image_classification/cpvt_gap.py:118-214 — the forward is CPVT's; ``gap`` exists for key/attribute parity only."""

    def _extra_modules(self):
        self.gap = nn.AdaptiveAvgPool1d(1)
