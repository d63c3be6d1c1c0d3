"""Drop-in for timm's ``VisionTransformerDistilled`` as the reference uses it (models/image_classification/deit.py:39-45,
65-70, 95-96, 162-166): same constructor keywords, ``set_distilled_training``, tuple output ``(head(x[:,0]),
head_dist(x[:,1]))`` in distilled training and their mean otherwise, timm state_dict key names.

timm is not importable in the build container, so the module structure below is restated from
timm's public semantics: ``cls_token, dist_token, pos_embed [1,N+2,D], patch_embed.proj, blocks.{i}.{norm1, attn.qkv,
attn.proj, norm2, mlp.fc1, mlp.fc2}, norm, head, head_dist``; LayerNorm eps 1e-6; erf-GELU; qkv_bias=True; token order
[cls, dist, patches].  The block arithmetic is identical to the ViT block (same kernels).
"""
import math

import torch
from torch import nn

from .engine import VitEngine, getstate_without_engine
from .vit import run_engine


class _PatchEmbed(nn.Module):
    def __init__(self, patch_size, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(3, embed_dim, kernel_size=patch_size, stride=patch_size)


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, mlp_hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, mlp_hidden)

    def roles(self):
        return {"ln1_w": self.norm1.weight, "ln1_b": self.norm1.bias, "qkv_w": self.attn.qkv.weight, "qkv_b": self.attn.qkv.bias,
                "proj_w": self.attn.proj.weight, "proj_b": self.attn.proj.bias, "ln2_w": self.norm2.weight, "ln2_b": self.norm2.bias,
                "fc1_w": self.mlp.fc1.weight, "fc1_b": self.mlp.fc1.bias, "fc2_w": self.mlp.fc2.weight, "fc2_b": self.mlp.fc2.bias}


class VisionTransformerDistilled(nn.Module):
    def __init__(self, img_size=224, patch_size=16, depth=12, num_heads=12, embed_dim=768, mlp_ratio=4.0, drop_rate=0.0,
                 attn_drop_rate=0.0, num_classes=1000, **kwargs):
        super().__init__()
        assert img_size % patch_size == 0, "Input shape indivisible by patch size!"
        self.img_size, self.patch_size, self.depth, self.num_heads = img_size, patch_size, depth, num_heads
        self.embed_dim = self.num_features = embed_dim
        self.mlp_ratio, self.drop_rate, self.attn_drop_rate, self.num_classes = mlp_ratio, drop_rate, attn_drop_rate, num_classes
        self.distilled_training = False
        n = (img_size // patch_size) ** 2
        self.patch_embed = _PatchEmbed(patch_size, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 2, embed_dim))
        self.blocks = nn.Sequential(*[_Block(embed_dim, int(embed_dim * mlp_ratio)) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Linear(embed_dim, num_classes)
        self.head_dist = nn.Linear(embed_dim, num_classes)
        # timm-style init: trunc_normal(.02) for tokens / pos_embed / Linear weights, zero biases, default LayerNorm
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.dist_token, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        self.__dict__["_engine"] = None

    __getstate__ = getstate_without_engine

    def set_distilled_training(self, enable=True):
        self.distilled_training = enable

    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            g = {"cls": self.cls_token, "dist": self.dist_token, "pos": self.pos_embed, "conv_w": self.patch_embed.proj.weight,
                 "conv_b": self.patch_embed.proj.bias, "lnf_w": self.norm.weight, "lnf_b": self.norm.bias,
                 "head_w": self.head.weight, "head_b": self.head.bias, "headd_w": self.head_dist.weight, "headd_b": self.head_dist.bias}
            eng = VitEngine(image_size=self.img_size, patch_size=self.patch_size, hidden_dim=self.embed_dim, num_heads=self.num_heads,
                            mlp_dim=int(self.embed_dim * self.mlp_ratio), num_layers=self.depth, num_classes=self.num_classes,
                            n_prefix=2, eps=1e-6, globals_=g, layers=[b.roles() for b in self.blocks])
            self.__dict__["_engine"] = eng
        return eng

    def forward_features(self, x):
        eng = self._get_engine()
        return run_engine(eng, x, "features", [p for _, p in eng._order], self.training, (self.drop_rate, self.attn_drop_rate))

    def forward(self, x):
        n, c, h, w = x.shape
        torch._assert(h == self.img_size and w == self.img_size, f"Input image size ({h}*{w}) doesn't match model ({self.img_size}).")
        eng = self._get_engine()
        out, out_dist = run_engine(eng, x, "logits", [p for _, p in eng._order], self.training, (self.drop_rate, self.attn_drop_rate))
        if self.distilled_training and self.training and not torch.jit.is_scripting():
            return out, out_dist
        return (out + out_dist) / 2
