"""CPU oracle: a functional PyTorch-fp32 restatement of the reference's encoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (vitb200/) may import this module; it is used by
tests/, by __graft_entry__.smoke() as the checker, and by bench.py for the `cpu_baseline` / `--impl reference`
legs.  It computes with the same ATen CPU operators the reference modules dispatch to (F.linear, F.layer_norm,
F.gelu, F.scaled_dot_product_attention, F.conv2d), driven by a plain ``state_dict`` with the reference's keys,
so that its results and its timing are those of the reference's own CPU path.

Parity status: PINNED against the reference itself.  tools/make_golden.py imports the unmodified modules from
/root/reference (in the build container), runs them on seeded inputs/weights and stores the results under
tests/golden/; tests/test_oracle.py checks this restatement against those fixtures (and, when /root/reference
is present, against the live reference).  Exception: the DeiT *model class* lives in timm, which is absent
 — its block arithmetic is the pinned ViT arithmetic, only key names / token order are restated
from timm's public semantics ("parity unpinned" for those names).  The DETR decoder layer cannot run as written
(transformer.py:122 registers ``multi_head_attn``, :148 calls ``self.multihead_attn``); its restatement is pinned against the live
reference run with exactly that one alias added (tools/make_golden.py::decoder_case).

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------------
# ViT (models/image_classification/vanilla_vit.py)
# ------------------------------------------------------------------------------------------------------------
def mha_batch_first(x, in_w, in_b, out_w, out_b, num_heads, attn_drop=None):
    """nn.MultiheadAttention(D, H, dropout=p_attn, batch_first=True)(x, x, x, need_weights=False)  — vanilla_vit.py:67,77.

    Packed in-projection (torch/nn/functional.py:5835-5847), per-head softmax(QK^T/sqrt(hd))V via SDPA
    (functional.py:6676-6682), out-projection (functional.py:6690).  ``attn_drop``: optional callable applied to the
    [B,H,S,S] attention probabilities (dropout with an explicit mask; SDPA's dropout_p acts at the same place).
    """
    B, S, D = x.shape
    hd = D // num_heads
    qkv = F.linear(x, in_w, in_b)
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, S, num_heads, hd).transpose(1, 2)
    k = k.view(B, S, num_heads, hd).transpose(1, 2)
    v = v.view(B, S, num_heads, hd).transpose(1, 2)
    if attn_drop is None:
        o = F.scaled_dot_product_attention(q, k, v)
    else:
        o = attn_drop(torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)) @ v
    o = o.transpose(1, 2).reshape(B, S, D)
    return F.linear(o, out_w, out_b)


class ExplicitDropout:
    """Dropout with caller-supplied keep masks (0/1 tensors): y = keep * x / (1 - p) — the arithmetic of nn.Dropout in train()
    mode (vanilla_vit.py:38,42,68,94) with the random stream made an input, so that the GPU kernels' masks can be replayed.
    ``masks[(layer, site)]``: site 0 = out-proj output, 1 = after GELU, 2 = MLP output, 3 = attention probabilities;
    ``masks["embed"]`` = Encoder.dropout."""

    def __init__(self, masks, p_hidden, p_attn):
        self.masks, self.p_hidden, self.p_attn = masks, p_hidden, p_attn

    def __call__(self, key, x):
        p = self.p_attn if (isinstance(key, tuple) and key[1] in (3, 4)) else self.p_hidden   # 3 / 4: attention weights (self / cross)
        if p == 0 or key not in self.masks:
            return x
        return x * self.masks[key].reshape(x.shape).to(x.dtype) / (1.0 - p)


def encoder_block(x, sd, prefix, num_heads, eps, drop=None, layer=0):
    """EncoderBlock.forward — vanilla_vit.py:73-83 (pre-norm); ``drop``: None (p = 0) or an ExplicitDropout."""
    dz = (lambda key, t: t) if drop is None else drop
    h = F.layer_norm(x, (x.shape[-1],), sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"], eps)
    a = mha_batch_first(h, sd[prefix + "self_attention.in_proj_weight"], sd[prefix + "self_attention.in_proj_bias"],
                        sd[prefix + "self_attention.out_proj.weight"], sd[prefix + "self_attention.out_proj.bias"], num_heads,
                        attn_drop=None if (drop is None or drop.p_attn == 0) else (lambda P: drop((layer, 3), P)))
    x = dz((layer, 0), a) + x                                                      # :78-79
    y = F.layer_norm(x, (x.shape[-1],), sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"], eps)
    y = F.linear(y, sd[prefix + "mlp.0.weight"], sd[prefix + "mlp.0.bias"])       # MLPBlock: vanilla_vit.py:33-34
    y = dz((layer, 1), F.gelu(y))                                                  # nn.GELU() (erf) :50, mlp.2 dropout :38
    y = F.linear(y, sd[prefix + "mlp.3.weight"], sd[prefix + "mlp.3.bias"])       # :41
    return x + dz((layer, 2), y)                                                   # mlp.4 dropout :42, residual :83


def vit_forward_features(sd, images, *, patch_size, num_layers, num_heads, eps=1e-6, drop=None):
    """ViT.forward_features + Encoder.forward — vanilla_vit.py:186-207, :102-106."""
    n = images.shape[0]
    D = sd["class_token"].shape[-1]
    x = F.conv2d(images, sd["conv_proj.weight"], sd["conv_proj.bias"], stride=patch_size)   # :196
    x = x.reshape(n, D, -1).permute(0, 2, 1)                                                   # :197-198
    x = torch.cat([sd["class_token"].expand(n, -1, -1), x], dim=1)                             # :202-203
    x = x + sd["encoder.pos_embedding"]                                                        # :104
    if drop is not None:
        x = drop("embed", x)                                                                   # Encoder.dropout :104
    for i in range(num_layers):
        x = encoder_block(x, sd, f"encoder.layers.encoder_layer_{i}.", num_heads, eps, drop=drop, layer=i)
    return F.layer_norm(x, (D,), sd["encoder.ln.weight"], sd["encoder.ln.bias"], eps)         # :106


def vit_forward(sd, images, **cfg):
    """ViT.forward — vanilla_vit.py:209-215."""
    x = vit_forward_features(sd, images, **cfg)
    return F.linear(x[:, 0], sd["heads.head.weight"], sd["heads.head.bias"])


def vit_param_shapes(image_size, patch_size, num_layers, num_heads, hidden_dim, mlp_dim, num_classes):
    """state_dict keys/shapes created by ViT.__init__ — vanilla_vit.py:109-151 (== torchvision vit_* keys)."""
    D, Fd = hidden_dim, mlp_dim
    S = (image_size // patch_size) ** 2 + 1
    shapes = {"class_token": (1, 1, D), "conv_proj.weight": (D, 3, patch_size, patch_size), "conv_proj.bias": (D,),
              "encoder.pos_embedding": (1, S, D)}
    for i in range(num_layers):
        p = f"encoder.layers.encoder_layer_{i}."
        shapes.update({p + "ln_1.weight": (D,), p + "ln_1.bias": (D,),
                       p + "self_attention.in_proj_weight": (3 * D, D), p + "self_attention.in_proj_bias": (3 * D,),
                       p + "self_attention.out_proj.weight": (D, D), p + "self_attention.out_proj.bias": (D,),
                       p + "ln_2.weight": (D,), p + "ln_2.bias": (D,),
                       p + "mlp.0.weight": (Fd, D), p + "mlp.0.bias": (Fd,),
                       p + "mlp.3.weight": (D, Fd), p + "mlp.3.bias": (D,)})
    shapes.update({"encoder.ln.weight": (D,), "encoder.ln.bias": (D,),
                   "heads.head.weight": (num_classes, D), "heads.head.bias": (num_classes,)})
    return shapes


# ------------------------------------------------------------------------------------------------------------
# CPE-ViT / CPVT (conditional positional encodings) — cpe_vit.py, cpvt.py, cpvt_gap.py
# ------------------------------------------------------------------------------------------------------------
def cond_pos_encoding(x, w, b):
    """ConditionalPositionalEncoding.forward — cpe_vit.py:21-30 == cpvt.py:21-30: depthwise 3x3 conv over the patch-token grid,
    class token re-attached."""
    B, S, D = x.shape
    cls, t = x[:, :1, :], x[:, 1:, :]
    G = int(math.sqrt(S - 1))
    assert G * G == S - 1
    t = t.transpose(1, 2).reshape(B, D, G, G)
    t = F.conv2d(t, w, b, padding=1, groups=D)
    t = t.reshape(B, D, S - 1).transpose(1, 2)
    return torch.cat((cls, t), dim=1)


def cpe_forward_features(sd, images, *, patch_size, num_layers, num_heads, peg_blocks, eps=1e-6, drop=None):
    """CPEViT.forward_features (cpe_vit.py:181-202; Encoder :110-115 adds the learned pos_embedding) when ``peg_blocks`` is False,
    CPVT.forward_features (cpvt.py:183-204; Encoder :111-113 has none; EncoderBlock :82-97 ends with
    ``x = x + y; x = peg(x); return x + y``) when True."""
    n = images.shape[0]
    D = sd["class_token"].shape[-1]
    dz = (lambda key, t: t) if drop is None else drop
    x = F.conv2d(images, sd["conv_proj.weight"], sd["conv_proj.bias"], stride=patch_size)
    x = x.reshape(n, D, -1).permute(0, 2, 1)
    x = torch.cat([sd["class_token"].expand(n, -1, -1), x], dim=1)
    x = cond_pos_encoding(x, sd["pos_embedding.conv.weight"], sd["pos_embedding.conv.bias"])
    if not peg_blocks:
        x = x + sd["encoder.pos_embedding"]
    x = dz("embed", x)
    for i in range(num_layers):
        p = f"encoder.layers.encoder_layer_{i}."
        if not peg_blocks:
            x = encoder_block(x, sd, p, num_heads, eps, drop=drop, layer=i)
            continue
        h = F.layer_norm(x, (D,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps)
        a = mha_batch_first(h, sd[p + "self_attention.in_proj_weight"], sd[p + "self_attention.in_proj_bias"],
                            sd[p + "self_attention.out_proj.weight"], sd[p + "self_attention.out_proj.bias"], num_heads,
                            attn_drop=None if (drop is None or drop.p_attn == 0) else (lambda P, i=i: drop((i, 3), P)))
        x = dz((i, 0), a) + x
        y = F.layer_norm(x, (D,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], eps)
        y = dz((i, 1), F.gelu(F.linear(y, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])))
        y = dz((i, 2), F.linear(y, sd[p + "mlp.3.weight"], sd[p + "mlp.3.bias"]))
        x = x + y                                                                   # cpvt.py:93
        x = cond_pos_encoding(x, sd[p + "peg.conv.weight"], sd[p + "peg.conv.bias"])  # :94
        x = x + y                                                                   # :96
    return F.layer_norm(x, (D,), sd["encoder.ln.weight"], sd["encoder.ln.bias"], eps)


def cpe_forward(sd, images, **cfg):
    x = cpe_forward_features(sd, images, **cfg)
    return F.linear(x[:, 0], sd["heads.head.weight"], sd["heads.head.bias"])


def cpe_param_shapes(image_size, patch_size, num_layers, num_heads, hidden_dim, mlp_dim, num_classes, *, peg_blocks):
    """state_dict keys/shapes of CPEViT (cpe_vit.py:117-166) or CPVT / CPVTGAP (cpvt.py:118-167)."""
    sh = vit_param_shapes(image_size, patch_size, num_layers, num_heads, hidden_dim, mlp_dim, num_classes)
    D = hidden_dim
    sh["pos_embedding.conv.weight"], sh["pos_embedding.conv.bias"] = (D, 1, 3, 3), (D,)
    if peg_blocks:
        del sh["encoder.pos_embedding"]
        for i in range(num_layers):
            p = f"encoder.layers.encoder_layer_{i}.peg.conv."
            sh[p + "weight"], sh[p + "bias"] = (D, 1, 3, 3), (D,)
    return sh


# ------------------------------------------------------------------------------------------------------------
# DeiT-style distilled ViT (timm VisionTransformerDistilled as used at deit.py:39-45,65,95-96)
# ------------------------------------------------------------------------------------------------------------
def deit_param_shapes(img_size, patch_size, depth, num_heads, embed_dim, mlp_ratio, num_classes):
    D = embed_dim
    Fd = int(D * mlp_ratio)
    N = (img_size // patch_size) ** 2
    shapes = {"cls_token": (1, 1, D), "dist_token": (1, 1, D), "pos_embed": (1, N + 2, D),
              "patch_embed.proj.weight": (D, 3, patch_size, patch_size), "patch_embed.proj.bias": (D,)}
    for i in range(depth):
        p = f"blocks.{i}."
        shapes.update({p + "norm1.weight": (D,), p + "norm1.bias": (D,),
                       p + "attn.qkv.weight": (3 * D, D), p + "attn.qkv.bias": (3 * D,),
                       p + "attn.proj.weight": (D, D), p + "attn.proj.bias": (D,),
                       p + "norm2.weight": (D,), p + "norm2.bias": (D,),
                       p + "mlp.fc1.weight": (Fd, D), p + "mlp.fc1.bias": (Fd,),
                       p + "mlp.fc2.weight": (D, Fd), p + "mlp.fc2.bias": (D,)})
    shapes.update({"norm.weight": (D,), "norm.bias": (D,), "head.weight": (num_classes, D), "head.bias": (num_classes,),
                   "head_dist.weight": (num_classes, D), "head_dist.bias": (num_classes,)})
    return shapes


def deit_forward(sd, images, *, patch_size, depth, num_heads, training, distilled_training, eps=1e-6):
    """Distilled ViT forward: tokens [cls, dist, patches] + pos_embed, pre-norm blocks (same arithmetic as
    encoder_block, timm key names), final norm, head(x[:,0]) / head_dist(x[:,1]); tuple iff training and
    distilled_training (deit.py:45,65,70), else their mean (deit.py:95-96)."""
    n = images.shape[0]
    D = sd["cls_token"].shape[-1]
    x = F.conv2d(images, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch_size)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd["cls_token"].expand(n, -1, -1), sd["dist_token"].expand(n, -1, -1), x], dim=1)
    x = x + sd["pos_embed"]
    for i in range(depth):
        p = f"blocks.{i}."
        h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
        a = mha_batch_first(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"], sd[p + "attn.proj.weight"],
                            sd[p + "attn.proj.bias"], num_heads)
        x = x + a
        y = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
        y = F.linear(F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                     sd[p + "mlp.fc2.bias"])
        x = x + y
    x = F.layer_norm(x, (D,), sd["norm.weight"], sd["norm.bias"], eps)
    out = F.linear(x[:, 0], sd["head.weight"], sd["head.bias"])
    out_dist = F.linear(x[:, 1], sd["head_dist.weight"], sd["head_dist.bias"])
    if training and distilled_training:
        return out, out_dist
    return (out + out_dist) / 2


def distillation_loss(outputs, outputs_kd, labels, teacher_logits, distillation_type, alpha, tau):
    """DistillationLoss.forward — utils/distillation_loss.py:30-75 (base criterion = mean cross-entropy)."""
    base = F.cross_entropy(outputs, labels)                                        # :43
    if distillation_type == "none":
        return base                                                                # :44-45
    if distillation_type == "soft":                                                # :55-67
        T = tau
        kd = F.kl_div(F.log_softmax(outputs_kd / T, dim=1), F.log_softmax(teacher_logits / T, dim=1), reduction="sum",
                      log_target=True) * (T * T) / outputs_kd.numel()
    else:                                                                          # hard, :71-72
        kd = F.cross_entropy(outputs_kd, teacher_logits.argmax(dim=1))
    return base * (1 - alpha) + kd * alpha                                         # :74


# ------------------------------------------------------------------------------------------------------------
# DETR transformer encoder (models/object_detection/transformer.py:98-115, 192-247)
# ------------------------------------------------------------------------------------------------------------
def detr_param_shapes(d_model, dim_feedforward, num_layers, normalize_before):
    D, Fd = d_model, dim_feedforward
    shapes = {}
    for i in range(num_layers):
        p = f"layers.{i}."
        shapes.update({p + "self_attn.in_proj_weight": (3 * D, D), p + "self_attn.in_proj_bias": (3 * D,),
                       p + "self_attn.out_proj.weight": (D, D), p + "self_attn.out_proj.bias": (D,),
                       p + "linear1.weight": (Fd, D), p + "linear1.bias": (Fd,),
                       p + "linear2.weight": (D, Fd), p + "linear2.bias": (D,),
                       p + "norm1.weight": (D,), p + "norm1.bias": (D,), p + "norm2.weight": (D,), p + "norm2.bias": (D,)})
    if normalize_before:
        shapes.update({"norm.weight": (D,), "norm.bias": (D,)})
    return shapes


def _detr_self_attn(qk_in, v_in, sd, p, nhead, key_padding_mask, attn_drop=None):
    """self_attn(q, k, value=src, key_padding_mask=...)[0] with q = k = src + pos — transformer.py:218-219.
    q is k but k is not v => three separate projections (torch/nn/functional.py:5866-5873); explicit
    softmax(QK^T/sqrt(hd) + mask)V (functional.py:6630-6666); sequence-first tensors [S, N, C]."""
    S, N, D = qk_in.shape
    hd = D // nhead
    w, b = sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"]
    q = F.linear(qk_in, w[:D], b[:D])
    k = F.linear(qk_in, w[D:2 * D], b[D:2 * D])
    v = F.linear(v_in, w[2 * D:], b[2 * D:])
    q = q.view(S, N, nhead, hd).permute(1, 2, 0, 3)
    k = k.view(S, N, nhead, hd).permute(1, 2, 0, 3)
    v = v.view(S, N, nhead, hd).permute(1, 2, 0, 3)
    mask = None
    if key_padding_mask is not None:
        mask = torch.zeros(N, 1, 1, S, dtype=q.dtype, device=q.device).masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    if attn_drop is None:
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    else:                                  # dropout on the [N,H,S,S] softmax weights (functional.py:6650-6652), explicit mask
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        o = attn_drop(torch.softmax(sc if mask is None else sc + mask, dim=-1)) @ v
    o = o.permute(2, 0, 1, 3).reshape(S, N, D)
    return F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])


def detr_encoder_forward(sd, src, *, nhead, num_layers, normalize_before=False, activation="relu", src_key_padding_mask=None,
                         pos=None, eps=1e-5, drop=None):
    """TransformerEncoder.forward over TransformerEncoderLayer.forward_post/pre — transformer.py:105-115, 213-241.
    src, pos: [S, N, C]; src_key_padding_mask: [N, S] bool, True = padding.  ``drop``: None (p = 0) or an ExplicitDropout whose
    masks[(layer, site)] replay dropout1 (site 0, :220/:236), dropout (1, :223/:239), dropout2 (2, :224/:240) and the attention
    dropout of nn.MultiheadAttention(dropout=p) (3, :195) — the reference uses ONE rate for all four."""
    act = F.relu if activation == "relu" else F.gelu
    D = src.shape[-1]
    x = src
    dz = (lambda key, t: t) if drop is None else drop
    for i in range(num_layers):
        p = f"layers.{i}."
        ad = None if (drop is None or drop.p_attn == 0) else (lambda P, i=i: drop((i, 3), P))
        if not normalize_before:                                                   # forward_post :213-226
            qk = x if pos is None else x + pos
            x = x + dz((i, 0), _detr_self_attn(qk, x, sd, p, nhead, src_key_padding_mask, ad))
            x = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
            y = F.linear(dz((i, 1), act(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))), sd[p + "linear2.weight"],
                         sd[p + "linear2.bias"])
            x = F.layer_norm(x + dz((i, 2), y), (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
        else:                                                                      # forward_pre :228-241
            h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
            qk = h if pos is None else h + pos
            x = x + dz((i, 0), _detr_self_attn(qk, h, sd, p, nhead, src_key_padding_mask, ad))
            h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
            x = x + dz((i, 2), F.linear(dz((i, 1), act(F.linear(h, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))),
                                        sd[p + "linear2.weight"], sd[p + "linear2.bias"]))
    if "norm.weight" in sd:                                                        # :112-113
        x = F.layer_norm(x, (D,), sd["norm.weight"], sd["norm.bias"], eps)
    return x


# ------------------------------------------------------------------------------------------------------------
# DETR: what sits between the backbone and the encoder — detr.py:33-63 (AbsolutePositionalEncoding), :125 (input_proj),
# transformer.py:47-53 (flatten to sequence-first), utils/coco/util/misc.py:307-332 (nested_tensor_from_tensor_list)
# ------------------------------------------------------------------------------------------------------------
def detr_abs_pos_encoding(row_w, col_w, n, h, w):
    """AbsolutePositionalEncoding.forward — detr.py:49-63: [n, 2*pf, h, w] with channels [col_embed[x] | row_embed[y]]."""
    i = torch.arange(w, device=row_w.device)
    j = torch.arange(h, device=row_w.device)
    x_emb = F.embedding(i, col_w)
    y_emb = F.embedding(j, row_w)
    return torch.cat([x_emb.unsqueeze(0).repeat(h, 1, 1), y_emb.unsqueeze(1).repeat(1, w, 1)], dim=-1).permute(2, 0, 1).unsqueeze(0).repeat(
        n, 1, 1, 1)


def detr_input_proj(features, w, b):
    """Detr.input_proj — detr.py:125: nn.Conv2d(C_in, hidden, kernel_size=1) on the [N, C_in, h, w] backbone feature map."""
    return F.conv2d(features, w, b)


def detr_flatten(src, pos_embed, mask):
    """Transformer.forward's first lines — transformer.py:49-53: NxCxHxW -> HWxNxC, mask [N, h, w] -> [N, h*w]."""
    return src.flatten(2).permute(2, 0, 1), pos_embed.flatten(2).permute(2, 0, 1), mask.flatten(1)


def nested_tensor_from_tensor_list(tensor_list):
    """utils/coco/util/misc.py:307-332: (zero-padded batch [b, c, H, W], mask [b, H, W] with True on padding)."""
    max_size = [max(img.shape[d] for img in tensor_list) for d in range(3)]
    tensor = torch.zeros([len(tensor_list)] + max_size, dtype=tensor_list[0].dtype)
    mask = torch.ones((len(tensor_list), max_size[1], max_size[2]), dtype=torch.bool)
    for img, pad_img, m in zip(tensor_list, tensor, mask):
        pad_img[: img.shape[0], : img.shape[1], : img.shape[2]].copy_(img)
        m[: img.shape[1], : img.shape[2]] = False
    return tensor, mask


# ------------------------------------------------------------------------------------------------------------
# DETR transformer decoder — transformer.py:66-95 (TransformerDecoder), :118-189 (TransformerDecoderLayer).
# The reference layer registers its cross-attention as ``multi_head_attn`` (:122) but calls ``self.multihead_attn`` (:148,:172), so it
# raises AttributeError as written; the restatement below is the forward the code spells out with that name resolved, and it is
# pinned against the live reference with exactly that alias added (tools/make_golden.py::decoder_case).
# ------------------------------------------------------------------------------------------------------------
def detr_decoder_param_shapes(d_model, dim_feedforward, num_layers, with_norm=True):
    D, Fd = d_model, dim_feedforward
    sh = {}
    for i in range(num_layers):
        p = f"layers.{i}."
        for a in ("self_attn", "multi_head_attn"):
            sh.update({p + a + ".in_proj_weight": (3 * D, D), p + a + ".in_proj_bias": (3 * D,),
                       p + a + ".out_proj.weight": (D, D), p + a + ".out_proj.bias": (D,)})
        sh.update({p + "linear1.weight": (Fd, D), p + "linear1.bias": (Fd,), p + "linear2.weight": (D, Fd), p + "linear2.bias": (D,)})
        for n in ("norm1", "norm2", "norm3"):
            sh.update({p + n + ".weight": (D,), p + n + ".bias": (D,)})
    if with_norm:
        sh.update({"norm.weight": (D,), "norm.bias": (D,)})
    return sh


def _mha_seq_first(q_in, k_in, v_in, sd, p, nhead, key_padding_mask, attn_drop=None):
    """nn.MultiheadAttention (sequence-first) with distinct query / key / value inputs: three projections from the packed
    in_proj_weight (torch/nn/functional.py:5866-5873), explicit softmax path (functional.py:6630-6666)."""
    Sq, N, D = q_in.shape
    Sk = k_in.shape[0]
    hd = D // nhead
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(q_in, w[:D], b[:D]).view(Sq, N, nhead, hd).permute(1, 2, 0, 3)
    k = F.linear(k_in, w[D:2 * D], b[D:2 * D]).view(Sk, N, nhead, hd).permute(1, 2, 0, 3)
    v = F.linear(v_in, w[2 * D:], b[2 * D:]).view(Sk, N, nhead, hd).permute(1, 2, 0, 3)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if key_padding_mask is not None:
        sc = sc.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    P = torch.softmax(sc, dim=-1)
    if attn_drop is not None:
        P = attn_drop(P)
    o = (P @ v).permute(2, 0, 1, 3).reshape(Sq, N, D)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def detr_decoder_forward(sd, tgt, memory, *, nhead, num_layers, activation="relu", memory_key_padding_mask=None, pos=None, query_pos=None,
                         return_intermediate=False, eps=1e-5, drop=None, normalize_before=False):
    """TransformerDecoder.forward over TransformerDecoderLayer.forward_post / forward_pre — transformer.py:74-95, 138-156, 158-178.  Returns [1, Q, N, D], or
    [L, Q, N, D] with return_intermediate.  ``drop``: ExplicitDropout with masks[(layer, site)]; sites 0 = dropout1, 1 = dropout,
    2 = dropout3, 3 = self-attention weights, 4 = cross-attention weights, 5 = dropout2."""
    act = F.relu if activation == "relu" else F.gelu
    D = tgt.shape[-1]
    dz = (lambda key, t: t) if drop is None else drop
    wp = lambda t, p_: t if p_ is None else t + p_
    has_norm = "norm.weight" in sd
    fnorm = lambda t: F.layer_norm(t, (D,), sd["norm.weight"], sd["norm.bias"], eps)
    x, inter = tgt, []
    for i in range(num_layers):
        p = f"layers.{i}."
        ad = (lambda site: None) if drop is None else (lambda site, i=i: (lambda P: drop((i, site), P)))
        if normalize_before:                                                                                    # forward_pre :158-178
            h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
            qk = wp(h, query_pos)
            x = x + dz((i, 0), _mha_seq_first(qk, qk, h, sd, p + "self_attn.", nhead, None, ad(3)))
            h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
            x = x + dz((i, 5), _mha_seq_first(wp(h, query_pos), wp(memory, pos), memory, sd, p + "multi_head_attn.", nhead,
                                              memory_key_padding_mask, ad(4)))
            h = F.layer_norm(x, (D,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], eps)
            x = x + dz((i, 2), F.linear(dz((i, 1), act(F.linear(h, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))),
                                        sd[p + "linear2.weight"], sd[p + "linear2.bias"]))
            if return_intermediate:
                inter.append(fnorm(x))
            continue
        qk = wp(x, query_pos)
        x = x + dz((i, 0), _mha_seq_first(qk, qk, x, sd, p + "self_attn.", nhead, None, ad(3)))                  # :142-144
        x = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
        c = _mha_seq_first(wp(x, query_pos), wp(memory, pos), memory, sd, p + "multi_head_attn.", nhead, memory_key_padding_mask, ad(4))
        x = x + dz((i, 5), c)                                                                                   # :145-150
        x = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
        y = F.linear(dz((i, 1), act(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))), sd[p + "linear2.weight"],
                     sd[p + "linear2.bias"])
        x = F.layer_norm(x + dz((i, 2), y), (D,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], eps)            # :152-155
        if return_intermediate:
            inter.append(fnorm(x))                                                                              # :84-85
    if return_intermediate:
        return torch.stack(inter)        # :87-93: the last entry is norm(output) either way
    return (fnorm(x) if has_norm else x).unsqueeze(0)


# ------------------------------------------------------------------------------------------------------------
# Seeded weights / inputs shared by the golden generator, the tests, smoke() and bench.py
# ------------------------------------------------------------------------------------------------------------
def seeded_state_dict(shapes, seed, std=0.02):
    """Deterministic (CPU generator) N(0, std^2) tensors; LayerNorm weights are 1 + N(0, std^2).
    A freshly constructed reference ViT has a zero head (vanilla_vit.py:149-151), which makes parity tests
    vacuous, hence seeded random weights everywhere."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        t = torch.randn(*shp, generator=g, dtype=torch.float32) * std
        is_norm_w = k.endswith("weight") and any(s in k for s in ("ln_1.", "ln_2.", "ln.", "norm1.", "norm2.", "norm3.", "norm."))
        if is_norm_w:
            t = t + 1.0
        if k.endswith("in_proj_weight") or k.endswith("qkv.weight") or k.endswith(".0.weight") or k.endswith(".3.weight") or \
                k.endswith("linear1.weight") or k.endswith("linear2.weight") or k.endswith("fc1.weight") or k.endswith("fc2.weight") or \
                k.endswith("proj.weight") or k.endswith("out_proj.weight"):
            t = t * (1.0 / std) * math.sqrt(1.0 / shp[-1]) if len(shp) == 2 else t
        if k.endswith("conv_proj.weight") or k.endswith("patch_embed.proj.weight"):
            t = t * (1.0 / std) * math.sqrt(1.0 / (shp[1] * shp[2] * shp[3]))
        if k.endswith("conv.weight") and tuple(shp[1:]) == (1, 3, 3):   # CPE / PEG depthwise conv: identity-ish, so the stream survives it
            t = t * (0.15 / std)
            t[:, 0, 1, 1] += 1.0
        if k.endswith("head.weight") or k.endswith("head_dist.weight"):
            t = t * (1.0 / std) * math.sqrt(1.0 / shp[-1])
        sd[k] = t
    return sd


def seeded_images(batch, image_size, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(batch, 3, image_size, image_size, generator=g, dtype=torch.float32)


def seeded_labels(batch, num_classes, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, num_classes, (batch,), generator=g)
