"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference, build container
only) on seeded inputs and weights.  The fixtures pin the CPU oracle (oracle/vit_oracle.py) and, through it, the
CUDA path.  Re-run:  python tools/make_golden.py
"""
import math
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("VITB200_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
for name in ("pycocotools", "pycocotools.coco"):      # utils/load_data.py:6 imports it at module top
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["pycocotools.coco"].COCO = object
sys.modules["pycocotools"].coco = sys.modules["pycocotools.coco"]

import torch  # noqa: E402

from models.image_classification.vanilla_vit import ViT as RefViT  # noqa: E402
from models.object_detection.transformer import TransformerEncoder as RefEnc, TransformerEncoderLayer as RefLayer  # noqa: E402
from utils.distillation_loss import DistillationLoss as RefDistill  # noqa: E402
from utils.args import get_args  # noqa: E402

from oracle import vit_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
SMALL = 4096


def grads_summary(named_params):
    norms, full = {}, {}
    for n, p in named_params:
        norms[n] = p.grad.norm().item()
        if p.grad.numel() <= SMALL:
            full[n] = p.grad.clone()
    return norms, full


def vit_case(name, cfg, batch, seed):
    m = RefViT(cfg["image_size"], cfg["patch_size"], cfg["num_layers"], cfg["num_heads"], cfg["hidden_dim"], cfg["mlp_dim"], 0.0, 0.0,
               cfg["num_classes"])
    sd = O.seeded_state_dict(O.vit_param_shapes(**cfg), seed)
    m.load_state_dict(sd)
    m.train()
    images, labels = O.seeded_images(batch, cfg["image_size"], seed + 1), O.seeded_labels(batch, cfg["num_classes"], seed + 2)
    logits = m(images)
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    with torch.no_grad():
        feats = m.forward_features(images)
    norms, full = grads_summary(m.named_parameters())
    torch.save({"cfg": cfg, "batch": batch, "seed": seed, "logits": logits.detach(), "loss": loss.item(),
                "features_cls": feats[:, 0].clone(), "features_norm": feats.norm().item(), "grad_norms": norms, "grads_small": full},
               os.path.join(OUT, name))
    print(name, "loss", loss.item())


def _ref_abs_pos_encoding_class():
    """models/object_detection/detr.py does not parse (its last line is an unfinished assignment), so the class cannot be imported:
    its source lines (the class statement up to the next top-level statement) are executed as they are."""
    lines = open(os.path.join(REF, "models", "object_detection", "detr.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.startswith("class AbsolutePositionalEncoding"))
    i1 = next(i for i in range(i0 + 1, len(lines)) if lines[i].startswith("def ") or lines[i].startswith("class "))
    ns = {"torch": torch, "nn": torch.nn, "NestedTensor": object}
    exec("\n".join(lines[i0:i1]), ns)
    return ns["AbsolutePositionalEncoding"]


def detr_front_case(name, seed, c_in=64, d_model=256, nhead=4, ffn=512, layers=2):
    """Two images of different sizes -> the reference's nested_tensor_from_tensor_list -> a fixed strided-conv stand-in for the backbone
    (not on the path) -> the reference's AbsolutePositionalEncoding + nn.Conv2d(C_in, hidden, 1) (detr.py:125) + the flatten / permute
    lines of Transformer.forward (transformer.py:49-53) -> the reference encoder."""
    from utils.coco.util.misc import nested_tensor_from_tensor_list as ref_nested
    RefPos = _ref_abs_pos_encoding_class()
    g = torch.Generator().manual_seed(seed)
    imgs = [torch.randn(3, 96, 128, generator=g), torch.randn(3, 112, 104, generator=g)]
    nt = ref_nested(imgs)
    stem = torch.randn(c_in, 3, 8, 8, generator=g) * 0.05
    feats = torch.nn.functional.conv2d(nt.tensors, stem, stride=8).detach().requires_grad_(True)          # [2, c_in, 14, 16]
    mask = torch.nn.functional.interpolate(nt.mask[None].float(), size=feats.shape[-2:]).to(torch.bool)[0]
    posm = RefPos(d_model // 2)
    with torch.no_grad():
        posm.row_embed.weight.copy_(torch.rand(posm.row_embed.weight.shape, generator=g))
        posm.col_embed.weight.copy_(torch.rand(posm.col_embed.weight.shape, generator=g))

    class _NT:
        tensors = feats
    pos = posm(_NT())
    proj = torch.nn.Conv2d(c_in, d_model, kernel_size=1)
    with torch.no_grad():
        proj.weight.copy_(torch.randn(proj.weight.shape, generator=g) * (1.0 / math.sqrt(c_in)))
        proj.bias.copy_(torch.randn(proj.bias.shape, generator=g) * 0.1)
    src = proj(feats)
    bs, c, h, w = src.shape
    s2 = src.flatten(2).permute(2, 0, 1)                      # transformer.py:49-53
    p2 = pos.flatten(2).permute(2, 0, 1)
    m2 = mask.flatten(1)
    enc = RefEnc(RefLayer(d_model, nhead, ffn, 0.0, "relu", False), layers, None)
    sd = O.seeded_state_dict(O.detr_param_shapes(d_model, ffn, layers, False), seed + 1)
    enc.load_state_dict(sd)
    enc.train()
    out = enc(s2, src_key_padding_mask=m2, pos=p2)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout)
    torch.save({"seed": seed, "c_in": c_in, "d_model": d_model, "nhead": nhead, "ffn": ffn, "layers": layers,
                "padded_shape": tuple(nt.tensors.shape), "padded_sum": nt.tensors.sum().item(), "padded_1_0_100": nt.tensors[1, 0, 100].clone(),
                "mask_full": nt.mask.clone(), "mask_feat": mask.clone(), "pos_sum": pos.sum().item(),
                "pos_00": pos[0, :, 0, 0].clone(), "pos_last": pos[1, :, -1, -1].clone(), "src_norm": src.norm().item(),
                "src_n0_c0": src[0, 0].detach().clone(), "out_norm": out.norm().item(), "out_row0": out[0].detach().clone(),
                "dfeat_norm": feats.grad.norm().item(), "dfeat_n1_c3": feats.grad[1, 3].clone(),
                "dW_norm": proj.weight.grad.norm().item(), "dW_row0": proj.weight.grad[0, :, 0, 0].clone(), "db": proj.bias.grad.clone(),
                "drow": posm.row_embed.weight.grad.clone(), "dcol": posm.col_embed.weight.grad.clone()}, os.path.join(OUT, name))
    print(name, "out norm", out.norm().item())


def detr_case(name, d_model, nhead, ffn, layers, S, N, seed, pre_norm=False):
    # transformer.py:32-33: the encoder gets a final LayerNorm iff normalize_before
    enc = RefEnc(RefLayer(d_model, nhead, ffn, 0.0, "relu", pre_norm), layers, torch.nn.LayerNorm(d_model) if pre_norm else None)
    sd = O.seeded_state_dict(O.detr_param_shapes(d_model, ffn, layers, pre_norm), seed)
    enc.load_state_dict(sd)
    enc.train()
    g = torch.Generator().manual_seed(seed + 1)
    src = torch.randn(S, N, d_model, generator=g, requires_grad=True)
    pos = torch.randn(S, N, d_model, generator=g, requires_grad=True)
    valid = torch.randint(S // 2, S + 1, (N,), generator=g)
    kpm = torch.arange(S)[None, :] >= valid[:, None]
    gout = torch.randn(S, N, d_model, generator=g)
    out = enc(src, src_key_padding_mask=kpm, pos=pos)
    out.backward(gout)
    norms, full = grads_summary(enc.named_parameters())
    torch.save({"d_model": d_model, "nhead": nhead, "ffn": ffn, "layers": layers, "S": S, "N": N, "seed": seed, "pre_norm": pre_norm,
                "out": out.detach(),
                "dsrc_norm": src.grad.norm().item(), "dpos_norm": pos.grad.norm().item(), "dsrc_row0": src.grad[0].clone(),
                "grad_norms": norms, "grads_small": full}, os.path.join(OUT, name))
    print(name, "out norm", out.norm().item())


def distill_case(name, seed):
    g = torch.Generator().manual_seed(seed)
    B, C = 16, 100
    out, kd = torch.randn(B, C, generator=g), torch.randn(B, C, generator=g)
    labels = torch.randint(0, C, (B,), generator=g)
    W = torch.randn(3 * 8 * 8, C, generator=g) * 0.05
    teacher = lambda x: x.flatten(1) @ W
    inputs = torch.randn(B, 3, 8, 8, generator=g)
    res = {"seed": seed}
    for kind in ("none", "soft", "hard"):
        crit = RefDistill(torch.nn.CrossEntropyLoss(), teacher, kind, 0.5, 5.0)
        res[kind] = crit(inputs, (out, kd), labels).item()
    torch.save(res, os.path.join(OUT, name))
    print(name, res)


class ReplayedDropout:
    """Context manager: torch.nn.functional.dropout (what nn.Dropout.forward and nn.MultiheadAttention's explicit-softmax path call)
    draws its keep masks, in call order, from a seeded generator: keep = rand(x.shape) >= p, y = keep * x / (1 - p).  The oracle test
    regenerates the same masks in the same order, which pins the SITES of the oracle's ExplicitDropout against the live reference."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.shapes = []

    def __enter__(self):
        import torch.nn.functional as Fn
        self._orig = Fn.dropout

        def fake(x, p=0.5, training=True, inplace=False):
            if not training or p == 0:
                return x
            keep = (torch.rand(x.shape, generator=self.g) >= p).to(x.dtype)
            self.shapes.append(tuple(x.shape))
            return x * keep / (1.0 - p)
        Fn.dropout = fake
        return self

    def __exit__(self, *exc):
        import torch.nn.functional as Fn
        Fn.dropout = self._orig


def detr_dropout_case(name, d_model, nhead, ffn, layers, S, N, seed, p, pre_norm=False):
    enc = RefEnc(RefLayer(d_model, nhead, ffn, p, "relu", pre_norm), layers, torch.nn.LayerNorm(d_model) if pre_norm else None)
    sd = O.seeded_state_dict(O.detr_param_shapes(d_model, ffn, layers, pre_norm), seed)
    enc.load_state_dict(sd)
    enc.train()
    g = torch.Generator().manual_seed(seed + 1)
    src = torch.randn(S, N, d_model, generator=g, requires_grad=True)
    pos = torch.randn(S, N, d_model, generator=g, requires_grad=True)
    valid = torch.randint(S // 2, S + 1, (N,), generator=g)
    kpm = torch.arange(S)[None, :] >= valid[:, None]
    gout = torch.randn(S, N, d_model, generator=g)
    with ReplayedDropout(seed + 2) as rd:
        out = enc(src, src_key_padding_mask=kpm, pos=pos)
    out.backward(gout)
    norms, full = grads_summary(enc.named_parameters())
    torch.save({"d_model": d_model, "nhead": nhead, "ffn": ffn, "layers": layers, "S": S, "N": N, "seed": seed, "pre_norm": pre_norm,
                "p": p, "mask_shapes": rd.shapes, "out": out.detach(), "dsrc_norm": src.grad.norm().item(),
                "dpos_norm": pos.grad.norm().item(), "grad_norms": norms, "grads_small": full}, os.path.join(OUT, name))
    print(name, "out norm", out.norm().item(), "dropout calls", len(rd.shapes))


def vit_hidden_dropout_case(name, cfg, batch, seed, p):
    """Hidden dropout only (attention_dropout = 0): the attention dropout of the ViT path lives inside SDPA and cannot be replayed."""
    m = RefViT(cfg["image_size"], cfg["patch_size"], cfg["num_layers"], cfg["num_heads"], cfg["hidden_dim"], cfg["mlp_dim"], p, 0.0,
               cfg["num_classes"])
    sd = O.seeded_state_dict(O.vit_param_shapes(**cfg), seed)
    m.load_state_dict(sd)
    m.train()
    images, labels = O.seeded_images(batch, cfg["image_size"], seed + 1), O.seeded_labels(batch, cfg["num_classes"], seed + 2)
    with ReplayedDropout(seed + 3) as rd:
        logits = m(images)
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    norms, full = grads_summary(m.named_parameters())
    torch.save({"cfg": cfg, "batch": batch, "seed": seed, "p": p, "mask_shapes": rd.shapes, "logits": logits.detach(), "loss": loss.item(),
                "grad_norms": norms, "grads_small": full}, os.path.join(OUT, name))
    print(name, "loss", loss.item(), "dropout calls", len(rd.shapes))


def cpe_case(name, which, cfg, batch, seed):
    """CPEViT (cpe_vit.py) / CPVT (cpvt.py) / CPVTGAP (cpvt_gap.py) forward + backward; identity-ish seeded PEG weights."""
    import importlib
    mod = importlib.import_module({"CPEViT": "models.image_classification.cpe_vit", "CPVT": "models.image_classification.cpvt",
                                   "CPVTGAP": "models.image_classification.cpvt_gap"}[which])
    m = getattr(mod, which)(cfg["image_size"], cfg["patch_size"], cfg["num_layers"], cfg["num_heads"], cfg["hidden_dim"], cfg["mlp_dim"],
                            0.0, 0.0, cfg["num_classes"])
    sd = O.seeded_state_dict(O.cpe_param_shapes(**cfg, peg_blocks=which != "CPEViT"), seed)
    assert set(sd) == set(m.state_dict()), set(sd) ^ set(m.state_dict())
    m.load_state_dict(sd)
    m.train()
    images, labels = O.seeded_images(batch, cfg["image_size"], seed + 1), O.seeded_labels(batch, cfg["num_classes"], seed + 2)
    logits = m(images)
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    with torch.no_grad():
        feats = m.forward_features(images)
    norms, full = grads_summary(m.named_parameters())
    torch.save({"which": which, "cfg": cfg, "batch": batch, "seed": seed, "logits": logits.detach(), "loss": loss.item(),
                "features_cls": feats[:, 0].clone(), "features_norm": feats.norm().item(), "grad_norms": norms, "grads_small": full},
               os.path.join(OUT, name))
    print(name, "loss", loss.item())


def decoder_case(name, d_model, nhead, ffn, layers, Q, S, N, seed, p=0.0, return_intermediate=False, pre_norm=False):
    """TransformerDecoder over TransformerDecoderLayer.forward_post.  The reference layer registers ``multi_head_attn``
    (transformer.py:122) and calls ``self.multihead_attn`` (:148): the alias below is the ONE change that lets the reference's own
    forward run; everything else is the unmodified code."""
    from models.object_detection.transformer import TransformerDecoder as RefDec, TransformerDecoderLayer as RefDecLayer
    if not hasattr(RefDecLayer, "multihead_attn"):
        RefDecLayer.multihead_attn = property(lambda self: self.multi_head_attn)
    dec = RefDec(RefDecLayer(d_model, nhead, ffn, p, "relu", pre_norm), layers, torch.nn.LayerNorm(d_model), return_intermediate=return_intermediate)
    sd = O.seeded_state_dict(O.detr_decoder_param_shapes(d_model, ffn, layers), seed)
    assert set(sd) == set(dec.state_dict()), set(sd) ^ set(dec.state_dict())
    dec.load_state_dict(sd)
    dec.train()
    g = torch.Generator().manual_seed(seed + 1)
    tgt = torch.randn(Q, N, d_model, generator=g, requires_grad=True)
    memory = torch.randn(S, N, d_model, generator=g, requires_grad=True)
    pos = torch.randn(S, N, d_model, generator=g, requires_grad=True)
    qpos = torch.randn(Q, N, d_model, generator=g, requires_grad=True)
    valid = torch.randint(S // 2, S + 1, (N,), generator=g)
    kpm = torch.arange(S)[None, :] >= valid[:, None]
    with ReplayedDropout(seed + 2) as rd:
        out = dec(tgt, memory, memory_key_padding_mask=kpm, pos=pos, query_pos=qpos)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout)
    norms, full = grads_summary(dec.named_parameters())
    torch.save({"d_model": d_model, "nhead": nhead, "ffn": ffn, "layers": layers, "Q": Q, "S": S, "N": N, "seed": seed, "p": p,
                "return_intermediate": return_intermediate, "pre_norm": pre_norm, "mask_shapes": rd.shapes, "out": out.detach(),
                "dtgt_norm": tgt.grad.norm().item(), "dmem_norm": memory.grad.norm().item(), "dpos_norm": pos.grad.norm().item(),
                "dqpos_norm": qpos.grad.norm().item(), "dmem_row0": memory.grad[0].clone(), "grad_norms": norms, "grads_small": full},
               os.path.join(OUT, name))
    print(name, "out", tuple(out.shape), out.norm().item(), "dropout calls", len(rd.shapes))


def kat_case(name):
    torch.manual_seed(123)
    args = dict(get_args("vit_tiny_cifar10"))
    m = RefViT(**args)
    init_sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.eval()
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    logits = m(x)
    loss = torch.nn.CrossEntropyLoss()(logits, torch.tensor([0, 1, 2, 3]))
    b16 = RefViT(224, 16, 12, 12, 768, 3072, 0.0, 0.0, 1000)
    res = {"tiny_args": args, "tiny_keys": list(init_sd.keys()), "tiny_params": sum(p.numel() for p in m.parameters()),
           "tiny_fresh_logits_absmax": logits.abs().max().item(), "tiny_fresh_loss": loss.item(), "ln10": math.log(10.0),
           "b16_keys": list(b16.state_dict().keys()), "b16_params": sum(p.numel() for p in b16.parameters()),
           "tiny_init_checksums": {k: (v.double().sum().item(), v.double().abs().sum().item()) for k, v in init_sd.items()},
           "init_seed": 123}
    torch.save(res, os.path.join(OUT, name))
    print(name, res["tiny_params"], res["b16_params"], res["tiny_fresh_loss"])


if __name__ == "__main__":
    TINY = dict(image_size=32, patch_size=4, num_layers=7, num_heads=4, hidden_dim=256, mlp_dim=512, num_classes=10)
    B16_2L = dict(image_size=224, patch_size=16, num_layers=2, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
    vit_case("vit_tiny_b4.pt", TINY, 4, 101)
    vit_case("vit_b16x2_b2.pt", B16_2L, 2, 111)
    detr_case("detr_enc_d256.pt", 256, 4, 512, 2, 70, 2, 121)
    detr_case("detr_enc_prenorm_d256.pt", 256, 4, 512, 2, 70, 2, 131, pre_norm=True)
    distill_case("distill_loss.pt", 131)
    kat_case("kat.pt")
    detr_dropout_case("detr_enc_dropout_d256.pt", 256, 4, 512, 2, 70, 2, 141, 0.1)
    detr_dropout_case("detr_enc_prenorm_dropout_d256.pt", 256, 4, 512, 2, 70, 2, 151, 0.2, pre_norm=True)
    vit_hidden_dropout_case("vit_tiny_hidden_dropout_b4.pt", dict(TINY, num_layers=3), 4, 161, 0.1)
    decoder_case("detr_dec_d256.pt", 256, 4, 512, 2, 20, 70, 2, 201, return_intermediate=True)
    decoder_case("detr_dec_dropout_d256.pt", 256, 4, 512, 2, 20, 70, 2, 211, p=0.1)
    decoder_case("detr_dec_prenorm_dropout_d256.pt", 256, 4, 512, 2, 20, 70, 2, 221, p=0.1, return_intermediate=True, pre_norm=True)
    cpe_case("cpe_vit_tiny_b4.pt", "CPEViT", dict(TINY, num_layers=3), 4, 171)
    cpe_case("cpvt_tiny_b4.pt", "CPVT", dict(TINY, num_layers=3), 4, 181)
    cpe_case("cpvt_gap_tiny_b4.pt", "CPVTGAP", dict(TINY, num_layers=2), 4, 191)
    detr_front_case("detr_front_d256.pt", 301)
