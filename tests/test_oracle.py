"""CPU tests: the oracle restatement against the golden fixtures generated from the unmodified reference
(tools/make_golden.py), and — when /root/reference is present (build container only) — against the live reference."""
import math
import os
import sys
import types

import pytest
import torch

from helpers import O, rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = os.environ.get("VITB200_REFERENCE", "/root/reference")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


@pytest.mark.parametrize("name", ["vit_tiny_b4.pt", "vit_b16x2_b2.pt"])
def test_oracle_vit_matches_reference_golden(name):
    gd = load(name)
    cfg, batch, seed = gd["cfg"], gd["batch"], gd["seed"]
    sd = {k: v.requires_grad_(True) for k, v in O.seeded_state_dict(O.vit_param_shapes(**cfg), seed).items()}
    images, labels = O.seeded_images(batch, cfg["image_size"], seed + 1), O.seeded_labels(batch, cfg["num_classes"], seed + 2)
    kw = dict(patch_size=cfg["patch_size"], num_layers=cfg["num_layers"], num_heads=cfg["num_heads"])
    logits = O.vit_forward(sd, images, **kw)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    assert rel_l2(logits, gd["logits"]) < 1e-5
    assert abs(loss.item() - gd["loss"]) < 1e-5
    feats = O.vit_forward_features(sd, images, **kw).detach()
    assert rel_l2(feats[:, 0], gd["features_cls"]) < 1e-5
    assert abs(feats.norm().item() - gd["features_norm"]) < 1e-3 * gd["features_norm"]
    for k, n in gd["grad_norms"].items():
        assert abs(sd[k].grad.norm().item() - n) <= 1e-4 * max(n, 1e-6), k
    for k, g in gd["grads_small"].items():
        assert rel_l2(sd[k].grad, g) < 1e-4, k


@pytest.mark.parametrize("name", ["detr_enc_d256.pt", "detr_enc_prenorm_d256.pt"])
def test_oracle_detr_matches_reference_golden(name):
    gd = load(name)
    d, h, ffn, L, S, N, seed = gd["d_model"], gd["nhead"], gd["ffn"], gd["layers"], gd["S"], gd["N"], gd["seed"]
    pre = bool(gd.get("pre_norm", False))     # forward_pre + final encoder norm (transformer.py:228-241, 32-33) vs forward_post
    sd = {k: v.requires_grad_(True) for k, v in O.seeded_state_dict(O.detr_param_shapes(d, ffn, L, pre), seed).items()}
    g = torch.Generator().manual_seed(seed + 1)
    src = torch.randn(S, N, d, generator=g, requires_grad=True)
    pos = torch.randn(S, N, d, generator=g, requires_grad=True)
    valid = torch.randint(S // 2, S + 1, (N,), generator=g)
    kpm = torch.arange(S)[None, :] >= valid[:, None]
    gout = torch.randn(S, N, d, generator=g)
    out = O.detr_encoder_forward(sd, src, nhead=h, num_layers=L, normalize_before=pre, src_key_padding_mask=kpm, pos=pos)
    out.backward(gout)
    assert rel_l2(out, gd["out"]) < 1e-5
    assert abs(src.grad.norm().item() - gd["dsrc_norm"]) < 1e-4 * gd["dsrc_norm"]
    assert abs(pos.grad.norm().item() - gd["dpos_norm"]) < 1e-4 * gd["dpos_norm"]
    assert rel_l2(src.grad[0], gd["dsrc_row0"]) < 1e-4
    for k, n in gd["grad_norms"].items():
        assert abs(sd[k].grad.norm().item() - n) <= 1e-4 * max(n, 1e-6), k
    for k, gr in gd["grads_small"].items():
        assert rel_l2(sd[k].grad, gr) < 1e-4, k


def test_oracle_distillation_loss_matches_reference_golden():
    gd = load("distill_loss.pt")
    g = torch.Generator().manual_seed(gd["seed"])
    B, C = 16, 100
    out, kd = torch.randn(B, C, generator=g), torch.randn(B, C, generator=g)
    labels = torch.randint(0, C, (B,), generator=g)
    W = torch.randn(3 * 8 * 8, C, generator=g) * 0.05
    inputs = torch.randn(B, 3, 8, 8, generator=g)
    teacher = inputs.flatten(1) @ W
    for kind in ("none", "soft", "hard"):
        v = O.distillation_loss(out, kd, labels, teacher, kind, 0.5, 5.0).item()
        assert abs(v - gd[kind]) < 1e-5, (kind, v, gd[kind])


def _replayed_masks(seed, shapes, p):
    """The masks tools/make_golden.py::ReplayedDropout drew inside the live reference, regenerated in the same call order."""
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(sh, generator=g) >= p).float() for sh in shapes]


@pytest.mark.parametrize("name", ["detr_enc_dropout_d256.pt", "detr_enc_prenorm_dropout_d256.pt"])
def test_oracle_detr_dropout_sites_match_reference_golden(name):
    """dropout > 0 in train(): the reference's four dropout sites per layer (transformer.py:195,220,223,224 / :236,239,240) with the
    masks made an input.  Call order inside a reference layer: attention weights (site 3), dropout1 (0), dropout (1), dropout2 (2)."""
    gd = load(name)
    d, h, ffn, L, S, N, seed, p, pre = gd["d_model"], gd["nhead"], gd["ffn"], gd["layers"], gd["S"], gd["N"], gd["seed"], gd["p"], gd["pre_norm"]
    drawn = _replayed_masks(seed + 2, gd["mask_shapes"], p)
    assert len(drawn) == 4 * L and gd["mask_shapes"][0] == (N * h, S, S)
    masks = {}
    for i in range(L):
        for j, site in enumerate((3, 0, 1, 2)):
            masks[(i, site)] = drawn[4 * i + j]
    sd = {k: v.requires_grad_(True) for k, v in O.seeded_state_dict(O.detr_param_shapes(d, ffn, L, pre), seed).items()}
    g = torch.Generator().manual_seed(seed + 1)
    src = torch.randn(S, N, d, generator=g, requires_grad=True)
    pos = torch.randn(S, N, d, generator=g, requires_grad=True)
    valid = torch.randint(S // 2, S + 1, (N,), generator=g)
    kpm = torch.arange(S)[None, :] >= valid[:, None]
    gout = torch.randn(S, N, d, generator=g)
    out = O.detr_encoder_forward(sd, src, nhead=h, num_layers=L, normalize_before=pre, src_key_padding_mask=kpm, pos=pos,
                                 drop=O.ExplicitDropout(masks, p, p))
    out.backward(gout)
    assert rel_l2(out, gd["out"]) < 1e-5
    assert abs(src.grad.norm().item() - gd["dsrc_norm"]) < 1e-4 * gd["dsrc_norm"]
    assert abs(pos.grad.norm().item() - gd["dpos_norm"]) < 1e-4 * gd["dpos_norm"]
    for k, n in gd["grad_norms"].items():
        assert abs(sd[k].grad.norm().item() - n) <= 1e-4 * max(n, 1e-6), k
    for k, gr in gd["grads_small"].items():
        assert rel_l2(sd[k].grad, gr) < 1e-4, k


def test_oracle_vit_hidden_dropout_sites_match_reference_golden():
    """ViT hidden dropout (vanilla_vit.py:104 Encoder.dropout, :78 after attention, :38 mlp.2, :42 mlp.4) with replayed masks; the
    attention dropout of this path sits inside SDPA and cannot be replayed from outside (attention_dropout = 0 in the fixture)."""
    gd = load("vit_tiny_hidden_dropout_b4.pt")
    cfg, batch, seed, p = gd["cfg"], gd["batch"], gd["seed"], gd["p"]
    drawn = _replayed_masks(seed + 3, gd["mask_shapes"], p)
    L = cfg["num_layers"]
    assert len(drawn) == 1 + 3 * L
    masks = {"embed": drawn[0]}
    for i in range(L):
        for j, site in enumerate((0, 1, 2)):
            masks[(i, site)] = drawn[1 + 3 * i + j]
    sd = {k: v.requires_grad_(True) for k, v in O.seeded_state_dict(O.vit_param_shapes(**cfg), seed).items()}
    images, labels = O.seeded_images(batch, cfg["image_size"], seed + 1), O.seeded_labels(batch, cfg["num_classes"], seed + 2)
    logits = O.vit_forward(sd, images, patch_size=cfg["patch_size"], num_layers=L, num_heads=cfg["num_heads"],
                           drop=O.ExplicitDropout(masks, p, 0.0))
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    assert rel_l2(logits, gd["logits"]) < 1e-5 and abs(loss.item() - gd["loss"]) < 1e-5
    for k, n in gd["grad_norms"].items():
        assert abs(sd[k].grad.norm().item() - n) <= 1e-4 * max(n, 1e-6), k


@pytest.mark.parametrize("name", ["cpe_vit_tiny_b4.pt", "cpvt_tiny_b4.pt", "cpvt_gap_tiny_b4.pt"])
def test_oracle_cpe_models_match_reference_golden(name):
    """CPEViT / CPVT / CPVTGAP (cpe_vit.py, cpvt.py, cpvt_gap.py): conditional positional encodings."""
    gd = load(name)
    cfg, batch, seed, peg = gd["cfg"], gd["batch"], gd["seed"], gd["which"] != "CPEViT"
    sd = {k: v.requires_grad_(True) for k, v in O.seeded_state_dict(O.cpe_param_shapes(**cfg, peg_blocks=peg), seed).items()}
    images, labels = O.seeded_images(batch, cfg["image_size"], seed + 1), O.seeded_labels(batch, cfg["num_classes"], seed + 2)
    kw = dict(patch_size=cfg["patch_size"], num_layers=cfg["num_layers"], num_heads=cfg["num_heads"], peg_blocks=peg)
    logits = O.cpe_forward(sd, images, **kw)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    assert rel_l2(logits, gd["logits"]) < 1e-5 and abs(loss.item() - gd["loss"]) < 1e-5
    feats = O.cpe_forward_features(sd, images, **kw).detach()
    assert rel_l2(feats[:, 0], gd["features_cls"]) < 1e-5
    assert abs(feats.norm().item() - gd["features_norm"]) < 1e-3 * gd["features_norm"]
    for k, n in gd["grad_norms"].items():
        assert abs(sd[k].grad.norm().item() - n) <= 1e-4 * max(n, 1e-6), k
    for k, g in gd["grads_small"].items():
        assert rel_l2(sd[k].grad, g) < 1e-4, k


def test_cpe_module_keys_and_seed_parity():
    """The drop-in CPE modules have the reference's state_dict keys / shapes; with the reference tree present, the same seed gives a
    bit-identical initial state_dict (sub-modules are constructed in the reference's order)."""
    from vitb200 import cpvt as ours
    cfg = dict(image_size=32, patch_size=4, num_layers=2, num_heads=4, hidden_dim=256, mlp_dim=512, num_classes=10)
    for cls, peg in ((ours.CPEViT, False), (ours.CPVT, True), (ours.CPVTGAP, True)):
        m = cls(32, 4, 2, 4, 256, 512, 0.1, 0.1, 10)
        sh = O.cpe_param_shapes(**cfg, peg_blocks=peg)
        assert set(sh) == set(m.state_dict().keys())
        assert all(tuple(m.state_dict()[k].shape) == tuple(v) for k, v in sh.items())
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 3, 32, 32))              # no CPU fallback
    if not os.path.isdir(os.path.join(REF, "models")):
        return
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("pycocotools", "pycocotools.coco"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools"].coco = sys.modules["pycocotools.coco"]
    import importlib
    for modname, clsname in (("cpe_vit", "CPEViT"), ("cpvt", "CPVT"), ("cpvt_gap", "CPVTGAP")):
        ref_cls = getattr(importlib.import_module("models.image_classification." + modname), clsname)
        torch.manual_seed(77)
        a = ref_cls(32, 4, 2, 4, 256, 512, 0.1, 0.1, 10).state_dict()
        torch.manual_seed(77)
        b = getattr(ours, clsname)(32, 4, 2, 4, 256, 512, 0.1, 0.1, 10).state_dict()
        assert list(a.keys()) == list(b.keys()), clsname
        for k in a:
            assert torch.equal(a[k], b[k]), (clsname, k)


@pytest.mark.parametrize("name", ["detr_dec_d256.pt", "detr_dec_dropout_d256.pt", "detr_dec_prenorm_dropout_d256.pt"])
def test_oracle_detr_decoder_matches_reference_golden(name):
    """TransformerDecoder / TransformerDecoderLayer.forward_post (transformer.py:66-95, 138-156) against the live
    reference run with the ``multihead_attn`` alias (tools/make_golden.py::decoder_case); with dropout the masks drawn inside the
    reference are replayed — call order inside a layer: self-attention weights (site 3), dropout1 (0), cross-attention weights (4),
    dropout2 (5), dropout (1), dropout3 (2)."""
    gd = load(name)
    d, h, ffn, L, Q, S, N, seed, p = gd["d_model"], gd["nhead"], gd["ffn"], gd["layers"], gd["Q"], gd["S"], gd["N"], gd["seed"], gd["p"]
    drop = None
    if p > 0:
        drawn = _replayed_masks(seed + 2, gd["mask_shapes"], p)
        assert len(drawn) == 6 * L and gd["mask_shapes"][0] == (N * h, Q, Q) and gd["mask_shapes"][2] == (N * h, Q, S)
        masks = {}
        for i in range(L):
            for j, site in enumerate((3, 0, 4, 5, 1, 2)):
                masks[(i, site)] = drawn[6 * i + j]
        drop = O.ExplicitDropout(masks, p, p)
    sd = {k: v.requires_grad_(True) for k, v in O.seeded_state_dict(O.detr_decoder_param_shapes(d, ffn, L), seed).items()}
    g = torch.Generator().manual_seed(seed + 1)
    tgt = torch.randn(Q, N, d, generator=g, requires_grad=True)
    memory = torch.randn(S, N, d, generator=g, requires_grad=True)
    pos = torch.randn(S, N, d, generator=g, requires_grad=True)
    qpos = torch.randn(Q, N, d, generator=g, requires_grad=True)
    valid = torch.randint(S // 2, S + 1, (N,), generator=g)
    kpm = torch.arange(S)[None, :] >= valid[:, None]
    out = O.detr_decoder_forward(sd, tgt, memory, nhead=h, num_layers=L, memory_key_padding_mask=kpm, pos=pos, query_pos=qpos,
                                 return_intermediate=gd["return_intermediate"], drop=drop, normalize_before=gd.get("pre_norm", False))
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout)
    assert out.shape == gd["out"].shape and rel_l2(out, gd["out"]) < 1e-5
    for t, key in ((tgt, "dtgt_norm"), (memory, "dmem_norm"), (pos, "dpos_norm"), (qpos, "dqpos_norm")):
        assert abs(t.grad.norm().item() - gd[key]) < 1e-4 * gd[key], key
    assert rel_l2(memory.grad[0], gd["dmem_row0"]) < 1e-4
    for k, n in gd["grad_norms"].items():
        assert abs(sd[k].grad.norm().item() - n) <= 1e-4 * max(n, 1e-6), k
    for k, gr in gd["grads_small"].items():
        assert rel_l2(sd[k].grad, gr) < 1e-4, k


def test_known_answers_from_reference_init():
    """KAT-1/3/4: zero head => logits 0 and loss ln(C); key sets and parameter counts."""
    gd = load("kat.pt")
    assert gd["tiny_fresh_logits_absmax"] == 0.0
    assert abs(gd["tiny_fresh_loss"] - math.log(10.0)) < 1e-6
    a = gd["tiny_args"]
    shapes = O.vit_param_shapes(a["image_size"], a["patch_size"], a["num_layers"], a["num_heads"], a["hidden_dim"], a["mlp_dim"],
                                a["num_classes"])
    assert list(shapes.keys()) == gd["tiny_keys"]
    assert sum(math.prod(s) for s in shapes.values()) == gd["tiny_params"] == 3722250
    b16 = O.vit_param_shapes(224, 16, 12, 12, 768, 3072, 1000)
    assert list(b16.keys()) == gd["b16_keys"] and len(b16) == 152
    assert sum(math.prod(s) for s in b16.values()) == gd["b16_params"] == 86567656


def test_dropin_module_contract_and_seed_parity():
    """The drop-in ViT has the reference's keys/shapes/attributes and (KAT-5) the same seed gives the same init."""
    from vitb200.vit import ViT
    gd = load("kat.pt")
    a = gd["tiny_args"]
    torch.manual_seed(gd["init_seed"])
    m = ViT(**a)
    sd = m.state_dict()
    assert list(sd.keys()) == gd["tiny_keys"]
    for k, (s1, s2) in gd["tiny_init_checksums"].items():
        assert abs(sd[k].double().sum().item() - s1) < 1e-9 + 1e-12 * abs(s1), k
        assert abs(sd[k].double().abs().sum().item() - s2) < 1e-9 + 1e-12 * abs(s2), k
    for attr in ("image_size", "patch_size", "hidden_dim", "mlp_dim", "attention_dropout", "dropout", "num_classes", "norm_layer",
                 "num_patches", "num_layers", "num_heads", "device", "conv_proj", "class_token", "encoder", "heads"):
        assert hasattr(m, attr), attr
    with pytest.raises(Exception):
        ViT(33, 4, 1, 4, 256, 512, 0.0, 0.0, 10)       # image_size % patch_size != 0 (vanilla_vit.py:115)
    b = ViT(224, 16, 12, 12, 768, 3072, 0.0, 0.0, 1000)
    assert list(b.state_dict().keys()) == gd["b16_keys"]
    # no CPU fallback: running on CPU must fail loudly
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 32, 32))


def test_deit_and_detr_module_keys():
    from vitb200.deit import VisionTransformerDistilled
    from vitb200.detr import TransformerEncoder, TransformerEncoderLayer
    d = VisionTransformerDistilled(img_size=32, patch_size=16, depth=2, num_heads=6, embed_dim=384, mlp_ratio=4.0, drop_rate=0.0,
                                   attn_drop_rate=0.0, num_classes=100)
    sh = O.deit_param_shapes(32, 16, 2, 6, 384, 4.0, 100)
    assert set(sh) == set(d.state_dict().keys())
    assert all(tuple(d.state_dict()[k].shape) == tuple(v) for k, v in sh.items())
    enc = TransformerEncoder(TransformerEncoderLayer(512, 8, 2048, 0.1, "relu", False), 6)
    sh = O.detr_param_shapes(512, 2048, 6, False)
    assert set(sh) == set(enc.state_dict().keys())
    assert enc.layers[0] is not enc.layers[1] and enc.layers[0].linear1.weight is not enc.layers[1].linear1.weight
    with pytest.raises(RuntimeError):
        TransformerEncoderLayer(512, 8, 2048, 0.1, "swish", False)
    from vitb200.detr import TransformerDecoder, TransformerDecoderLayer
    dec = TransformerDecoder(TransformerDecoderLayer(512, 8, 2048, 0.1, "relu", False), 6, torch.nn.LayerNorm(512), return_intermediate=True)
    sh = O.detr_decoder_param_shapes(512, 2048, 6)
    assert set(sh) == set(dec.state_dict().keys())
    assert all(tuple(dec.state_dict()[k].shape) == tuple(v) for k, v in sh.items())
    assert dec.layers[0].multihead_attn is dec.layers[0].multi_head_attn


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("pycocotools", "pycocotools.coco"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools"].coco = sys.modules["pycocotools.coco"]
    from models.image_classification.vanilla_vit import ViT as RefViT
    cfg = dict(image_size=32, patch_size=4, num_layers=3, num_heads=4, hidden_dim=256, mlp_dim=512, num_classes=10)
    sd = O.seeded_state_dict(O.vit_param_shapes(**cfg), 7)
    m = RefViT(32, 4, 3, 4, 256, 512, 0.0, 0.0, 10)
    m.load_state_dict(sd)
    x = O.seeded_images(5, 32, 8)
    y = O.seeded_labels(5, 10, 9)
    ref = m(x)
    torch.nn.functional.cross_entropy(ref, y).backward()
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.vit_forward(osd, x, patch_size=4, num_layers=3, num_heads=4)
    torch.nn.functional.cross_entropy(out, y).backward()
    assert rel_l2(out, ref) < 1e-5
    for n, p in m.named_parameters():
        assert rel_l2(osd[n].grad, p.grad) < 1e-4, n


def _detr_front_oracle(gd):
    """The oracle's restatement of the DETR front end on the fixture's seeded inputs; returns everything the fixture records."""
    import torch.nn.functional as F
    seed, c_in, D = gd["seed"], gd["c_in"], gd["d_model"]
    g = torch.Generator().manual_seed(seed)
    imgs = [torch.randn(3, 96, 128, generator=g), torch.randn(3, 112, 104, generator=g)]
    padded, mask_full = O.nested_tensor_from_tensor_list(imgs)
    stem = torch.randn(c_in, 3, 8, 8, generator=g) * 0.05
    feats = F.conv2d(padded, stem, stride=8).detach().requires_grad_(True)
    mask = F.interpolate(mask_full[None].float(), size=feats.shape[-2:]).to(torch.bool)[0]
    row_w = torch.rand(50, D // 2, generator=g).requires_grad_(True)
    col_w = torch.rand(50, D // 2, generator=g).requires_grad_(True)
    pw = (torch.randn(D, c_in, 1, 1, generator=g) * (1.0 / math.sqrt(c_in))).requires_grad_(True)
    pb = (torch.randn(D, generator=g) * 0.1).requires_grad_(True)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.seeded_state_dict(O.detr_param_shapes(D, gd["ffn"], gd["layers"], False), seed + 1).items()}
    return dict(imgs=imgs, padded=padded, mask_full=mask_full, feats=feats, mask=mask, row_w=row_w, col_w=col_w, pw=pw, pb=pb, sd=sd, g=g)


def test_oracle_detr_front_end_matches_reference_golden():
    """nested_tensor_from_tensor_list (misc.py:307-332), AbsolutePositionalEncoding (detr.py:33-63), input_proj (detr.py:125), the
    flatten / permute of Transformer.forward (transformer.py:49-53) and the encoder behind them, against the fixture generated from
    the reference's own code (tools/make_golden.py::detr_front_case)."""
    gd = load("detr_front_d256.pt")
    t = _detr_front_oracle(gd)
    assert tuple(t["padded"].shape) == tuple(gd["padded_shape"]) and abs(t["padded"].sum().item() - gd["padded_sum"]) < 1e-2
    assert torch.equal(t["padded"][1, 0, 100], gd["padded_1_0_100"]) and torch.equal(t["mask_full"], gd["mask_full"])
    assert torch.equal(t["mask"], gd["mask_feat"])
    n, _, h, w = t["feats"].shape
    import torch.nn.functional as F  # noqa: F401
    pos = O.detr_abs_pos_encoding(t["row_w"], t["col_w"], n, h, w)
    assert abs(pos.sum().item() - gd["pos_sum"]) < 1e-2 and torch.allclose(pos[0, :, 0, 0], gd["pos_00"]) and torch.allclose(pos[1, :, -1, -1], gd["pos_last"])
    src = O.detr_input_proj(t["feats"], t["pw"], t["pb"])
    assert abs(src.norm().item() - gd["src_norm"]) < 1e-3 * gd["src_norm"] and torch.allclose(src[0, 0], gd["src_n0_c0"], atol=1e-5)
    s2, p2, m2 = O.detr_flatten(src, pos, t["mask"])
    out = O.detr_encoder_forward(t["sd"], s2, nhead=gd["nhead"], num_layers=gd["layers"], src_key_padding_mask=m2, pos=p2)
    gout = torch.randn(out.shape, generator=t["g"])
    out.backward(gout)
    assert abs(out.norm().item() - gd["out_norm"]) < 1e-4 * gd["out_norm"] and torch.allclose(out[0], gd["out_row0"], atol=1e-4)
    assert abs(t["feats"].grad.norm().item() - gd["dfeat_norm"]) < 1e-4 * gd["dfeat_norm"]
    assert torch.allclose(t["feats"].grad[1, 3], gd["dfeat_n1_c3"], atol=1e-5)
    assert abs(t["pw"].grad.norm().item() - gd["dW_norm"]) < 1e-4 * gd["dW_norm"] and torch.allclose(t["pw"].grad[0, :, 0, 0], gd["dW_row0"], atol=1e-4)
    assert torch.allclose(t["pb"].grad, gd["db"], atol=1e-4)
    assert torch.allclose(t["row_w"].grad, gd["drow"], atol=1e-4) and torch.allclose(t["col_w"].grad, gd["dcol"], atol=1e-4)
