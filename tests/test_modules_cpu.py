"""CPU checks of the drop-in modules' host-side behaviour: pickling / copying never drags the engine (flat-buffer bookkeeping,
workspaces, captured CUDA graphs) along, and state survives the round trip."""
import copy
import io

import torch


def _models():
    from vitb200.cpvt import CPEViT, CPVT
    from vitb200.deit import VisionTransformerDistilled
    from vitb200.detr import TransformerDecoder, TransformerDecoderLayer, TransformerEncoder, TransformerEncoderLayer
    from vitb200.vit import ViT
    return [ViT(32, 4, 2, 4, 256, 512, 0.1, 0.1, 10), CPEViT(32, 4, 2, 4, 256, 512, 0.0, 0.0, 10), CPVT(32, 4, 2, 4, 256, 512, 0.0, 0.0, 10),
            VisionTransformerDistilled(img_size=32, patch_size=16, depth=2, num_heads=6, embed_dim=384, num_classes=10),
            TransformerEncoder(TransformerEncoderLayer(256, 4, 512, 0.1, "relu", True), 2, torch.nn.LayerNorm(256)),
            TransformerDecoder(TransformerDecoderLayer(256, 4, 512, 0.1, "relu", False), 2, torch.nn.LayerNorm(256), return_intermediate=True)]


def test_modules_pickle_and_deepcopy_without_engine():
    for m in _models():
        eng = m._get_engine()
        eng.__dict__["_train_graphs"] = {"k": object()}        # stands in for captured CUDA graphs (not picklable)
        buf = io.BytesIO()
        torch.save(m, buf)
        buf.seek(0)
        m2 = torch.load(buf, weights_only=False)
        assert m2.__dict__.get("_engine") is None and m.__dict__.get("_engine") is eng
        sd, sd2 = m.state_dict(), m2.state_dict()
        assert list(sd) == list(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)
        m3 = copy.deepcopy(m)
        assert m3.__dict__.get("_engine") is None
        assert all(a is not b and torch.equal(a, b) for a, b in zip(m.parameters(), m3.parameters()))
        assert m3._get_engine() is not eng


def test_flat_layout_segments_alignment_and_split_k():
    """Host logic of the engines: parameters are laid out in gradient-production order, 256-byte aligned, in contiguous segments that
    tile the flat buffer (what the data-parallel buckets are cut from); the split-K heuristic fills whole waves."""
    from vitb200.engine import pick_split_k
    for m in _models():
        eng = m._get_engine()
        seen, end = set(), 0
        for key, p in eng._order:
            o = eng.offsets[key]
            assert o % 64 == 0 and o >= end, key          # 64 fp32 elements = 256 bytes; no overlap, production order
            end = o + p.numel()
            assert id(p) not in seen
            seen.add(id(p))
        assert seen == {id(p) for p in m.parameters()}, type(m).__name__      # every nn.Parameter is in the flat buffer exactly once
        assert eng.total >= end and eng.total % 4 == 0
        b = eng.segment_bounds
        assert b[0][0] == 0 and b[-1][1] == eng.total and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
    # the wgrad shapes of ViT-B/16 at batch 256 (788 k-blocks of 64 tokens) on 74 CTA pairs: >= 95 % of whole waves, >= 8 k-blocks a unit
    for tiles in (9, 27, 36):
        s_ = pick_split_k(tiles, 788, 74)
        units = tiles * s_
        assert units / (-(-units // 74) * 74) >= 0.95 and 788 // s_ >= 8, (tiles, s_)
    assert pick_split_k(3, 16, 74) == 2 and pick_split_k(3, 7, 74) == 1                  # never fewer than 8 k-blocks per unit


def test_dropout_sites_are_distinct():
    from vitb200.detr import DetrDecoderEngine, DetrEngine
    from vitb200.engine import VitEngine
    sites = {VitEngine.drop_site(li, s) for li in range(24) for s in range(4)} | {VitEngine.EMBED_SITE}
    assert len(sites) == 24 * 4 + 1
    assert len({DetrDecoderEngine.drop_site(li, s) for li in range(6) for s in range(6)}) == 36
    assert DetrEngine.drop_site(2, 3) == DetrDecoderEngine.drop_site(2, 3)


def test_fused_trainer_refuses_frozen_parameters():
    """The fused flat Adam updates every parameter; frozen ones must fail loudly instead of being trained silently."""
    import pytest
    from vitb200.trainer import FusedAdam
    from vitb200.vit import ViT
    m = ViT(32, 4, 2, 4, 256, 512, 0.0, 0.0, 10)
    m.encoder.pos_embedding.requires_grad_(False)
    with pytest.raises(NotImplementedError):
        FusedAdam(m)._ensure_state()
