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
