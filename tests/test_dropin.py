"""Drop-in installation behind the reference's import paths (needs the reference tree: build container only)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VITB200_REFERENCE", "/root/reference")

CODE = r"""
import sys
sys.path.insert(0, %r)
import vitb200.dropin as dropin
classes = dropin.install(%r)
from models.image_classification.vanilla_vit import ViT
from models.image_classification.base import BaseTransformer
from models.object_detection.transformer import TransformerEncoder, TransformerEncoderLayer
from utils.args import get_args
import vitb200.vit, vitb200.detr
assert issubclass(ViT, BaseTransformer) and issubclass(ViT, vitb200.vit.ViT)
args = dict(get_args("vit_tiny_cifar10")); args["dropout"] = 0.0; args["attention_dropout"] = 0.0
m = ViT(**args)
assert hasattr(m, "train_model") and m.train_model.__func__.__module__ == "models.image_classification.vanilla_vit"
assert m.device in ("cuda", "cpu", "mps")
assert TransformerEncoder is vitb200.detr.TransformerEncoder
from models.object_detection.transformer import Transformer, TransformerDecoder
assert TransformerDecoder is vitb200.detr.TransformerDecoder
t = Transformer(d_model=256, nhead=4, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=512)   # the reference's own wrapper
assert type(t.encoder) is vitb200.detr.TransformerEncoder and type(t.decoder) is vitb200.detr.TransformerDecoder
assert "decoder.layers.0.multi_head_attn.in_proj_weight" in t.state_dict() and "decoder.norm.weight" in t.state_dict()
from models.image_classification import deit   # imports timm.models.deit.VisionTransformerDistilled through the shim
assert deit.VisionTransformerDistilled.__module__ == "vitb200.deit"
from models.image_classification import t2t_vit
assert t2t_vit.Encoder is vitb200.vit.Encoder and t2t_vit.EncoderBlock is vitb200.vit.EncoderBlock
from models.image_classification import cpvt, cpe_vit, cpvt_gap
import vitb200.cpvt
for mod, name in ((cpe_vit, "CPEViT"), (cpvt, "CPVT"), (cpvt_gap, "CPVTGAP")):
    cls = getattr(mod, name)
    assert issubclass(cls, getattr(vitb200.cpvt, name)) and cls.__module__ == mod.__name__
    assert cls.train_model.__module__ == mod.__name__          # the reference's own loop
    c = cls(**args)
    assert c.device == "cuda" and len(c.state_dict()) > 90
print("DROPIN_OK", len(m.state_dict()))
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present (GPU box)")
def test_install_rebinds_reference_modules():
    r = subprocess.run([sys.executable, "-c", CODE % (ROOT, REF)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "DROPIN_OK 92" in r.stdout


def test_transformer_wrapper_keys_match_reference_constructor():
    """vitb200.detr_front.Transformer vs the reference's own Transformer.__init__ (transformer.py:25-46): same parameter names and shapes
    (the forward differs only by the reference's typos), and AbsolutePositionalEncoding / InputProjection carry the keys of detr.py:41-42,125."""
    if not os.path.isdir(REF):
        import pytest
        pytest.skip("reference tree not present")
    sys.path.insert(0, REF)
    src = open(os.path.join(REF, "models", "object_detection", "transformer.py")).read()
    ns = {}
    exec(compile(src, "ref_transformer", "exec"), ns)          # a private copy of the unmodified module (dropin may have rebound the live one)
    ref = ns["Transformer"](d_model=256, nhead=4, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=512)
    from vitb200.detr_front import AbsolutePositionalEncoding, InputProjection, Transformer
    ours = Transformer(d_model=256, nhead=4, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=512)
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert a == b
    assert set(AbsolutePositionalEncoding(128).state_dict()) == {"row_embed.weight", "col_embed.weight"}
    assert {k: tuple(v.shape) for k, v in InputProjection(2048, 256, kernel_size=1).state_dict().items()} == {"weight": (256, 2048, 1, 1), "bias": (256,)}
