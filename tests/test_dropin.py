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
