"""Installs the B200 classes behind the reference's own import paths, so the reference's unmodified training loops,
``DistillationLoss`` and notebook drive the hand-written kernels.

    import vitb200.dropin as dropin
    dropin.install("/path/to/vision-transformers")          # the reference checkout
    from models.image_classification.vanilla_vit import ViT  # now the sm_100a implementation
    model = ViT(**get_args("vit_tiny_cifar10")); model.train_model(model, train_loader, test_loader, 50, val_loader)

What is rebound:
  models.image_classification.vanilla_vit.{ViT, Encoder, EncoderBlock, MLPBlock, MLP}   (vanilla_vit.py:22-215)
  models.object_detection.transformer.{TransformerEncoderLayer, TransformerEncoder}      (transformer.py:98-115,192-247)
  models.object_detection.transformer.{TransformerDecoderLayer, TransformerDecoder}      (transformer.py:66-95,118-189)
  timm.models.deit.VisionTransformerDistilled (a shim module, since timm is what deit.py:4 imports)
  models.image_classification.{cpe_vit.CPEViT, cpvt.CPVT, cpvt_gap.CPVTGAP}
The rebound ``ViT`` subclasses the reference's ``BaseTransformer`` (base.py:12) and borrows the reference's
``ViT.train_model`` function object (vanilla_vit.py:217), so the training loop that runs is the reference's own code.
Nothing is copied from the reference tree.
"""
import importlib
import sys
import types


def _stub(name, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def install(reference_root, stub_missing=True):
    """Returns a dict of the rebound classes."""
    from . import deit as our_deit
    from . import detr as our_detr
    from . import vit as our_vit

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    sys.dont_write_bytecode = True
    if stub_missing:
        try:
            importlib.import_module("pycocotools.coco")
        except Exception:   # utils/load_data.py:6 imports pycocotools at module top; only the COCO loader needs it
            _stub("pycocotools")
            _stub("pycocotools.coco", COCO=object)
            sys.modules["pycocotools"].coco = sys.modules["pycocotools.coco"]
    ref_vit = importlib.import_module("models.image_classification.vanilla_vit")
    ref_base = importlib.import_module("models.image_classification.base")
    ref_tr = importlib.import_module("models.object_detection.transformer")

    ref_train_model = ref_vit.ViT.__dict__.get("train_model")

    class ViT(our_vit.ViT, ref_base.BaseTransformer):
        __doc__ = our_vit.ViT.__doc__

    if ref_train_model is not None:
        ViT.train_model = ref_train_model
    ViT.__module__ = ref_vit.__name__
    ViT.__qualname__ = "ViT"
    ref_vit.ViT = ViT
    ref_vit.Encoder = our_vit.Encoder
    ref_vit.EncoderBlock = our_vit.EncoderBlock
    ref_vit.MLPBlock = our_vit.MLPBlock
    ref_vit.MLP = our_vit.MLP
    ref_tr.TransformerEncoderLayer = our_detr.TransformerEncoderLayer
    ref_tr.TransformerEncoder = our_detr.TransformerEncoder
    # the decoder: the reference layer cannot run as written (transformer.py:122 vs :148); ours resolves the name
    ref_tr.TransformerDecoderLayer = our_detr.TransformerDecoderLayer
    ref_tr.TransformerDecoder = our_detr.TransformerDecoder
    # Transformer (transformer.py:25-63): the reference's constructor; its forward ends in two typos (memory.permte, hs.transpose(1, 1)),
    # ours is the forward the code spells out.  AbsolutePositionalEncoding / input_proj / NestedTensor live in detr.py, which does not
    # parse (:155) and therefore cannot be rebound: import them from vitb200.detr_front.
    from . import detr_front as our_front
    ref_tr.Transformer = our_front.Transformer

    # timm shim for models/image_classification/deit.py:4-5
    have_timm = True
    try:
        importlib.import_module("timm.models.deit")
    except Exception:
        have_timm = False
    if have_timm:
        sys.modules["timm.models.deit"].VisionTransformerDistilled = our_deit.VisionTransformerDistilled
    elif stub_missing:
        def create_model(name, *a, **k):
            raise RuntimeError(f"timm is not installed: cannot create teacher model '{name}'; pass your own teacher nn.Module "
                               "to utils.distillation_loss.DistillationLoss")
        import torch

        class DropPath(torch.nn.Module):
            """Stochastic depth per sample (timm.models.layers.DropPath semantics), needed by token_transformer.py:7."""

            def __init__(self, drop_prob=0.0, scale_by_keep=True):
                super().__init__()
                self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1.0 - self.drop_prob
                mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
                if keep > 0.0 and self.scale_by_keep:
                    mask.div_(keep)
                return x * mask

        _stub("timm")
        _stub("timm.models", create_model=create_model)
        sys.modules["timm.models"].__path__ = []   # a package, so that `timm.models.layers` / `.deit` resolve through sys.modules
        _stub("timm.models.deit", VisionTransformerDistilled=our_deit.VisionTransformerDistilled)
        _stub("timm.models.layers", DropPath=DropPath)
        sys.modules["timm"].models = sys.modules["timm.models"]
        sys.modules["timm.models"].deit = sys.modules["timm.models.deit"]
        sys.modules["timm.models"].layers = sys.modules["timm.models.layers"]
    # T2T_ViT carries a verbatim copy of the encoder classes (t2t_vit.py:25-110) and calls self.encoder(x) on its own tokens
    # (SURVEY.md §8 a13): the stand-alone Encoder.forward serves it.
    try:
        ref_t2t = importlib.import_module("models.image_classification.t2t_vit")
        ref_t2t.Encoder, ref_t2t.EncoderBlock = our_vit.Encoder, our_vit.EncoderBlock
        ref_t2t.MLPBlock, ref_t2t.MLP = our_vit.MLPBlock, our_vit.MLP
    except Exception:   # optional: its other imports (token_performer, load_data) may be unavailable
        pass
    # CPE-ViT / CPVT / CPVT-GAP: the model classes are rebound and keep the reference's own train_model functions
    # (cpe_vit.py:214, cpvt.py:215, cpvt_gap.py:216); the reference's helper classes (its own EncoderBlock / PEG) are left untouched.
    from . import cpvt as our_cpvt
    rebound = {}
    for modname, clsname in (("cpe_vit", "CPEViT"), ("cpvt", "CPVT"), ("cpvt_gap", "CPVTGAP")):
        try:
            ref_mod = importlib.import_module("models.image_classification." + modname)
        except Exception:   # optional siblings: their data-loader imports may be unavailable
            continue
        ref_cls = getattr(ref_mod, clsname)
        new_cls = type(clsname, (getattr(our_cpvt, clsname),), {"__doc__": getattr(our_cpvt, clsname).__doc__})
        tm = ref_cls.__dict__.get("train_model")
        if tm is not None:
            new_cls.train_model = tm
        new_cls.__module__ = ref_mod.__name__
        setattr(ref_mod, clsname, new_cls)
        rebound[clsname] = new_cls
    return {**rebound, "ViT": ViT, "TransformerEncoder": our_detr.TransformerEncoder, "TransformerEncoderLayer": our_detr.TransformerEncoderLayer,
            "VisionTransformerDistilled": our_deit.VisionTransformerDistilled}
