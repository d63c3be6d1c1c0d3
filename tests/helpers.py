"""Shared helpers for the parity tests: relative-L2 comparison and oracle runs (tests may import oracle/)."""
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import vit_oracle as O  # noqa: E402


@contextlib.contextmanager
def strict_fp32():
    """Runs the oracle's PyTorch operators on the GPU in true fp32 (no TF32 in matmuls or cuDNN convolutions) — used for the parity
    cases at benchmark size, where the CPU would need minutes."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.get_float32_matmul_precision())
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old[0], old[1]
        torch.set_float32_matmul_precision(old[2])


def rel_l2(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def oracle_vit_run(cfg, sd, images, labels=None, want="logits", grad_out=None):
    """Runs the CPU oracle forward (+ backward). Returns (output, loss, grads dict)."""
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    kw = dict(patch_size=cfg["patch_size"], num_layers=cfg["num_layers"], num_heads=cfg["num_heads"])
    if want == "logits":
        out = O.vit_forward(sd, images, **kw)
    else:
        out = O.vit_forward_features(sd, images, **kw)
    loss = None
    if labels is not None:
        loss = torch.nn.functional.cross_entropy(out, labels)
        loss.backward()
    elif grad_out is not None:
        out.backward(grad_out)
    grads = {k: v.grad for k, v in sd.items() if v.grad is not None}
    return out.detach(), (loss.detach() if loss is not None else None), grads
