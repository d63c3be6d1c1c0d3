"""Quick end-to-end check on the GPU (same as tests/test_vit_parity_gpu.py but prints every tensor's error)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from helpers import O, oracle_vit_run, rel_l2
from test_vit_parity_gpu import TINY, B16_2L, build

for cfg, batch in ((TINY, 8), (B16_2L, 3)):
    m, sd = build(cfg, 1)
    images = O.seeded_images(batch, cfg["image_size"], 2)
    labels = O.seeded_labels(batch, cfg["num_classes"], 3)
    ref_logits, ref_loss, ref_grads = oracle_vit_run(cfg, sd, images, labels)
    m.train()
    logits = m(images.cuda())
    loss = torch.nn.functional.cross_entropy(logits, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    print(f"logits rel {rel_l2(logits, ref_logits):.3e}  loss {loss.item():.6f} ref {ref_loss.item():.6f}", flush=True)
    for n, p in m.named_parameters():
        print(f"  {n:70s} {rel_l2(p.grad, ref_grads[n]):.3e}  |ref|={ref_grads[n].norm().item():.3e}")
