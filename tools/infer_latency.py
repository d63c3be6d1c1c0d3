"""ViT-L/16 eval latency at small batch, default vs VITB200_INFER_SPLIT_K=1 (run once per setting: the switch is read at import).
Usage: python tools/infer_latency.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vitb200.vit import ViT

m = ViT(224, 16, 24, 16, 1024, 4096, 0.0, 0.0, 1000)
with torch.no_grad():
    m.heads.head.weight.normal_(std=0.02)
m = m.cuda().eval()
ref = None
for B in (1, 2, 4, 8):
    x = torch.randn(B, 3, 224, 224, device="cuda", generator=torch.Generator(device="cuda").manual_seed(B))
    with torch.no_grad():
        for _ in range(4):
            y = m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            y = m(x)
        e1.record()
        torch.cuda.synchronize()
    print(f"ViT-L/16 eval batch {B}: {e0.elapsed_time(e1) / 20:6.3f} ms  (split-K {os.environ.get('VITB200_INFER_SPLIT_K', '0')}); logits checksum {y.double().sum().item():.6f}",
          flush=True)
