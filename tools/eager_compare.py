"""The "library bar" of SURVEY.md §8d: the same ViT-B/16 training step under PyTorch eager on the same B200 (fp32 and
torch.autocast(bf16); cuBLASLt GEMMs + SDPA flash kernels), next to the hand-written path.  The eager model is
torchvision.models.vit_b_16(weights=None) — the class the reference's vanilla_vit.py restates key for key
 — driven by the reference's own loop body (zero_grad, CE, backward, Adam: vanilla_vit.py:235-239).
Usage: python tools/eager_compare.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda"
torch.manual_seed(0)
images = torch.randn(B, 3, 224, 224, device=dev)
labels = torch.randint(0, 1000, (B,), device=dev)


def timed(step, warm=3, iters=8):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def eager(autocast):
    m = torchvision.models.vit_b_16(weights=None).to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)

    def step():
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = torch.nn.functional.cross_entropy(m(images), labels)
        loss.backward()
        opt.step()
    ms = timed(step)
    del m, opt
    torch.cuda.empty_cache()
    return ms


def ours():
    from vitb200.trainer import Trainer
    from vitb200.vit import ViT
    m = ViT(224, 16, 12, 12, 768, 3072, 0.0, 0.0, 1000)
    with torch.no_grad():
        m.heads.head.weight.normal_(std=0.02)
    tr = Trainer(m.to(dev).train(), lr=1e-4)
    return timed(lambda: tr.step(images, labels), warm=5, iters=20)


for name, fn in (("torch eager fp32 (TF32 off)", lambda: eager(False)), ("torch eager autocast bf16", lambda: eager(True)), ("vitb200 (this repo)", ours)):
    ms = fn()
    print(f"{name:32s} batch {B}: {ms:8.2f} ms/step  {B / ms * 1e3:9.1f} images/s", flush=True)
