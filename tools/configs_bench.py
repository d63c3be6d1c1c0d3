"""Throughput of the other BASELINE.json configs through the drop-in modules (synthetic data, random-init weights):
cfg3 DeiT-S/16 distilled training (student step; teacher excluded), cfg4 ViT-L/16 inference batch sweep, cfg5 DETR encoder
training at S = 1050.  CUDA-event timed after warm-up.  Usage: python tools/configs_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

dev = "cuda"


def timed(step, warm=3, iters=10):
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


def deit_s():
    from vitb200.deit import VisionTransformerDistilled
    B = 256
    m = VisionTransformerDistilled(img_size=224, patch_size=16, depth=12, num_heads=6, embed_dim=384, mlp_ratio=4, num_classes=1000).to(dev).train()
    m.set_distilled_training(True)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    x = torch.randn(B, 3, 224, 224, device=dev)
    y = torch.randint(0, 1000, (B,), device=dev)
    t = torch.randint(0, 1000, (B,), device=dev)   # hard teacher labels (argmax of a frozen teacher; teacher forward excluded)

    def step():
        opt.zero_grad()
        out, out_kd = m(x)
        loss = 0.5 * torch.nn.functional.cross_entropy(out, y) + 0.5 * torch.nn.functional.cross_entropy(out_kd, t)   # distillation_loss.py:57-73 (hard)
        loss.backward()
        opt.step()
    ms = timed(step)
    print(f"cfg3 DeiT-S/16 distilled training  batch {B}: {ms:7.2f} ms/step {B / ms * 1e3:9.1f} images/s  ({B / ms * 27.745:.0f} TFLOP/s model FLOPs)", flush=True)


def vit_l_inference():
    from vitb200.vit import ViT
    m = ViT(224, 16, 24, 16, 1024, 4096, 0.0, 0.0, 1000).to(dev).eval()
    for B in (1, 8, 64, 256, 1024):
        x = torch.randn(B, 3, 224, 224, device=dev)
        with torch.no_grad():
            ms = timed(lambda: m(x), warm=3, iters=10 if B < 1024 else 4)
        print(f"cfg4 ViT-L/16 inference           batch {B:4d}: {ms:7.2f} ms {B / ms * 1e3:9.1f} images/s  ({B / ms * 123.109:.0f} TFLOP/s)", flush=True)
        del x


def detr_encoder():
    from vitb200.detr import TransformerEncoder, TransformerEncoderLayer
    S, N = 1050, 4
    enc = TransformerEncoder(TransformerEncoderLayer(512, 8, 2048, 0.0, "relu", False), 6).to(dev).train()
    src = torch.randn(S, N, 512, device=dev, requires_grad=True)
    pos = torch.randn(S, N, 512, device=dev)
    mask = torch.zeros(N, S, dtype=torch.bool, device=dev)
    mask[:, 900:] = True
    opt = torch.optim.Adam(enc.parameters(), lr=1e-4)

    def step():
        opt.zero_grad()
        out = enc(src, src_key_padding_mask=mask, pos=pos)
        out.float().square().mean().backward()
        opt.step()
    ms = timed(step)
    print(f"cfg5 DETR encoder training S=1050 batch {N}: {ms:7.2f} ms/step {N / ms * 1e3:9.1f} images/s  ({N / ms * 159.55:.0f} TFLOP/s)", flush=True)


for f in (deit_s, vit_l_inference, detr_encoder):
    try:
        f()
    except Exception as e:  # keep going: each config is independent
        print(f"{f.__name__}: FAILED {type(e).__name__}: {e}", flush=True)
    torch.cuda.empty_cache()
