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


def tiny_reference_loop():
    """cfg1's model (utils/args.py:6-7: 32x32, patch 4, 7 layers, 256 wide) driven the way the reference's own loop drives it
    (base.py:51-57: zero_grad / model(images) / CrossEntropyLoss / backward / torch.optim.Adam), with and without the CUDA-graph replay
    of the autograd node."""
    from vitb200.vit import ViT
    B = 256
    for mode in ("0", "1"):
        os.environ["VITB200_AUTOGRAD_GRAPH"] = mode
        m = ViT(32, 4, 7, 4, 256, 512, 0.1, 0.1, 10).to(dev).train()
        with torch.no_grad():
            m.heads.head.weight.normal_(std=0.02)
        opt = torch.optim.Adam(m.parameters(), lr=1e-4)
        crit = torch.nn.CrossEntropyLoss()
        x, y = torch.randn(B, 3, 32, 32, device=dev), torch.randint(0, 10, (B,), device=dev)

        def step():
            opt.zero_grad()
            loss = crit(m(x), y)
            loss.backward()
            opt.step()
        ms = timed(step, warm=4, iters=20)
        print(f"cfg1 tiny ViT (CIFAR shape, dropout 0.1) reference-style loop, batch {B}, autograd graphs {'on ' if mode == '1' else 'off'}: "
              f"{ms:6.2f} ms/step {B / ms * 1e3:9.1f} images/s", flush=True)
    os.environ.pop("VITB200_AUTOGRAD_GRAPH", None)


def detr_decoder():
    from vitb200.detr import TransformerDecoder, TransformerDecoderLayer
    Q, S, N = 100, 1050, 4
    dec = TransformerDecoder(TransformerDecoderLayer(512, 8, 2048, 0.0, "relu", False), 6, torch.nn.LayerNorm(512), return_intermediate=True).to(dev).train()
    tgt = torch.zeros(Q, N, 512, device=dev)
    mem = torch.randn(S, N, 512, device=dev, requires_grad=True)
    pos, qpos = torch.randn(S, N, 512, device=dev), torch.randn(Q, N, 512, device=dev, requires_grad=True)
    mask = torch.zeros(N, S, dtype=torch.bool, device=dev)
    mask[:, 900:] = True
    opt = torch.optim.Adam(dec.parameters(), lr=1e-4)

    def step():
        opt.zero_grad()
        out = dec(tgt, mem, memory_key_padding_mask=mask, pos=pos, query_pos=qpos)
        out.float().square().mean().backward()
        opt.step()
    ms = timed(step)
    print(f"DETR decoder training Q=100 S=1050 batch {N} (6 layers, intermediates): {ms:7.2f} ms/step {N / ms * 1e3:9.1f} images/s", flush=True)


only = sys.argv[1:]
for f in (tiny_reference_loop, deit_s, vit_l_inference, detr_encoder, detr_decoder):
    if only and f.__name__ not in only:
        continue
    try:
        f()
    except Exception as e:  # keep going: each config is independent
        print(f"{f.__name__}: FAILED {type(e).__name__}: {e}", flush=True)
    torch.cuda.empty_cache()
