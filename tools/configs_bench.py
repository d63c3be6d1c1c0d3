"""The other BASELINE.json configs through the drop-in modules (synthetic data of the named shape, random-init weights):
  cfg3  deit_s_distill  DeiT-S/16 with distillation token + DistillationLoss('hard', 0.5, 5.0), 224^2, training step
                        (student forward/backward + Adam; the frozen random-init teacher's logits are an input)
  cfg4  vit_l_infer     ViT-L/16 224^2 inference throughput, batch sweep
  cfg5  detr_enc        DETR TransformerEncoder (512 wide, 8 heads, 2048 FFN, 6 layers) on S = 25 x 42 = 1050 tokens with COCO-like
                        padding masks and positional encodings, training step, data-parallel over N GPUs
plus cfg1's model in a reference-style loop and the DETR decoder.  CUDA-event timed after warm-up.

  python tools/configs_bench.py [names...]                      # human-readable lines (1 GPU)
  python bench.py --config {deit_s_distill,vit_l_infer,detr_enc} [--gpus N ...]   # one bench-contract JSON line
bench.py's default run embeds `summary()` as "other_configs".
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

GFLOP = {"deit_s_distill": 27.745, "vit_l_infer": 123.109, "detr_enc": 159.55}   # per image, SURVEY.md §8(d)


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


# ---------------------------------------------------------------------------------------------------- cfg3
def deit_s_step(dev, B=256, kind="hard"):
    """Returns (step(), model): deit.py:57-70's loop body with the fused trainer (Trainer(distillation=...)); the teacher is a frozen
    random-init linear probe on 8x8-pooled pixels (its cost is negligible: the number is the student's step)."""
    from vitb200.deit import VisionTransformerDistilled
    from vitb200.trainer import Trainer
    m = VisionTransformerDistilled(img_size=224, patch_size=16, depth=12, num_heads=6, embed_dim=384, mlp_ratio=4, num_classes=1000).to(dev).train()
    m.set_distilled_training(True)
    teacher = torch.nn.Sequential(torch.nn.AvgPool2d(28), torch.nn.Flatten(1), torch.nn.Linear(3 * 8 * 8, 1000)).to(dev).eval()
    for p in teacher.parameters():
        p.requires_grad_(False)
    tr = Trainer(m, lr=1e-4, distillation=dict(teacher=teacher, type=kind, alpha=0.5, tau=5.0))
    x = torch.randn(B, 3, 224, 224, device=dev)
    y = torch.randint(0, 1000, (B,), device=dev)
    return (lambda: tr.step(x, y)), m


def deit_s(dev="cuda", B=256):
    step, _ = deit_s_step(dev, B)
    ms = timed(step)
    return {"ms_per_step": ms, "images_per_sec": B / ms * 1e3, "tflops": B / ms * GFLOP["deit_s_distill"], "batch": B,
            "what": "DeiT-S/16 distilled training step (fwd + DistillationLoss('hard') + bwd + Adam, fused Trainer)"}


# ---------------------------------------------------------------------------------------------------- cfg4
def vit_l_inference(dev="cuda", batches=(1, 8, 64, 256, 1024)):
    from vitb200.vit import ViT
    m = ViT(224, 16, 24, 16, 1024, 4096, 0.0, 0.0, 1000).to(dev).eval()
    out = []
    for B in batches:
        x = torch.randn(B, 3, 224, 224, device=dev)
        with torch.no_grad():
            ms = timed(lambda: m(x), warm=3, iters=10 if B < 1024 else 4)
        out.append({"batch": B, "ms": ms, "images_per_sec": B / ms * 1e3, "tflops": B / ms * GFLOP["vit_l_infer"]})
        del x
    return out


# ---------------------------------------------------------------------------------------------------- cfg5
def coco_like_mask(N, gen, dev):
    """True = padding: each image occupies the top-left (h, w) in [0.6, 1] x (25, 42) of the 25 x 42 feature grid (misc.py:307-332)."""
    hv = (torch.rand(N, generator=gen) * 0.4 + 0.6) * 25
    wv = (torch.rand(N, generator=gen) * 0.4 + 0.6) * 42
    yy, xx = torch.meshgrid(torch.arange(25), torch.arange(42), indexing="ij")
    return ((yy[None] >= hv[:, None, None]) | (xx[None] >= wv[:, None, None])).reshape(N, 1050).to(dev)


def detr_encoder_step(dev, N=4, reducer=None, seed=0):
    from vitb200.detr import TransformerEncoder, TransformerEncoderLayer
    S = 1050
    torch.manual_seed(0)
    enc = TransformerEncoder(TransformerEncoderLayer(512, 8, 2048, 0.0, "relu", False), 6).to(dev).train()
    g = torch.Generator().manual_seed(1234 + seed)
    src = torch.randn(S, N, 512, generator=g).to(dev).requires_grad_(True)
    pos = torch.randn(S, N, 512, generator=g).to(dev)
    mask = coco_like_mask(N, g, dev)
    opt = torch.optim.Adam(enc.parameters(), lr=1e-4)
    if reducer is not None:
        reducer.attach(enc._get_engine())

    def step():
        opt.zero_grad()
        out = enc(src, src_key_padding_mask=mask, pos=pos)
        loss = out.float().square().mean()
        if reducer is not None:
            loss = loss / reducer.world_size        # SUM all-reduce of the ranks' gradients = gradient of the global mean
            reducer.begin_step()
        loss.backward()
        if reducer is not None:
            reducer.finish_step()
        opt.step()
        return loss
    return step, enc


def detr_encoder(dev="cuda", N=4):
    step, _ = detr_encoder_step(dev, N)
    ms = timed(step)
    return {"ms_per_step": ms, "images_per_sec": N / ms * 1e3, "tflops": N / ms * GFLOP["detr_enc"], "batch": N,
            "what": "DETR encoder training step S=1050 (fwd + bwd through autograd + torch.optim.Adam), masks + pos"}


def tiny_reference_loop(dev="cuda"):
    """cfg1's model (utils/args.py:6-7: 32x32, patch 4, 7 layers, 256 wide) driven the way the reference's own loop drives it
    (base.py:51-57: zero_grad / model(images) / CrossEntropyLoss / backward / torch.optim.Adam), with and without the CUDA-graph replay
    of the autograd node."""
    from vitb200.vit import ViT
    B, res = 256, {}
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
        res["graphs_on" if mode == "1" else "graphs_off"] = {"ms_per_step": ms, "images_per_sec": B / ms * 1e3}
    os.environ.pop("VITB200_AUTOGRAD_GRAPH", None)
    return res


def detr_decoder(dev="cuda"):
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
    return {"ms_per_step": ms, "images_per_sec": N / ms * 1e3, "batch": N}


def summary(dev, peak_tflops):
    """Bounded measurements of cfg3 / cfg4 / cfg5 for bench.py's default line ("other_configs")."""
    out = {}
    r = deit_s(dev)
    r["frac_of_burst_peak"] = r["tflops"] / peak_tflops
    out["cfg3_deit_s16_distilled_train"] = r
    torch.cuda.empty_cache()
    sweep = vit_l_inference(dev, batches=(1, 64, 256))
    for e in sweep:
        e["frac_of_burst_peak"] = e["tflops"] / peak_tflops
    out["cfg4_vit_l16_inference"] = sweep
    torch.cuda.empty_cache()
    r = detr_encoder(dev)
    r["frac_of_burst_peak"] = r["tflops"] / peak_tflops
    out["cfg5_detr_encoder_s1050_train"] = r
    torch.cuda.empty_cache()
    return out


def bench_line(args):
    """`bench.py --config X --gpus N --steps K --warmup W`: one bench-contract JSON line for configs[2] / [3] / [4]."""
    import torch.distributed as dist
    sys.stdout.flush()
    json_fd = os.dup(1)      # stdout carries exactly one line; C-level prints of libraries (NCCL's banner) go to stderr
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    reducer = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group(backend="nccl", device_id=dev)
    W, K = max(3, args.warmup), args.steps
    name = args.config
    if name == "deit_s_distill":
        B = 256
        if world > 1:
            from vitb200.dp import GradReducer
            reducer = GradReducer()
        from vitb200.deit import VisionTransformerDistilled
        from vitb200.trainer import Trainer
        torch.manual_seed(0)
        m = VisionTransformerDistilled(img_size=224, patch_size=16, depth=12, num_heads=6, embed_dim=384, mlp_ratio=4, num_classes=1000).to(dev).train()
        m.set_distilled_training(True)
        teacher = torch.nn.Sequential(torch.nn.AvgPool2d(28), torch.nn.Flatten(1), torch.nn.Linear(3 * 8 * 8, 1000)).to(dev).eval()
        tr = Trainer(m, lr=1e-4, distillation=dict(teacher=teacher, type="hard", alpha=0.5, tau=5.0), reducer=reducer)
        g = torch.Generator().manual_seed(1234 + rank)
        hx, hy = torch.randn(B, 3, 224, 224, generator=g).pin_memory(), torch.randint(0, 1000, (B,), generator=g).pin_memory()
        x, y = hx.to(dev), hy.to(dev)
        step = lambda: tr.step(x, y)
        e2e_step = lambda: tr.step(hx, hy)
        h2d = hx.numel() * 4 + hy.numel() * 8
        metric, unit, workload = "DeiT-S/16 distilled train images/sec", "images/sec", \
            "DeiT-S/16 + distillation token, 224x224, DistillationLoss('hard', 0.5, 5.0) with a frozen random-init teacher (BASELINE.json configs[2])"
    elif name == "detr_enc":
        B = 4
        if world > 1:
            from vitb200.dp import GradReducer
            reducer = GradReducer(overlap=False)      # launch-bound model: graph-replayed forward / backward + one all-reduce
        step, enc = detr_encoder_step(dev, B, reducer=reducer, seed=rank)
        e2e_step, h2d = None, 0
        metric, unit, workload = "DETR encoder train images/sec", "images/sec", \
            "DETR TransformerEncoder 6 x (512, 8 heads, 2048), S = 1050 tokens of an 800x1333 image at stride 32, masks + pos (BASELINE.json configs[4])"
    else:
        B = 256
        from vitb200.vit import ViT
        m = ViT(224, 16, 24, 16, 1024, 4096, 0.0, 0.0, 1000).to(dev).eval()
        hx = torch.randn(B, 3, 224, 224).pin_memory()
        x = hx.to(dev)

        def step():
            with torch.no_grad():
                return m(x)

        def e2e_step():
            with torch.no_grad():
                return m(hx.to(dev, non_blocking=True))
        h2d = hx.numel() * 4
        metric, unit, workload = "ViT-L/16 inference images/sec", "images/sec", "ViT-L/16 224x224 eval forward, batch 256 per GPU (BASELINE.json configs[3]); replicas only"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(fn, n):
        for _ in range(W):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), r
    ms, last = run(step, K)
    value = world * B * K / (ms / 1e3)
    e2e = None
    if e2e_step is not None:
        host = torch.zeros(1).pin_memory()

        def e2e_fn():
            r = e2e_step()
            host.copy_(r.flatten()[:1].float(), non_blocking=True)
            return r
        ms2, _ = run(e2e_fn, K)
        e2e = {"value": world * B * K / (ms2 / 1e3), "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
    if rank == 0:
        sys.path.insert(0, ROOT)
        from bench import measured_peaks
        peak, peak_sus, _, src = measured_peaks()
        tf = value / world * GFLOP[name] / 1e3
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2_policy": "activations exceed the 126 MB L2; no explicit flush"},
                "e2e": e2e, "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                                         "kernel": "whole step, model FLOPs of SURVEY.md §8(d)", "peak_kind": f"bf16_tflops burst ({src})"}}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        import threading
        sys.stdout.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()      # a finished measurement must not become a hung job at teardown
        if name == "deit_s_distill":
            tr.close()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    only = sys.argv[1:]
    for f in (tiny_reference_loop, deit_s, vit_l_inference, detr_encoder, detr_decoder):
        if only and f.__name__ not in only:
            continue
        try:
            print(f.__name__, json.dumps(f()), flush=True)
        except Exception as e:  # keep going: each config is independent
            print(f"{f.__name__}: FAILED {type(e).__name__}: {e}", flush=True)
        torch.cuda.empty_cache()
