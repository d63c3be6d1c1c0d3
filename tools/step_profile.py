"""Per-op GPU time of one eager ViT-B/16 training step (CUDA events around every C-ABI call), after warm-up.
Usage: python tools/step_profile.py [batch] [vit_b16 | deit_s | vit_l16]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from vitb200 import ops
from vitb200.trainer import Trainer
from vitb200.vit import ViT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else "vit_b16"
torch.manual_seed(0)
if which == "detr":      # BASELINE.json configs[4]: DETR encoder, S = 1050, autograd path + torch.optim.Adam
    import configs_bench
    step_fn, model = configs_bench.detr_encoder_step("cuda", B if B <= 16 else 4)
elif which == "deit_s":     # BASELINE.json configs[2]: DeiT-S/16 with distillation token (student step, plain CE on the class head here)
    from vitb200.deit import VisionTransformerDistilled
    model = VisionTransformerDistilled(img_size=224, patch_size=16, depth=12, num_heads=6, embed_dim=384, mlp_ratio=4, num_classes=1000)
elif which == "vit_l16":
    model = ViT(224, 16, 24, 16, 1024, 4096, 0.0, 0.0, 1000)
    with torch.no_grad():
        model.heads.head.weight.normal_(std=0.02)
else:
    model = ViT(224, 16, 12, 12, 768, 3072, 0.0, 0.0, 1000)
    with torch.no_grad():
        model.heads.head.weight.normal_(std=0.02)
if which == "detr":
    os.environ["VITB200_AUTOGRAD_GRAPH"] = "0"
    step = step_fn
else:
    model = model.cuda().train()
    tr = Trainer(model, use_cuda_graph=False)
    images = torch.randn(B, 3, 224, 224, device="cuda")
    labels = torch.randint(0, 1000, (B,), device="cuda")
    step = lambda: tr.step(images, labels)
for _ in range(5):
    step()
torch.cuda.synchronize()

rec = []
names = ["gemm", "layernorm_fwd", "layernorm_bwd", "attention_fwd", "attention_bwd", "cast_bf16", "patchify", "token_rows", "colsum_bf16",
         "embed_bwd", "cross_entropy", "adam_step", "add_cast_bf16", "add3", "dropout_f32"]
orig = {n: getattr(ops, n) for n in names}


def wrap(n):
    f = orig[n]

    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = f(*a, **k)
        e1.record()
        tag = n
        if n == "gemm":
            tag = f"gemm maj=({k.get('a_major', 0)},{k.get('b_major', 0)}) epi={k.get('epilogue', 0)}"
        rec.append((tag, e0, e1))
        return r
    return g


for n in names:
    setattr(ops, n, wrap(n))
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    rec.clear()
    s0.record()
    step()
    s1.record()
    torch.cuda.synchronize()
tot = collections.Counter()
cnt = collections.Counter()
for tag, a, b in rec:
    tot[tag] += a.elapsed_time(b)
    cnt[tag] += 1
total = s0.elapsed_time(s1)
ksum = sum(tot.values())
print(f"{which}: step {total:.2f} ms, sum of bracketed ops {ksum:.2f} ms, batch {B}")
if os.environ.get("STEP_PROFILE_LIST"):
    want = os.environ["STEP_PROFILE_LIST"]
    print(want, "per call (us):", " ".join(f"{a.elapsed_time(b) * 1e3:.0f}" for tag, a, b in rec if tag.startswith(want)))
for k, v in tot.most_common():
    print(f"  {k:34s} n={cnt[k]:3d} {v:8.3f} ms {100 * v / total:5.1f}%")
