"""Stability soak: many back-to-back launches of the tcgen05 attention kernels at random shapes (hang / race detection), then a long
training run.  Every iteration is bounded by a watchdog (faulthandler) so a hang prints where it is stuck instead of burning the slot."""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops
faulthandler.dump_traceback_later(150, exit=True)
torch.manual_seed(0)
t0 = time.time()
n = 0
while time.time() - t0 < 25:
    B = int(torch.randint(1, 40, (1,)))
    H = int(torch.randint(1, 13, (1,)))
    S = int(torch.randint(1, 209, (1,)))
    D, M = H * 64, B * S
    qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
    do = torch.randn(M, D, device="cuda").bfloat16()
    o = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device="cuda")
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, S, device="cuda")
    cs = torch.zeros(3 * D, device="cuda")
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    for _ in range(3):
        ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
        ops.attention_bwd(q, k, v, o, lse, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], delta, B=B, H=H, S=S, tok_stride=1, batch_stride=S,
                          dqkv_colsum=cs)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all() and torch.isfinite(o.float()).all(), (B, H, S)
    n += 1
print(f"attention soak: {n} random shapes x 3 fwd+bwd OK", flush=True)
from vitb200.trainer import Trainer
from vitb200.vit import ViT
m = ViT(224, 16, 12, 12, 768, 3072, 0.0, 0.0, 1000)
with torch.no_grad():
    m.heads.head.weight.normal_(std=0.02)
tr = Trainer(m.cuda().train(), lr=1e-4)
x = torch.randn(64, 3, 224, 224, device="cuda")
y = torch.randint(0, 1000, (64,), device="cuda")
for i in range(400):
    loss = tr.step(x, y)
torch.cuda.synchronize()
print(f"training soak: 400 steps at batch 64 OK, final loss {loss.item():.4f}", flush=True)
