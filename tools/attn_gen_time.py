"""Timing of the general attention path (DETR shapes: sequence-first, key-padding masks, S = 1050; decoder cross-attention 100 x 1050).
Usage: python tools/attn_gen_time.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, D = 8, 512
torch.manual_seed(0)


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (Sq, Sk, name) in ((1050, 1050, "encoder self-attention"), (100, 1050, "decoder cross-attention"), (4200, 4200, "encoder S=4200")):
    if Sq == 4200 and N > 2:
        continue
    q = torch.randn(Sq * N, D, device="cuda").bfloat16()
    k = torch.randn(Sk * N, D, device="cuda").bfloat16()
    v = torch.randn(Sk * N, D, device="cuda").bfloat16()
    o = torch.empty_like(q)
    do = torch.randn_like(q)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    lse = torch.empty(N, H, Sq, device="cuda")
    delta = torch.empty(N, H, Sq, device="cuda")
    kpm = torch.zeros(N, Sk, dtype=torch.uint8, device="cuda")
    kpm[:, int(Sk * 0.85):] = 1
    kw = dict(B=N, H=H, S=Sq, tok_stride=N, batch_stride=1, key_padding_mask=kpm, S_kv=Sk)
    tf = timeit(lambda: ops.attention_fwd(q, k, v, o, lse, **kw))
    tb = timeit(lambda: ops.attention_bwd(q, k, v, o, lse, do, dq, dk, dv, delta, **kw))
    fl = 4.0 * Sq * Sk * 64 * N * H
    print(f"{name:26s} N={N} Sq={Sq} Sk={Sk}: fwd {tf:7.1f} us ({fl / tf / 1e6:6.0f} TF/s)  bwd {tb:7.1f} us ({2.5 * fl / tb / 1e6:6.0f} TF/s)")
