"""Isolated timing of the attention forward / backward kernels at the ViT-B/16 bench shape (CUDA events, L2-flushed inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 12
S = int(sys.argv[3]) if len(sys.argv) > 3 else 197
D = H * 64
M = B * S
torch.manual_seed(0)
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
o = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device="cuda")
do = torch.randn(M, D, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
delta = torch.empty(B, H, S, device="cuda")
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timeit(f, n=10):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


fwd = lambda: ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
cs = torch.zeros(3 * D, device="cuda") if os.environ.get("ATTN_TIME_COLSUM") else None
bwd = lambda: ops.attention_bwd(q, k, v, o, lse, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], delta, B=B, H=H, S=S, tok_stride=1, batch_stride=S,
                                dqkv_colsum=cs)
tf, tb = timeit(fwd), timeit(bwd)
fl = 4.0 * S * S * 64 * B * H
print(f"B={B} H={H} S={S} : fwd {tf*1e3:.1f} us ({fl/tf/1e9:.0f} TF/s)  bwd {tb*1e3:.1f} us ({2.5*fl/tb/1e9:.0f} TF/s)")
