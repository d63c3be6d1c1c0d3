"""Small attention-only workload for ncu: H=12, S=197, forward + backward, 3 iterations.  Usage: attn_prof.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops
B, H, S = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 12, 197
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
for _ in range(3):
    ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
    ops.attention_bwd(q, k, v, o, lse, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], delta, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
torch.cuda.synchronize()
print("done")
