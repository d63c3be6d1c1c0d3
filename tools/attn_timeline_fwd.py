"""In-kernel cycle stamps of the tcgen05 attention forward (CTA 0, warp 4 lane 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops, _lib
B, H, S = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 12, 197
D = H * 64; M = B * S
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
o = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device="cuda")
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
dbg = torch.zeros(64 * 16, device="cuda", dtype=torch.int64)
lib = _lib.load()
for rep in range(2):
    dbg.zero_()
    lib.vb_debug_set_attn_timeline(dbg.data_ptr())
    ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
    torch.cuda.synchronize()
lib.vb_debug_set_attn_timeline(None)
t = dbg.view(64, 16).cpu()
t0 = t[0, 0].item()
print("tile: start | wait S | load + row max | exchange | exp, pack, store   (cycles; stamps of warp 4 lane 0)")
for i in range(24):
    r = [x.item() - t0 for x in t[i, :5]]
    print(f"{i:2d}: start {r[0]:7d} | waitS {r[1]-r[0]:5d} | pass1 {r[2]-r[1]:5d} | exch {r[3]-r[2]:5d} | pass2 {r[4]-r[3]:5d} | total {r[4]-r[0]:6d}")
