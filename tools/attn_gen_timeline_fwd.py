"""In-kernel cycle stamps of the general tcgen05 attention forward (CTA 0): DETR encoder shape, sequence-first, masked."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops, _lib
N, H, S = 4, 8, 1050
D = H * 64
q = torch.randn(S * N, D, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
o = torch.empty_like(q)
lse = torch.empty(N, H, S, device="cuda")
kpm = torch.zeros(N, S, dtype=torch.uint8, device="cuda"); kpm[:, 900:] = 1
dbg = torch.zeros(64 * 16, device="cuda", dtype=torch.int64)
lib = _lib.load()
kw = dict(B=N, H=H, S=S, tok_stride=N, batch_stride=1, key_padding_mask=kpm)
for rep in range(2):
    dbg.zero_()
    lib.vb_debug_set_attn_timeline(dbg.data_ptr())
    ops.attention_fwd(q, k, v, o, lse, **kw)
    torch.cuda.synchronize()
lib.vb_debug_set_attn_timeline(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention_fwd(q, k, v, o, lse, **kw)
e1.record(); torch.cuda.synchronize()
print("fwd total per call us:", e0.elapsed_time(e1) * 100)
t = dbg.view(64, 16).cpu()
t0 = t[0, 0].item()
for i in range(14):
    r = [x.item() - t0 for x in t[i, :5]]
    print(f"{i:2d}: start {r[0]:7d} | waitS {r[1]-r[0]:5d} | pass1 {r[2]-r[1]:5d} | exch {r[3]-r[2]:5d} | pass2 {r[4]-r[3]:5d} | total {r[4]-r[0]:6d}")
