"""In-kernel cycle stamps of the tcgen05 attention backward (CTA 0): where does an item's time go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops, _lib
import sys
B, H, S = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 12, 197
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
ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
dbg = torch.zeros(64 * 16, device="cuda", dtype=torch.int64)
lib = _lib.load()
for rep in range(2):
    dbg.zero_()
    lib.vb_debug_set_attn_timeline(dbg.data_ptr())
    ops.attention_bwd(q, k, v, o, lse, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], delta, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
    torch.cuda.synchronize()
lib.vb_debug_set_attn_timeline(None)
t = dbg.view(64, 16).cpu()
t0 = t[0, 0].item()
print("item: mma[tile_free->scores issued | ->p_full seen | ->out issued]  ew[s_full seen | elementwise done | o_full seen | readout done]  (cycles from start)")
prev_end = 0
for i in range(42):
    r = [(x.item() - t0) for x in t[i, :8]]
    gap = r[0] - prev_end; prev_end = r[7]
    if i % 2 == 0 and t[i, 13].item() > 0:
        print(f"    stats warp for head {i // 2}: stage free {t[i,13].item()-t0:7d} | dO landed {t[i,14].item()-t0:7d} | delta done {t[i,15].item()-t0:7d}   (head's first tile: s_full seen at {r[4]:7d})")
    if i < 6: print("    group ends rel. to s_full seen:", [(x.item() - t0) - r[4] for x in t[i, 8:13]])
    print(f"{i:2d}: gap {gap:6d} mma {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} | ew {r[4]:7d} {r[5]:7d} {r[6]:7d} {r[7]:7d}   item total {r[7]-r[0]:6d}  scoreMMA+wake {r[4]-r[1]:5d} elem {r[5]-r[4]:5d} p_full->mma {r[2]-r[5]:5d} outMMA+wake {r[6]-r[3]:5d} readout {r[7]-r[6]:5d}")
