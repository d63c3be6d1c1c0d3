"""Attention forward: tcgen05 kernel vs an fp32 PyTorch reference (output and lse), several shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops
ok = True
for (B, H, S) in [(2, 3, 197), (3, 2, 198), (2, 4, 65), (1, 2, 128), (2, 2, 129), (1, 3, 208), (2, 1, 16), (1, 1, 192), (2, 2, 144), (37, 12, 197), (256, 12, 197)]:
    D = H * 64; M = B * S
    g = torch.Generator(device="cuda").manual_seed(S * 131 + B)
    qkv = (torch.randn(M, 3 * D, device="cuda", generator=g) * 1.5).bfloat16()
    o = torch.full((M, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((B, H, S), float("nan"), device="cuda")
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
    torch.cuda.synchronize()
    qf, kf, vf = [t.float().view(B, S, H, 64).transpose(1, 2) for t in (q, k, v)]
    sc = (qf @ kf.transpose(-1, -2)) * 0.125
    ref = (torch.softmax(sc, -1) @ vf).transpose(1, 2).reshape(M, D)
    lse_ref = torch.logsumexp(sc, -1) * 1.4426950408889634
    e = ((o.float() - ref).norm() / ref.norm()).item()
    el = (lse - lse_ref).abs().max().item()
    bad = not (e < 1e-2 and el < 1e-3)
    ok &= not bad
    print(f"B={B} H={H} S={S}: out rel-L2 {e:.2e}  lse max-abs {el:.2e}" + ("  FAIL" if bad else ""))
print("ALL OK" if ok else "FAILED")
