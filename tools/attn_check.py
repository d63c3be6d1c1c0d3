"""Attention backward: the tcgen05 five-product kernel (and its store-warp column sums) vs an fp32 PyTorch reference, several shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops


def run(B, H, S):
    D = H * 64
    M = B * S
    g = torch.Generator(device="cuda").manual_seed(S * 131 + B)
    qkv = (torch.randn(M, 3 * D, device="cuda", generator=g) * 1.5).bfloat16()
    do = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    o = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device="cuda")
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty(B, H, S, device="cuda")
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
    cs = torch.zeros(3 * D, device="cuda")
    ops.attention_bwd(q, k, v, o, lse, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], delta, B=B, H=H, S=S, tok_stride=1, batch_stride=S,
                      dqkv_colsum=cs)
    torch.cuda.synchronize()
    cs_ref = dqkv.float().sum(0)
    cs_ref[D:2 * D] = 0   # the kernel leaves the (mathematically zero) key-bias part untouched
    run.colsum_err = ((cs - cs_ref).norm() / cs_ref.norm()).item()
    # fp32 reference
    qf, kf, vf = [t.float().view(B, S, H, 64).transpose(1, 2).requires_grad_(True) for t in (q, k, v)]
    of = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf)
    of.backward(do.float().view(B, S, H, 64).transpose(1, 2))
    ref = torch.cat([t.grad.transpose(1, 2).reshape(M, D) for t in (qf, kf, vf)], dim=1)
    return dqkv.float(), ref


ok = True
for (B, H, S) in [(2, 3, 197), (3, 2, 198), (2, 4, 65), (1, 2, 128), (2, 2, 129), (1, 3, 208), (2, 1, 16), (1, 1, 192), (2, 2, 144), (37, 12, 197)]:
    a, ref = run(B, H, S)
    D = H * 64
    errs = []
    for i, name in enumerate("qkv"):
        sl = slice(i * D, (i + 1) * D)
        errs.append((name, ((a[:, sl] - ref[:, sl]).norm() / ref[:, sl].norm()).item()))
    bad = any(not (ea < 2e-2) for _, ea in errs) or not (run.colsum_err < 2e-3)
    errs.append(("colsum", run.colsum_err))
    ok &= not bad
    print(f"B={B} H={H} S={S}: " + "  ".join(f"d{n}: {ea:.2e}" for n, ea in errs) + ("  FAIL" if bad else ""))
print("ALL OK" if ok else "FAILED")
