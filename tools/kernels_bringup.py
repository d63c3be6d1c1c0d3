"""GPU bring-up sweep for the non-GEMM kernels (LayerNorm, attention, helpers) vs torch fp32 references.
Run on a B200:  timeout 300 python tools/kernels_bringup.py [--perf]
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from vitb200 import ops

torch.manual_seed(0)
dev = "cuda"
ALL_OK = True


def rel(a, b):
    a, b = a.float(), b.float()
    if b.norm().item() < 1e-6:          # exact-zero reference (e.g. dQ, dK of a one-key softmax): absolute error
        return (a - b).norm().item()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def report(name, errs, tol):
    global ALL_OK
    ok = all((e == e) and e < tol for e in errs.values())
    ALL_OK &= ok
    print(f"{'OK  ' if ok else 'FAIL'} {name}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)


def test_ln(rows, D, eps, dy_bf16):
    x = torch.randn(rows, D, device=dev) * 2 + 0.5
    g = torch.randn(D, device=dev)
    b = torch.randn(D, device=dev)
    yb = torch.empty(rows, D, device=dev, dtype=torch.bfloat16)
    yf = torch.empty(rows, D, device=dev)
    mean = torch.empty(rows, device=dev)
    rstd = torch.empty(rows, device=dev)
    ops.layernorm_fwd(x, g, b, eps, y_bf16=yb, y_f32=yf, mean=mean, rstd=rstd)
    xr = x.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), gr, br, eps)
    errs = {"y_f32": rel(yf, ref), "y_bf16": rel(yb, ref), "mean": rel(mean, x.mean(1)),
            "rstd": rel(rstd, 1 / torch.sqrt(x.var(1, unbiased=False) + eps))}
    report(f"ln_fwd rows={rows} D={D}", errs, 5e-3)
    dy = torch.randn(rows, D, device=dev)
    dy_in = dy.bfloat16() if dy_bf16 else dy
    dres = torch.randn(rows, D, device=dev)
    ref.backward(dy_in.float())
    dx = torch.empty(rows, D, device=dev)
    dxb = torch.empty(rows, D, device=dev, dtype=torch.bfloat16)
    dg = torch.zeros(D, device=dev)
    db = torch.zeros(D, device=dev)
    cs = torch.zeros(D, device=dev)
    ops.layernorm_bwd(dy_in, x, mean, rstd, g, dres=dres, dx=dx, dx_bf16=dxb, dgamma=dg, dbeta=db, dx_colsum=cs)
    dx_ref = xr.grad + dres
    errs = {"dx": rel(dx, dx_ref), "dx_bf16": rel(dxb, dx_ref), "dgamma": rel(dg, gr.grad), "dbeta": rel(db, br.grad),
            "colsum": rel(cs, dx_ref.sum(0))}
    report(f"ln_bwd rows={rows} D={D} dy_bf16={dy_bf16}", errs, 5e-3)


def test_attn(B, H, S, seq_first, masked):
    D = H * 64
    qkv = (torch.randn(B, S, 3 * D, device=dev) * 1.5).bfloat16()
    if seq_first:
        buf = qkv.transpose(0, 1).contiguous()        # [S, B, 3D]
        flat = buf.view(S * B, 3 * D)
        tok_stride, batch_stride = B, 1
    else:
        flat = qkv.view(B * S, 3 * D)
        tok_stride, batch_stride = 1, S
    q, k, v = flat[:, :D], flat[:, D:2 * D], flat[:, 2 * D:]
    kpm = None
    if masked:
        valid = torch.randint(S // 2, S + 1, (B,), device=dev)
        kpm_b = torch.arange(S, device=dev)[None, :] >= valid[:, None]
        kpm = kpm_b.to(torch.uint8).contiguous()
    o = torch.empty(flat.shape[0], D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device=dev)
    ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=tok_stride, batch_stride=batch_stride, key_padding_mask=kpm)
    # reference in fp32 on the same bf16 inputs
    qr, kr, vr = [t.float().view(B, S, H, 64).transpose(1, 2).detach().requires_grad_(True) for t in qkv.split(D, dim=2)]
    sc = qr @ kr.transpose(-1, -2) / 8.0
    if masked:
        sc = sc.masked_fill(kpm_b[:, None, None, :], float("-inf"))
    pr = torch.softmax(sc, -1)
    oref = pr @ vr                                     # [B,H,S,64]
    lse_ref = torch.logsumexp(sc, -1) * math.log2(math.e)
    o_bsd = (o.view(S, B, D).transpose(0, 1) if seq_first else o.view(B, S, D))
    o_ref_bsd = oref.transpose(1, 2).reshape(B, S, D)
    errs = {"o": rel(o_bsd, o_ref_bsd), "lse": rel(lse, lse_ref)}
    report(f"attn_fwd B={B} H={H} S={S} seq_first={seq_first} masked={masked}", errs, 1e-2)
    do_bsd = torch.randn(B, S, D, device=dev).bfloat16()
    oref.backward(do_bsd.float().view(B, S, H, 64).transpose(1, 2))
    do_flat = (do_bsd.transpose(0, 1).contiguous().view(S * B, D) if seq_first else do_bsd.view(B * S, D))
    dqkv = torch.full_like(flat, float("nan"))
    delta = torch.empty(B, H, S, device=dev)
    # backward consumes the kernel's own o / lse, as the engine does
    ops.attention_bwd(q, k, v, o, lse, do_flat, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], delta, B=B, H=H, S=S,
                      tok_stride=tok_stride, batch_stride=batch_stride, key_padding_mask=kpm)
    dqkv_bsd = dqkv.view(S, B, 3 * D).transpose(0, 1) if seq_first else dqkv.view(B, S, 3 * D)
    refs = [t.grad.transpose(1, 2).reshape(B, S, D) for t in (qr, kr, vr)]
    errs = {n: rel(dqkv_bsd[:, :, i * D:(i + 1) * D], r) for i, (n, r) in enumerate(zip(("dq", "dk", "dv"), refs))}
    report(f"attn_bwd B={B} H={H} S={S} seq_first={seq_first} masked={masked}", errs, 1.5e-2)


def test_helpers():
    B, C, H, p, D = 5, 3, 224, 16, 768
    img = torch.randn(B, C, H, H, device=dev)
    out = torch.empty(B, (H // p) ** 2, C * p * p, device=dev, dtype=torch.bfloat16)
    ops.patchify(img, out, p)
    ref = F.unfold(img, kernel_size=p, stride=p).transpose(1, 2)  # [B, P, C*p*p], k = c*p*p + i*p + j
    report("patchify p=16", {"err": rel(out, ref.bfloat16())}, 1e-6)
    img = torch.randn(7, 3, 32, 32, device=dev)
    out = torch.empty(7, 64, 48, device=dev, dtype=torch.bfloat16)
    ops.patchify(img, out, 4)
    report("patchify p=4", {"err": rel(out, F.unfold(img, 4, stride=4).transpose(1, 2).bfloat16())}, 1e-6)
    S = 198
    x = torch.full((B, S, D), float("nan"), device=dev)
    t0, t1, pos = torch.randn(D, device=dev), torch.randn(D, device=dev), torch.randn(S, D, device=dev)
    ops.token_rows(x, t0, t1, pos, 2)
    ok_rest = bool(torch.isnan(x[:, 2:]).all().item())
    report("token_rows", {"r0": rel(x[:, 0], (t0 + pos[0]).expand(B, D)), "r1": rel(x[:, 1], (t1 + pos[1]).expand(B, D)),
                          "rest_touched": 0.0 if ok_rest else 1.0}, 1e-6)
    xs = torch.randn(3000, 2304, device=dev).bfloat16()
    acc = torch.ones(2304, device=dev)
    ops.colsum_bf16(xs, acc)
    report("colsum", {"err": rel(acc, xs.float().sum(0) + 1)}, 1e-5)
    src = torch.randn(1 << 20, device=dev)
    dst = torch.empty(1 << 20, device=dev, dtype=torch.bfloat16)
    ops.cast_bf16(src, dst)
    report("cast", {"err": rel(dst, src.bfloat16())}, 1e-7)
    Bq, S, D, npref = 9, 197, 768, 1
    dx = torch.randn(Bq, S, D, device=dev)
    possum = torch.empty(S, D, device=dev)
    dxp = torch.empty(Bq * (S - npref), D, device=dev, dtype=torch.bfloat16)
    dpos = torch.ones(S, D, device=dev)
    dtok = torch.ones(D, device=dev)
    dbias = torch.ones(D, device=dev)
    ops.embed_bwd(dx, possum, dxp, dpos, dtok, None, dbias, npref)
    report("embed_bwd", {"dpos": rel(dpos, dx.sum(0) + 1), "dtok": rel(dtok, dx[:, 0].sum(0) + 1),
                         "dbias": rel(dbias, dx[:, 1:].sum((0, 1)) + 1),
                         "dxp": rel(dxp, dx[:, 1:].reshape(-1, D).bfloat16())}, 1e-5)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def perf():
    B, S, H = 256, 197, 12
    D = H * 64
    M = B * S
    x = torch.randn(M, D, device=dev)
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    yb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    ms = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-6, y_bf16=yb, mean=mean, rstd=rstd))
    print(f"PERF ln_fwd   {ms:.3f} ms  {M * D * 6 / ms / 1e6:.0f} GB/s (4B in + 2B out per element)", flush=True)
    dy = torch.randn(M, D, device=dev).bfloat16()
    dres = torch.randn(M, D, device=dev)
    dx = torch.empty(M, D, device=dev)
    dxb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    dg, db, cs = torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    ms = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dres=dres, dx=dx, dx_bf16=dxb, dgamma=dg, dbeta=db, dx_colsum=cs))
    print(f"PERF ln_bwd   {ms:.3f} ms  {M * D * 16 / ms / 1e6:.0f} GB/s (2+4+4 B in, 4+2 B out per element)", flush=True)
    qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
    o = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device=dev)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    ms = timeit(lambda: ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=1, batch_stride=S))
    fl = 4.0 * S * S * 64 * H * B
    print(f"PERF attn_fwd {ms:.3f} ms  {fl / ms / 1e9:.1f} TF/s algorithmic", flush=True)
    do = torch.randn(M, D, device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, S, device=dev)
    ms = timeit(lambda: ops.attention_bwd(q, k, v, o, lse, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], delta, B=B, H=H, S=S,
                                          tok_stride=1, batch_stride=S))
    print(f"PERF attn_bwd {ms:.3f} ms  {2 * fl / ms / 1e9:.1f} TF/s algorithmic (2x fwd FLOPs)", flush=True)
    qh = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous()
    ms = timeit(lambda: F.scaled_dot_product_attention(qh[0], qh[1], qh[2]))
    print(f"PERF torch sdpa fwd {ms:.3f} ms  {fl / ms / 1e9:.1f} TF/s", flush=True)
    xs = torch.randn(M, 3072, device=dev).bfloat16()
    acc = torch.zeros(3072, device=dev)
    ms = timeit(lambda: ops.colsum_bf16(xs, acc))
    print(f"PERF colsum [M,3072] {ms:.3f} ms  {M * 3072 * 2 / ms / 1e6:.0f} GB/s", flush=True)
    img = torch.randn(B, 3, 224, 224, device=dev)
    pat = torch.empty(B, 196, 768, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.patchify(img, pat, 16))
    print(f"PERF patchify {ms:.3f} ms  {img.numel() * 6 / ms / 1e6:.0f} GB/s", flush=True)


def correctness():
    for D in (64, 192, 256, 320, 384, 512, 768, 960, 1024):
        test_ln(1003, D, 1e-6, True)
    test_ln(50, 768, 1e-5, False)
    for (B, H, S) in ((3, 4, 65), (2, 12, 197), (2, 6, 198), (1, 8, 1050), (4, 2, 64), (2, 2, 1)):
        test_attn(B, H, S, False, False)
    test_attn(2, 8, 1050, True, True)
    test_attn(3, 8, 300, True, False)
    test_attn(3, 4, 197, False, True)
    test_helpers()
    return ALL_OK


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    correctness()
    print("ALL OK" if ALL_OK else "SOME FAILED", flush=True)
    if "--perf" in sys.argv:
        perf()
    return 0 if ALL_OK else 1


if __name__ == "__main__":
    sys.exit(main())
