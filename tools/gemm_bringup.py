"""GPU bring-up sweep for the tcgen05 GEMM: correctness of every (major, epilogue) variant vs a torch fp32
reference, through both the bring-up direct-store path and the TMA-store epilogue, then timings.
Run on a B200:  timeout 300 python tools/gemm_bringup.py [--perf]
"""
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vitb200 import ops

torch.manual_seed(0)
dev = "cuda"


def rel(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def gelu(x):
    return 0.5 * x * (1 + torch.erf(x / math.sqrt(2)))


def dgelu(x):
    return 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)


def case(name, M, N, K, a_major, b_major, epi, cdt, *, bias=False, split=1, direct=False, tol=1e-2):
    A = torch.randn(M, K, device=dev).bfloat16()
    Bm = torch.randn(N, K, device=dev).bfloat16()
    ref = A.float() @ Bm.float().t()
    A_in = A if a_major == 0 else A.t().contiguous()
    B_in = Bm if b_major == 0 else Bm.t().contiguous()
    bvec = torch.randn(N, device=dev) if bias else None
    if bias:
        ref = ref + bvec
    odt = torch.bfloat16 if cdt == ops.BF16 else torch.float32
    ldc = (N + 15) // 16 * 16
    Cfull = torch.full((M, ldc), float("nan"), device=dev, dtype=odt)
    C = Cfull[:, :N]
    C2 = aux = None
    refs = []
    if epi == ops.EPI_STORE:
        refs = [ref]
    elif epi == ops.EPI_GELU:
        C2 = torch.full((M, ldc), float("nan"), device=dev, dtype=odt)[:, :N]
        refs = [dgelu(ref), gelu(ref)]
    elif epi == ops.EPI_RESIDUAL:
        aux = torch.randn(M, ldc, device=dev, dtype=odt)[:, :N]
        refs = [ref + aux.float()]
    elif epi == ops.EPI_RELU:
        refs = [torch.relu(ref)]
    elif epi == ops.EPI_DGELU:
        aux = torch.randn(M, ldc, device=dev).to(odt)[:, :N]
        refs = [ref * aux.float()]
    elif epi == ops.EPI_DRELU:
        aux = torch.randn(M, ldc, device=dev).to(odt)[:, :N]
        refs = [torch.where(aux.float() > 0, ref, torch.zeros_like(ref))]
    elif epi == ops.EPI_ACCUM:
        init = torch.randn(M, N, device=dev)
        C.copy_(init)
        refs = [ref + init]
    ops.gemm(A_in, B_in, C, a_major=a_major, b_major=b_major, epilogue=epi, bias=bvec, aux=aux, C2=C2, split_k=split,
             direct=direct)
    torch.cuda.synchronize()
    outs = [C] + ([C2] if C2 is not None else [])
    errs = [rel(o, r) for o, r in zip(outs, refs)]
    nan = any(torch.isnan(o.float()).any().item() for o in outs)
    pad_ok = True
    # TMA stores clip at 16-byte granularity: columns beyond round_up(N * elem_size, 16) must stay untouched
    # (the engine always allocates row pitches that are multiples of 16 bytes, so the clipped tail is padding).
    es = 2 if cdt == ops.BF16 else 4
    n_clip = (N * es + 15) // 16 * 16 // es
    if ldc > n_clip:
        pad = Cfull[:, n_clip:]
        pad_ok = bool(torch.isnan(pad.float()).all().item())
    ok = (not nan) and all(e < tol for e in errs) and pad_ok
    print(f"{'OK  ' if ok else 'FAIL'} {name:34s} M={M:6d} N={N:5d} K={K:5d} maj=({a_major},{b_major}) epi={epi} cdt={cdt} "
          f"split={split} direct={int(direct)} err={['%.2e' % e for e in errs]} nan={nan} pad_untouched={pad_ok}", flush=True)
    return ok


def batched_case(direct):
    # patch-embed geometry: per image 196 patch rows -> rows 1..196 of a 197-row output, plus broadcast aux (pos emb)
    nb, P, S, K, N = 3, 196, 197, 768, 768
    A = torch.randn(nb, P, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    pos = torch.randn(1, S, N, device=dev)
    out = torch.full((nb, S, N), float("nan"), device=dev)
    ops.gemm(A, W, out, epilogue=ops.EPI_RESIDUAL, bias=bias, aux=pos, c_row_offset=1, aux_broadcast=True, direct=direct)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias + pos[:, 1:]
    e = rel(out[:, 1:], ref)
    row0_untouched = bool(torch.isnan(out[:, 0]).all().item())
    ok = e < 1e-2 and row0_untouched
    print(f"{'OK  ' if ok else 'FAIL'} batched patch-embed geometry direct={int(direct)} err={e:.2e} row0_untouched={row0_untouched}", flush=True)
    return ok


def perf(M, N, K, a_major, b_major, epi, cdt, split=1, iters=20, name=""):
    A = torch.randn(M, K, device=dev).bfloat16()
    Bm = torch.randn(N, K, device=dev).bfloat16()
    A_in = A if a_major == 0 else A.t().contiguous()
    B_in = Bm if b_major == 0 else Bm.t().contiguous()
    odt = torch.bfloat16 if cdt == ops.BF16 else torch.float32
    C = torch.zeros(M, N, device=dev, dtype=odt)
    C2 = torch.empty_like(C) if epi == ops.EPI_GELU else None
    aux = torch.randn(M, N, device=dev).to(odt) if epi in (ops.EPI_RESIDUAL, ops.EPI_DGELU, ops.EPI_DRELU) else None
    bias = torch.randn(N, device=dev) if epi in (ops.EPI_STORE, ops.EPI_GELU, ops.EPI_RESIDUAL) and a_major == 0 and b_major == 0 else None
    def run():
        ops.gemm(A_in, B_in, C, a_major=a_major, b_major=b_major, epilogue=epi, bias=bias, aux=aux, C2=C2, split_k=split)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS comparator (plain matmul, no epilogue)
    X = A if a_major == 0 else A_in.t()
    Y = Bm.t() if b_major == 0 else B_in
    for _ in range(3):
        torch.matmul(X, Y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(X, Y)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    tf2 = 2.0 * M * N * K / ms2 / 1e9
    print(f"PERF {name:22s} M={M} N={N} K={K} maj=({a_major},{b_major}) epi={epi} split={split}: {ms:.3f} ms {tf:7.1f} TF/s | cuBLAS {ms2:.3f} ms {tf2:7.1f} TF/s", flush=True)


def correctness():
    ok = True
    for direct in (True, False):
        ok &= case("single tile 1 k-block", 128, 256, 64, 0, 0, ops.EPI_STORE, ops.BF16, direct=direct)
        ok &= case("single tile K=768", 128, 256, 768, 0, 0, ops.EPI_STORE, ops.BF16, direct=direct)
        ok &= case("tails M/N", 1000, 2304, 768, 0, 0, ops.EPI_STORE, ops.BF16, bias=True, direct=direct)
        ok &= case("fp32 out N=1000", 300, 1000, 768, 0, 0, ops.EPI_STORE, ops.F32, bias=True, direct=direct)
        ok &= case("tiny head N=10", 64, 10, 256, 0, 0, ops.EPI_STORE, ops.F32, bias=True, direct=direct)
        ok &= case("K=48 (p=4 patch)", 640, 256, 48, 0, 0, ops.EPI_STORE, ops.BF16, bias=True, direct=direct)
        ok &= case("gelu", 1000, 3072, 768, 0, 0, ops.EPI_GELU, ops.BF16, bias=True, direct=direct)
        ok &= case("residual fp32", 1000, 768, 3072, 0, 0, ops.EPI_RESIDUAL, ops.F32, bias=True, direct=direct)
        ok &= case("relu", 1000, 2048, 512, 0, 0, ops.EPI_RELU, ops.BF16, bias=True, direct=direct)
        ok &= case("dgrad K,MN", 1000, 768, 2304, 0, 1, ops.EPI_STORE, ops.BF16, direct=direct)
        ok &= case("dgrad dgelu", 1000, 3072, 768, 0, 1, ops.EPI_DGELU, ops.BF16, direct=direct)
        ok &= case("dgrad drelu", 1000, 2048, 512, 0, 1, ops.EPI_DRELU, ops.BF16, direct=direct)
        ok &= case("wgrad MN,MN store", 768, 768, 1000, 1, 1, ops.EPI_STORE, ops.F32, direct=direct, tol=2e-2)
        ok &= case("wgrad MN,MN accum", 3072, 768, 4000, 1, 1, ops.EPI_ACCUM, ops.F32, direct=direct)
        ok &= case("wgrad split 4", 1000, 768, 4000, 1, 1, ops.EPI_ACCUM, ops.F32, split=4, direct=direct)
        ok &= case("many tiles (persistence)", 20000, 768, 768, 0, 0, ops.EPI_STORE, ops.BF16, bias=True, direct=direct)
        ok &= batched_case(direct)
    # ViT-L/16 (D = 1024, F = 4096: BASELINE.json configs[3]), DeiT-S (D = 384) and DeiT-tiny (D = 192, an odd multiple of 64:
    # utils/args.py:43-45) shapes through the production (TMA-store) epilogues
    for D, Fd in ((1024, 4096), (384, 1536), (192, 768)):
        M = 8 * 197 if D > 192 else 64 * 6
        ok &= case(f"D={D} qkv", M, 3 * D, D, 0, 0, ops.EPI_STORE, ops.BF16, bias=True)
        ok &= case(f"D={D} out-proj residual", M, D, D, 0, 0, ops.EPI_RESIDUAL, ops.F32, bias=True)
        ok &= case(f"D={D} fc1 gelu", M, Fd, D, 0, 0, ops.EPI_GELU, ops.BF16, bias=True)
        ok &= case(f"D={D} fc2 residual", M, D, Fd, 0, 0, ops.EPI_RESIDUAL, ops.F32, bias=True)
        ok &= case(f"D={D} fc2 dgrad dgelu", M, Fd, D, 0, 1, ops.EPI_DGELU, ops.BF16)
        ok &= case(f"D={D} fc1 dgrad", M, D, Fd, 0, 1, ops.EPI_STORE, ops.BF16)
        ok &= case(f"D={D} qkv dgrad", M, D, 3 * D, 0, 1, ops.EPI_STORE, ops.BF16)
        ok &= case(f"D={D} fc1 wgrad split 2", Fd, D, M, 1, 1, ops.EPI_ACCUM, ops.F32, split=2)
        ok &= case(f"D={D} qkv wgrad", 3 * D, D, M, 1, 1, ops.EPI_ACCUM, ops.F32)
    ok &= case("ViT-L fc1 at batch-64 rows", 64 * 197, 4096, 1024, 0, 0, ops.EPI_GELU, ops.BF16, bias=True)
    return ok


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    ok = correctness()
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    if "--perf" in sys.argv:
        M = 256 * 197
        perf(M, 2304, 768, 0, 0, ops.EPI_STORE, ops.BF16, name="qkv fwd")
        perf(M, 768, 768, 0, 0, ops.EPI_RESIDUAL, ops.F32, name="out-proj fwd")
        perf(M, 3072, 768, 0, 0, ops.EPI_GELU, ops.BF16, name="fc1 fwd")
        perf(M, 768, 3072, 0, 0, ops.EPI_RESIDUAL, ops.F32, name="fc2 fwd")
        perf(M, 768, 3072, 0, 1, ops.EPI_STORE, ops.BF16, name="fc1 dgrad")
        perf(M, 3072, 768, 0, 1, ops.EPI_DGELU, ops.BF16, name="fc2 dgrad")
        perf(M, 768, 2304, 0, 1, ops.EPI_STORE, ops.BF16, name="qkv dgrad")
        perf(3072, 768, M, 1, 1, ops.EPI_ACCUM, ops.F32, split=2, name="fc1 wgrad")
        perf(768, 3072, M, 1, 1, ops.EPI_ACCUM, ops.F32, split=2, name="fc2 wgrad")
        perf(2304, 768, M, 1, 1, ops.EPI_ACCUM, ops.F32, split=8, name="qkv wgrad")
        perf(768, 768, M, 1, 1, ops.EPI_ACCUM, ops.F32, split=8, name="out-proj wgrad")
        perf(8192, 8192, 8192, 0, 0, ops.EPI_STORE, ops.BF16, name="8192^3")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
