"""Times the ViT-B/16 weight-gradient GEMMs (batch 256: 50432 tokens) with and without VbGemmDesc::a_colsum, and the stand-alone
column-sum kernel the fused form replaces.  CUDA events, L2 flushed between launches, isolated kernels (clocks higher than in-step)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops
from vitb200.engine import pick_split_k


def timed(fn, reps=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    evs = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    return ts[len(ts) // 2]


def main():
    torch.manual_seed(0)
    M = 50432
    for name, n_out, k_in in (("fc1", 3072, 768), ("qkv", 2304, 768), ("fc2", 768, 3072), ("proj", 768, 768)):
        dy = torch.randn(M, n_out, device="cuda").bfloat16()
        x = torch.randn(M, k_in, device="cuda").bfloat16()
        dW = torch.zeros(n_out, k_in, device="cuda")
        db = torch.zeros(n_out, device="cuda")
        tiles = (-(-(-(-n_out // 128)) // 2)) * (-(-k_in // 256))
        split = pick_split_k(tiles, -(-M // 64), 74)
        t0 = timed(lambda: ops.gemm(dy, x, dW, a_major=1, b_major=1, epilogue=ops.EPI_ACCUM, split_k=split))
        t1 = timed(lambda: ops.gemm(dy, x, dW, a_major=1, b_major=1, epilogue=ops.EPI_ACCUM, split_k=split, a_colsum=db))
        t2 = timed(lambda: ops.colsum_bf16(dy, db))
        fl = 2.0 * M * n_out * k_in
        print(f"{name}: split {split}  wgrad {t0:.1f} us ({fl / t0 * 1e-6:.0f} TF/s)  wgrad+a_colsum {t1:.1f} us (+{t1 - t0:.1f})  colsum kernel {t2:.1f} us")


if __name__ == "__main__":
    main()
