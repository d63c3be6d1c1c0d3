"""Split-K sweep of the weight-gradient GEMMs (CUDA events, L2 flushed, isolated): which split the engine's pick_split_k should choose."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops
from vitb200.engine import pick_split_k
from wgrad_colsum_time import timed


def main():
    torch.manual_seed(0)
    shapes = [("vitb qkv", 50432, 2304, 768), ("vitb proj", 50432, 768, 768), ("vitb fc1", 50432, 3072, 768), ("vitb fc2", 50432, 768, 3072),
              ("deits qkv", 50688, 1152, 384), ("deits proj", 50688, 384, 384), ("deits fc1", 50688, 1536, 384), ("deits fc2", 50688, 384, 1536),
              ("detr in_qk", 4200, 1024, 512), ("detr lin1", 4200, 2048, 512)]
    for name, M, n_out, k_in in shapes:
        dy = torch.randn(M, n_out, device="cuda").bfloat16()
        x = torch.randn(M, k_in, device="cuda").bfloat16()
        dW = torch.zeros(n_out, k_in, device="cuda")
        tiles = (-(-(-(-n_out // 128)) // 2)) * (-(-k_in // 256))
        picked = pick_split_k(tiles, -(-M // 64), 74)
        res = []
        for s in sorted(set([1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 30, picked])):
            if s > 1 and (-(-M // 64)) // s < 8:
                continue
            t = timed(lambda: ops.gemm(dy, x, dW, a_major=1, b_major=1, epilogue=ops.EPI_ACCUM, split_k=s), reps=10)
            res.append((s, t))
        best = min(res, key=lambda r: r[1])
        print(f"{name}: tiles {tiles} picked {picked}; " + " ".join(f"{s}:{t:.0f}" + ("*" if s == best[0] else "") for s, t in res))


if __name__ == "__main__":
    main()
