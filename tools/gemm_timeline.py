"""Cycle accounting of the GEMM's MMA issuer (cluster 0): how much of a launch it waits for operand stages (TMA / L2), for a free
TMEM accumulator (epilogue) and for the tile scheduler.  ViT-B/16 shapes at batch 256."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops, _lib
M = 256 * 197
dev = "cuda"
lib = _lib.load()
dbg = torch.zeros(8, device=dev, dtype=torch.int64)


def run(name, Mr, N, K, a_major, b_major, epi, cdt, split=1, direct=False):
    A = torch.randn(Mr, K, device=dev).bfloat16()
    Bm = torch.randn(N, K, device=dev).bfloat16()
    A_in = A if a_major == 0 else A.t().contiguous()
    B_in = Bm if b_major == 0 else Bm.t().contiguous()
    odt = torch.bfloat16 if cdt == ops.BF16 else torch.float32
    C = torch.zeros(Mr, N, device=dev, dtype=odt)
    C2 = torch.empty_like(C) if epi == ops.EPI_GELU else None
    aux = torch.randn(Mr, N, device=dev).to(odt) if epi in (ops.EPI_RESIDUAL, ops.EPI_DGELU) else None
    bias = torch.randn(N, device=dev) if epi in (ops.EPI_STORE, ops.EPI_GELU, ops.EPI_RESIDUAL) and a_major == 0 and b_major == 0 else None
    for rep in range(3):
        dbg.zero_()
        lib.vb_debug_set_gemm_timeline(dbg.data_ptr())
        ops.gemm(A_in, B_in, C, a_major=a_major, b_major=b_major, epilogue=epi, bias=bias, aux=aux, C2=C2, split_k=split, direct=direct)
        torch.cuda.synchronize()
    lib.vb_debug_set_gemm_timeline(None)
    t, wf, wt, ws, n = [x.item() for x in dbg[:5]]
    kblocks = -(-K // 64) // split
    print(f"{name:24s} tiles {n:3d}  total {t:8d} clk = {t / max(n, 1):7.0f}/tile (MMA floor {kblocks * 4 * 130:6d})  wait operands {100 * wf / t:5.1f}%  "
          f"wait accumulator {100 * wt / t:5.1f}%  wait scheduler {100 * ws / t:5.1f}%")


run("qkv fwd", M, 2304, 768, 0, 0, ops.EPI_STORE, ops.BF16)
run("qkv fwd NO EPILOGUE", M, 2304, 768, 0, 0, ops.EPI_STORE, ops.BF16, direct=2)
run("qkv fwd TMEM LD ONLY", M, 2304, 768, 0, 0, ops.EPI_STORE, ops.BF16, direct=3)
run("fc1 dgrad NO EPILOGUE", M, 768, 3072, 0, 1, ops.EPI_STORE, ops.BF16, direct=2)
run("out-proj fwd", M, 768, 768, 0, 0, ops.EPI_RESIDUAL, ops.F32)
run("fc1 fwd (GELU)", M, 3072, 768, 0, 0, ops.EPI_GELU, ops.BF16)
run("fc2 fwd", M, 768, 3072, 0, 0, ops.EPI_RESIDUAL, ops.F32)
run("fc1 dgrad", M, 768, 3072, 0, 1, ops.EPI_STORE, ops.BF16)
run("fc2 dgrad (DGELU)", M, 3072, 768, 0, 1, ops.EPI_DGELU, ops.BF16)
run("qkv dgrad", M, 768, 2304, 0, 1, ops.EPI_STORE, ops.BF16)
run("fc1 wgrad", 3072, 768, M, 1, 1, ops.EPI_ACCUM, ops.F32, split=2)
