"""Times DeiT-S (D = 384) forward / dgrad GEMMs whose N = 384 leaves a half-empty second n-block (CUDA events, L2 flushed).
Run twice: VITB200_GEMM_TAIL128=1 (default: the tail n-block is a 256x128 MMA) and =0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops
from wgrad_colsum_time import timed

M = 256 * 198
torch.manual_seed(0)
for name, N, K, epi, f32 in (("proj fwd", 384, 384, ops.EPI_RESIDUAL, True), ("fc2 fwd", 384, 1536, ops.EPI_RESIDUAL, True),
                             ("qkv dgrad", 384, 1152, ops.EPI_STORE, False), ("fc1 dgrad", 384, 1536, ops.EPI_STORE, False),
                             ("qkv fwd", 1152, 384, ops.EPI_STORE, False)):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    C = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    aux = torch.randn(M, N, device="cuda") if epi == ops.EPI_RESIDUAL else None
    t = timed(lambda: ops.gemm(A, W, C, epilogue=epi, aux=aux), reps=10)
    print(f"{name}: {t:.1f} us  {2.0 * M * N * K / t * 1e-6:.0f} TF/s")
