"""Small GEMM workload for ncu: the four forward variants + dgelu dgrad + a wgrad at ViT-B shapes, B = 64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitb200 import ops
torch.manual_seed(0)
M = int(os.environ.get("VB_PROF_B", "256")) * 197
dev = "cuda"
x768 = torch.randn(M, 768, device=dev).bfloat16()
x3072 = torch.randn(M, 3072, device=dev).bfloat16()
w_qkv = torch.randn(2304, 768, device=dev).bfloat16()
w_o = torch.randn(768, 768, device=dev).bfloat16()
w_1 = torch.randn(3072, 768, device=dev).bfloat16()
w_2 = torch.randn(768, 3072, device=dev).bfloat16()
bias = {n: torch.randn(n, device=dev) for n in (768, 2304, 3072)}
res = torch.randn(M, 768, device=dev)
qkv = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
o32 = torch.empty(M, 768, device=dev)
a = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
g = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
da = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(3072, 768, device=dev)
for _ in range(2):
    ops.gemm(x768, w_qkv, qkv, bias=bias[2304])
    ops.gemm(x768, w_o, o32, epilogue=ops.EPI_RESIDUAL, bias=bias[768], aux=res)
    ops.gemm(x768, w_1, a, C2=g, epilogue=ops.EPI_GELU, bias=bias[3072])
    ops.gemm(x3072, w_2, o32, epilogue=ops.EPI_RESIDUAL, bias=bias[768], aux=res)
    ops.gemm(x768, w_2, da, b_major=1, epilogue=ops.EPI_DGELU, aux=a)
    ops.gemm(x3072, x768, dw, a_major=1, b_major=1, epilogue=ops.EPI_ACCUM, split_k=2)
torch.cuda.synchronize()
print("done")
