"""qkv / out-proj wgrad GEMM time vs split-K factor (ViT-B/16, B=256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemm_bringup as g
from vitb200 import ops
M = 256 * 197
for s in (1, 2, 3, 4, 5, 6, 8, 11, 16):
    g.perf(2304, 768, M, 1, 1, ops.EPI_ACCUM, ops.F32, split=s, name=f"qkv wgrad s={s}")
for s in (2, 4, 8, 12, 16):
    g.perf(768, 768, M, 1, 1, ops.EPI_ACCUM, ops.F32, split=s, name=f"proj wgrad s={s}")
