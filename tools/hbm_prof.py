"""HBM-bound kernels of the ViT-B/16 step at batch 256 (M = 50432 tokens), one launch each after a warm-up, for ncu:
LayerNorm forward / backward, bias-gradient column sum, attention delta, flat Adam, patchify, embedding backward, PEG depthwise conv.
Usage (see profiles/r1d_hbm_kernels_ncu_summary.txt):
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
      --clock-control none --csv --log-file gpurun_out/hbm_r1d.csv python tools/hbm_prof.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vitb200 import ops

B, S, D, Fd, H = 256, 197, 768, 3072, 12
M = B * S
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, D, device=dev)
g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
y_bf = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
dy = torch.randn(M, D, device=dev).bfloat16()
dres, dx = torch.randn(M, D, device=dev), torch.empty(M, D, device=dev)
dx_bf = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
dg, db, cs = torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
da = torch.randn(M, Fd, device=dev).bfloat16()
csf = torch.zeros(Fd, device=dev)
n = 86_567_680
p, gr, m1, m2 = (torch.zeros(n, device=dev) for _ in range(4))
p_bf = torch.empty(n, device=dev, dtype=torch.bfloat16)
step = torch.zeros(1, device=dev, dtype=torch.int32)
img = torch.randn(B, 3, 224, 224, device=dev)
pat = torch.empty(B, 196, 768, device=dev, dtype=torch.bfloat16)
x3, o3 = x.view(B, S, D), torch.empty(B, S, D, device=dev)
w9, b9 = torch.randn(D, 1, 3, 3, device=dev), torch.randn(D, device=dev)
dw9, db9 = torch.zeros(D * 9, device=dev), torch.zeros(D, device=dev)
o_bf, do_bf = torch.randn(M, D, device=dev).bfloat16(), torch.randn(M, D, device=dev).bfloat16()
qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
lse, delta = torch.zeros(B, H, S, device=dev), torch.empty(B, H, S, device=dev)
dqkv = torch.empty_like(qkv)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def run():
    for f in (lambda: ops.layernorm_fwd(x, g, b, 1e-6, y_bf16=y_bf, mean=mean, rstd=rstd),
              lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dres=dres, dx=dx, dx_bf16=dx_bf, dgamma=dg, dbeta=db, dx_colsum=cs),
              lambda: ops.colsum_bf16(da, csf),
              lambda: ops.adam_step(p, gr, m1, m2, p_bf, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=0, step_counter=step),
              lambda: ops.patchify(img, pat, 16),
              lambda: ops.dwconv_fwd(x3, w9, b9, o3, n_prefix=1),
              lambda: ops.dwconv_bwd_data(x3, w9, n_prefix=1, dx=o3, sum_bf16=y_bf),
              lambda: ops.dwconv_bwd_weight(x3, o3, dw9, db9, n_prefix=1)):
        flush.zero_()          # evict the 126 MB L2 so that every kernel streams from HBM
        f()


run()
torch.cuda.synchronize()
run()
torch.cuda.synchronize()
print("done")
