"""-m gpu unit tests of every kernel family against torch fp32 references on random tensors (through the C ABI):
GEMM variants (operand majors x epilogues x TMA/direct store, ragged M/N/K, batched patch-embed geometry, split-K),
LayerNorm fwd/bwd, attention fwd/bwd (S = 1, 64, 65, 197, 198, 300, 1050; batch-first and sequence-first; key-padding
masks), and the helper kernels.  The sweeps live in tools/ so they can also be run stand-alone with timings."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.gpu
def test_gemm_variants():
    import gemm_bringup
    assert gemm_bringup.correctness()


@pytest.mark.gpu
def test_layernorm_attention_helpers():
    import kernels_bringup
    assert kernels_bringup.correctness()


@pytest.mark.gpu
def test_fused_cross_entropy_and_adam():
    import torch
    from vitb200 import ops
    torch.manual_seed(0)
    B, C = 37, 1000
    ld = 1008
    logits = torch.randn(B, ld, device="cuda")[:, :C]
    labels = torch.randint(0, C, (B,), device="cuda")
    loss = torch.zeros(1, device="cuda")
    dz = torch.zeros(B, ld, device="cuda", dtype=torch.bfloat16)
    dzf = torch.zeros(B, ld, device="cuda")
    correct = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.cross_entropy(logits, labels, loss, weight=1.0 / B, dlogits_bf16=dz[:, :C], dlogits_f32=dzf[:, :C], correct_accum=correct)
    lr = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, labels)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item())              # fp32-accumulated: 1e-4
    assert ((dzf[:, :C] - lr.grad).norm() / lr.grad.norm()).item() < 1e-5
    assert ((dz[:, :C].float() - lr.grad).norm() / lr.grad.norm()).item() < 5e-3
    assert correct.item() == (logits.argmax(1) == labels).sum().item()
    n = 4096 * 3
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda")
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    for step in range(1, 4):
        ops.adam_step(p, g, m, v, pb, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=step)
        pr.grad = g.clone()
        opt.step()
    assert ((p - pr.detach()).abs().max()).item() < 1e-6
    assert torch.equal(pb, p.bfloat16())


@pytest.mark.gpu
def test_gemm_dynamic_and_static_schedules_agree_bitwise():
    """The dynamic tile scheduler only changes WHICH cluster computes a tile: outputs must be bit-identical to the static
    round-robin schedule (store epilogues) for a many-tile problem, repeated so that the self-re-arming counters are reused."""
    import subprocess
    code = r'''
import os, sys, torch
sys.path.insert(0, %r)
from vitb200 import ops
torch.manual_seed(0)
M, N, K = 20000, 2304, 768
A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16(); b = torch.randn(N, device="cuda")
outs = []
for rep in range(300):                       # > the 256-entry counter pool
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, C, bias=b)
    if rep in (0, 299): outs.append(C)
torch.cuda.synchronize()
assert torch.equal(outs[0], outs[1])
torch.save(outs[0].cpu(), sys.argv[1])
''' % ROOT
    import tempfile
    import torch
    res = []
    for dyn in ("1", "0"):
        with tempfile.NamedTemporaryFile(suffix=".pt") as fh:
            env = dict(os.environ, VITB200_GEMM_DYNAMIC=dyn)
            subprocess.check_call([sys.executable, "-c", code, fh.name], env=env)
            res.append(torch.load(fh.name))
    assert torch.equal(res[0], res[1])


@pytest.mark.gpu
@pytest.mark.parametrize("tokens,n_out,k_in,split", [(50432, 3072, 768, 2), (50432, 2304, 768, 3), (5000, 200, 304, 1), (777, 1000, 64, 4),
                                                     (256, 3072, 768, 1), (12608, 384, 1536, 6)])
def test_gemm_wgrad_bias_gradient_from_operand_stages(tokens, n_out, k_in, split):
    """VbGemmDesc::a_colsum: the bias gradient (column sums of dY) taken from the A-operand stages of the weight-gradient GEMM equals
    the fp32 sum of the same bf16 values, for every split-K / n-block partition (also with dynamic and static scheduling, ragged M,
    N, K, accumulation onto a non-zero vector, and repeated launches), and leaves dW itself bit-identical."""
    import torch
    from vitb200 import ops
    torch.manual_seed(tokens + n_out)
    dy = (torch.randn(tokens, n_out, device="cuda") * 0.5 + 0.1).bfloat16()
    x = torch.randn(tokens, k_in, device="cuda").bfloat16()
    ref = dy.double().sum(0)
    scale = dy.double().abs().sum(0)
    dW0 = torch.zeros(n_out, k_in, device="cuda")
    ops.gemm(dy, x, dW0, a_major=1, b_major=1, epilogue=ops.EPI_ACCUM, split_k=split)
    for rep in range(3):
        db = torch.full((n_out,), 3.0, device="cuda")
        dW = torch.zeros(n_out, k_in, device="cuda")
        ops.gemm(dy, x, dW, a_major=1, b_major=1, epilogue=ops.EPI_ACCUM, split_k=split, a_colsum=db)
        torch.cuda.synchronize()
        err = ((db.double() - 3.0 - ref).abs() / scale).max().item()
        assert err < 2e-6, err                                    # fp32 accumulation of exactly the same addends: 2e-6 of sum |dy|
        if split == 1:
            assert torch.equal(dW, dW0)                           # one unit per tile: same order of accumulation
        else:
            assert ((dW - dW0).norm() / dW0.norm()).item() < 1e-6  # split-K partials land through fp32 reduce-adds in any order


@pytest.mark.gpu
def test_attention_backward_fused_bias_gradient():
    """dqkv_colsum of vb_attention_bwd (in-projection bias gradient summed inside the tcgen05 backward) against the column sums
    of the stored dq | dk | dv; the key-bias part is mathematically zero and is left untouched by the tcgen05 path."""
    import torch
    from vitb200 import ops
    B, H, S = 5, 6, 197
    D, M = H * 64, B * S
    torch.manual_seed(1)
    qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
    do = torch.randn(M, D, device="cuda").bfloat16()
    o = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device="cuda")
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, S, device="cuda")
    cs = torch.zeros(3 * D, device="cuda")
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    ops.attention_fwd(q, k, v, o, lse, B=B, H=H, S=S, tok_stride=1, batch_stride=S)
    ops.attention_bwd(q, k, v, o, lse, do, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], delta, B=B, H=H, S=S, tok_stride=1,
                      batch_stride=S, dqkv_colsum=cs)
    ref = dqkv.float().sum(0)
    for sl in (slice(0, D), slice(2 * D, 3 * D)):
        assert ((cs[sl] - ref[sl]).norm() / ref[sl].norm()).item() < 3e-3     # fp32 sums of the unrounded accumulators vs bf16 outputs
    assert cs[D:2 * D].abs().max().item() == 0.0
    assert ref[D:2 * D].norm().item() < 2e-2 * ref[:D].norm().item()          # and it is (numerically) zero in the outputs too


@pytest.mark.gpu
@pytest.mark.parametrize("Sq,Sk,N,H,masked,p", [(100, 300, 3, 8, True, 0.0), (100, 1050, 2, 4, True, 0.0), (37, 70, 2, 4, False, 0.0),
                                                (100, 200, 2, 4, True, 0.15), (70, 70, 2, 4, True, 0.15)])
def test_cross_attention_kernels(Sq, Sk, N, H, masked, p):
    """vb_attention_fwd/bwd with S_kv != S (the DETR decoder's cross-attention, transformer.py:145-147: query = tgt + query_pos against
    key = memory + pos, value = memory, memory_key_padding_mask) on sequence-first layouts, against the explicit softmax path in fp32 on
    the same bf16-rounded inputs; with attention dropout the kernels' own keep mask is replayed.  Tolerance: bf16 outputs, 2e-2."""
    import math
    import torch
    from vitb200 import ops
    D = H * 64
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(Sq, N, D, generator=g) * 0.7).bfloat16()
    k = (torch.randn(Sk, N, D, generator=g) * 0.7).bfloat16()
    v = torch.randn(Sk, N, D, generator=g).bfloat16()
    do = torch.randn(Sq, N, D, generator=g).bfloat16()
    kpm = None
    if masked:
        valid = torch.randint(Sk // 2, Sk + 1, (N,), generator=g)
        kpm = torch.arange(Sk)[None, :] >= valid[:, None]
    qc, kc, vc, doc = (t.cuda().view(-1, D) for t in (q, k, v, do))
    o = torch.empty(Sq * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(N, H, Sq, device="cuda")
    delta = torch.empty(N, H, Sq, device="cuda")
    dq, dk, dv = torch.empty_like(qc), torch.empty_like(kc), torch.empty_like(vc)
    seed = torch.full((1,), 1234, device="cuda", dtype=torch.int32)
    drop = (p, seed, 77) if p > 0 else None
    kw = dict(B=N, H=H, S=Sq, S_kv=Sk, tok_stride=N, batch_stride=1, key_padding_mask=kpm.cuda().to(torch.uint8) if masked else None,
              dropout=drop)
    ops.attention_fwd(qc, kc, vc, o, lse, **kw)
    ops.attention_bwd(qc, kc, vc, o, lse, doc, dq, dk, dv, delta, **kw)
    keep = None
    if p > 0:
        keep = ops.dropout_mask(N * H * Sq * Sk, p, seed, 77, "cuda").view(N, H, Sq, Sk).cpu().float()
        assert abs(keep.mean().item() - (1 - p)) < 0.02
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    qh = qf.view(Sq, N, H, 64).permute(1, 2, 0, 3)
    kh = kf.view(Sk, N, H, 64).permute(1, 2, 0, 3)
    vh = vf.view(Sk, N, H, 64).permute(1, 2, 0, 3)
    sc = (qh @ kh.transpose(-1, -2)) / math.sqrt(64)
    if masked:
        sc = sc.masked_fill(kpm[:, None, None, :], float("-inf"))
    P = torch.softmax(sc, dim=-1)
    if keep is not None:
        P = P * keep / (1 - p)
    ref = (P @ vh).permute(2, 0, 1, 3).reshape(Sq, N, D)
    ref.backward(do.float())

    def err(a, b):
        return ((a.float().cpu().view(b.shape) - b).norm() / b.norm()).item()
    assert err(o, ref.detach()) < 2e-2
    assert err(dq, qf.grad) < 2e-2 and err(dk, kf.grad) < 2e-2 and err(dv, vf.grad) < 2e-2
    lse_ref = torch.logsumexp(sc, dim=-1) / math.log(2.0)
    assert (lse.cpu() - lse_ref).abs().max().item() < 2e-2
