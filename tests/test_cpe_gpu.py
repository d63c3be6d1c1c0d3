"""-m gpu: conditional positional encodings — the depthwise-conv kernels against torch's conv2d on the same inputs
(fp32 accumulations: 1e-4, stated below) and the CPEViT / CPVT / CPVTGAP drop-ins against the CPU oracle (bf16 tensor-core path vs
fp32 oracle: logits 1.5e-2, gradients 3e-2 relative L2, as for the ViT)."""
import pytest
import torch

from helpers import O, rel_l2


def _ref_cpe(x, w, b, n_prefix=1):
    return O.cond_pos_encoding(x, w, b) if n_prefix == 1 else None


@pytest.mark.gpu
@pytest.mark.parametrize("B,G,D", [(3, 8, 256), (2, 14, 768), (5, 2, 128), (1, 1, 128)])
def test_dwconv_kernels_match_torch(B, G, D):
    from vitb200 import ops
    g = torch.Generator().manual_seed(5)
    S = G * G + 1
    x = torch.randn(B, S, D, generator=g)
    w = torch.randn(D, 1, 3, 3, generator=g) * 0.3
    b = torch.randn(D, generator=g)
    pos = torch.randn(S, D, generator=g)
    sub = torch.randn(B, S, D, generator=g)
    dy = torch.randn(B, S, D, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = O.cond_pos_encoding(xr, wr, br)
    ref.backward(dy)
    xc, wc, bc, dyc = x.cuda(), w.cuda(), b.cuda(), dy.cuda()
    out = torch.empty_like(xc)
    ops.dwconv_fwd(xc, wc, bc, out, n_prefix=1)
    assert rel_l2(out, ref) < 1e-5
    ops.dwconv_fwd(xc, wc, bc, out, n_prefix=1, pos=pos.cuda(), sub=sub.cuda())
    assert rel_l2(out, ref.detach() + pos + (x - sub)) < 1e-5
    dx, s32, sbf = torch.empty_like(xc), torch.empty_like(xc), torch.empty_like(xc, dtype=torch.bfloat16)
    ops.dwconv_bwd_data(dyc, wc, n_prefix=1, dx=dx, sum_f32=s32, sum_bf16=sbf)
    assert rel_l2(dx, xr.grad) < 1e-5
    assert rel_l2(s32, xr.grad + dy) < 1e-5 and rel_l2(sbf, xr.grad + dy) < 1e-2
    dw, db = torch.zeros_like(wc), torch.zeros_like(bc)
    ops.dwconv_bwd_weight(dyc, xc, dw.view(-1), db, n_prefix=1)
    assert rel_l2(dw, wr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4      # fp32 accumulation over B*G*G terms (atomics): 1e-4


def _model_case(which, layers, p_hidden=0.0, p_attn=0.0, B=6):
    from vitb200 import cpvt as ours
    cfg = dict(image_size=32, patch_size=4, num_layers=layers, num_heads=4, hidden_dim=256, mlp_dim=512, num_classes=10)
    peg = which != "CPEViT"
    sd = O.seeded_state_dict(O.cpe_param_shapes(**cfg, peg_blocks=peg), 51)
    m = getattr(ours, which)(32, 4, layers, 4, 256, 512, p_hidden, p_attn, 10)
    m.load_state_dict(sd)
    m = m.cuda().train()
    images, labels = O.seeded_images(B, 32, 52), O.seeded_labels(B, 10, 53)
    logits = m(images.cuda())
    torch.nn.functional.cross_entropy(logits, labels.cuda()).backward()
    drop = None
    if p_hidden > 0 or p_attn > 0:
        from test_dropout_gpu import _masks
        eng = m._get_engine()
        drop = O.ExplicitDropout(_masks(eng, eng.workspace(B, True), B, p_hidden, p_attn), p_hidden, p_attn)
    kw = dict(patch_size=4, num_layers=layers, num_heads=4, peg_blocks=peg)
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.cpe_forward(ref_sd, images, drop=drop, **kw)
    torch.nn.functional.cross_entropy(ref, labels).backward()
    assert rel_l2(logits, ref) < 1.5e-2, rel_l2(logits, ref)
    worst = max(((rel_l2(p.grad, ref_sd[n].grad), n) for n, p in m.named_parameters()))
    assert worst[0] < 3e-2, worst
    return m, sd, images, kw


@pytest.mark.gpu
@pytest.mark.parametrize("which,layers", [("CPEViT", 3), ("CPVT", 3), ("CPVTGAP", 2)])
def test_cpe_models_match_oracle(which, layers):
    m, sd, images, kw = _model_case(which, layers)
    # features (dense backward path) and eval
    feats = m.forward_features(images.cuda())
    gout = torch.randn(feats.shape, generator=torch.Generator().manual_seed(54))
    m.zero_grad()
    feats.backward(gout.cuda())
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rf = O.cpe_forward_features(ref_sd, images, **kw)
    rf.backward(gout)
    assert rel_l2(feats, rf) < 1.5e-2
    worst = max(((rel_l2(p.grad, ref_sd[n].grad), n) for n, p in m.named_parameters() if ref_sd[n].grad is not None))
    assert worst[0] < 3e-2, worst
    m.eval()
    with torch.no_grad():
        a, b = m(images.cuda()), m(images.cuda())      # second call replays the captured inference graph
    assert torch.equal(a, b) and rel_l2(a, O.cpe_forward(sd, images, **kw)) < 1.5e-2


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["CPEViT", "CPVT"])
def test_cpe_models_dropout_replayed_masks(which):
    _model_case(which, 2, p_hidden=0.1, p_attn=0.1)


@pytest.mark.gpu
def test_cpvt_trainer_step_decreases_loss():
    from vitb200 import cpvt as ours
    from vitb200.trainer import Trainer
    torch.manual_seed(0)
    m = ours.CPVT(32, 4, 2, 4, 256, 512, 0.0, 0.0, 10)
    with torch.no_grad():
        m.heads.head.weight.normal_(std=0.02)
    m = m.cuda().train()
    tr = Trainer(m, lr=1e-3)
    images, labels = torch.randn(16, 3, 32, 32, device="cuda"), torch.randint(0, 10, (16,), device="cuda")
    losses = [tr.step(images, labels).item() for _ in range(8)]
    assert losses[-1] < losses[0], losses
