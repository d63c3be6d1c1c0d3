"""-m gpu parity tests for the DeiT-style distilled ViT and the DETR transformer encoder against the CPU oracle."""
import pytest
import torch

from helpers import O, rel_l2

LOGIT_TOL = 1.5e-2
GRAD_TOL = 3e-2


@pytest.mark.gpu
@pytest.mark.parametrize("distill", ["hard", "soft"])
def test_deit_distilled_training(distill):
    from vitb200.deit import VisionTransformerDistilled
    cfg = dict(img_size=32, patch_size=16, depth=3, num_heads=6, embed_dim=384, mlp_ratio=4, num_classes=100)
    sd = O.seeded_state_dict(O.deit_param_shapes(**cfg), 21)
    m = VisionTransformerDistilled(drop_rate=0.0, attn_drop_rate=0.0, **cfg)
    m.load_state_dict(sd)
    m = m.cuda().train()
    m.set_distilled_training(True)
    B = 6
    images, labels = O.seeded_images(B, 32, 22), O.seeded_labels(B, 100, 23)
    teacher = torch.randn(B, 100, generator=torch.Generator().manual_seed(24))
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ro, rd = O.deit_forward(ref_sd, images, patch_size=16, depth=3, num_heads=6, training=True, distilled_training=True)
    ref_loss = O.distillation_loss(ro, rd, labels, teacher, distill, 0.5, 5.0)
    ref_loss.backward()
    out, out_dist = m(images.cuda())
    loss = O.distillation_loss(out, out_dist, labels.cuda(), teacher.cuda(), distill, 0.5, 5.0)  # same loss formula on the GPU tensors
    loss.backward()
    assert rel_l2(out, ro) < LOGIT_TOL and rel_l2(out_dist, rd) < LOGIT_TOL
    assert abs(loss.item() - ref_loss.item()) < 1e-2 * max(1.0, abs(ref_loss.item()))
    worst = max(((rel_l2(p.grad, ref_sd[n].grad), n) for n, p in m.named_parameters()))
    assert worst[0] < GRAD_TOL, worst
    # eval: single averaged output (deit.py:95-96)
    m.eval()
    with torch.no_grad():
        avg = m(images.cuda())
    ref_avg = O.deit_forward(sd, images, patch_size=16, depth=3, num_heads=6, training=False, distilled_training=True)
    assert avg.shape == (B, 100) and rel_l2(avg, ref_avg) < LOGIT_TOL


def _detr_masks(eng, ws, S, N, p):
    """The keep masks the kernels used in the last training forward (regenerated from the forward's seed), keyed like the oracle's
    ExplicitDropout: (layer, site) with site 0 = dropout1, 1 = dropout, 2 = dropout2, 3 = attention weights."""
    from vitb200 import ops
    seed, dev, D, Fd, H = ws["drop_seed"], eng.flat.device, eng.D, eng.F, eng.H
    masks = {}
    for li in range(eng.L):
        masks[(li, 0)] = ops.dropout_mask(S * N * D, p, seed, eng.drop_site(li, 0), dev).view(S, N, D).cpu().float()
        masks[(li, 1)] = ops.dropout_mask(S * N * Fd, p, seed, eng.drop_site(li, 1), dev).view(S, N, Fd).cpu().float()
        masks[(li, 2)] = ops.dropout_mask(S * N * D, p, seed, eng.drop_site(li, 2), dev).view(S, N, D).cpu().float()
        masks[(li, 3)] = ops.dropout_mask(N * H * S * S, p, seed, eng.drop_site(li, 3), dev).view(N, H, S, S).cpu().float()
    return masks


def _detr_run(S, N, d_model, nhead, ffn, layers, masked, with_pos, pre_norm=False, p_drop=0.0, activation="relu"):
    from vitb200.detr import TransformerEncoder, TransformerEncoderLayer
    sd = O.seeded_state_dict(O.detr_param_shapes(d_model, ffn, layers, pre_norm), 31)
    enc = TransformerEncoder(TransformerEncoderLayer(d_model, nhead, ffn, p_drop, activation, pre_norm), layers,
                             torch.nn.LayerNorm(d_model) if pre_norm else None)   # transformer.py:32-33
    enc.load_state_dict(sd)
    enc = enc.cuda().train()
    g = torch.Generator().manual_seed(32)
    src = torch.randn(S, N, d_model, generator=g)
    pos = torch.randn(S, N, d_model, generator=g) if with_pos else None
    kpm = None
    if masked:
        valid = torch.randint(S // 2, S + 1, (N,), generator=g)
        kpm = torch.arange(S)[None, :] >= valid[:, None]
    gout = torch.randn(S, N, d_model, generator=g)
    csrc = src.cuda().requires_grad_(True)
    cpos = pos.cuda().requires_grad_(True) if with_pos else None
    out = enc(csrc, src_key_padding_mask=kpm.cuda() if masked else None, pos=cpos)
    out.backward(gout.cuda())
    drop = None
    if p_drop > 0:   # replay the kernels' own masks in the oracle (the random stream differs from PyTorch's Philox by design)
        eng = enc._get_engine()
        masks = _detr_masks(eng, eng.workspace(S, N, True), S, N, p_drop)
        for k, v in masks.items():
            assert abs(v.mean().item() - (1 - p_drop)) < 0.02, (k, v.mean().item())
        drop = O.ExplicitDropout(masks, p_drop, p_drop)
    okw = dict(nhead=nhead, num_layers=layers, normalize_before=pre_norm, activation=activation, src_key_padding_mask=kpm, drop=drop)
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rsrc = src.clone().requires_grad_(True)
    rpos = pos.clone().requires_grad_(True) if with_pos else None
    ref = O.detr_encoder_forward(ref_sd, rsrc, pos=rpos, **okw)
    ref.backward(gout)
    # Calibration (BASELINE.md §6 rule): the bf16 path must be no worse than max(1e-2, the reference's own
    # autocast-bf16 error against the same fp32 truth) -- measured here on the oracle, with 25 % head-room.
    # Post-norm layers fed with unit-variance src/pos put that floor at ~4e-2 for d_src and ~5e-2 for weights.
    ac_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    asrc = src.clone().requires_grad_(True)
    apos = pos.clone().requires_grad_(True) if with_pos else None
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ac = O.detr_encoder_forward(ac_sd, asrc, pos=apos, **okw)
    ac.float().backward(gout)
    floor_in = max(1e-2, rel_l2(asrc.grad, rsrc.grad))
    floor_w = max(1e-2, max(rel_l2(ac_sd[k].grad, ref_sd[k].grad) for k in sd))
    assert rel_l2(out, ref) < LOGIT_TOL, rel_l2(out, ref)
    assert rel_l2(csrc.grad, rsrc.grad) < 1.25 * floor_in, (rel_l2(csrc.grad, rsrc.grad), floor_in)
    if with_pos:
        assert rel_l2(cpos.grad, rpos.grad) < 1.25 * floor_in, (rel_l2(cpos.grad, rpos.grad), floor_in)
    worst = max(((rel_l2(p.grad, ref_sd[n].grad), n) for n, p in enc.named_parameters()))
    assert worst[0] < 1.25 * floor_w, (worst, floor_w)


@pytest.mark.gpu
def test_detr_encoder_masked_with_pos():
    _detr_run(S=300, N=3, d_model=512, nhead=8, ffn=2048, layers=2, masked=True, with_pos=True)


@pytest.mark.gpu
def test_detr_encoder_short_no_mask_no_pos():
    _detr_run(S=70, N=2, d_model=256, nhead=4, ffn=512, layers=3, masked=False, with_pos=False)


@pytest.mark.gpu
def test_detr_encoder_pre_norm_masked_with_pos():
    """normalize_before=True: TransformerEncoderLayer.forward_pre + the final encoder LayerNorm (transformer.py:228-241, 32-33, 112-113)."""
    _detr_run(S=300, N=3, d_model=512, nhead=8, ffn=2048, layers=2, masked=True, with_pos=True, pre_norm=True)


@pytest.mark.gpu
def test_detr_encoder_pre_norm_no_pos():
    _detr_run(S=70, N=2, d_model=256, nhead=4, ffn=512, layers=3, masked=False, with_pos=False, pre_norm=True)


@pytest.mark.gpu
@pytest.mark.parametrize("S,N,d_model,nhead,ffn,masked,with_pos,pre_norm,p,act", [
    (300, 3, 512, 8, 2048, True, True, False, 0.1, "relu"),    # the reference's default rate and layer type (transformer.py:27-28)
    (70, 2, 256, 4, 512, False, False, False, 0.1, "relu"),    # short, un-masked: dropout routes it to the 64x64-tile kernels
    (300, 2, 512, 8, 2048, True, True, True, 0.2, "relu"),     # pre-norm
    (130, 2, 256, 4, 512, True, True, False, 0.1, "gelu"),
])
def test_detr_encoder_dropout_replayed_masks(S, N, d_model, nhead, ffn, masked, with_pos, pre_norm, p, act):
    """dropout > 0 in train() mode (SURVEY.md §8 f1 for the DETR layer): attention-weight dropout inside the mma.sync attention kernels
    (any S, key-padding masks, sequence-first), dropout1 / dropout / dropout2 as in transformer.py:220-224, 236-240."""
    _detr_run(S=S, N=N, d_model=d_model, nhead=nhead, ffn=ffn, layers=2, masked=masked, with_pos=with_pos, pre_norm=pre_norm, p_drop=p,
              activation=act)


@pytest.mark.gpu
def test_detr_encoder_dropout_eval_and_new_masks():
    from vitb200.detr import TransformerEncoder, TransformerEncoderLayer
    torch.manual_seed(0)
    enc = TransformerEncoder(TransformerEncoderLayer(256, 4, 512, 0.1, "relu", False), 2).cuda().train()
    src = torch.randn(90, 2, 256, device="cuda")
    a, b = enc(src), enc(src)
    assert (a - b).abs().max().item() > 1e-3          # new masks every forward
    enc.eval()
    with torch.no_grad():
        c, d = enc(src), enc(src)
    assert torch.equal(c, d)                           # eval(): dropout is the identity


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hard", "soft"])
def test_fused_distillation_loss_kernel(kind):
    """vb_distill_loss against the oracle's DistillationLoss on the inputs of the reference-generated golden fixture
    (tests/golden/distill_loss.pt): loss to 1e-5 of the reference's own value, both logits gradients to 1e-4 (fp32 accumulations)."""
    import os
    from vitb200 import ops
    gd = torch.load(os.path.join(os.path.dirname(__file__), "golden", "distill_loss.pt"))
    g = torch.Generator().manual_seed(gd["seed"])
    B, C = 16, 100
    out, kd = torch.randn(B, C, generator=g), torch.randn(B, C, generator=g)
    labels = torch.randint(0, C, (B,), generator=g)
    W = torch.randn(3 * 8 * 8, C, generator=g) * 0.05
    teacher = torch.randn(B, 3, 8, 8, generator=g).flatten(1) @ W
    if kind == "hard":
        teacher[3, 7] = teacher[3, 50] = teacher[3].max() + 1.0   # tie: torch.argmax takes the first maximal index
    ro, rk = out.clone().requires_grad_(True), kd.clone().requires_grad_(True)
    ref = O.distillation_loss(ro, rk, labels, teacher, kind, 0.5, 5.0)
    ref.backward()
    loss = torch.zeros(1, device="cuda")
    correct = torch.zeros(1, device="cuda", dtype=torch.int32)
    dz, dk = torch.empty(B, C, device="cuda"), torch.empty(B, C, device="cuda")
    dzb, dkb = torch.empty(B, C, device="cuda", dtype=torch.bfloat16), torch.empty(B, C, device="cuda", dtype=torch.bfloat16)
    ops.distill_loss(out.cuda(), kd.cuda(), teacher.cuda(), labels.cuda(), loss, kind=kind, alpha=0.5, tau=5.0, dlogits_f32=dz,
                     dlogits_kd_f32=dk, dlogits_bf16=dzb, dlogits_kd_bf16=dkb, correct_accum=correct)
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    if kind == "soft":
        assert abs(loss.item() - gd["soft"]) < 1e-5      # the unmodified reference's value on the same inputs
    assert rel_l2(dz, ro.grad) < 1e-4 and rel_l2(dk, rk.grad) < 1e-4
    assert rel_l2(dzb, ro.grad) < 1e-2 and rel_l2(dkb, rk.grad) < 1e-2
    assert correct.item() == (out.argmax(1) == labels).sum().item()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hard", "soft"])
def test_distillation_trainer_step(kind):
    """Trainer(distillation=...) (deit.py:57-70's loop body: teacher under no_grad, DistillationLoss, backward, Adam) against the
    oracle: loss, and the first Adam update's direction through the flat gradient buffer."""
    from vitb200.deit import VisionTransformerDistilled
    from vitb200.trainer import Trainer
    cfg = dict(img_size=32, patch_size=16, depth=2, num_heads=6, embed_dim=384, mlp_ratio=4, num_classes=100)
    sd = O.seeded_state_dict(O.deit_param_shapes(**cfg), 41)
    m = VisionTransformerDistilled(drop_rate=0.0, attn_drop_rate=0.0, **cfg)
    m.load_state_dict(sd)
    m = m.cuda().train()
    m.set_distilled_training(True)
    B = 8
    images, labels = O.seeded_images(B, 32, 42), O.seeded_labels(B, 100, 43)
    Wt = torch.randn(3 * 32 * 32, 100, generator=torch.Generator().manual_seed(44)) * 0.02
    teacher = torch.nn.Linear(3 * 32 * 32, 100, bias=False)
    teacher.weight.data.copy_(Wt.t())
    teacher_mod = torch.nn.Sequential(torch.nn.Flatten(1), teacher).cuda()
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ro, rd = O.deit_forward(ref_sd, images, patch_size=16, depth=2, num_heads=6, training=True, distilled_training=True)
    ref_loss = O.distillation_loss(ro, rd, labels, images.flatten(1) @ Wt, kind, 0.5, 5.0)
    ref_loss.backward()
    tr = Trainer(m, lr=1e-4, distillation=dict(teacher=teacher_mod, type=kind, alpha=0.5, tau=5.0), use_cuda_graph=False)
    loss = tr.step(images.cuda(), labels.cuda())
    assert abs(loss.item() - ref_loss.item()) < 1e-2 * max(1.0, abs(ref_loss.item()))
    eng = m._get_engine()
    worst = max(((rel_l2(p.grad, ref_sd[n].grad), n) for n, p in m.named_parameters()))
    assert worst[0] < GRAD_TOL, worst
    # graph path: three more steps replay the captured step (teacher eager, ahead of the replay) and the loss goes down
    tr2 = Trainer(m, lr=1e-3, distillation=dict(teacher=teacher_mod, type=kind, alpha=0.5, tau=5.0))
    losses = [tr2.step(images.cuda(), labels.cuda()).item() for _ in range(6)]
    assert losses[-1] < losses[0], losses
    assert eng is m._get_engine()
