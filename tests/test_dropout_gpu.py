"""-m gpu: dropout.  The kernels' masks are counter-based hashes (vitb200/csrc/dropout.cuh), not PyTorch's
Philox stream, so parity is checked by replaying the SAME masks in the oracle: every site's keep mask is dumped through
vb_dropout_mask_u8 with the seed the forward used, and the oracle applies keep * x / (1 - p) at the reference's dropout sites
(vanilla_vit.py:38,42,67-68,78,94,104).  Tolerances as in test_vit_parity_gpu.py (bf16 tensor-core path vs fp32 oracle)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _masks(eng, ws, B, p_hidden, p_attn):
    from vitb200 import ops
    S, D, Fd, H, L = eng.S, eng.D, eng.F, eng.H, eng.L
    seed, dev = ws["drop_seed"], eng.flat.device
    masks = {}
    if p_hidden > 0:
        masks["embed"] = ops.dropout_mask(B * S * D, p_hidden, seed, eng.EMBED_SITE, dev).view(B, S, D).cpu().float()
    for li in range(L):
        if p_hidden > 0:
            masks[(li, 0)] = ops.dropout_mask(B * S * D, p_hidden, seed, eng.drop_site(li, 0), dev).view(B, S, D).cpu().float()
            masks[(li, 1)] = ops.dropout_mask(B * S * Fd, p_hidden, seed, eng.drop_site(li, 1), dev).view(B, S, Fd).cpu().float()
            masks[(li, 2)] = ops.dropout_mask(B * S * D, p_hidden, seed, eng.drop_site(li, 2), dev).view(B, S, D).cpu().float()
        if p_attn > 0:
            masks[(li, 3)] = ops.dropout_mask(B * H * S * S, p_attn, seed, eng.drop_site(li, 3), dev).view(B, H, S, S).cpu().float()
    return masks


@pytest.mark.gpu
@pytest.mark.parametrize("p_hidden,p_attn", [(0.1, 0.0), (0.0, 0.1), (0.1, 0.1), (0.25, 0.2)])
def test_vit_dropout_matches_oracle_with_replayed_masks(p_hidden, p_attn):
    from oracle import vit_oracle as O
    from vitb200.vit import ViT
    cfg = dict(image_size=32, patch_size=4, num_layers=3, num_heads=4, hidden_dim=256, mlp_dim=512, num_classes=10)
    B = 6
    sd = O.seeded_state_dict(O.vit_param_shapes(**cfg), 21)
    m = ViT(32, 4, 3, 4, 256, 512, p_hidden, p_attn, 10)
    m.load_state_dict(sd)
    m = m.cuda().train()
    images, labels = O.seeded_images(B, 32, 22), O.seeded_labels(B, 10, 23)
    logits = m(images.cuda())
    loss = torch.nn.functional.cross_entropy(logits, labels.cuda())
    loss.backward()
    eng = m._get_engine()
    ws = eng.workspace(B, True)
    masks = _masks(eng, ws, B, p_hidden, p_attn)
    for k, v in masks.items():   # keep fraction ~ 1 - p
        p = p_attn if (isinstance(k, tuple) and k[1] == 3) else p_hidden
        assert abs(v.mean().item() - (1 - p)) < 0.02, (k, v.mean().item())
    drop = O.ExplicitDropout(masks, p_hidden, p_attn)
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.vit_forward(ref_sd, images, patch_size=4, num_layers=3, num_heads=4, drop=drop)
    torch.nn.functional.cross_entropy(ref, labels).backward()
    err = ((logits.float().cpu() - ref).norm() / ref.norm()).item()
    assert err < 1.5e-2, f"logits rel-L2 {err}"            # bf16 operands vs fp32 oracle: 1.5e-2
    worst = max(((p.grad.float().cpu() - ref_sd[n].grad).norm() / (ref_sd[n].grad.norm() + 1e-30)).item() for n, p in m.named_parameters())
    assert worst < 3e-2, f"worst gradient rel-L2 {worst}"  # gradients: 3e-2
    # a second forward draws new masks; eval() ignores dropout
    logits2 = m(images.cuda())
    assert (logits2 - logits).abs().max().item() > 1e-4
    m.eval()
    with torch.no_grad():
        e1, e2 = m(images.cuda()), m(images.cuda())
    assert torch.equal(e1, e2)
    ref_eval = O.vit_forward(sd, images, patch_size=4, num_layers=3, num_heads=4)
    assert ((e1.float().cpu() - ref_eval).norm() / ref_eval.norm()).item() < 1.5e-2


@pytest.mark.gpu
def test_trainer_with_dropout_changes_masks_every_graph_replay():
    from vitb200.trainer import Trainer
    from vitb200.vit import ViT
    torch.manual_seed(0)
    m = ViT(32, 4, 2, 4, 256, 512, 0.1, 0.1, 10)
    with torch.no_grad():
        m.heads.head.weight.normal_(std=0.02)
    m = m.cuda().train()
    tr = Trainer(m, lr=0.0)      # lr 0: parameters stay put, so the loss changes only through the dropout masks
    images = torch.randn(8, 3, 32, 32, device="cuda")
    labels = torch.randint(0, 10, (8,), device="cuda")
    losses = [tr.step(images, labels).item() for _ in range(5)]   # eager, capture, 3 replays
    assert len({round(l, 6) for l in losses}) == 5, losses
