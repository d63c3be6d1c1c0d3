"""-m gpu parity tests: the CUDA path (through the public nn.Module API -> C ABI) against the fp32 CPU oracle on
the same seeded weights and inputs.

Tolerance (BASELINE.json north_star; calibration in BASELINE.md §6): the bf16 tensor-core path is compared with
the fp32 oracle by per-tensor relative L2; logits and gradients must be within 1e-2 ... 2e-2 (the reference's own
autocast-bf16 run sits at 0.8-1.2e-2 against fp32 on these shapes).  The loss is a function of the bf16-computed
logits, so it inherits their tolerance (1e-2 relative); fp32-accumulated quantities given identical inputs (LayerNorm
statistics, softmax LSE, column sums) are checked to 1e-4 or better in tests/test_kernels_gpu.py.
"""
import pytest
import torch

from helpers import O, oracle_vit_run, rel_l2

TINY = dict(image_size=32, patch_size=4, num_layers=7, num_heads=4, hidden_dim=256, mlp_dim=512, num_classes=10)
B16_2L = dict(image_size=224, patch_size=16, num_layers=2, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)

LOGIT_TOL = 1.5e-2
GRAD_TOL = 3e-2


def build(cfg, seed):
    from vitb200.vit import ViT
    m = ViT(cfg["image_size"], cfg["patch_size"], cfg["num_layers"], cfg["num_heads"], cfg["hidden_dim"], cfg["mlp_dim"], 0.0, 0.0,
            cfg["num_classes"])
    sd = O.seeded_state_dict(O.vit_param_shapes(**cfg), seed)
    m.load_state_dict(sd)
    return m.cuda(), sd


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,batch", [(TINY, 8), (TINY, 37), (B16_2L, 3)], ids=["tiny-b8", "tiny-b37", "b16x2-b3"])
def test_vit_logits_and_grads(cfg, batch):
    m, sd = build(cfg, seed=1)
    images = O.seeded_images(batch, cfg["image_size"], seed=2)
    labels = O.seeded_labels(batch, cfg["num_classes"], seed=3)
    ref_logits, ref_loss, ref_grads = oracle_vit_run(cfg, sd, images, labels)
    m.train()
    logits = m(images.cuda())
    loss = torch.nn.functional.cross_entropy(logits, labels.cuda())
    loss.backward()
    assert logits.shape == ref_logits.shape
    e = rel_l2(logits, ref_logits)
    assert e < LOGIT_TOL, f"logits rel-L2 {e:.3e}"
    assert abs(loss.item() - ref_loss.item()) < 1e-2 * max(1.0, abs(ref_loss.item())), (loss.item(), ref_loss.item())
    worst = ("", 0.0)
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        ge = rel_l2(p.grad, ref_grads[name])
        if ge > worst[1]:
            worst = (name, ge)
    assert worst[1] < GRAD_TOL, f"worst gradient {worst[0]}: rel-L2 {worst[1]:.3e}"


@pytest.mark.gpu
def test_vit_forward_features_and_eval():
    cfg, batch = TINY, 5
    m, sd = build(cfg, seed=4)
    images = O.seeded_images(batch, cfg["image_size"], seed=5)
    gout = torch.randn(batch, (cfg["image_size"] // cfg["patch_size"]) ** 2 + 1, cfg["hidden_dim"], generator=torch.Generator().manual_seed(6))
    ref_feat, _, ref_grads = oracle_vit_run(cfg, sd, images, want="features", grad_out=gout)
    m.train()
    feat = m.forward_features(images.cuda())
    feat.backward(gout.cuda())
    assert rel_l2(feat, ref_feat) < LOGIT_TOL
    worst = max(rel_l2(p.grad, ref_grads[n]) for n, p in m.named_parameters() if n in ref_grads)
    assert worst < GRAD_TOL, worst
    # eval / no_grad path gives the same logits as the training path
    m.eval()
    with torch.no_grad():
        l_eval = m(images.cuda())
    ref_logits, _, _ = oracle_vit_run(cfg, sd, images)
    assert rel_l2(l_eval, ref_logits) < LOGIT_TOL


@pytest.mark.gpu
def test_grad_accumulation_and_zero_grad():
    cfg, batch = TINY, 4
    m, sd = build(cfg, seed=7)
    images = O.seeded_images(batch, cfg["image_size"], seed=8).cuda()
    labels = O.seeded_labels(batch, cfg["num_classes"], seed=9).cuda()
    m.train()
    torch.nn.functional.cross_entropy(m(images), labels).backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
    torch.nn.functional.cross_entropy(m(images), labels).backward()   # accumulates, like autograd
    for n, p in m.named_parameters():
        assert rel_l2(p.grad, 2 * g1[n]) < 1e-3, n
    m.zero_grad(set_to_none=True)
    torch.nn.functional.cross_entropy(m(images), labels).backward()
    for n, p in m.named_parameters():
        assert rel_l2(p.grad, g1[n]) < 1e-3, n


@pytest.mark.gpu
def test_device_prefetcher_delivers_every_batch_in_order():
    """vitb200.data.DevicePrefetcher: double-buffered pinned staging must hand over exactly the loader's batches,
    including a ragged last one, while the consumer keeps the previous batch busy on the compute stream."""
    import torch
    from vitb200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(8 if i < 6 else 3, 3, 32, 32, generator=g), torch.randint(0, 10, (8 if i < 6 else 3,), generator=g)) for i in range(7)]
    seen = []
    for img, lab in DevicePrefetcher(batches):
        assert img.is_cuda and lab.is_cuda
        busy = img.float() @ torch.randn(32, 32, device="cuda")   # keep the consumer stream busy with this batch
        seen.append((img.clone(), lab.clone(), busy.sum()))
    torch.cuda.synchronize()
    assert len(seen) == len(batches)
    for (img, lab, _), (himg, hlab) in zip(seen, batches):
        assert torch.equal(img.cpu(), himg) and torch.equal(lab.cpu(), hlab)


@pytest.mark.gpu
def test_device_prefetcher_host_running_ahead_of_the_gpu():
    """with a sync-free consumer the host runs several batches ahead of the GPU; the pinned staging buffer of a
    slot must not be overwritten before its previous (GPU-delayed) H2D copy has left it.  The consumer queues ~20 ms of GPU work
    per batch and never synchronises, the loader is instantaneous."""
    from vitb200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(6)
    batches = [(torch.full((64, 3, 64, 64), float(i)) + torch.randn(64, 3, 64, 64, generator=g) * 0.01, torch.full((64,), i)) for i in range(12)]
    heavy = torch.randn(8192, 8192, device="cuda")
    sums, labs = [], []
    for img, lab in DevicePrefetcher(batches):
        for _ in range(6):
            heavy = torch.tanh(heavy @ heavy * 1e-4)       # keeps the compute stream far behind the host
        sums.append(img.mean() + 0 * heavy[0, 0])
        labs.append(lab.float().mean())
    torch.cuda.synchronize()
    for i, (s, l) in enumerate(zip(sums, labs)):
        assert abs(s.item() - i) < 0.01 and l.item() == i, (i, s.item(), l.item())


@pytest.mark.gpu
def test_standalone_encoder_matches_oracle():
    """vitb200.vit.Encoder used on its own (vanilla_vit.py:88-106; T2T_ViT's identical copy t2t_vit.py:89-110): tokens in,
    normalised tokens out, gradients for the tokens and every parameter."""
    import torch
    import torch.nn.functional as F
    from oracle import vit_oracle as O
    from vitb200.vit import Encoder
    B, S, D, H, Fd, L = 5, 65, 256, 4, 512, 3
    enc = Encoder(S, L, H, D, Fd, 0.0, 0.0)
    g = torch.Generator().manual_seed(41)
    with torch.no_grad():
        for p in enc.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05 + (1.0 if p.dim() == 1 and "ln" in "" else 0.0))
        for n, p in enc.named_parameters():
            if n.endswith("ln_1.weight") or n.endswith("ln_2.weight") or n == "ln.weight":
                p.add_(1.0)
    sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    enc = enc.cuda().train()
    x = torch.randn(B, S, D, generator=g)
    gout = torch.randn(B, S, D, generator=g)
    xc = x.cuda().requires_grad_(True)
    out = enc(xc)
    out.backward(gout.cuda())
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    h = xr + ref_sd["pos_embedding"]
    for i in range(L):
        h = O.encoder_block(h, ref_sd, f"layers.encoder_layer_{i}.", H, 1e-6)
    ref = F.layer_norm(h, (D,), ref_sd["ln.weight"], ref_sd["ln.bias"], 1e-6)
    ref.backward(gout)
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(out, ref) < 1.5e-2
    assert rel(xc.grad, xr.grad) < 3e-2
    worst = max((rel(p.grad, ref_sd[n].grad), n) for n, p in enc.named_parameters())
    assert worst[0] < 3e-2, worst
    enc.eval()
    with torch.no_grad():
        assert rel(enc(x.cuda()), ref) < 1.5e-2


@pytest.mark.gpu
@pytest.mark.parametrize("S,D,H,Fd", [(65, 256, 4, 512), (197, 768, 12, 3072)])
def test_standalone_encoder_block_matches_oracle(S, D, H, Fd):
    """vitb200.vit.EncoderBlock called on its own (vanilla_vit.py:73-83 — what nn.Sequential does inside the reference's Encoder): tokens
    in, tokens out (no position embedding, no final norm), gradients for the tokens and every parameter; a second sequence length
    through the same module; then the same block inside an Encoder still works (its engine re-binds the parameters)."""
    import torch
    from oracle import vit_oracle as O
    from vitb200.vit import EncoderBlock
    B = 3
    blk = EncoderBlock(H, D, Fd, 0.0, 0.0)
    g = torch.Generator().manual_seed(43)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            if n in ("ln_1.weight", "ln_2.weight"):
                p.add_(1.0)
    sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    blk = blk.cuda().train()
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    for s_len in (S, S - 7):
        x = torch.randn(B, s_len, D, generator=g)
        gout = torch.randn(B, s_len, D, generator=g)
        blk.zero_grad()
        xc = x.cuda().requires_grad_(True)
        out = blk(xc)
        out.backward(gout.cuda())
        ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xr = x.clone().requires_grad_(True)
        ref = O.encoder_block(xr, ref_sd, "", H, 1e-6)
        ref.backward(gout)
        assert rel(out, ref) < 1.5e-2                       # one block of bf16 GEMMs against fp32: same flat bounds as the small-shape tests
        assert rel(xc.grad, xr.grad) < 3e-2
        worst = max((rel(p.grad, ref_sd[n].grad), n) for n, p in blk.named_parameters())
        assert worst[0] < 3e-2, worst
    blk.eval()
    with torch.no_grad():
        assert rel(blk(x.cuda()), ref) < 1.5e-2
    with pytest.raises(RuntimeError):
        blk(x)                                              # CPU tensor: no fallback


@pytest.mark.gpu
def test_small_batch_inference_graph_replay_matches_eager(monkeypatch):
    """Eval forwards at small batch are replayed from a CUDA graph (engine.forward_inference): bit-identical to the eager kernel
    sequence for fresh inputs, and parameter edits between calls are seen (the bf16 cast is part of the graph)."""
    from vitb200.vit import ViT
    torch.manual_seed(3)
    m = ViT(32, 4, 3, 4, 256, 512, 0.1, 0.1, 10)
    with torch.no_grad():
        m.heads.head.weight.normal_(std=0.05)
    m = m.cuda().eval()
    xs = [torch.randn(3, 3, 32, 32, device="cuda") for _ in range(5)]
    with torch.no_grad():
        got = [m(x) for x in xs]                     # eager, capture + replay, replay, ...
        feats = [m.forward_features(x) for x in xs[:3]]
    eng = m._get_engine()
    assert any(isinstance(v, tuple) for v in eng._infer_graphs.values()), "no inference graph was captured"
    monkeypatch.setenv("VITB200_INFER_GRAPH", "0")
    with torch.no_grad():
        want = [m(x) for x in xs]
        want_f = [m.forward_features(x) for x in xs[:3]]
    for a, b in zip(got + feats, want + want_f):
        assert torch.equal(a, b)
    monkeypatch.setenv("VITB200_INFER_GRAPH", "1")
    with torch.no_grad():
        m.heads.head.bias.add_(1.0)
        shifted = m(xs[0])
    assert torch.allclose(shifted, want[0] + 1.0, atol=1e-5)
    # a training step in between (other workspaces, fused optimizer) does not disturb the captured graph
    m.train()
    loss = torch.nn.functional.cross_entropy(m(xs[1]), torch.randint(0, 10, (3,), device="cuda"))
    loss.backward()
    m.eval()
    with torch.no_grad():
        again = m(xs[0])
    assert torch.equal(again, shifted)


def _loop(monkeypatch, mode, steps=5, p=0.0):
    """The reference's loop body (base.py:53-57) with torch.optim.SGD; returns losses and the final parameters."""
    from vitb200.vit import ViT
    monkeypatch.setenv("VITB200_AUTOGRAD_GRAPH", mode)
    torch.manual_seed(11)
    m = ViT(32, 4, 2, 4, 256, 512, p, p, 10)
    with torch.no_grad():
        m.heads.head.weight.normal_(std=0.05)
    m = m.cuda().train()
    init = [p_.detach().clone() for p_ in m.parameters()]
    opt = torch.optim.SGD(m.parameters(), lr=0.002)
    g = torch.Generator().manual_seed(12)
    losses = []
    for _ in range(steps):
        x, y = torch.randn(8, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (8,), generator=g).cuda()
        opt.zero_grad()                      # set_to_none=True: p.grad is rebuilt as a view of the flat buffer every step
        loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return losses, [p_.detach().clone() for p_ in m.parameters()], m, init


@pytest.mark.gpu
def test_autograd_path_graph_replay_matches_eager(monkeypatch):
    """Forward and backward of the autograd node are replayed from CUDA graphs for launch-bound sizes (engine.forward_train /
    backward_train): the same kernels on the same buffers as the eager sequence."""
    le, pe, _, init = _loop(monkeypatch, "0")
    lg, pg, m, _ = _loop(monkeypatch, "1")
    # same kernels on the same buffers; split-K reduce-adds and atomic column sums make fp32 sums order-dependent, hence 1e-4 not 0
    assert all(abs(a - b) <= 1e-3 * max(1.0, abs(a)) for a, b in zip(le, lg)), (le, lg)
    # the accumulated UPDATES agree to bf16-path resolution (the two trajectories see order-dependent fp32 sums; the key bias, whose
    # gradient is exactly zero in exact arithmetic, is pure rounding noise and gets an absolute floor)
    bad = [(n, (a - b).norm().item(), (a - i).norm().item()) for (n, _), a, b, i in zip(m.named_parameters(), pe, pg, init)
           if (a - b).norm().item() > 5e-2 * (a - i).norm().item() + 1e-6 * a.numel() ** 0.5]   # gross errors (wrong buffers) are O(1)
    assert not bad, bad
    kinds = {k[0] for k, v in m._get_engine()._train_graphs.items() if isinstance(v, tuple)}
    assert kinds == {"fwd", "bwd"}, kinds
    # with dropout every replay draws new masks (the seed counter is bumped inside the graph)
    ld = _loop(monkeypatch, "1", steps=4, p=0.2)[0]
    assert len({round(v, 6) for v in ld}) == 4
