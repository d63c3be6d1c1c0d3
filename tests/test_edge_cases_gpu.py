"""-m gpu: edge cases of the module contract: batch 1, ragged last batches, non-contiguous / non-fp32 inputs, the
reference's shape assertions, fully padded DETR samples."""
import pytest
import torch

from helpers import O, rel_l2

CFG = dict(image_size=32, patch_size=4, num_layers=2, num_heads=4, hidden_dim=256, mlp_dim=512, num_classes=10)
KW = dict(patch_size=4, num_layers=2, num_heads=4)


def _model(train=True):
    from vitb200.vit import ViT
    sd = O.seeded_state_dict(O.vit_param_shapes(**CFG), 71)
    m = ViT(32, 4, 2, 4, 256, 512, 0.0, 0.0, 10)
    m.load_state_dict(sd)
    m = m.cuda()
    return (m.train() if train else m.eval()), sd


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 2, 5])
def test_tiny_batches_train(batch):
    """A loader's ragged last batch (len(dataset) % batch_size; base.py:51 does not drop it)."""
    m, sd = _model()
    x, y = O.seeded_images(batch, 32, 72), O.seeded_labels(batch, 10, 73)
    out = m(x.cuda())
    torch.nn.functional.cross_entropy(out, y.cuda()).backward()
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.vit_forward(ref_sd, x, **KW)
    torch.nn.functional.cross_entropy(ref, y).backward()
    # 1.5e-2 is the bound for tensors with hundreds of entries; ten bf16-path logits of ONE sample scatter more (measured 1.7e-2)
    assert out.shape == (batch, 10) and rel_l2(out, ref) < (2.5e-2 if batch == 1 else 1.5e-2)
    worst = max(((rel_l2(p.grad, ref_sd[n].grad), n) for n, p in m.named_parameters()))
    assert worst[0] < (4e-2 if batch == 1 else 3e-2), worst


@pytest.mark.gpu
def test_alternating_batch_sizes_reuse_their_workspaces():
    m, sd = _model()
    for b in (8, 3, 8, 3, 1, 8):
        x = O.seeded_images(b, 32, 80 + b)
        out = m(x.cuda())
        out.sum().backward()
        assert rel_l2(out, O.vit_forward(sd, x, **KW)) < 1.5e-2


@pytest.mark.gpu
def test_input_layouts_and_dtypes():
    m, sd = _model(train=False)
    x = O.seeded_images(4, 32, 74)
    ref = O.vit_forward(sd, x, **KW)
    with torch.no_grad():
        a = m(x.cuda().to(memory_format=torch.channels_last))                 # non-contiguous NCHW view
        b = m(x.cuda().double())                                             # fp64 batch
        c = m(x.cuda()[:, :, :, :].expand(4, 3, 32, 32))
        h = m(x.cuda().half())                                               # fp16 batch: rounded inputs, same path
    assert rel_l2(a, ref) < 1.5e-2 and rel_l2(b, ref) < 1.5e-2 and rel_l2(c, ref) < 1.5e-2
    assert rel_l2(h, O.vit_forward(sd, x.half().float(), **KW)) < 1.5e-2


@pytest.mark.gpu
def test_reference_shape_assertions():
    m, _ = _model(train=False)
    with pytest.raises(Exception):
        m(torch.zeros(2, 3, 16, 16, device="cuda"))                           # vanilla_vit.py:190-191
    with pytest.raises(Exception):
        m(torch.zeros(2, 3, 32, 16, device="cuda"))
    with pytest.raises(Exception):
        m(torch.zeros(2, 1, 32, 32, device="cuda"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 3, 32, 32))                                          # host tensor into a CUDA model: loud error, no fallback
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.train()(torch.zeros(2, 3, 32, 32))
    m.eval()
    from vitb200.detr import TransformerDecoder, TransformerDecoderLayer, TransformerEncoder, TransformerEncoderLayer
    enc2 = TransformerEncoder(TransformerEncoderLayer(256, 4, 512, 0.0, "relu", False), 1).cuda()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc2(torch.zeros(10, 2, 256))
    with pytest.raises(RuntimeError):
        enc2(torch.zeros(10, 2, 128, device="cuda"))                          # wrong d_model
    with pytest.raises(RuntimeError):
        enc2(torch.zeros(10, 2, 256, device="cuda"), src_key_padding_mask=torch.zeros(2, 9, dtype=torch.bool, device="cuda"))
    dec2 = TransformerDecoder(TransformerDecoderLayer(256, 4, 512, 0.0, "relu", False), 1).cuda()
    with pytest.raises(RuntimeError):
        dec2(torch.zeros(5, 2, 256, device="cuda"), torch.zeros(10, 3, 256, device="cuda"))     # batch mismatch
    torch.cuda.synchronize()                                                  # nothing illegal was launched
    enc = m.encoder
    with pytest.raises(Exception):
        enc(torch.zeros(2, 65, device="cuda"))                                # vanilla_vit.py:103: 3-D input expected


@pytest.mark.gpu
def test_detr_encoder_fully_padded_sample_is_finite_elsewhere():
    """A sample whose keys are ALL padding: the reference's softmax over an all -inf row is NaN for that sample only; here that
    sample's attention output is defined as zero (l = 0 -> 0), and the other samples are unaffected either way."""
    from vitb200.detr import TransformerEncoder, TransformerEncoderLayer
    sd = O.seeded_state_dict(O.detr_param_shapes(256, 512, 2, False), 75)
    enc = TransformerEncoder(TransformerEncoderLayer(256, 4, 512, 0.0, "relu", False), 2)
    enc.load_state_dict(sd)
    enc = enc.cuda().eval()
    g = torch.Generator().manual_seed(76)
    src, pos = torch.randn(70, 3, 256, generator=g), torch.randn(70, 3, 256, generator=g)
    kpm = torch.zeros(3, 70, dtype=torch.bool)
    kpm[1, :] = True
    kpm[2, 50:] = True
    with torch.no_grad():
        out = enc(src.cuda(), src_key_padding_mask=kpm.cuda(), pos=pos.cuda())
    ref = O.detr_encoder_forward(sd, src, nhead=4, num_layers=2, src_key_padding_mask=kpm, pos=pos)
    assert torch.isfinite(out).all()
    for b in (0, 2):
        assert rel_l2(out[:, b], ref[:, b]) < 1.5e-2


@pytest.mark.gpu
@pytest.mark.parametrize("graphs", ["0", "1"])
def test_two_forwards_before_backward_and_eval_in_between(monkeypatch, graphs):
    """Saved activations survive later forwards (PyTorch saved-tensor semantics): out1 = m(x1); out2 = m(x2); an evaluation forward;
    then ONE backward over both — the gradients are the sum of the two single-batch gradients."""
    monkeypatch.setenv("VITB200_AUTOGRAD_GRAPH", graphs)
    m, sd = _model()
    x1, x2 = O.seeded_images(4, 32, 91), O.seeded_images(4, 32, 92)
    y1, y2 = O.seeded_labels(4, 10, 93), O.seeded_labels(4, 10, 94)
    ce = torch.nn.functional.cross_entropy
    for _ in range(3):                 # the third round runs with captured graphs when they are on
        m.zero_grad()
        o1 = m(x1.cuda())
        o2 = m(x2.cuda())              # same shape: must not overwrite what o1's backward needs
        with torch.no_grad():
            m(x1.cuda())               # train-mode forward without grad in between
        dropped = m(x2.cuda())         # and a grad-enabled forward whose output is simply dropped
        (ce(o1, y1.cuda()) + ce(o2, y2.cuda())).backward()
    del dropped, o1, o2                # the last lease goes away with the last reference to that forward's autograd node
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    (ce(O.vit_forward(ref_sd, x1, **KW), y1) + ce(O.vit_forward(ref_sd, x2, **KW), y2)).backward()
    worst = max(((rel_l2(p.grad, ref_sd[n].grad), n) for n, p in m.named_parameters()))
    assert worst[0] < 3e-2, worst
    eng = m._get_engine()
    assert len(eng._ws[(4, True)]) <= 4 and not any(w.get("leased") for w in eng._ws[(4, True)])


@pytest.mark.gpu
def test_detr_two_forwards_before_backward():
    from vitb200.detr import TransformerEncoder, TransformerEncoderLayer
    sd = O.seeded_state_dict(O.detr_param_shapes(256, 512, 2, False), 95)
    enc = TransformerEncoder(TransformerEncoderLayer(256, 4, 512, 0.0, "relu", False), 2)
    enc.load_state_dict(sd)
    enc = enc.cuda().train()
    g = torch.Generator().manual_seed(96)
    a, b = torch.randn(50, 2, 256, generator=g), torch.randn(50, 2, 256, generator=g)
    for _ in range(3):
        enc.zero_grad()
        oa, ob = enc(a.cuda()), enc(b.cuda())
        (oa.float().square().mean() + ob.float().square().mean()).backward()
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    (O.detr_encoder_forward(ref_sd, a, nhead=4, num_layers=2).square().mean()
     + O.detr_encoder_forward(ref_sd, b, nhead=4, num_layers=2).square().mean()).backward()
    worst = max(((rel_l2(p.grad, ref_sd[n].grad), n) for n, p in enc.named_parameters()))
    assert worst[0] < 6e-2, worst      # post-norm DETR layers: the encoder tests' calibrated bound is ~5e-2 for the weights


@pytest.mark.gpu
def test_frozen_parameters_get_no_grad_and_are_not_updated():
    """Fine-tuning the head on a frozen encoder: requires_grad=False parameters keep p.grad = None (as under autograd), so
    optim.Adam(model.parameters()) (vanilla_vit.py:221) does not move them; the trainable ones still match the oracle."""
    m, sd = _model()
    for n, p in m.named_parameters():
        if not n.startswith("heads."):
            p.requires_grad_(False)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    x, y = O.seeded_images(6, 32, 97), O.seeded_labels(6, 10, 98)
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    for _ in range(2):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(m(x.cuda()), y.cuda()).backward()
        assert all((p.grad is None) == (not p.requires_grad) for p in m.parameters())
        if _ == 0:
            ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            torch.nn.functional.cross_entropy(O.vit_forward(ref_sd, x, **KW), y).backward()
            for n in ("heads.head.weight", "heads.head.bias"):
                assert rel_l2(dict(m.named_parameters())[n].grad, ref_sd[n].grad) < 3e-2
        opt.step()
    for n, p in m.named_parameters():
        assert torch.equal(p, before[n]) == (not n.startswith("heads.")), n


@pytest.mark.gpu
def test_opt_in_split_k_inference_matches_default(monkeypatch):
    """VITB200_INFER_SPLIT_K=1: residual GEMMs of small-batch eval forwards are split along K (reduce-add into residual + bias).  Same
    numbers up to the fp32 summation order (and the bf16 re-quantisation that follows it)."""
    import vitb200.engine as E
    m, sd = _model(train=False)
    x = O.seeded_images(3, 32, 99)
    monkeypatch.setenv("VITB200_INFER_GRAPH", "0")
    with torch.no_grad():
        a = m(x.cuda())
        fa = m.forward_features(x.cuda())
        monkeypatch.setattr(E, "_INFER_SPLIT_K", True)
        b = m(x.cuda())
        fb = m.forward_features(x.cuda())
    # a different fp32 summation order moves a few LayerNorm outputs across a bf16 rounding boundary: bf16-level, not 1e-6-level, agreement
    assert rel_l2(b, a) < 1e-2 and rel_l2(fb, fa) < 1e-2
    assert rel_l2(b, O.vit_forward(sd, x, **KW)) < 1.5e-2


@pytest.mark.gpu
@pytest.mark.parametrize("frozen", [False, True])
def test_input_image_gradients(frozen):
    """d loss / d images (what autograd gives the reference's conv_proj when the batch requires grad: saliency maps, adversarial
    examples), also with every parameter frozen."""
    m, sd = _model()
    if frozen:
        for p in m.parameters():
            p.requires_grad_(False)
    x, y = O.seeded_images(5, 32, 101), O.seeded_labels(5, 10, 102)
    xc = x.cuda().requires_grad_(True)
    for _ in range(3):                      # eager, captured, replayed
        xc.grad = None
        torch.nn.functional.cross_entropy(m(xc), y.cuda()).backward()
    xr = x.clone().requires_grad_(True)
    ref_sd = {k: v.clone().requires_grad_(not frozen) for k, v in sd.items()}
    torch.nn.functional.cross_entropy(O.vit_forward(ref_sd, xr, **KW), y).backward()
    assert xc.grad is not None and xc.grad.shape == x.shape
    assert rel_l2(xc.grad, xr.grad) < 3e-2, rel_l2(xc.grad, xr.grad)
    if frozen:
        assert all(p.grad is None for p in m.parameters())


@pytest.mark.gpu
def test_out_of_range_label_gives_nan_loss_not_an_illegal_read():
    from vitb200 import ops
    z = torch.randn(4, 10, device="cuda")
    y = torch.tensor([1, 2, 10, 3], device="cuda")            # 10 is out of range for 10 classes
    loss = torch.zeros(1, device="cuda")
    dz = torch.empty(4, 10, device="cuda")
    ops.cross_entropy(z, y, loss, weight=0.25, dlogits_f32=dz)
    torch.cuda.synchronize()
    assert torch.isnan(loss).all()
    loss.zero_()
    ops.distill_loss(z, z.clone(), z.clone(), y, loss, kind="hard", alpha=0.5, tau=1.0, dlogits_f32=dz, dlogits_kd_f32=dz.clone())
    torch.cuda.synchronize()
    assert torch.isnan(loss).all()
    y[2] = 9
    loss.zero_()
    ops.cross_entropy(z, y, loss, weight=0.25, dlogits_f32=dz)
    assert abs(loss.item() - torch.nn.functional.cross_entropy(z, y).item()) < 1e-5
