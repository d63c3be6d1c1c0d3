"""-m gpu: DETR transformer decoder against the CPU oracle (pinned to the reference by tests/golden/detr_dec*.pt).
Tolerances are self-calibrated like the encoder's (tests/test_deit_detr_gpu.py): the bf16 tensor-core path must be no worse than
1.25 x max(1e-2, the oracle's own autocast-bf16 error against the same fp32 truth); outputs within 1.5e-2 relative L2."""
import pytest
import torch

from helpers import O, rel_l2


def _masks(eng, ws, Q, S, N, p):
    from vitb200 import ops
    seed, dev, D, Fd, H = ws["drop_seed"], eng.flat.device, eng.D, eng.F, eng.H
    m = {}
    for li in range(eng.L):
        for site, n, shape in ((0, Q * N * D, (Q, N, D)), (1, Q * N * Fd, (Q, N, Fd)), (2, Q * N * D, (Q, N, D)), (5, Q * N * D, (Q, N, D)),
                               (3, N * H * Q * Q, (N, H, Q, Q)), (4, N * H * Q * S, (N, H, Q, S))):
            m[(li, site)] = ops.dropout_mask(n, p, seed, eng.drop_site(li, site), dev).view(*shape).cpu().float()
    return m


def _run(Q, S, N, d_model, nhead, ffn, layers, *, inter, masked=True, with_pos=True, p=0.0, act="relu", norm=True, pre=False):
    from vitb200.detr import TransformerDecoder, TransformerDecoderLayer
    sd = O.seeded_state_dict(O.detr_decoder_param_shapes(d_model, ffn, layers, with_norm=norm), 61)
    dec = TransformerDecoder(TransformerDecoderLayer(d_model, nhead, ffn, p, act, pre), layers, torch.nn.LayerNorm(d_model) if norm else None,
                             return_intermediate=inter)
    dec.load_state_dict(sd)
    dec = dec.cuda().train()
    g = torch.Generator().manual_seed(62)
    tgt = torch.randn(Q, N, d_model, generator=g) * 0.5
    mem = torch.randn(S, N, d_model, generator=g)
    pos = torch.randn(S, N, d_model, generator=g) if with_pos else None
    qpos = torch.randn(Q, N, d_model, generator=g) if with_pos else None
    kpm = None
    if masked:
        valid = torch.randint(S // 2, S + 1, (N,), generator=g)
        kpm = torch.arange(S)[None, :] >= valid[:, None]
    leaf = lambda t, dev: None if t is None else t.clone().to(dev).requires_grad_(True)
    c = [leaf(t, "cuda") for t in (tgt, mem, pos, qpos)]
    out = dec(c[0], c[1], memory_key_padding_mask=kpm.cuda() if masked else None, pos=c[2], query_pos=c[3])
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout.cuda())
    drop = None
    if p > 0:
        eng = dec._get_engine()
        masks = _masks(eng, eng.workspace(Q, S, N, True), Q, S, N, p)
        for k, v in masks.items():
            assert abs(v.mean().item() - (1 - p)) < 0.03, (k, v.mean().item())
        drop = O.ExplicitDropout(masks, p, p)
    kw = dict(nhead=nhead, num_layers=layers, activation=act, memory_key_padding_mask=kpm, return_intermediate=inter, drop=drop,
              normalize_before=pre)

    def oracle(autocast):
        osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        r = [leaf(t, "cpu") for t in (tgt, mem, pos, qpos)]
        if autocast:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                o = O.detr_decoder_forward(osd, r[0], r[1], pos=r[2], query_pos=r[3], **kw)
        else:
            o = O.detr_decoder_forward(osd, r[0], r[1], pos=r[2], query_pos=r[3], **kw)
        o.float().backward(gout)
        return o.detach().float(), osd, r
    ref, ref_sd, r = oracle(False)
    _, ac_sd, a = oracle(True)
    floor_in = max(1e-2, max(rel_l2(x.grad, y.grad) for x, y in zip(a, r) if x is not None))
    floor_w = max(1e-2, max(rel_l2(ac_sd[k].grad, ref_sd[k].grad) for k in sd))
    assert out.shape == ref.shape and rel_l2(out, ref) < 1.5e-2, rel_l2(out, ref)
    for name, x, y in zip(("tgt", "memory", "pos", "query_pos"), c, r):
        if x is not None:
            assert rel_l2(x.grad, y.grad) < 1.25 * floor_in, (name, rel_l2(x.grad, y.grad), floor_in)
    worst = max(((rel_l2(pp.grad, ref_sd[n].grad), n) for n, pp in dec.named_parameters()))
    assert worst[0] < 1.25 * floor_w, (worst, floor_w)
    return dec


@pytest.mark.gpu
def test_decoder_intermediate_masked_with_pos():
    """DETR's configuration: 100 object queries against a masked memory, decoder norm, return_intermediate (detr.py:139, transformer.py:36-37)."""
    _run(Q=100, S=300, N=3, d_model=512, nhead=8, ffn=2048, layers=2, inter=True)


@pytest.mark.gpu
def test_decoder_small_variants():
    _run(Q=20, S=70, N=2, d_model=256, nhead=4, ffn=512, layers=3, inter=False)
    _run(Q=20, S=70, N=2, d_model=256, nhead=4, ffn=512, layers=2, inter=False, masked=False, with_pos=False, norm=False)
    _run(Q=33, S=130, N=2, d_model=256, nhead=4, ffn=512, layers=2, inter=True, act="gelu")


@pytest.mark.gpu
@pytest.mark.parametrize("inter,act", [(False, "relu"), (True, "relu"), (False, "gelu")])
def test_decoder_dropout_replayed_masks(inter, act):
    dec = _run(Q=40, S=130, N=2, d_model=256, nhead=4, ffn=512, layers=2, inter=inter, p=0.1, act=act)
    dec.eval()
    with torch.no_grad():
        t, m = torch.randn(40, 2, 256, device="cuda"), torch.randn(130, 2, 256, device="cuda")
        a, b = dec(t, m), dec(t, m)
    assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("inter,norm,p,act", [(True, True, 0.0, "relu"), (False, True, 0.0, "gelu"), (False, False, 0.0, "relu"),
                                              (True, True, 0.1, "relu"), (False, True, 0.1, "relu")])
def test_decoder_pre_norm(inter, norm, p, act):
    """normalize_before=True: TransformerDecoderLayer.forward_pre (transformer.py:158-178)."""
    _run(Q=40, S=130, N=2, d_model=256, nhead=4, ffn=512, layers=3, inter=inter, norm=norm, p=p, act=act, pre=True)


@pytest.mark.gpu
def test_decoder_layer_standalone_and_errors():
    from vitb200.detr import TransformerDecoder, TransformerDecoderLayer
    layer = TransformerDecoderLayer(256, 4, 512, 0.0, "relu", False).cuda()
    t, m = torch.randn(10, 2, 256, device="cuda"), torch.randn(50, 2, 256, device="cuda")
    y = layer(t, m)
    assert y.shape == (10, 2, 256)
    sd = {"layers.0." + k: v.cpu() for k, v in layer.state_dict().items()}
    ref = O.detr_decoder_forward(sd, t.cpu(), m.cpu(), nhead=4, num_layers=1)[0]
    assert rel_l2(y, ref) < 1.5e-2
    with pytest.raises(NotImplementedError):
        TransformerDecoder(layer, 1).cuda()(t, m, tgt_mask=torch.zeros(10, 10, device="cuda"))


@pytest.mark.gpu
def test_detr_autograd_graph_replay_matches_eager(monkeypatch):
    """Encoder and decoder autograd nodes replayed from CUDA graphs (FlatParams.graphed) against the eager kernel sequence: same kernels,
    same buffers; fp32 split-K / atomic sums are order-dependent, hence 1e-4 rather than bit equality."""
    from vitb200.detr import TransformerDecoder, TransformerDecoderLayer, TransformerEncoder, TransformerEncoderLayer

    def run(mode):
        monkeypatch.setenv("VITB200_AUTOGRAD_GRAPH", mode)
        torch.manual_seed(5)
        enc = TransformerEncoder(TransformerEncoderLayer(256, 4, 512, 0.0, "relu", False), 2).cuda().train()
        dec = TransformerDecoder(TransformerDecoderLayer(256, 4, 512, 0.0, "relu", False), 2, torch.nn.LayerNorm(256), True).cuda().train()
        opt = torch.optim.SGD(list(enc.parameters()) + list(dec.parameters()), lr=0.01)
        g = torch.Generator().manual_seed(6)
        qpos = torch.randn(30, 2, 256, generator=g).cuda().requires_grad_(True)
        losses = []
        for _ in range(4):
            src, pos = torch.randn(90, 2, 256, generator=g).cuda(), torch.randn(90, 2, 256, generator=g).cuda()
            kpm = (torch.arange(90)[None, :] >= torch.tensor([[90], [70]])).cuda()
            opt.zero_grad()
            mem = enc(src, src_key_padding_mask=kpm, pos=pos)
            hs = dec(torch.zeros_like(qpos), mem, memory_key_padding_mask=kpm, pos=pos, query_pos=qpos)
            loss = hs.float().square().mean()
            loss.backward()
            opt.step()
            losses.append(loss.item())
        captured = {k[0][0] for e in (enc, dec) for k, v in e._get_engine().__dict__.get("_train_graphs", {}).items() if isinstance(v, tuple)}
        return losses, [p.detach().clone() for p in list(enc.parameters()) + list(dec.parameters())], qpos.grad.clone(), captured
    le, pe, qe, ce = run("0")
    lg, pg, qg, cg = run("1")
    assert ce == set() and cg == {"enc_fwd", "enc_bwd", "dec_fwd", "dec_bwd"}, (ce, cg)
    assert all(abs(a - b) <= 1e-3 * max(1.0, abs(a)) for a, b in zip(le, lg)), (le, lg)
    # zero-initialised biases whose gradient is rounding noise (the key bias: exactly zero in exact arithmetic) need an absolute floor
    assert all((a - b).norm().item() <= 5e-4 * a.norm().item() + 5e-6 * a.numel() ** 0.5 for a, b in zip(pe, pg))
    assert rel_l2(qg, qe) < 2e-2      # a bf16-path gradient after four steps of (order-dependent) fp32 sums: bf16 resolution, not 1e-4
