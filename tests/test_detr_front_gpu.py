"""-m gpu parity tests for what sits between the DETR backbone and the encoder (SURVEY.md §8 f3 remainder): NestedTensor batching
(utils/coco/util/misc.py:307-332), the learned 2-D position embedding (detr.py:33-63), input_proj as a GEMM on the NCHW feature
map (detr.py:125) and Transformer.forward's flatten / permute plumbing (transformer.py:47-63), against the oracle and the
reference-generated fixture tests/golden/detr_front_d256.pt."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import O, rel_l2, strict_fp32
from test_oracle import _detr_front_oracle, load


def _tol(floor):
    return max(1e-2, 1.25 * floor)


@pytest.mark.gpu
def test_detr_front_end_matches_oracle_and_reference_fixture():
    from vitb200.detr import TransformerEncoder, TransformerEncoderLayer
    from vitb200.detr_front import AbsolutePositionalEncoding, InputProjection, mask_at, nested_tensor_from_tensor_list
    gd = load("detr_front_d256.pt")
    t = _detr_front_oracle(gd)
    D, c_in = gd["d_model"], gd["c_in"]
    # ---- oracle (fp32 truth + its own autocast-bf16 error as the tolerance floor) ----
    n, _, h, w = t["feats"].shape
    res = {}
    for ac in (False, True):
        leaves = {k: t[k].detach().clone().requires_grad_(True) for k in ("feats", "row_w", "col_w", "pw", "pb")}
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in t["sd"].items()}
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=ac):
            pos = O.detr_abs_pos_encoding(leaves["row_w"], leaves["col_w"], n, h, w)
            src = O.detr_input_proj(leaves["feats"], leaves["pw"], leaves["pb"])
            s2, p2, m2 = O.detr_flatten(src, pos, t["mask"])
            out = O.detr_encoder_forward(sd, s2.float(), nhead=gd["nhead"], num_layers=gd["layers"], src_key_padding_mask=m2, pos=p2.float())
        gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
        out.float().backward(gout)
        res[ac] = (out.detach().float(), {**{k: v.grad for k, v in leaves.items()}, **{"enc." + k: v.grad for k, v in sd.items()}})
    ref_out, ref_g = res[False]
    floor_o = rel_l2(res[True][0], ref_out)
    floor_g = max(rel_l2(res[True][1][k], ref_g[k]) for k in ref_g)
    # ---- the CUDA path through the module API ----
    nt = nested_tensor_from_tensor_list([i.cuda() for i in t["imgs"]])
    assert torch.equal(nt.tensors.cpu(), t["padded"]) and torch.equal(nt.mask.cpu(), t["mask_full"])
    feats = t["feats"].detach().cuda().requires_grad_(True)
    mask = mask_at(nt.mask, feats.shape[-2:])
    assert torch.equal(mask.cpu(), gd["mask_feat"])
    posm = AbsolutePositionalEncoding(D // 2)
    proj = InputProjection(c_in, D, kernel_size=1)
    with torch.no_grad():
        posm.row_embed.weight.copy_(t["row_w"]); posm.col_embed.weight.copy_(t["col_w"])
        proj.weight.copy_(t["pw"]); proj.bias.copy_(t["pb"])
    posm, proj = posm.cuda(), proj.cuda()
    enc = TransformerEncoder(TransformerEncoderLayer(D, gd["nhead"], gd["ffn"], 0.0, "relu", False), gd["layers"], None)
    enc.load_state_dict({k: v.detach() for k, v in t["sd"].items()})
    enc = enc.cuda().train()
    pos = posm(nt.__class__(feats, mask))
    assert pos.shape == (n, D, h, w)
    assert abs(pos.sum().item() - gd["pos_sum"]) < 1e-2 and torch.allclose(pos[0, :, 0, 0].cpu(), gd["pos_00"]) \
        and torch.allclose(pos[1, :, -1, -1].cpu(), gd["pos_last"])                      # the reference's own values: exact gathers
    src = proj(feats)
    assert src.shape == (n, D, h, w)
    s2, p2, m2 = src.flatten(2).permute(2, 0, 1), pos.flatten(2).permute(2, 0, 1), mask.flatten(1)     # transformer.py:49-53
    assert s2.is_contiguous() and p2.is_contiguous(), "the NCHW results must be views of the sequence-first buffers"
    out = enc(s2, src_key_padding_mask=m2, pos=p2)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
    out.backward(gout.cuda())
    assert rel_l2(out, ref_out) < _tol(floor_o), (rel_l2(out, ref_out), floor_o)
    got = {"feats": feats.grad, "row_w": posm.row_embed.weight.grad, "col_w": posm.col_embed.weight.grad, "pw": proj.weight.grad,
           "pb": proj.bias.grad, **{"enc." + k: p.grad for k, p in enc.named_parameters()}}
    errs = {k: rel_l2(got[k], ref_g[k]) for k in ref_g}
    worst = max(errs, key=errs.get)
    assert errs[worst] < _tol(floor_g), (worst, errs[worst], floor_g)
    # against the unmodified reference's recorded forward results (fp32 on CPU): same bound (its gradients belong to another gout;
    # the oracle above is pinned on them by tests/test_oracle.py::test_oracle_detr_front_end_matches_reference_golden)
    assert abs(out.norm().item() - gd["out_norm"]) < _tol(floor_o) * gd["out_norm"]
    assert rel_l2(out[0], gd["out_row0"]) < _tol(floor_o) and rel_l2(src[0, 0], gd["src_n0_c0"]) < 1e-2


@pytest.mark.gpu
def test_input_projection_and_pos_embedding_at_cfg5_size():
    """ResNet-50 layer4 -> hidden: C_in = 2048, hidden = 512, 25 x 42 feature map (H*W = 1050 is not a multiple of 8: padded operand
    pitch), N = 2; forward and the three gradients of the 1x1 convolution, and the embedding-table gradients."""
    from vitb200.detr_front import AbsolutePositionalEncoding, InputProjection
    N, Cin, D, h, w = 2, 2048, 512, 25, 42
    g = torch.Generator().manual_seed(77)
    x = torch.randn(N, Cin, h, w, generator=g)
    W = torch.randn(D, Cin, 1, 1, generator=g) / math.sqrt(Cin)
    b = torch.randn(D, generator=g) * 0.1
    gout = torch.randn(N, D, h, w, generator=g)
    xr, Wr, br = (v.cuda().requires_grad_(True) for v in (x, W, b))
    with strict_fp32():
        ref = F.conv2d(xr, Wr, br)
        ref.backward(gout.cuda())
    proj = InputProjection(Cin, D, kernel_size=1)
    with torch.no_grad():
        proj.weight.copy_(W); proj.bias.copy_(b)
    proj = proj.cuda()
    xc = x.cuda().requires_grad_(True)
    out = proj(xc)
    out.backward(gout.cuda())
    assert rel_l2(out, ref) < 1e-2
    assert rel_l2(xc.grad, xr.grad) < 1e-2 and rel_l2(proj.weight.grad, Wr.grad) < 1e-2 and rel_l2(proj.bias.grad, br.grad) < 1e-2
    posm = AbsolutePositionalEncoding(D // 2).cuda()
    pos = posm(torch.empty(N, D, h, w, device="cuda"))
    refp = O.detr_abs_pos_encoding(posm.row_embed.weight.detach().cpu().requires_grad_(True), posm.col_embed.weight.detach().cpu().requires_grad_(True), N, h, w)
    assert torch.equal(pos.cpu(), refp.detach())
    gp = torch.randn(N, D, h, w, generator=g)
    pos.backward(gp.cuda())
    rw, cw = posm.row_embed.weight.detach().cpu().requires_grad_(True), posm.col_embed.weight.detach().cpu().requires_grad_(True)
    O.detr_abs_pos_encoding(rw, cw, N, h, w).backward(gp)
    assert rel_l2(posm.row_embed.weight.grad, rw.grad) < 1e-5 and rel_l2(posm.col_embed.weight.grad, cw.grad) < 1e-5


@pytest.mark.gpu
def test_transformer_module_end_to_end():
    """vitb200.detr_front.Transformer (transformer.py:25-63 with the forward's typos resolved): NCHW src / pos, mask, query embedding in,
    decoder states [L, N, Q, D]-transposed and the memory back in NCHW out; gradients reach every parameter."""
    from vitb200.detr_front import Transformer
    torch.manual_seed(5)
    tr = Transformer(d_model=256, nhead=4, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=512, dropout=0.0,
                     return_intermediate_dec=True).cuda().train()
    N, h, w, Q = 2, 9, 12, 20
    src = torch.randn(N, 256, h, w, device="cuda", requires_grad=True)
    pos = torch.randn(N, 256, h, w, device="cuda")
    mask = torch.zeros(N, h, w, dtype=torch.bool, device="cuda")
    mask[1, :, 9:] = True
    qe = torch.randn(Q, 256, device="cuda", requires_grad=True)
    hs, mem = tr(src, mask, qe, pos)
    assert hs.shape == (2, N, Q, 256) and mem.shape == (N, 256, h, w)
    (hs.float().square().mean() + mem.float().square().mean()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in tr.parameters())
    assert src.grad is not None and qe.grad is not None
    # oracle: encoder + decoder restatement on the same weights
    esd = {k[len("encoder."):]: v.detach().cpu() for k, v in tr.state_dict().items() if k.startswith("encoder.")}
    dsd = {k[len("decoder."):]: v.detach().cpu() for k, v in tr.state_dict().items() if k.startswith("decoder.")}
    s2, p2, m2 = O.detr_flatten(src.detach().cpu(), pos.cpu(), mask.cpu())
    memr = O.detr_encoder_forward(esd, s2, nhead=4, num_layers=2, src_key_padding_mask=m2, pos=p2)
    q2 = qe.detach().cpu().unsqueeze(1).repeat(1, N, 1)
    hsr = O.detr_decoder_forward(dsd, torch.zeros_like(q2), memr, nhead=4, num_layers=2, memory_key_padding_mask=m2, pos=p2, query_pos=q2,
                                 return_intermediate=True)
    assert rel_l2(mem, memr.permute(1, 2, 0).view(N, 256, h, w)) < 2e-2
    assert rel_l2(hs, hsr.transpose(1, 2)) < 3e-2
