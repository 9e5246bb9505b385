"""GPU parity of the fused scatter + LayerNorm BACKWARD (mbev_scatter_layernorm_backward) against
torch autograd through `nn.LayerNorm([C, ny, nx], eps=1e-3)` applied to the scatter output (mask_bev_encoders.py:75,
91-92; the reference has no backward code of its own, autograd derives it). The arbiter is float64 autograd on a dense
canvas built from the same rows; tolerance 1e-5 of max|ref| (fp32). Through the C ABI."""
import pytest
import torch

from helpers import assert_close, encoder_pair, ref_test_kwargs, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _random_problem(B, C, ny, nx, counts, extra_rows, seed):
    """Rows of B frames on random distinct cells; `extra_rows` capacity rows beyond the pillar count."""
    g = torch.Generator().manual_seed(seed)
    G = ny * nx
    P = sum(counts)
    coors = torch.zeros((P + extra_rows, 4), dtype=torch.int32)
    base = [0]
    for b, n in enumerate(counts):
        cells = torch.randperm(G, generator=g)[:n]
        coors[base[-1]:base[-1] + n, 0] = b
        coors[base[-1]:base[-1] + n, 2] = (cells // nx).to(torch.int32)
        coors[base[-1]:base[-1] + n, 3] = (cells % nx).to(torch.int32)
        base.append(base[-1] + n)
    feats = torch.randn((P + extra_rows, C), generator=g).relu_() * 2.0   # post-ReLU-like rows (exact zeros inside)
    weight = 1.0 + 0.3 * torch.randn((C, ny, nx), generator=g)
    bias = 0.3 * torch.randn((C, ny, nx), generator=g)
    dout = torch.randn((B, C, ny, nx), generator=g)
    return coors, torch.tensor(base, dtype=torch.int32), feats, weight, bias, dout, P


def _float64_reference(coors, feats, weight, bias, dout, B, C, ny, nx, P, eps):
    f = feats[:P].double().to(DEV).requires_grad_(True)
    w = weight.double().to(DEV).requires_grad_(True)
    bi = bias.double().to(DEV).requires_grad_(True)
    c = coors[:P].long().to(DEV)
    canvas = torch.zeros((B, C, ny * nx), dtype=torch.float64, device=DEV)
    lin = c[:, 2] * nx + c[:, 3]
    canvas[c[:, 0], :, lin] = f   # canvas[b, :, cell] = feats[p]
    out = torch.nn.functional.layer_norm(canvas.view(B, C, ny, nx), (C, ny, nx), w, bi, eps)
    out.backward(dout.double().to(DEV))
    return out.detach(), f.grad, w.grad, bi.grad


@pytest.mark.parametrize("B,C,ny,nx,counts,extra", [
    (3, 64, 40, 52, (300, 0, 517), 5),        # G = 2080: last 128-cell run is ragged; an empty frame; spare capacity
    (2, 128, 100, 100, (1500, 2300), 0),
    (1, 4, 8, 8, (20,), 3),                   # one channel chunk, G < one run
    (5, 12, 36, 36, (100, 200, 1, 1296, 40), 0),  # 3 chunks per run (CTA tasks straddle runs); a FULL frame
    (16, 32, 64, 64, tuple(200 + 10 * i for i in range(16)), 0),
    (70, 8, 16, 16, tuple(3 + (i % 5) for i in range(70)), 0),   # batch > 64: the 4-stage ring
    (2, 16, 20, 20, (30, 50), 0),             # fewer frames than ring stages
])
def test_scatter_layernorm_backward_matches_float64_autograd(B, C, ny, nx, counts, extra):
    from mask_bev_b200 import functional as F_
    eps = 1e-3
    coors, base, feats, weight, bias, dout, P = _random_problem(B, C, ny, nx, counts, extra, seed=B * 1000 + C)
    rows = P + extra
    coors_d, base_d, feats_d = coors.to(DEV), base.to(DEV), feats.to(DEV)
    weight_d, bias_d, dout_d = weight.to(DEV), bias.to(DEV), dout.to(DEV)
    assert F_.scatter_layernorm_backward_supported(B, C, ny, nx)
    table = F_.build_cell_table(coors_d, base_d[B:], rows, B, ny, nx)
    res = F_.scatter_layernorm_forward(feats_d, table, base_d, B, ny, nx, weight_d, bias_d, eps)
    assert res is not None
    out, stats = res
    dfeats, dweight, dbias = F_.scatter_layernorm_backward(dout_d, feats_d, table, coors_d, base_d[B:], weight_d, stats)
    ref_out, ref_df, ref_dw, ref_db = _float64_reference(coors, feats, weight, bias, dout, B, C, ny, nx, P, eps)
    assert_close(out.cpu().numpy(), ref_out.cpu().numpy(), what="forward")
    assert_close(dbias.cpu().numpy(), ref_db.cpu().numpy(), what="dbias")
    assert_close(dweight.cpu().numpy(), ref_dw.cpu().numpy(), what="dweight")
    assert_close(dfeats[:P].cpu().numpy(), ref_df.cpu().numpy(), what="dfeats")
    if extra:
        assert torch.count_nonzero(dfeats[P:]).item() == 0   # rows beyond the pillar count get zeros
    # fixed-order reductions: a second run is bit-identical
    again = F_.scatter_layernorm_backward(dout_d, feats_d, table, coors_d, base_d[B:], weight_d, stats)
    assert all(torch.equal(a, b) for a, b in zip((dfeats, dweight, dbias), again))


def test_rows_missing_from_the_cell_table_get_zero_gradient():
    """A hand-made coors list may name a cell twice (PointPillarsScatter keeps one of the rows): the row that did not
    reach the canvas has no gradient."""
    from mask_bev_b200 import functional as F_
    B, C, ny, nx = 1, 8, 8, 16
    coors, base, feats, weight, bias, dout, P = _random_problem(B, C, ny, nx, (10,), 0, seed=3)
    coors[9] = coors[0]
    coors_d, base_d, feats_d = coors.to(DEV), base.to(DEV), feats.to(DEV)
    table = F_.build_cell_table(coors_d, base_d[B:], P, B, ny, nx)
    winner = int(table[0, int(coors[0, 2]) * nx + int(coors[0, 3])].item())
    loser = 9 if winner == 0 else 0
    _, stats = F_.scatter_layernorm_forward(feats_d, table, base_d, B, ny, nx, weight.to(DEV), bias.to(DEV), 1e-3)
    dfeats, _, _ = F_.scatter_layernorm_backward(dout.to(DEV), feats_d, table, coors_d, base_d[B:], weight.to(DEV), stats)
    assert torch.count_nonzero(dfeats[loser]).item() == 0
    assert torch.count_nonzero(dfeats[winner]).item() > 0


def test_unsupported_shapes_are_reported():
    from mask_bev_b200 import functional as F_
    assert not F_.scatter_layernorm_backward_supported(2, 6, 16, 16)     # C % 4
    assert not F_.scatter_layernorm_backward_supported(2, 64, 25, 25)    # ny*nx % 4
    assert F_.scatter_layernorm_backward_supported(16, 128, 800, 800)


@pytest.mark.parametrize("train_bn", [False, True])
def test_encoder_training_step_fused_layernorm_matches_unfused(train_bn):
    """MaskBevEncoder.forward under autograd: K1 -> K2 -> (K3+LN) with the fused backward, against K3 followed by
    torch's LayerNorm (its fp32 CUDA statistics over millions of elements are themselves ~1e-5..1e-4 from float64, hence the
    looser bound here; the float64 test above is the tight one)."""
    from mask_bev_b200.synthetic import gen_frame
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32, vs=0.32)   # 250 x 250
    enc, _ = encoder_pair(kw, seed=8)
    g = torch.Generator().manual_seed(21)
    with torch.no_grad():
        enc._layer_norm.weight.copy_(1.0 + 0.3 * torch.randn(enc._layer_norm.weight.shape, generator=g))
        enc._layer_norm.bias.copy_(0.3 * torch.randn(enc._layer_norm.bias.shape, generator=g))
    enc = enc.to(DEV)
    enc.train(train_bn)
    pcs = [torch.from_numpy(gen_frame(30000, 4, s)).to(DEV) for s in (1, 2, 3)]
    r = torch.randn((3, 128, 250, 250), generator=g).to(DEV)
    state = {k: v.clone() for k, v in enc.state_dict().items()}

    def step(fused):
        enc.load_state_dict(state)   # train-mode BN updates its running statistics
        enc.fuse_layer_norm_autograd = fused
        enc.zero_grad(set_to_none=True)
        out = enc(pcs)
        assert out.requires_grad
        (out * r).sum().backward()
        return out.detach(), {n: p.grad.detach().clone() for n, p in enc.named_parameters() if p.grad is not None}

    out_f, gr_f = step(True)
    out_u, gr_u = step(False)
    assert set(gr_f) == set(gr_u) and "_layer_norm.weight" in gr_f and "_voxel_encoder.pfn_layers.0.linear.weight" in gr_f
    assert rel_err(out_f.cpu().numpy(), out_u.cpu().numpy()) <= 2e-4
    for n in gr_f:
        e = rel_err(gr_f[n].cpu().numpy(), gr_u[n].cpu().numpy())
        # PFN gradients see the two LayerNorm backwards through K2' (train-mode BN gradients are ill-conditioned)
        tol = 2e-4 if n.startswith("_layer_norm") else (2e-2 if train_bn else 1e-3)
        assert e <= tol, f"{n}: {e:.3e}"
