"""F2: Swin patch embedding consuming pillars. CPU: the algebra (LayerNorm pushed through the strided
convolution) against the dense upstream op sequence; GPU: csrc/patch_embed.cu against the dense oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from helpers import assert_close

CASES = [  # C, E, ny, nx, patch, batch, patch_norm
    (128, 192, 48, 64, 4, 3, True),
    (64, 96, 50, 44, 4, 2, True),    # 50 x 44: corner padding on both axes
    (32, 64, 36, 30, 6, 2, False),   # patch 6 (semantic_kitti/04:23): 36 cells per patch, no patch norm
    (128, 32, 24, 32, 4, 1, True),
]


def _problem(C, E, ny, nx, ps, B, patch_norm, seed=0, empty_frame=True):
    rng = np.random.default_rng(seed)
    coors, feats = [], []
    for b in range(B):
        n = 0 if (empty_frame and b == 1) else int(0.12 * ny * nx)
        cells = rng.choice(ny * nx, size=n, replace=False)
        c = np.zeros((n, 4), np.int32)
        c[:, 0], c[:, 2], c[:, 3] = b, cells // nx, cells % nx
        coors.append(c)
        feats.append(np.maximum(rng.standard_normal((n, C)), 0).astype(np.float32))
    coors, feats = np.concatenate(coors), np.concatenate(feats)
    p = dict(coors=coors, feats=feats,
             lw=(1 + 0.3 * rng.standard_normal((C, ny, nx))).astype(np.float32),
             lb=(0.3 * rng.standard_normal((C, ny, nx))).astype(np.float32),
             W=(rng.standard_normal((E, C, ps, ps)) / np.sqrt(C * ps * ps)).astype(np.float32),
             bias=(0.1 * rng.standard_normal(E)).astype(np.float32),
             nw=(1 + 0.2 * rng.standard_normal(E)).astype(np.float32) if patch_norm else None,
             nb=(0.2 * rng.standard_normal(E)).astype(np.float32) if patch_norm else None)
    return p


def _dense(p, B, ny, nx, ps, dtype):
    t = lambda a: None if a is None else torch.from_numpy(a).to(dtype)  # noqa: E731
    canvas = torch.from_numpy(O.scatter_np(p["feats"], p["coors"], B, ny, nx)).to(dtype)
    x = torch.nn.functional.layer_norm(canvas, canvas.shape[1:], t(p["lw"]), t(p["lb"]), 1e-3)
    y, size = O.patch_embed_dense(x, t(p["W"]), t(p["bias"]), ps, t(p["nw"]), t(p["nb"]), 1e-5)
    return y.numpy(), size


@pytest.mark.parametrize("C,E,ny,nx,ps,B,pn", CASES)
def test_layernorm_pushed_through_the_convolution_equals_the_dense_sequence(C, E, ny, nx, ps, B, pn):
    p = _problem(C, E, ny, nx, ps, B, pn)
    ref, size = _dense(p, B, ny, nx, ps, torch.float64)
    got, size2 = O.patch_embed_on_pillars_np(p["feats"], p["coors"], B, ny, nx, p["lw"], p["lb"], 1e-3, p["W"], p["bias"], ps,
                                             p["nw"], p["nb"], 1e-5)
    assert size == size2 == (-(-ny // ps), -(-nx // ps))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9 * np.abs(ref).max())


def test_module_mirrors_patch_embed_state_dict_and_refuses_cpu():
    import mask_bev_b200 as M
    pe = M.PillarPatchEmbed(in_channels=128, embed_dims=192, conv_type='Conv2d', kernel_size=4, stride=4,
                            norm_cfg=dict(type='LN'), init_cfg=None)
    assert sorted(pe.state_dict()) == ['norm.bias', 'norm.weight', 'projection.bias', 'projection.weight']
    assert tuple(pe.projection.weight.shape) == (192, 128, 4, 4) and pe.out_size(50, 44) == (13, 11)
    with pytest.raises(M.MbevError):
        pe(torch.zeros(1, 128, 8, 8))
    with pytest.raises(M.MbevError):
        pe.forward_pillars(torch.zeros(4, 128), None, None, None, 1, 8, 8, torch.nn.LayerNorm([128, 8, 8]))
    with pytest.raises(M.MbevError):
        M.PillarPatchEmbed(in_channels=128, embed_dims=192, kernel_size=4, stride=2)


@pytest.mark.gpu
@pytest.mark.parametrize("C,E,ny,nx,ps,B,pn", CASES)
def test_patch_embed_on_pillars_vs_dense_oracle(C, E, ny, nx, ps, B, pn):
    import mask_bev_b200 as M
    from mask_bev_b200 import functional as F_
    dev = torch.device("cuda:0")
    p = _problem(C, E, ny, nx, ps, B, pn)
    ref32, size = _dense(p, B, ny, nx, ps, torch.float32)
    ref64, _ = _dense(p, B, ny, nx, ps, torch.float64)
    pe = M.PillarPatchEmbed(in_channels=C, embed_dims=E, kernel_size=ps, stride=ps, norm_cfg=dict(type='LN') if pn else None).to(dev)
    ln = torch.nn.LayerNorm([C, ny, nx], eps=1e-3).to(dev)
    with torch.no_grad():
        pe.projection.weight.copy_(torch.from_numpy(p["W"]))
        pe.projection.bias.copy_(torch.from_numpy(p["bias"]))
        if pn:
            pe.norm.weight.copy_(torch.from_numpy(p["nw"]))
            pe.norm.bias.copy_(torch.from_numpy(p["nb"]))
        ln.weight.copy_(torch.from_numpy(p["lw"]))
        ln.bias.copy_(torch.from_numpy(p["lb"]))
    coors = torch.from_numpy(p["coors"]).to(dev)
    feats = torch.from_numpy(p["feats"]).to(dev)
    P = coors.shape[0]
    counts = torch.bincount(coors[:, 0].long(), minlength=B)
    base = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    base[1:] = torch.cumsum(counts, 0).int()
    table = F_.build_cell_table(coors, base[B:], P, B, ny, nx)
    with torch.no_grad():
        tok, size2 = pe.forward_pillars(feats, coors, table, base, B, ny, nx, ln)
        tok2, _ = pe.forward_pillars(feats, coors, table, base, B, ny, nx, ln)
    assert size2 == size and tuple(tok.shape) == ref32.shape
    assert torch.equal(tok, tok2)  # run-to-run identical
    from helpers import assert_close_arbitrated
    assert_close_arbitrated(tok.cpu().numpy(), ref32, ref64, 1e-5, f"patch embed C={C} E={E} {ny}x{nx} ps={ps}")


@pytest.mark.gpu
def test_encoder_forward_patch_tokens_equals_dense_chain():
    """K1 -> K2 -> f2 against the product's own canvas route: forward() (fused K3 + LayerNorm) followed by the dense
    torch convolution + LayerNorm, on LiDAR-shaped frames."""
    import mask_bev_b200 as M
    from mask_bev_b200.synthetic import encoder_kwargs, gen_batch
    dev = torch.device("cuda:0")
    kw = encoder_kwargs("semkitti_b1")
    enc = M.MaskBevEncoder(**kw).to(dev).eval()
    torch.manual_seed(1)
    with torch.no_grad():
        enc._layer_norm.weight.normal_(1.0, 0.3)
        enc._layer_norm.bias.normal_(0.0, 0.3)
    pe = M.PillarPatchEmbed(in_channels=kw["feat_channels"][-1], embed_dims=192, kernel_size=4, stride=4,
                            norm_cfg=dict(type='LN')).to(dev)
    frames = [torch.from_numpy(f).to(dev) for f in gen_batch("semkitti_b1", batch=2, n=30000)]
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        tok, size = enc.forward_patch_tokens(frames, pe)
        x = enc(frames)
        ref = pe.norm(pe.projection(x).flatten(2).transpose(1, 2))
    assert size == (x.shape[2] // 4, x.shape[3] // 4)
    assert_close(tok.cpu().numpy(), ref.cpu().numpy(), 2e-5, "forward_patch_tokens vs canvas + conv")
