"""GPU parity of the fused scatter + LayerNorm (mbev_scatter_layernorm_forward) against
nn.LayerNorm([C, ny, nx], eps=1e-3) applied to the scatter output (mask_bev_encoders.py:75, 91-92): the torch op on the
GPU canvas and the CPU oracle with layer_norm=True. Tolerance 1e-5 of max|ref| (fp32). Through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import O, assert_close, encoder_pair, ref_test_kwargs, rel_err  # noqa: F401

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _frames(n, C, seeds):
    from mask_bev_b200.synthetic import gen_frame
    return [gen_frame(n, C, s) for s in seeds]


def _randomise_ln(ln, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        ln.weight.copy_(1.0 + 0.3 * torch.randn(ln.weight.shape, generator=g))
        ln.bias.copy_(0.3 * torch.randn(ln.bias.shape, generator=g))


@pytest.mark.parametrize("chans,C,rng,vs,n,seeds", [
    ((128, 128, 128), 4, (-40, 40), 0.16, 40000, (1, 2, 3)),
    ((128, 128, 128), 4, (-40, 40), 0.1, 50000, (4, 5)),
    ((128, 64, 128), 5, (-75.2, 75.2), 0.32, 50000, (6, 7, 8)),
    ((64,), 3, (-40, 40), 0.16, 30000, (9,)),
])
def test_fused_scatter_layernorm_matches_torch_layernorm(chans, C, rng, vs, n, seeds):
    kw = ref_test_kwargs(feat_channels=chans, T=32, C=C, x_range=rng, y_range=rng, vs=vs)
    enc, _ = encoder_pair(kw, seed=5)
    _randomise_ln(enc._layer_norm, seed=11)
    enc = enc.to(DEV).eval()
    pcs = [torch.from_numpy(f).to(DEV) for f in _frames(n, C, seeds)]
    ln = enc._layer_norm
    with torch.no_grad():
        fused = enc(pcs)                                  # K1, K2, K3+LN
        canvas = enc.encode_batch(pcs)                    # K1, K2, K3
        ref32 = ln(canvas)                                # torch's fp32 LayerNorm (what the reference runs)
        ref64 = torch.nn.functional.layer_norm(canvas.double(), ln.normalized_shape, ln.weight.double(),
                                               ln.bias.double(), ln.eps)  # the arbiter
    assert fused.shape == ref32.shape
    f, r32, r64 = fused.cpu().numpy(), ref32.cpu().numpy(), ref64.cpu().numpy()
    e64 = assert_close(f, r64, what="fused scatter+LN vs float64 LayerNorm")
    # torch's own fp32 CUDA LayerNorm reduces C*ny*nx (up to 82 M) elements per row in float: it sits further from
    # the float64 result than the fused kernel does (fp64 statistics), so parity against it is bounded by ITS error
    e32 = rel_err(r32, r64)
    print(f"fused vs fp64 {e64:.2e}; torch fp32 LN vs fp64 {e32:.2e}; fused vs torch fp32 {rel_err(f, r32):.2e}")
    assert rel_err(f, r32) <= max(1e-5, 2.0 * e32 + 1e-6)
    with torch.enable_grad():                             # the autograd path stays unfused (FMA PFN, K3, torch LN)
        unfused = enc(pcs)
    assert unfused.requires_grad
    assert rel_err(unfused.detach().cpu().numpy(), r64) <= max(1e-5, 2.0 * e32 + 1e-6)


def test_fused_scatter_layernorm_vs_cpu_oracle_and_empty_frame():
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32, vs=0.16)
    enc, orc0 = encoder_pair(kw, seed=6)
    _randomise_ln(enc._layer_norm, seed=12)
    orc = O.MaskBevEncoderOracle(feat_channels=kw["feat_channels"], x_range=kw["x_range"], y_range=kw["y_range"],
                                 z_range=kw["z_range"], voxel_size_x=kw["voxel_size_x"], voxel_size_y=kw["voxel_size_y"],
                                 voxel_size_z=kw["voxel_size_z"], max_num_points=32, pc_point_dim=4, with_distance=True,
                                 layer_norm=True)
    orc.pfn.load_state_dict(orc0.pfn.state_dict())
    orc.pfn.eval()
    orc.layer_norm.load_state_dict(enc._layer_norm.state_dict())
    enc = enc.to(DEV).eval()
    frames = [_frames(20000, 4, (3,))[0], np.full((10, 4), 1000.0, np.float32)]  # second frame: nothing in range
    with torch.no_grad():
        out = enc([torch.from_numpy(f).to(DEV) for f in frames])
        ref = orc.forward(frames).numpy()
    assert_close(out.cpu().numpy(), ref, what="fused scatter+LN vs CPU oracle")
    # an empty frame normalises to the bias exactly: x = 0, mean = 0
    assert torch.equal(out[1].cpu(), enc._layer_norm.bias.detach().cpu())


def test_scatter_layernorm_unsupported_shape_falls_back_to_torch():
    kw = ref_test_kwargs(feat_channels=(64,), T=32, x_range=(-31.25, 31.25), y_range=(-31.25, 31.25), vs=2.5)  # 25 x 25
    enc, _ = encoder_pair(kw, seed=7)
    _randomise_ln(enc._layer_norm, seed=13)
    enc = enc.to(DEV).eval()
    pcs = [torch.from_numpy(f).to(DEV) for f in _frames(5000, 4, (1, 2))]
    with torch.no_grad():
        a = enc(pcs)
        b = enc._layer_norm(enc.encode_batch(pcs))
    assert torch.equal(a, b)


@pytest.mark.parametrize("B,C,ny,nx,P", [(3, 128, 100, 100, 900), (16, 64, 50, 70, 2000), (1, 32, 30, 26, 0), (5, 8, 64, 64, 4096)])
def test_layernorm_forward_walks_are_bit_identical(B, C, ny, nx, P):
    """MBEV_LN_WALK_FRAMES (a warp walks the frames with weight / bias in registers) computes the very expression of
    MBEV_LN_WALK_RUNS per element: same bits, incl. ragged last runs (G % 128 != 0), an empty batch and P = every cell."""
    from mask_bev_b200 import _lib
    from mask_bev_b200 import functional as F_
    rng = np.random.default_rng(B * 1000 + C)
    lin = np.sort(rng.choice(B * ny * nx, P, replace=False))
    coors = np.stack([lin // (ny * nx), np.zeros(P, np.int64), (lin // nx) % ny, lin % nx], 1).astype(np.int32).reshape(P, 4)
    base = np.concatenate([[0], np.cumsum(np.bincount(coors[:, 0], minlength=B))]).astype(np.int32)
    rows = max(P, 1)
    feats = torch.from_numpy(rng.normal(size=(rows, C)).astype(np.float32)).to(DEV)
    coors_d = torch.zeros((rows, 4), dtype=torch.int32, device=DEV)
    coors_d[:P] = torch.from_numpy(coors).to(DEV)
    base_d = torch.from_numpy(base).to(DEV)
    table = F_.build_cell_table(coors_d, base_d[B:], rows, B, ny, nx)
    w = torch.from_numpy(rng.normal(1.0, 0.3, size=(C, ny, nx)).astype(np.float32)).to(DEV)
    b = torch.from_numpy(rng.normal(0.0, 0.3, size=(C, ny, nx)).astype(np.float32)).to(DEV)
    a_out, a_st = F_.scatter_layernorm_forward(feats, table, base_d, B, ny, nx, w, b, 1e-3, walk=_lib.LN_WALK_RUNS)
    f_out, f_st = F_.scatter_layernorm_forward(feats, table, base_d, B, ny, nx, w, b, 1e-3, walk=_lib.LN_WALK_FRAMES)
    assert torch.equal(a_st, f_st)
    assert torch.equal(a_out, f_out)
    canvas = F_.scatter_forward(feats, table, B, ny, nx)
    ref = torch.nn.functional.layer_norm(canvas.double(), (C, ny, nx), w.double(), b.double(), 1e-3)
    assert_close(f_out.cpu().numpy(), ref.cpu().numpy(), what="frame-walking K3+LN vs float64 LayerNorm")
