"""GPU parity against outputs of the reference's OWN code, executed in the build container and committed as fixtures:
tests/golden/encoder_reference.npz (the reference's `MaskBevEncoder` class run end to end, generator
make_golden_encoder.py) and tests/golden/scatter_fossil.npz (its commented scatter + gather, make_golden_scatter.py).
Integer outputs bit-exact, the normalised canvas within 1e-5 of max|ref|. Through the C ABI.
(Named to run last: these tests were written after the last B200 run of round 1.)"""
import numpy as np
import pytest
import torch

from helpers import assert_close, encoder_reference

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _product_encoder():
    import mask_bev_b200 as M
    frames, weights, out, kw = encoder_reference()
    enc = M.MaskBevEncoder(**kw)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=True)
    return enc.to(DEV).eval(), [torch.from_numpy(f).to(DEV) for f in frames], out


def test_voxelize_equals_the_reference_encoder_run():
    enc, pcs, out = _product_encoder()
    voxels, num_points, coors = enc.voxelize(pcs)
    assert np.array_equal(coors.cpu().numpy(), out["coors"]), "pillar coordinates (b, z, y, x)"
    assert np.array_equal(num_points.cpu().numpy(), out["num_points"]), "points per pillar"
    assert np.array_equal(voxels.cpu().numpy(), out["voxels"]), "kept points, first T per pillar in input order"


@pytest.mark.parametrize("grad", [False, True])
def test_forward_equals_the_reference_encoder_run(grad):
    """no_grad: K1 -> K2 (tcgen05 or FMA by stack) -> K3+LayerNorm; with autograd: the FMA PFN and the fused pair."""
    enc, pcs, out = _product_encoder()
    with torch.set_grad_enabled(grad):
        img = enc(pcs)
    assert img.requires_grad == grad
    assert tuple(img.shape) == out["pseudo_img"].shape
    assert_close(img.detach().cpu().numpy(), out["pseudo_img"], what="pseudo image vs the reference's own run")
    # frame 1 has nothing in range: LayerNorm of an all-zero canvas is its bias
    assert torch.equal(img[1].detach().cpu(), enc._layer_norm.bias.detach().cpu())


def test_scatter_and_gather_on_the_reference_fossil_inputs():
    """K3 / K3' against outputs of the reference's own (commented) scatter + gather code on the committed inputs
    (tests/golden/scatter_fossil.npz, generator make_golden_scatter.py): canvas read back at the reference's query
    coordinates equals its `center_per_point` bit for bit, and the backward is that same gather."""
    import mask_bev_b200 as M
    from helpers import check_canvas_against_scatter_fossil, scatter_fossil
    g = scatter_fossil()
    B, C, ny, nx = (int(v) for v in g["shape"])
    sc = M.PointPillarsScatter(C, [ny, nx])
    f = torch.from_numpy(g["voxel_mean"]).to(DEV).requires_grad_(True)
    out = sc(f, torch.from_numpy(g["voxel_coors"]).to(DEV), B)
    check_canvas_against_scatter_fossil(out.detach().cpu().numpy(), g)
    out.backward(out.detach().clone())           # K3': dfeats[p] = dcanvas[b, :, y, x] — here the canvas itself
    assert np.array_equal(f.grad.cpu().numpy(), g["voxel_mean"])
    pc = g["pts_coors"].astype(np.int64)         # and the reference's own gather, through the backward kernel
    q = M.PointPillarsScatter(C, [ny, nx])
    uniq, first = np.unique(pc[:, [0, 2, 3]], axis=0, return_index=True)
    qc = g["pts_coors"][np.sort(first)]          # query cells, each once (a coors list names a cell once)
    z = torch.zeros((len(qc), C), device=DEV, requires_grad=True)
    q(z, torch.from_numpy(qc).to(DEV), B).backward(out.detach().clone())
    assert np.array_equal(z.grad.cpu().numpy(), g["center_per_point"][np.sort(first)])
