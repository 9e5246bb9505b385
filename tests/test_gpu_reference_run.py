"""GPU parity against outputs of the reference's OWN code, executed in the build container and committed as fixtures:
tests/golden/encoder_reference.npz (the reference's `MaskBevEncoder` class run end to end, generator
make_golden_encoder.py) and tests/golden/scatter_fossil.npz (its commented scatter + gather, make_golden_scatter.py).
Integer outputs bit-exact, the normalised canvas within 1e-5 of max|ref|. Through the C ABI.
The fixture holds the outputs of `MaskBevEncoder.voxelize` / `.forward` of the reference file
(mask_bev_encoders.py:77-111). Three routes of the product must reproduce them: the fused batch path, the per-frame
MODULE-LEVEL path sequenced exactly as the reference's forward sequences it (filter -> Voxelization per frame -> pad the
batch column -> cat -> PillarFeatureNet -> PointPillarsScatter -> LayerNorm), and — where the reference tree is
mounted next to a GPU (MASK_BEV_REFERENCE, default /root/reference) — the reference's own file imported unchanged
through mask_bev_b200/shims and run on the device."""
import importlib
import os
import sys

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


def _reference_forward_sequence(enc, pcs):
    """The statements of mask_bev_encoders.py:77-111 on the product's module-level API (one Voxelization call per
    frame, host-side concatenation), i.e. what the reference file does once its three imports resolve to the shims."""
    import torch.nn.functional as F
    voxels, coors, num_points = [], [], []
    for res in pcs:
        res = enc._filter_in_range(res)
        v, c, n = enc._voxel_layer(res)
        voxels.append(v)
        coors.append(c)
        num_points.append(n)
    voxels = torch.cat(voxels, dim=0)
    num_points = torch.cat(num_points, dim=0)
    coors_batch = torch.cat([F.pad(c, (1, 0), mode='constant', value=i) for i, c in enumerate(coors)], dim=0)
    feats = enc.encode(voxels, num_points, coors_batch)
    img = enc.middle_encode(feats, coors_batch, len(pcs))   # batch_size = len(point_clouds), :83
    return voxels, num_points, coors_batch, img


def test_per_frame_module_api_sequence_equals_the_reference_encoder_run():
    enc, pcs, out = _product_encoder()
    with torch.no_grad():
        voxels, num_points, coors, img = _reference_forward_sequence(enc, pcs)
        img = enc._layer_norm(img)
    assert np.array_equal(coors.cpu().numpy(), out["coors"])
    assert np.array_equal(num_points.cpu().numpy(), out["num_points"])
    assert np.array_equal(voxels.cpu().numpy(), out["voxels"])
    assert_close(img.cpu().numpy(), out["pseudo_img"], what="per-frame module path vs the reference's own run")
    with torch.no_grad():
        fused = enc(pcs)
    assert_close(fused.cpu().numpy(), img.cpu().numpy(), what="fused batch path vs per-frame module path")


def test_reference_file_itself_runs_on_the_device_through_the_shims():
    ref_root = os.environ.get("MASK_BEV_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref_root, "mask_bev", "models", "encoders", "mask_bev_encoders.py")):
        pytest.skip(f"reference tree not mounted at {ref_root} (it cannot travel to the GPU box: sources are not copied)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    frames, weights, out, kw = encoder_reference()
    saved = list(sys.path)
    try:
        sys.path.insert(0, os.path.join(root, "mask_bev_b200", "shims"))
        sys.path.insert(0, ref_root)
        for m in [m for m in sys.modules if m.split(".")[0] in ("mmcv", "mmdet3d", "mask_bev")]:
            del sys.modules[m]
        mod = importlib.import_module("mask_bev.models.encoders.mask_bev_encoders")
        enc = mod.MaskBevEncoder(**kw)
        enc.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=True)
        enc = enc.to(DEV).eval()
        pcs = [torch.from_numpy(f).to(DEV) for f in frames]
        with torch.no_grad():
            voxels, num_points, coors = enc.voxelize(pcs)
            img = enc(pcs)
        assert np.array_equal(coors.cpu().numpy(), out["coors"])
        assert np.array_equal(num_points.cpu().numpy(), out["num_points"])
        assert np.array_equal(voxels.cpu().numpy(), out["voxels"])
        assert_close(img.cpu().numpy(), out["pseudo_img"], what="the reference file on the device vs its CPU run")
    finally:
        sys.path[:] = saved
        for m in [m for m in sys.modules if m.split(".")[0] in ("mmcv", "mmdet3d", "mask_bev")]:
            del sys.modules[m]
