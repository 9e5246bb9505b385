"""Shared test helpers: matched oracle / product modules and tolerances.

Tolerances (BASELINE.json north_star): indices, counts, coordinates, occupancy bit-exact; pillar features and
canvas within 1e-5 relative in fp32. "Relative" is taken against the largest magnitude of the reference tensor
(max|a-b| <= tol * max|ref|): post-ReLU features contain exact and near zeros for which an element-wise ratio is
meaningless under any change of summation order.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402

FP32_REL_TOL = 1e-5


def rel_err(a, ref) -> float:
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


def assert_close(a, ref, tol=FP32_REL_TOL, what=""):
    e = rel_err(a, ref)
    assert e <= tol, f"{what}: max|a-ref|/max|ref| = {e:.3e} > {tol:g}"
    return e


def encoder_pair(kwargs, seed=0, dtype=torch.float32):
    """(product MaskBevEncoder on cuda:0, oracle MaskBevEncoderOracle) with identical randomised PFN weights."""
    import mask_bev_b200 as M
    okw = dict(feat_channels=kwargs["feat_channels"], x_range=kwargs["x_range"], y_range=kwargs["y_range"],
               z_range=kwargs["z_range"], voxel_size_x=kwargs["voxel_size_x"], voxel_size_y=kwargs["voxel_size_y"],
               voxel_size_z=kwargs["voxel_size_z"], max_num_points=kwargs["max_num_points"],
               max_voxels=kwargs.get("max_voxels", 500 * 500), pc_point_dim=kwargs.get("pc_point_dim", 4),
               with_distance=kwargs.get("encoder_params", {}).get("with_distance", False), dtype=dtype)
    orc = O.MaskBevEncoderOracle(**okw)
    O.randomise_pfn(orc.pfn, seed=seed)
    enc = M.MaskBevEncoder(**kwargs)
    sd = {"_voxel_encoder." + k: v.float() for k, v in orc.pfn.state_dict().items()}
    missing, unexpected = enc.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("_layer_norm") for k in missing), missing
    return enc, orc


def ref_test_kwargs(feat_channels=(16, 32, 64), T=100, C=4, x_range=(-40, 40), y_range=(-40, 40), z_range=(-20, 20),
                vs=0.16, max_voxels=500 * 500):
    """The reference tests' configuration (mask_bev_test/models/semantic_kitti/test_point_mask_encoders.py:16-35)."""
    return dict(feat_channels=list(feat_channels), x_range=x_range, y_range=y_range, z_range=z_range,
                voxel_size_x=vs, voxel_size_y=vs, voxel_size_z=z_range[1] - z_range[0], max_num_points=T,
                encoding_type="vanilla", fourier_enc_group=1, max_voxels=max_voxels,
                encoder_params=dict(with_distance=True), pc_point_dim=C)


def scatter_fossil():
    """Golden vectors of the scatter's index arithmetic from the reference's own (commented) code —
    tests/golden/make_golden_scatter.py. Returns the dict of arrays."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "scatter_fossil.npz")))


def check_canvas_against_scatter_fossil(canvas, g):
    """`canvas` (B, C, ny, nx) built from g['voxel_mean'] / g['voxel_coors'] must, read back at g['pts_coors'], give the
    reference's `center_per_point` bit for bit; every voxel row must sit at its own cell; everything else is zero."""
    B, C, ny, nx = (int(v) for v in g["shape"])
    canvas = np.asarray(canvas)
    assert canvas.shape == (B, C, ny, nx) and canvas.dtype == np.float32
    pc, vc = g["pts_coors"].astype(np.int64), g["voxel_coors"].astype(np.int64)
    assert np.array_equal(canvas[pc[:, 0], :, pc[:, 2], pc[:, 3]], g["center_per_point"])
    assert np.array_equal(canvas[vc[:, 0], :, vc[:, 2], vc[:, 3]], g["voxel_mean"])
    occ = np.zeros((B, ny, nx), bool)
    occ[vc[:, 0], vc[:, 2], vc[:, 3]] = True
    assert np.count_nonzero(canvas.transpose(0, 2, 3, 1)[~occ]) == 0


def encoder_reference():
    """Inputs, weights and outputs of the reference's OWN MaskBevEncoder (tests/golden/make_golden_encoder.py):
    (frames, state_dict of numpy arrays, dict of outputs, constructor kwargs)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "encoder_reference.npz"))
    frames = [g[f"frame{i}"] for i in range(sum(1 for k in g.files if k.startswith("frame")))]
    weights = {k[2:]: g[k] for k in g.files if k.startswith("w:")}
    out = {k: g[k] for k in ("voxels", "num_points", "coors", "pseudo_img", "canvas_shape")}
    kw = dict(feat_channels=[16, 32], x_range=(-8, 8), y_range=(-6, 6), z_range=(-2, 2), voxel_size_x=0.5,
              voxel_size_y=0.5, voxel_size_z=4, max_num_points=8, encoding_type='vanilla', fourier_enc_group=1,
              max_voxels=250000, encoder_params=dict(with_distance=True), pc_point_dim=4)
    return frames, weights, out, kw
