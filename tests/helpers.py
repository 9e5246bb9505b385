"""Shared test helpers: matched oracle / product modules and tolerances.

Tolerances (BASELINE.json north_star): indices, counts, coordinates, occupancy bit-exact; pillar features and
canvas within 1e-5 relative in fp32. Two checks, both must hold (assert_close):
  * norm-wise:     max|a - ref| <= tol * max|ref|
  * element-wise:  |a - ref| <= tol * |ref| + ATOL_FRAC * tol * max|ref|   for EVERY element.
The absolute term exists because a feature is a sum of up to 128 products that cancel (BatchNorm shift, ReLU at 0):
its rounding error scales with the magnitude of the TERMS, not of the result, so a near-zero output has no meaningful
element-wise ratio under any change of summation order — torch's own fp32 result sits ~1e-6 of max|ref| from float64
on these tensors. ATOL_FRAC = 0.5 keeps that floor at 2.5e-6 of max|ref| for tol = 1e-5. Every call appends how many
elements needed the absolute term to gpurun_out/parity_elementwise.jsonl (summarised in DESIGN.md).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402

FP32_REL_TOL = 1e-5
ATOL_FRAC = 0.5
_REPORT = os.path.join(ROOT, "gpurun_out", "parity_elementwise.jsonl")


def rel_err(a, ref) -> float:
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


def elementwise_report(a, ref, tol=FP32_REL_TOL, atol_frac=ATOL_FRAC) -> dict:
    """Element-wise comparison: how many elements pass on the relative term alone, how many need the absolute term,
    how many fail, and the worst excess over the bound."""
    a = np.asarray(a, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    if ref.size == 0:
        return dict(n=0, rel_only=0, need_abs=0, fail=0, worst_ratio=0.0)
    mx = max(float(np.abs(ref).max()), 1e-30)
    d = np.abs(a - ref)
    rel_ok = d <= tol * np.abs(ref)
    bound = tol * np.abs(ref) + atol_frac * tol * mx
    ok = d <= bound
    return dict(n=int(ref.size), rel_only=int(rel_ok.sum()), need_abs=int((ok & ~rel_ok).sum()), fail=int((~ok).sum()),
                worst_ratio=float((d / bound).max()), max_abs_over_max_ref=float(d.max() / mx))


def assert_close(a, ref, tol=FP32_REL_TOL, what="", elementwise=True):
    e = rel_err(a, ref)
    assert e <= tol, f"{what}: max|a-ref|/max|ref| = {e:.3e} > {tol:g}"
    if elementwise:
        r = elementwise_report(a, ref, tol)
        try:
            os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
            with open(_REPORT, "a") as f:
                f.write(json.dumps(dict(what=what, tol=tol, norm_wise=e, **r)) + "\n")
        except OSError:
            pass
        assert r["fail"] == 0, (f"{what}: {r['fail']} of {r['n']} elements exceed tol*|ref| + {ATOL_FRAC}*tol*max|ref| "
                                f"(worst {r['worst_ratio']:.2f}x the bound; {r['need_abs']} needed the absolute term)")
    return e


def assert_close_arbitrated(a, ref32, ref64, tol=FP32_REL_TOL, what=""):
    """Train-mode BatchNorm divides every activation by a batch standard deviation that torch computes in float32
    over P*T slots, so the float32 reference itself can sit further than `tol` from the float64 result. Float64
    arbitrates: `a` must be within tol of the float32 reference (norm-wise and element-wise) — or, failing that, no
    further from ref64 than max(tol, 1.5x the float32 reference's own distance from ref64). All three distances go to
    the element-wise report."""
    e32, e64, spread = rel_err(a, ref32), rel_err(a, ref64), rel_err(ref32, ref64)
    if e32 <= tol:
        # the usual case: within tol of the float32 reference itself, norm-wise and element-wise
        rep, against, ok = elementwise_report(a, ref32, tol), "f32_reference", True
        ok = rep["fail"] == 0
    else:
        bound = max(tol, 1.5 * spread)
        rep, against, ok = elementwise_report(a, ref64, bound), "f64", e64 <= bound
    try:
        os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
        with open(_REPORT, "a") as f:
            f.write(json.dumps(dict(what=what, tol=tol, elementwise_against=against, vs_f32_reference=e32, vs_f64=e64,
                                    f32_reference_vs_f64=spread, **rep)) + "\n")
    except OSError:
        pass
    assert ok, (f"{what}: {e32:.3e} from the float32 reference, {e64:.3e} from float64 (the float32 reference itself: "
                f"{spread:.3e}); element-wise against {against}: {rep['fail']} of {rep['n']} fail")
    return e64


def oracle64_of(orc, kwargs):
    """float64 twin of a float32 oracle encoder (same weights, same BatchNorm buffers, same train / eval mode)."""
    o64 = O.MaskBevEncoderOracle(
        feat_channels=kwargs["feat_channels"], x_range=kwargs["x_range"], y_range=kwargs["y_range"],
        z_range=kwargs["z_range"], voxel_size_x=kwargs["voxel_size_x"], voxel_size_y=kwargs["voxel_size_y"],
        voxel_size_z=kwargs["voxel_size_z"], max_num_points=kwargs["max_num_points"],
        max_voxels=kwargs.get("max_voxels", 500 * 500), pc_point_dim=kwargs.get("pc_point_dim", 4),
        with_distance=kwargs.get("encoder_params", {}).get("with_distance", False), dtype=torch.float64)
    o64.pfn.load_state_dict({k: v.double() if v.is_floating_point() else v.clone() for k, v in orc.pfn.state_dict().items()})
    o64.pfn.train(orc.pfn.training)
    return o64


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
