#!/usr/bin/env python
"""Golden vectors of the whole encoder, produced by EXECUTING the reference's own `MaskBevEncoder`.

/root/reference/mask_bev/models/encoders/mask_bev_encoders.py:21-123 is live code: geometry (`int((hi - lo) / v)`),
`_filter_in_range` (strict bounds), the per-frame `voxelize` loop with its concatenation and `F.pad` batch column,
`encode`, `middle_encode` and the trailing `nn.LayerNorm([C, ny, nx], eps=1e-3)`. Only the three upstream classes it
imports (`mmcv.ops.Voxelization`, `mmdet3d.models.PillarFeatureNet` / `PointPillarsScatter`) are absent from this
machine. This script registers CPU stand-ins for exactly those three names (built on the restatements in oracle/, same
constructor arguments and return conventions as upstream), imports the reference module unchanged from where it lies,
builds ITS `MaskBevEncoder`, runs ITS `voxelize` and `forward` on seeded frames (points exactly on the range bounds,
a frame with nothing in range, pillars with more than T points) and stores inputs, weights and outputs under
tests/golden/encoder_reference.npz. What this pins: every line of the reference file on the path. What it does not:
the arithmetic inside the three upstream classes (see oracle/oracle.py's header). Nothing of the reference is copied.

    python tests/golden/make_golden_encoder.py        # needs /root/reference (this container only)
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF_ROOT = "/root/reference"
OUT = os.path.join(HERE, "encoder_reference.npz")

KW = dict(feat_channels=[16, 32], x_range=(-8, 8), y_range=(-6, 6), z_range=(-2, 2), voxel_size_x=0.5,
          voxel_size_y=0.5, voxel_size_z=4, max_num_points=8, encoding_type='vanilla', fourier_enc_group=1,
          max_voxels=250000, encoder_params=dict(with_distance=True), pc_point_dim=4)


def upstream_standins():
    """(mmcv.ops, mmdet3d.models) module objects exporting the three names the reference imports."""

    class Voxelization(nn.Module):
        def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, deterministic=True):
            super().__init__()
            self.voxel_size, self.point_cloud_range = voxel_size, point_cloud_range
            self.max_num_points = max_num_points
            self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else (max_voxels, max_voxels)

        def forward(self, points):
            mv = self.max_voxels[0] if self.training else self.max_voxels[1]
            v, c, n, _ = O.hard_voxelize_np(points.detach().numpy(), self.voxel_size, self.point_cloud_range,
                                            self.max_num_points, mv)
            return torch.from_numpy(v), torch.from_numpy(c.astype(np.int32)), torch.from_numpy(n.astype(np.int32))

    def PillarFeatureNet(in_channels=4, feat_channels=(64,), with_distance=False, with_cluster_center=True,
                         with_voxel_center=True, voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1),
                         legacy=True, **_):
        return O.make_pfn_oracle(in_channels=in_channels, feat_channels=feat_channels, with_distance=with_distance,
                                 with_cluster_center=with_cluster_center, with_voxel_center=with_voxel_center,
                                 voxel_size=voxel_size, point_cloud_range=point_cloud_range, legacy=legacy)

    class PointPillarsScatter(nn.Module):
        def __init__(self, in_channels, output_shape):
            super().__init__()
            self.ny, self.nx = output_shape

        def forward(self, voxel_features, coors, batch_size=None):
            c = coors.numpy()
            if batch_size is None:
                c, batch_size = c.copy(), 1
                c[:, 0] = 0
            return torch.from_numpy(O.scatter_np(voxel_features.detach().numpy(), c, batch_size, self.ny, self.nx))

    ops = types.ModuleType("mmcv.ops")
    ops.Voxelization = Voxelization
    mmcv = types.ModuleType("mmcv")
    mmcv.ops = ops
    models = types.ModuleType("mmdet3d.models")
    models.PillarFeatureNet, models.PointPillarsScatter = PillarFeatureNet, PointPillarsScatter
    mmdet3d = types.ModuleType("mmdet3d")
    mmdet3d.models = models
    return {"mmcv": mmcv, "mmcv.ops": ops, "mmdet3d": mmdet3d, "mmdet3d.models": models}


def reference_encoder_class():
    """The reference's MaskBevEncoder, imported unchanged with the stand-ins registered for its two absent imports."""
    saved_path = list(sys.path)
    doomed = [m for m in sys.modules if m.split(".")[0] in ("mmcv", "mmdet3d", "mask_bev")]
    saved_mods = {m: sys.modules.pop(m) for m in doomed}
    try:
        sys.modules.update(upstream_standins())
        sys.path.insert(0, REF_ROOT)
        return importlib.import_module("mask_bev.models.encoders.mask_bev_encoders").MaskBevEncoder
    finally:
        sys.path[:] = saved_path
        for m in [m for m in sys.modules if m.split(".")[0] in ("mmcv", "mmdet3d", "mask_bev")]:
            del sys.modules[m]
        sys.modules.update(saved_mods)


def make_frames():
    rng = np.random.default_rng(20261019)

    def cloud(n):
        p = np.empty((n, 4), np.float32)
        p[:, 0] = rng.uniform(-9, 9, n)
        p[:, 1] = rng.uniform(-7, 7, n)
        p[:, 2] = rng.uniform(-2.5, 2.5, n)
        p[:, 3] = rng.uniform(0, 1, n)
        return p
    f0 = cloud(1500)
    f0[:6, :3] = [[-8, 0, 0], [8, 0, 0], [0, -6, 0], [0, 6, 0], [0, 0, -2], [0, 0, 2]]   # ON the bounds: dropped
    f0[6:10, :3] = [[-7.999, 0, 0], [7.999, 5.999, 1.999], [-7.75, -5.75, -1.999], [0, 0, 0]]  # just inside: kept
    f0[10:30, :2] = np.array([3.1, 2.1]) + rng.uniform(0, 0.3, (20, 2))              # 20 points in one pillar (> T = 8)
    f0[10:30, 2] = rng.uniform(-1, 1, 20)
    f1 = cloud(40)
    f1[:, 0] += 100.0                                                                # nothing in range: empty frame
    f2 = cloud(900)
    return [f0, f1, f2]


def randomise(enc, seed=7):
    O.randomise_pfn(enc._voxel_encoder, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        enc._layer_norm.weight.copy_(1.0 + 0.3 * torch.randn(enc._layer_norm.weight.shape, generator=g))
        enc._layer_norm.bias.copy_(0.3 * torch.randn(enc._layer_norm.bias.shape, generator=g))


def run_reference(frames):
    enc = reference_encoder_class()(**KW)
    randomise(enc)
    enc.eval()
    pcs = [torch.from_numpy(f) for f in frames]
    with torch.no_grad():
        voxels, num_points, coors = enc.voxelize(pcs)
        img = enc(pcs)
    out = dict(voxels=voxels.numpy(), num_points=num_points.numpy(), coors=coors.numpy(), pseudo_img=img.numpy(),
               canvas_shape=np.array([enc._num_voxel_y, enc._num_voxel_x], np.int64))
    weights = {"w:" + k: v.numpy() for k, v in enc.state_dict().items()}
    return out, weights


def main():
    frames = make_frames()
    out, weights = run_reference(frames)
    np.savez_compressed(OUT, **{f"frame{i}": f for i, f in enumerate(frames)}, **out, **weights)
    print("wrote", OUT, {k: v.shape for k, v in out.items()}, "pillars:", len(out["num_points"]),
          "max points/pillar:", int(out["num_points"].max()), "weights:", len(weights))


if __name__ == "__main__":
    main()
