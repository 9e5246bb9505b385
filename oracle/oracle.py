"""ORACLE — CPU restatement of MaskBEV's point-cloud -> BEV front end.

TEST INFRASTRUCTURE ONLY. Nothing under ``mask_bev_b200/`` imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs do, and there
only as the checker or as the timed CPU baseline.

PARITY UNPINNED, EXCEPT THE POINT DECORATION AND THE SCATTER / GATHER INDEX ARITHMETIC (see the last two bullets). The reference path (``/root/reference/mask_bev/models/encoders/mask_bev_encoders.py:21-123``)
delegates its arithmetic to ``mmcv==2.0.0`` (``mmcv.ops.Voxelization``) and ``mmdet3d==1.1.0``
(``PillarFeatureNet`` / ``PFNLayer`` / ``PointPillarsScatter``), pinned in ``Dockerfile:25,28``; neither
is vendored, installed or installable offline, and the reference's own tests
(``mask_bev_test/models/*/test_*_encoders.py``) assert shapes and ranges only. This file restates the
published upstream algorithms and is pinned by
  * the hand-derivable golden vectors of SURVEY.md A.6 (``tests/golden/a6_*.json``),
  * three independent voxelizer restatements that must agree (pure-Python loop, numpy stable-sort, C),
  * two independent PFN restatements that must agree (dense upstream op sequence with torch CPU ops vs
    sparse "virtual row" float64 numpy form),
  * the invariants of SURVEY.md A.5 under hypothesis,
  * and, for the decoration only, outputs of the reference's OWN code: the commented mmdet3d-0.x ``forward`` kept in
    ``mask_bev_encoders.py:270-317`` is executed by ``tests/golden/make_golden_decoration.py`` where the reference
    lies; ``decorate(voxel_center_dims=2)`` reproduces the committed vectors bit for bit (tests/test_oracle.py),
  * and, for the scatter and its gather backward, the commented ``map_voxel_center_to_point`` of the same file
    (``canvas[:, b*ny*nx + y*nx + x] = rows.t()`` then a per-coordinate gather), executed by
    ``tests/golden/make_golden_scatter.py``; ``scatter_np`` reproduces ``scatter_fossil.npz`` bit for bit.
  * and, for everything the reference file itself does (geometry, strict filter, per-frame loop, concatenation, batch
    column, LayerNorm), the reference's OWN ``MaskBevEncoder`` class executed end to end by
    ``tests/golden/make_golden_encoder.py`` with stand-ins for the three absent upstream classes;
    ``MaskBevEncoderOracle`` reproduces ``encoder_reference.npz`` bit for bit.
  The voxelizer (mmcv's compiled op) and the PFN layers (Linear / BN1d / max / concat of mmdet3d's ``PFNLayer``) have no
  code on disk to execute: they stay pinned by the restatement agreements above only.

Reference call sites each function follows are cited in its docstring.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


# --------------------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------------------
def encoder_geometry(x_range, y_range, z_range, voxel_size_x, voxel_size_y, voxel_size_z):
    """mask_bev_encoders.py:63-68 — grid (int() of python-float division), range and voxel size lists."""
    nx = int((x_range[1] - x_range[0]) / voxel_size_x)
    ny = int((y_range[1] - y_range[0]) / voxel_size_y)
    pc_range = [x_range[0], y_range[0], z_range[0], x_range[1], y_range[1], z_range[1]]
    voxel_size = [voxel_size_x, voxel_size_y, voxel_size_z]
    return dict(nx=nx, ny=ny, nz=1, point_cloud_range=pc_range, voxel_size=voxel_size)


def grid_size(pc_range: Sequence[float], voxel_size: Sequence[float]) -> Tuple[int, int, int]:
    """mmcv/ops/voxelize.py Voxelization.__init__: round((hi-lo)/vs) evaluated in float32. -> (nx, ny, nz)."""
    r = np.asarray(pc_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    g = np.round((r[3:] - r[:3]) / v)
    return int(g[0]), int(g[1]), int(g[2])


# --------------------------------------------------------------------------------------------------
# A.1 range filter — mask_bev_encoders.py:113-117
# --------------------------------------------------------------------------------------------------
def filter_in_range(points: np.ndarray, x_range, y_range, z_range) -> Tuple[np.ndarray, np.ndarray]:
    """Strict ``lo < v < hi`` on x, y, z in float32 (python-scalar bounds are compared as float32).
    Returns (filtered points, source indices), input order preserved."""
    p = np.asarray(points, dtype=np.float32)
    f = np.float32
    m = ((f(x_range[0]) < p[:, 0]) & (p[:, 0] < f(x_range[1]))
         & (f(y_range[0]) < p[:, 1]) & (p[:, 1] < f(y_range[1]))
         & (f(z_range[0]) < p[:, 2]) & (p[:, 2] < f(z_range[1])))
    idx = np.nonzero(m)[0]
    return p[idx], idx


# --------------------------------------------------------------------------------------------------
# A.2 hard voxelization — mmcv hard_voxelize_forward (call site mask_bev_encoders.py:100)
# --------------------------------------------------------------------------------------------------
def _cell_coords(points: np.ndarray, voxel_size, pc_range):
    """dynamic_voxelize: c_j = floor((p_j - lo_j) / vs_j) in float32; valid iff 0 <= c_j < grid_j."""
    p = np.asarray(points, dtype=np.float32)
    lo = np.asarray(pc_range[:3], dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    g = np.asarray(grid_size(pc_range, voxel_size), dtype=np.int64)
    with np.errstate(invalid="ignore", over="ignore"):
        q = np.floor((p[:, :3] - lo) / vs)  # float32 subtract, float32 divide
        ok = np.isfinite(q).all(axis=1)
        q = np.where(np.isfinite(q), q, -1.0)
        q = np.clip(q, -2.0, 2.0e9)
    c = q.astype(np.int64)
    ok &= ((c >= 0) & (c < g)).all(axis=1)
    return c, ok, g


def hard_voxelize_py(points, voxel_size, pc_range, max_num_points: int, max_voxels: int):
    """Literal pure-Python transcription of SURVEY.md A.2 (small inputs only)."""
    p = np.asarray(points, dtype=np.float32)
    n, cdim = p.shape
    c, ok, g = _cell_coords(p, voxel_size, pc_range)
    T = max_num_points
    table = {}
    coors: List[Tuple[int, int, int]] = []
    nump: List[int] = []
    kept: List[List[int]] = []
    for i in range(n):
        if not ok[i]:
            continue
        key = (int(c[i, 2]), int(c[i, 1]), int(c[i, 0]))
        v = table.get(key, -1)
        if v == -1:
            if len(coors) >= max_voxels:
                continue
            v = len(coors)
            table[key] = v
            coors.append(key)
            nump.append(0)
            kept.append([])
        if nump[v] < T:
            kept[v].append(i)
            nump[v] += 1
    P = len(coors)
    kept_idx = np.full((P, T), -1, dtype=np.int64)
    for v in range(P):
        kept_idx[v, :len(kept[v])] = kept[v]
    return _materialise(p, np.asarray(coors, dtype=np.int32).reshape(P, 3),
                        np.asarray(nump, dtype=np.int32), kept_idx)


def hard_voxelize_np(points, voxel_size, pc_range, max_num_points: int, max_voxels: int):
    """Independent vectorised restatement (stable sort on the linear cell id)."""
    p = np.asarray(points, dtype=np.float32)
    c, ok, g = _cell_coords(p, voxel_size, pc_range)
    T = max_num_points
    src = np.nonzero(ok)[0]
    cell = (c[src, 2] * g[1] + c[src, 1]) * g[0] + c[src, 0]
    order = np.argsort(cell, kind="stable")
    cs = cell[order]
    head = np.ones(len(cs), dtype=bool)
    head[1:] = cs[1:] != cs[:-1]
    seg_id = np.cumsum(head) - 1                      # segment (cell) of each sorted slot
    seg_start = np.nonzero(head)[0]
    rank = np.arange(len(cs)) - seg_start[seg_id]     # input-order rank inside the cell
    first_src = src[order][seg_start]                 # first point (input order) of each cell
    pill_order = np.argsort(first_src, kind="stable")  # pillars in first-appearance order
    pid_of_seg = np.empty(len(seg_start), dtype=np.int64)
    pid_of_seg[pill_order] = np.arange(len(seg_start))
    pid = pid_of_seg[seg_id]
    keep = (rank < T) & (pid < max_voxels)
    P = int(min(len(seg_start), max_voxels))
    kept_idx = np.full((P, T), -1, dtype=np.int64)
    kept_idx[pid[keep], rank[keep]] = src[order][keep]
    counts = np.bincount(seg_id, minlength=len(seg_start))
    nump = np.minimum(counts, T)[pill_order][:P].astype(np.int32)
    fc = c[first_src[pill_order][:P]]
    coors = np.stack([fc[:, 2], fc[:, 1], fc[:, 0]], axis=1).astype(np.int32).reshape(P, 3)
    return _materialise(p, coors, nump, kept_idx)


def _materialise(p, coors, nump, kept_idx):
    P, T = kept_idx.shape
    voxels = np.zeros((P, T, p.shape[1]), dtype=np.float32)
    m = kept_idx >= 0
    voxels[m] = p[kept_idx[m]]
    return voxels, coors, nump, kept_idx


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/hard_voxelize.c (gcc) if the .so is missing or stale."""
    src = os.path.join(_HERE, "hard_voxelize.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/liboracle.so"])
    return _LIB_PATH


def _load_c():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build_c_oracle()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.mbev_oracle_hard_voxelize.restype = ctypes.c_int64
        lib.mbev_oracle_hard_voxelize.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p]
        lib.mbev_oracle_filter_in_range.restype = ctypes.c_int64
        lib.mbev_oracle_filter_in_range.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                                    ctypes.c_void_p, ctypes.c_void_p]
        lib.mbev_oracle_scatter.restype = None
        lib.mbev_oracle_scatter.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib = lib
    return _lib


def hard_voxelize_c(points, voxel_size, pc_range, max_num_points: int, max_voxels: int,
                    want_kept: bool = True):
    """C restatement (oracle/hard_voxelize.c) — serial like mmcv's CPU kernel; the timed CPU baseline."""
    lib = _load_c()
    p = np.ascontiguousarray(points, dtype=np.float32)
    n, cdim = p.shape
    T = max_num_points
    cap = int(min(max_voxels, max(n, 1)))
    voxels = np.zeros((cap, T, cdim), dtype=np.float32)
    coors = np.zeros((cap, 3), dtype=np.int32)
    nump = np.zeros((cap,), dtype=np.int32)
    kept = np.full((cap, T), -1, dtype=np.int64) if want_kept else None
    vs = np.asarray(voxel_size, dtype=np.float32)
    rg = np.asarray(pc_range, dtype=np.float32)
    P = lib.mbev_oracle_hard_voxelize(
        p.ctypes.data, n, cdim, vs.ctypes.data, rg.ctypes.data, T, int(max_voxels),
        voxels.ctypes.data, coors.ctypes.data, nump.ctypes.data,
        kept.ctypes.data if kept is not None else None)
    if P < 0:
        raise MemoryError("oracle table allocation failed")
    return voxels[:P], coors[:P], nump[:P], (kept[:P] if kept is not None else None)


def filter_in_range_c(points, x_range, y_range, z_range):
    lib = _load_c()
    p = np.ascontiguousarray(points, dtype=np.float32)
    rg = np.asarray([x_range[0], y_range[0], z_range[0], x_range[1], y_range[1], z_range[1]],
                    dtype=np.float32)
    idx = np.empty((p.shape[0],), dtype=np.int64)
    m = lib.mbev_oracle_filter_in_range(p.ctypes.data, p.shape[0], p.shape[1], rg.ctypes.data,
                                        idx.ctypes.data)
    idx = idx[:m]
    return p[idx], idx


# --------------------------------------------------------------------------------------------------
# A.3 / A.4 pillar feature net — mmdet3d PillarFeatureNet / PFNLayer (call site mask_bev_encoders.py:70-72,120)
# Dense restatement with torch CPU ops: the upstream op sequence verbatim in structure
# (Linear(bias=False) -> BatchNorm1d(eps 1e-3, momentum 0.01) over (P,units,T) -> ReLU -> max over T -> concat).
# --------------------------------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def make_pfn_oracle(in_channels=4, feat_channels=(64,), with_distance=False, with_cluster_center=True,
                    with_voxel_center=True, voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1),
                    legacy=True, voxel_center_dims=3, dtype=None):
    """Build the dense oracle module; state-dict keys equal upstream's (pfn_layers.{l}.linear/norm.*)."""
    torch = _torch()
    nn = torch.nn
    F = torch.nn.functional

    class PFNLayerOracle(nn.Module):
        def __init__(self, cin, cout, last_layer):
            super().__init__()
            self.last_vfe = last_layer
            if not last_layer:
                cout = cout // 2
            self.units = cout
            self.norm = nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01)
            self.linear = nn.Linear(cin, cout, bias=False)

        def forward(self, inputs):
            x = self.linear(inputs)
            x = self.norm(x.permute(0, 2, 1).contiguous()).permute(0, 2, 1).contiguous()
            x = F.relu(x)
            x_max = torch.max(x, dim=1, keepdim=True)[0]
            if self.last_vfe:
                return x_max
            return torch.cat([x, x_max.repeat(1, inputs.shape[1], 1)], dim=2)

    class PillarFeatureNetOracle(nn.Module):
        def __init__(self):
            super().__init__()
            cin = in_channels
            if with_cluster_center:
                cin += 3
            if with_voxel_center:
                cin += voxel_center_dims
            if with_distance:
                cin += 1
            self.in_channels = cin
            chans = [cin] + list(feat_channels)
            self.pfn_layers = nn.ModuleList(
                [PFNLayerOracle(chans[i], chans[i + 1], last_layer=(i == len(chans) - 2))
                 for i in range(len(chans) - 1)])
            self.vx, self.vy, self.vz = voxel_size[0], voxel_size[1], voxel_size[2]
            self.x_offset = self.vx / 2 + point_cloud_range[0]
            self.y_offset = self.vy / 2 + point_cloud_range[1]
            self.z_offset = self.vz / 2 + point_cloud_range[2]

        def decorate(self, features, num_points, coors):
            features = features.clone()  # upstream mutates the caller's tensor in legacy mode; values identical
            fl = [features]
            if with_cluster_center:
                mean = features[:, :, :3].sum(dim=1, keepdim=True) / num_points.type_as(features).view(-1, 1, 1)
                fl.append(features[:, :, :3] - mean)
            if with_voxel_center:
                cx = coors[:, 3].type_as(features).unsqueeze(1) * self.vx + self.x_offset
                cy = coors[:, 2].type_as(features).unsqueeze(1) * self.vy + self.y_offset
                cz = coors[:, 1].type_as(features).unsqueeze(1) * self.vz + self.z_offset
                if legacy:
                    # a VIEW of the first `voxel_center_dims` raw channels: the in-place writes alias them. With the
                    # 2-channel centre of mmdet3d 0.x only x, y are touched — pinned against the reference's own
                    # (commented) forward, mask_bev_encoders.py:270-317, through tests/golden/decoration_vcd2.npz.
                    f_center = features[:, :, :voxel_center_dims]
                    f_center[:, :, 0] = f_center[:, :, 0] - cx
                    f_center[:, :, 1] = f_center[:, :, 1] - cy
                    if voxel_center_dims > 2:
                        f_center[:, :, 2] = f_center[:, :, 2] - cz
                else:
                    f_center = torch.zeros_like(features[:, :, :3])
                    f_center[:, :, 0] = features[:, :, 0] - cx
                    f_center[:, :, 1] = features[:, :, 1] - cy
                    f_center[:, :, 2] = features[:, :, 2] - cz
                fl.append(f_center[:, :, :voxel_center_dims])
            if with_distance:
                fl.append(torch.norm(features[:, :, :3], 2, 2, keepdim=True))
            out = torch.cat(fl, dim=-1)
            T = out.shape[1]
            mask = (torch.arange(T, device=out.device).view(1, -1) < num_points.view(-1, 1).int())
            return out * mask.unsqueeze(-1).type_as(out)

        def forward(self, features, num_points, coors):
            x = self.decorate(features, num_points, coors)
            for pfn in self.pfn_layers:
                x = pfn(x)
            return x.squeeze(1)

    m = PillarFeatureNetOracle()
    if dtype is not None:
        m = m.to(dtype)
    return m


def randomise_pfn(module, seed: int = 0, warm_input=None):
    """SURVEY.md §8d weights: default init under manual_seed(seed), then BN gamma~N(1,.3), beta~N(0,.3);
    running stats from one train-mode pass over ``warm_input`` (features, num_points, coors) when given,
    else running_mean~N(0,.2), running_var~U(.5,1.5). Makes ReLU(BN(0)) non-trivial."""
    torch = _torch()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for layer in module.pfn_layers:
            bound = 1.0 / (layer.linear.in_features ** 0.5)
            layer.linear.weight.copy_((torch.rand(layer.linear.weight.shape, generator=g) * 2 - 1) * bound)
            layer.norm.weight.copy_(1.0 + 0.3 * torch.randn(layer.units, generator=g))
            layer.norm.bias.copy_(0.3 * torch.randn(layer.units, generator=g))
            layer.norm.running_mean.copy_(0.2 * torch.randn(layer.units, generator=g))
            layer.norm.running_var.copy_(0.5 + torch.rand(layer.units, generator=g))
    if warm_input is not None:
        was = module.training
        module.train()
        for layer in module.pfn_layers:
            layer.norm.momentum = 1.0
        with torch.no_grad():
            module(*warm_input)
        for layer in module.pfn_layers:
            layer.norm.momentum = 0.01
            layer.norm.num_batches_tracked.zero_()
        module.train(was)
    return module


def pfn_sparse_np(voxels, num_points, coors4, weights, bn, voxel_size, pc_range, T: int, training: bool,
                  eps: float = 1e-3, voxel_center_dims: int = 3, with_distance: bool = True):
    """Independent second restatement in float64 numpy using the virtual-row identity.

    voxels (P,T,C) zero padded; weights[l] (units_l, in_l); bn[l] = dict(weight,bias,running_mean,running_var).
    Works on N_k real rows + one weighted virtual row per pillar instead of P*T rows.
    Returns (features (P, C_out) float64, stats list[(mean, biased_var)] per layer).
    """
    v = np.asarray(voxels, dtype=np.float64)
    n = np.asarray(num_points).astype(np.int64)
    P, Tt, C = v.shape
    assert Tt == T
    f32 = np.float32
    vx, vy, vz = voxel_size
    xo, yo, zo = vx / 2 + pc_range[0], vy / 2 + pc_range[1], vz / 2 + pc_range[2]
    slot = np.arange(T)[None, :] < n[:, None]                                  # (P,T)
    mean = v[:, :, :3].sum(axis=1) / n[:, None]
    # centre computed as upstream does, in float32: c*vx + off
    c32 = np.asarray(coors4)
    ctr = np.stack([(c32[:, 3].astype(f32) * f32(vx) + f32(xo)),
                    (c32[:, 2].astype(f32) * f32(vy) + f32(yo)),
                    (c32[:, 1].astype(f32) * f32(vz) + f32(zo))], axis=1).astype(np.float64)
    cluster = v[:, :, :3] - mean[:, None, :]
    centre = v[:, :, :3] - ctr[:, None, :]
    aliased = centre.copy()  # legacy in-place write: only the first `voxel_center_dims` raw channels are overwritten
    if voxel_center_dims < 3:
        aliased[:, :, 2] = v[:, :, 2]
    parts = [aliased, v[:, :, 3:], cluster, centre[:, :, :voxel_center_dims]]
    if with_distance:
        parts.append(np.linalg.norm(aliased, axis=2, keepdims=True))
    dec = np.concatenate(parts, axis=2)
    pid, tt = np.nonzero(slot)
    x = dec[pid, tt]                                                            # (N_k, D) real rows
    w_virt = (T - n).astype(np.float64)                                         # multiplicity of the virtual row
    xv = np.zeros((P, dec.shape[2]))                                            # v^(0) = 0
    M = float(P * T)
    stats = []
    L = len(weights)
    for l in range(L):
        W = np.asarray(weights[l], dtype=np.float64)
        y = x @ W.T
        yv = xv @ W.T
        if training:
            mu = (y.sum(0) + (w_virt[:, None] * yv).sum(0)) / M
            var = (((y - mu) ** 2).sum(0) + (w_virt[:, None] * (yv - mu) ** 2).sum(0)) / M
        else:
            mu = np.asarray(bn[l]["running_mean"], dtype=np.float64)
            var = np.asarray(bn[l]["running_var"], dtype=np.float64)
        stats.append((mu, var))
        g = np.asarray(bn[l]["weight"], dtype=np.float64)
        b = np.asarray(bn[l]["bias"], dtype=np.float64)
        a = np.maximum((y - mu) / np.sqrt(var + eps) * g + b, 0.0)
        av = np.maximum((yv - mu) / np.sqrt(var + eps) * g + b, 0.0)
        m = np.full((P, W.shape[0]), -np.inf)
        np.maximum.at(m, pid, a)
        m = np.where((w_virt > 0)[:, None], np.maximum(m, av), m)
        if l == L - 1:
            return m, stats
        x = np.concatenate([a, m[pid]], axis=1)
        xv = np.concatenate([av, m], axis=1)
    raise AssertionError


# --------------------------------------------------------------------------------------------------
# A.5 scatter — mmdet3d PointPillarsScatter.forward_batch (call site mask_bev_encoders.py:123)
# --------------------------------------------------------------------------------------------------
def scatter_np(feat: np.ndarray, coors4: np.ndarray, batch_size: int, ny: int, nx: int) -> np.ndarray:
    feat = np.asarray(feat)
    C = feat.shape[1]
    canvas = np.zeros((batch_size, C, ny * nx), dtype=feat.dtype)
    for b in range(batch_size):
        sel = coors4[:, 0] == b
        idx = coors4[sel, 2].astype(np.int64) * nx + coors4[sel, 3].astype(np.int64)
        canvas[b][:, idx] = feat[sel].T
    return canvas.reshape(batch_size, C, ny, nx)


def occupancy_np(coors4: np.ndarray, batch_size: int, ny: int, nx: int) -> np.ndarray:
    occ = np.zeros((batch_size, ny, nx), dtype=bool)
    occ[coors4[:, 0], coors4[:, 2], coors4[:, 3]] = True
    return occ


# --------------------------------------------------------------------------------------------------
# F2 — the first consumer of the pseudo image: mmdet PatchEmbed as the reference's Swin builds and calls it
# (/root/reference/mask_bev/models/networks/swin/swin.py:13 import, :578-586 construction, :745-746 call).
# mmdet==3.x is not on disk (Dockerfile pins it next to mmcv / mmdet3d): PARITY UNPINNED — this restates the published
# mmdet.models.layers.PatchEmbed.forward with padding='corner' (AdaptivePadding: zeros at the bottom / right so that
# H, W become multiples of the stride), Conv2d(kernel = stride = patch, bias), flatten(2).transpose(1, 2), then the
# optional LayerNorm(E). Dense torch CPU ops, exactly the sequence upstream executes.
# --------------------------------------------------------------------------------------------------
def patch_embed_dense(x, conv_weight, conv_bias, patch: int, norm_weight=None, norm_bias=None, norm_eps: float = 1e-5):
    """x (B, C, H, W) torch CPU tensor (the LayerNorm-ed pseudo image) -> (tokens (B, Hp*Wp, E), (Hp, Wp))."""
    torch = _torch()
    import math
    import torch.nn.functional as F
    H, W = x.shape[-2:]
    pad_h = max((math.ceil(H / patch) - 1) * patch + (patch - 1) + 1 - H, 0)
    pad_w = max((math.ceil(W / patch) - 1) * patch + (patch - 1) + 1 - W, 0)
    if pad_h > 0 or pad_w > 0:
        x = F.pad(x, [0, pad_w, 0, pad_h])
    y = F.conv2d(x, conv_weight, conv_bias, stride=patch)
    out_size = (y.shape[2], y.shape[3])
    y = y.flatten(2).transpose(1, 2)
    if norm_weight is not None:
        y = F.layer_norm(y, (y.shape[-1],), norm_weight, norm_bias, norm_eps)
    return y, out_size


def patch_embed_on_pillars_np(feat, coors4, batch_size, ny, nx, ln_weight, ln_bias, ln_eps, conv_weight, conv_bias, patch,
                              norm_weight=None, norm_bias=None, norm_eps: float = 1e-5):
    """The same tokens from the pillars alone, in float64 numpy — the algebra csrc/patch_embed.cu implements (LayerNorm
    pushed through the convolution), used to check the identity against ``patch_embed_dense`` on CPU."""
    import math
    f = np.asarray(feat, dtype=np.float64)
    C = f.shape[1]
    lw, lb = np.asarray(ln_weight, np.float64), np.asarray(ln_bias, np.float64)
    W = np.asarray(conv_weight, np.float64)
    E = W.shape[0]
    Hp, Wp = math.ceil(ny / patch), math.ceil(nx / patch)
    pad = lambda a: np.pad(a, ((0, 0), (0, Hp * patch - ny), (0, Wp * patch - nx)))  # noqa: E731
    conv = lambda a: np.einsum('cyixj,ecij->yxe', pad(a).reshape(C, Hp, patch, Wp, patch), W)  # noqa: E731
    p0 = conv(lb) + (0.0 if conv_bias is None else np.asarray(conv_bias, np.float64))
    p1 = conv(lw)
    out = np.zeros((batch_size, Hp, Wp, E))
    M = C * ny * nx
    for b in range(batch_size):
        sel = coors4[:, 0] == b
        fb, yb, xb = f[sel], coors4[sel, 2].astype(np.int64), coors4[sel, 3].astype(np.int64)
        mu = fb.sum() / M
        var = max((fb * fb).sum() / M - mu * mu, 0.0)
        rstd = 1.0 / np.sqrt(var + ln_eps)
        sparse = np.zeros((Hp, Wp, E))
        g = fb * lw[:, yb, xb].T
        z = np.einsum('pc,pec->pe', g, W[:, :, yb % patch, xb % patch].transpose(2, 0, 1))
        np.add.at(sparse, (yb // patch, xb // patch), z)
        out[b] = p0 - mu * rstd * p1 + rstd * sparse
    y = out.reshape(batch_size, Hp * Wp, E)
    if norm_weight is not None:
        m = y.mean(-1, keepdims=True)
        v = ((y - m) ** 2).mean(-1, keepdims=True)
        y = (y - m) / np.sqrt(v + norm_eps) * np.asarray(norm_weight, np.float64) + np.asarray(norm_bias, np.float64)
    return y, (Hp, Wp)


# --------------------------------------------------------------------------------------------------
# F4 — point side of the reference's augmentations (mask_bev/augmentations/semantic_kitti_mask_augmentations.py),
# in the order train_mask_bev.py:71 composes the configured list: RandomDropPoints (:152-162), Flip (:44-56),
# RandomRotate (:73-101), JitterPoints (:116-149). PINNED to executed reference code: tests/golden/augment_reference.npz
# holds outputs of the reference's own classes (make_golden_augment.py); this restatement, driven by the decisions the
# product's sampler replays from the same numpy seed, reproduces them bit for bit (tests/test_augment.py).
# --------------------------------------------------------------------------------------------------
def augment_points_np(points: np.ndarray, keep=None, flip_x=False, flip_y=False, theta_deg=None, noise=None):
    """points (N, 4) float32; keep: bool (N,) or None; noise: (N, 4) float64 in ORIGINAL row indexing or None.
    Returns the augmented cloud with the dropped rows removed, as the reference leaves it."""
    pc = np.array(points, dtype=np.float32, copy=True)
    if keep is not None:
        pc = pc[keep]                                           # :160
        if noise is not None:
            noise = noise[keep]
    if flip_x:
        pc[:, 0] = -pc[:, 0]                                    # :51
    if flip_y:
        pc[:, 1] = -pc[:, 1]                                    # :54
    if theta_deg is not None:                                   # :88-97: float64 R @ (x, y, z, 1), stored into float32
        c, s = np.cos(np.deg2rad(theta_deg)), np.sin(np.deg2rad(theta_deg))
        x, y = pc[:, 0].astype(np.float64), pc[:, 1].astype(np.float64)
        pc[:, 0] = (c * x + (-s) * y) + 0.0                     # the matrix row's 0 * z + 0 * 1 terms: -0 becomes +0
        pc[:, 1] = (s * x + c * y) + 0.0
        pc[:, 2] = pc[:, 2].astype(np.float64) + 0.0
    if noise is not None:
        pc += noise                                             # :145 (float64 sum, stored into float32)
        np.clip(pc[:, 3], 0, 1, pc[:, 3])                       # :146
    return pc


# --------------------------------------------------------------------------------------------------
# The whole path — mirrors MaskBevEncoder (mask_bev_encoders.py:21-123) on CPU
# --------------------------------------------------------------------------------------------------
class MaskBevEncoderOracle:
    """CPU restatement of ``MaskBevEncoder`` minus the Fourier branch (never configured) and, by default,
    minus the trailing LayerNorm (row f1 of SURVEY.md §8; enable with ``layer_norm=True``)."""

    def __init__(self, feat_channels, x_range, y_range, z_range, voxel_size_x, voxel_size_y, voxel_size_z,
                 max_num_points, max_voxels=500 * 500, pc_point_dim=4, with_distance=True, dtype=None,
                 voxelizer: str = "c", layer_norm: bool = False):
        torch = _torch()
        self.geo = encoder_geometry(x_range, y_range, z_range, voxel_size_x, voxel_size_y, voxel_size_z)
        self.x_range, self.y_range, self.z_range = x_range, y_range, z_range
        self.T = max_num_points
        self.max_voxels = max_voxels
        self.C = pc_point_dim
        self.C_out = feat_channels[-1]
        self.dtype = dtype or torch.float32
        self.pfn = make_pfn_oracle(in_channels=pc_point_dim, feat_channels=feat_channels,
                                   with_distance=with_distance, voxel_size=self.geo["voxel_size"],
                                   point_cloud_range=self.geo["point_cloud_range"], dtype=self.dtype)
        self._vox = dict(c=hard_voxelize_c, np=hard_voxelize_np, py=hard_voxelize_py)[voxelizer]
        self.layer_norm = (torch.nn.LayerNorm([self.C_out, self.geo["ny"], self.geo["nx"]], eps=1e-3).to(self.dtype)
                           if layer_norm else None)

    def voxelize(self, point_clouds):
        """mask_bev_encoders.py:95-111. Returns voxels (P,T,C), num_points (P,), coors (P,4)=(b,z,y,x),
        kept_idx (P,T) = row in the UNFILTERED frame of every stored point (-1 padded)."""
        vs, cs, ns, ks = [], [], [], []
        for b, pc in enumerate(point_clouds):
            pc = np.asarray(pc, dtype=np.float32)
            f, src = filter_in_range(pc, self.x_range, self.y_range, self.z_range)
            v, c, n, k = self._vox(f, self.geo["voxel_size"], self.geo["point_cloud_range"], self.T,
                                   self.max_voxels)
            k = np.where(k >= 0, src[np.clip(k, 0, None)] if len(src) else k, -1)
            vs.append(v)
            ns.append(n)
            ks.append(k)
            cs.append(np.concatenate([np.full((len(c), 1), b, dtype=np.int32), c], axis=1))
        return (np.concatenate(vs, 0), np.concatenate(ns, 0), np.concatenate(cs, 0).astype(np.int32),
                np.concatenate(ks, 0))

    def encode(self, voxels, num_points, coors):
        torch = _torch()
        return self.pfn(torch.from_numpy(np.ascontiguousarray(voxels)).to(self.dtype),
                        torch.from_numpy(np.ascontiguousarray(num_points)),
                        torch.from_numpy(np.ascontiguousarray(coors)))

    def middle_encode(self, feats, coors, batch_size):
        torch = _torch()
        canvas = scatter_np(feats.detach().numpy(), coors, batch_size, self.geo["ny"], self.geo["nx"])
        return torch.from_numpy(canvas)

    def forward(self, point_clouds):
        voxels, num_points, coors, _ = self.voxelize(point_clouds)
        feats = self.encode(voxels, num_points, coors)
        img = self.middle_encode(feats, coors, len(point_clouds))
        if self.layer_norm is not None:
            img = self.layer_norm(img)
        return img

    def forward_autograd(self, point_clouds):
        """The same forward with the scatter written in torch (``canvas[b, :, y, x] = feats[p]``, upstream's
        ``canvas[:, indices] = voxels.t()`` per frame), so that autograd reaches the PFN and LayerNorm parameters — the
        reference trains through exactly this graph. Returns the same values as ``forward``."""
        torch = _torch()
        voxels, num_points, coors, _ = self.voxelize(point_clouds)
        feats = self.encode(voxels, num_points, coors)
        c = torch.from_numpy(np.ascontiguousarray(coors)).long()
        img = torch.zeros((len(point_clouds), self.C_out, self.geo["ny"], self.geo["nx"]), dtype=feats.dtype)
        img[c[:, 0], :, c[:, 2], c[:, 3]] = feats
        if self.layer_norm is not None:
            img = self.layer_norm(img)
        return img
