"""Thin functional layer over the C ABI: torch tensors in, torch tensors out, everything enqueued on the
current CUDA stream. PyTorch is used for device memory and streams only; all arithmetic of the path runs
in libmask_bev_b200.so. CPU tensors are rejected (no CPU fallback).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import MAX_LAYERS, MAX_UNITS, MbevGeometry, MbevPfnParams, check, ptr, ptr_array


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.MbevError(f"{name} is on {t.device}: mask_bev_b200 has no CPU path — move the input to a CUDA "
                             f"device (the reference's CPU use is covered by the oracle in tests only)")


# ------------------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------------------
def make_geometry(voxel_size: Sequence[float], point_cloud_range: Sequence[float], max_num_points: int,
                  max_voxels: int, num_feats: int, strict_filter: bool) -> MbevGeometry:
    """float32 views of the voxel layer's arguments (mmcv hands `torch.tensor(voxel_size)` to the op);
    grid = round((hi - lo) / vs) in float32 (mmcv/ops/voxelize.py Voxelization.__init__)."""
    g = MbevGeometry()
    r = np.asarray(point_cloud_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    grid = np.round((r[3:] - r[:3]) / v).astype(np.int64)
    for j in range(6):
        g.range[j] = float(r[j])
    for j in range(3):
        g.voxel[j] = float(v[j])
        g.grid[j] = int(grid[j])
    g.max_points = int(max_num_points)
    g.max_voxels = int(max_voxels)
    g.num_feats = int(num_feats)
    g.strict_filter = 1 if strict_filter else 0
    return g


def _offsets(frame_sizes: Sequence[int]):
    n = len(frame_sizes)
    arr = (ctypes.c_int64 * (n + 1))()
    acc = 0
    for i, s in enumerate(frame_sizes):
        arr[i] = acc
        acc += int(s)
    arr[n] = acc
    return arr, acc


def pillar_capacity(geo: MbevGeometry, frame_sizes: Sequence[int]) -> int:
    """Rows every per-pillar buffer of a batch needs: sum over frames of min(points, max_voxels, cells) — the one
    bound the library checks (mbev_pillar_capacity), at least 1 so that empty batches still own valid pointers."""
    off, _ = _offsets(frame_sizes)
    cap = int(_lib.load().mbev_pillar_capacity(off, len(frame_sizes), ctypes.byref(geo)))
    if cap < 0:
        raise _lib.MbevError("pillar_capacity: bad geometry or frame sizes")
    return max(1, cap)


@dataclass
class VoxelBatch:
    """Device-resident result of K1. `pillar_base[-1]` is the total pillar count (still on the device)."""
    cell_table: torch.Tensor   # (B, nz*ny*nx) int32, pillar id or -1  == occupancy / inverse map
    coors: torch.Tensor        # (cap, 4) int32 (b, z, y, x)
    num_points: torch.Tensor   # (cap,) int32
    kept_idx: torch.Tensor     # (cap, T) int32 rows into the concatenated point tensor
    pillar_base: torch.Tensor  # (B+1,) int32
    capacity: int
    batch: int
    points: Optional[torch.Tensor] = None  # the augmented cloud when K1 ran with augmentations (kept_idx indexes it)

    @property
    def num_pillars_dev(self) -> torch.Tensor:
        return self.pillar_base[self.batch:]

    def occupancy(self) -> torch.Tensor:
        return self.cell_table >= 0


def voxelize_batch(points: torch.Tensor, frame_sizes: Sequence[int], geo: MbevGeometry,
                   capacity: Optional[int] = None, augment=None) -> VoxelBatch:
    """K1 over a batch of concatenated frames. No host synchronisation.
    augment: an ``augment.BatchAugment`` — the augmentations run in K1's load stage and the returned
    batch carries ``points`` = the augmented cloud, which is what ``kept_idx`` indexes (the PFN must gather from it)."""
    _need_cuda(points, "points")
    lib = _lib.load()
    if points.dtype != torch.float32:
        raise _lib.MbevError(f"points must be float32, got {points.dtype}")
    points = points.contiguous()
    B = len(frame_sizes)
    off, total = _offsets(frame_sizes)
    if points.shape[0] != total or (points.dim() != 2) or points.shape[1] != geo.num_feats:
        raise _lib.MbevError(f"points shape {tuple(points.shape)} does not match frame sizes (sum {total}) x C={geo.num_feats}")
    cap = pillar_capacity(geo, frame_sizes) if capacity is None else int(capacity)
    dev = points.device
    cells = geo.grid[0] * geo.grid[1] * geo.grid[2]
    nbytes = ctypes.c_size_t()
    check(lib.mbev_voxelize_workspace_bytes(ctypes.byref(geo), B, total, ctypes.byref(nbytes)), "voxelize_workspace_bytes")
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
    out = VoxelBatch(
        cell_table=torch.empty((B, cells), dtype=torch.int32, device=dev),
        coors=torch.empty((cap, 4), dtype=torch.int32, device=dev),
        num_points=torch.empty((cap,), dtype=torch.int32, device=dev),
        kept_idx=torch.empty((cap, geo.max_points), dtype=torch.int32, device=dev),
        pillar_base=torch.empty((B + 1,), dtype=torch.int32, device=dev),
        capacity=cap, batch=B)
    with torch.cuda.device(dev):
        if augment is None:
            check(lib.mbev_voxelize(ptr(points), off, B, ctypes.byref(geo), ptr(out.cell_table), ptr(out.coors),
                                    ptr(out.num_points), ptr(out.kept_idx), ptr(out.pillar_base), cap, ptr(ws),
                                    ws.numel(), _stream()), "voxelize")
        else:
            aug, keep_alive = augment.to_struct(dev, B, total, points.shape[1])
            out.points = torch.empty_like(points)
            aug.points_out = out.points.data_ptr()
            check(lib.mbev_voxelize_augmented(ptr(points), off, B, ctypes.byref(geo), ctypes.byref(aug),
                                              ptr(out.cell_table), ptr(out.coors), ptr(out.num_points),
                                              ptr(out.kept_idx), ptr(out.pillar_base), cap, ptr(ws), ws.numel(),
                                              _stream()), "voxelize_augmented")
            out._keep_alive = keep_alive
    return out


def gather_voxels(points: torch.Tensor, vb: VoxelBatch, num_pillars: int, T: int) -> torch.Tensor:
    """Zero-padded (P, T, C) voxel tensor, as mmcv's op returns it."""
    lib = _lib.load()
    C = points.shape[1]
    voxels = torch.empty((num_pillars, T, C), dtype=torch.float32, device=points.device)
    if num_pillars == 0:
        return voxels
    with torch.cuda.device(points.device):
        check(lib.mbev_gather_voxels(ptr(points), ptr(vb.kept_idx), ptr(vb.num_points), ptr(vb.num_pillars_dev),
                                     num_pillars, T, C, ptr(voxels), _stream()), "gather_voxels")
    return voxels


# ------------------------------------------------------------------------------------------------------
# PFN
# ------------------------------------------------------------------------------------------------------
@dataclass
class PfnConfig:
    in_channels: int          # raw point features C
    units: List[int]          # PFNLayer.units per layer
    in_dims: List[int]        # Linear.in_features per layer
    with_cluster_center: bool
    with_voxel_center: bool
    with_distance: bool
    legacy: bool
    voxel_center_dims: int
    vx: float
    vy: float
    vz: float
    x_offset: float
    y_offset: float
    z_offset: float
    eps: float = 1e-3
    gemm_path: int = 0        # _lib.GEMM_AUTO / GEMM_FMA / GEMM_TCGEN05 / GEMM_TCGEN05_BF16 (forward Linear layers)


def _pfn_struct(cfg: PfnConfig, weights, scales, shifts) -> MbevPfnParams:
    p = MbevPfnParams()
    L = len(cfg.units)
    if L > MAX_LAYERS:
        raise _lib.MbevError(f"{L} PFN layers > {MAX_LAYERS}")
    p.num_layers = L
    for l in range(L):
        p.in_dim[l] = cfg.in_dims[l]
        p.units[l] = cfg.units[l]
        p.weight[l] = weights[l].data_ptr() if weights[l] is not None else None
        p.scale[l] = scales[l].data_ptr() if scales is not None else None
        p.shift[l] = shifts[l].data_ptr() if shifts is not None else None
    p.with_cluster_center = int(cfg.with_cluster_center)
    p.with_voxel_center = int(cfg.with_voxel_center)
    p.with_distance = int(cfg.with_distance)
    p.legacy = int(cfg.legacy)
    p.voxel_center_dims = int(cfg.voxel_center_dims)
    p.vx, p.vy, p.vz = cfg.vx, cfg.vy, cfg.vz
    p.x_offset, p.y_offset, p.z_offset = cfg.x_offset, cfg.y_offset, cfg.z_offset
    p.gemm_path = int(cfg.gemm_path)
    return p


def pfn_path(cfg: PfnConfig, T: int) -> str:
    """Which device implementation a forward with this stack takes: 'tcgen05', 'tcgen05_bf16' or 'fma'."""
    lib = _lib.load()
    params = _pfn_struct(cfg, [None] * len(cfg.units), None, None)
    r = lib.mbev_pfn_path(ctypes.byref(params), int(T))
    if r < 0:
        check(r, "pfn_path")
    return {_lib.GEMM_FMA: "fma", _lib.GEMM_TCGEN05: "tcgen05", _lib.GEMM_TCGEN05_BF16: "tcgen05_bf16"}[r]


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def pfn_forward_eval(rows, kept_idx, num_points, coors, num_pillars_dev, capacity, T, cfg: PfnConfig,
                     weights, scales, shifts, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(rows, "features")
    dev = rows.device
    weights = [_f32c(w) for w in weights]
    scales = [_f32c(s) for s in scales]
    shifts = [_f32c(s) for s in shifts]
    params = _pfn_struct(cfg, weights, scales, shifts)
    nbytes = ctypes.c_size_t()
    check(lib.mbev_pfn_workspace_bytes(ctypes.byref(params), T, capacity, 0, ctypes.byref(nbytes)), "pfn_workspace_bytes")
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
    feats = out if out is not None else torch.empty((capacity, cfg.units[-1]), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.mbev_pfn_forward(ptr(rows), cfg.in_channels, ptr(kept_idx), ptr(num_points), ptr(coors),
                                   ptr(num_pillars_dev), capacity, T, ctypes.byref(params), ptr(feats), ptr(ws),
                                   ws.numel(), _stream()), "pfn_forward")
    return feats


def pfn_forward_train(rows, kept_idx, num_points, coors, num_pillars_dev, capacity, T, cfg: PfnConfig,
                      weights, gammas, betas):
    """Returns (feats, scale_shift (L,2,MAX_UNITS), batch_stats (L,2,MAX_UNITS) = mean / biased var)."""
    lib = _lib.load()
    _need_cuda(rows, "features")
    dev = rows.device
    L = len(cfg.units)
    weights = [_f32c(w) for w in weights]
    gammas = [_f32c(g) for g in gammas]
    betas = [_f32c(b) for b in betas]
    params = _pfn_struct(cfg, weights, None, None)
    nbytes = ctypes.c_size_t()
    check(lib.mbev_pfn_workspace_bytes(ctypes.byref(params), T, capacity, 1, ctypes.byref(nbytes)), "pfn_workspace_bytes")
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
    feats = torch.empty((capacity, cfg.units[-1]), dtype=torch.float32, device=dev)
    scale_shift = torch.zeros((L, 2, MAX_UNITS), dtype=torch.float32, device=dev)
    batch_stats = torch.zeros((L, 2, MAX_UNITS), dtype=torch.float32, device=dev)
    g_arr, b_arr = ptr_array(gammas), ptr_array(betas)
    with torch.cuda.device(dev):
        check(lib.mbev_pfn_forward_train(ptr(rows), cfg.in_channels, ptr(kept_idx), ptr(num_points), ptr(coors),
                                         ptr(num_pillars_dev), capacity, T, ctypes.byref(params),
                                         ctypes.byref(g_arr), ctypes.byref(b_arr), cfg.eps, ptr(feats),
                                         ptr(scale_shift), ptr(batch_stats), ptr(ws), ws.numel(), _stream()),
              "pfn_forward_train")
    return feats, scale_shift, batch_stats


def pfn_backward(rows, kept_idx, num_points, coors, num_pillars_dev, capacity, T, cfg: PfnConfig, weights,
                 scale_shift, batch_stats, train: bool, dfeats, rows_capacity: int = 0):
    """Parameter gradients of the PFN: lists (dweight[l], dgamma[l], dbeta[l]).
    rows_capacity: host-side bound of the compact row count (real rows + one virtual row per padded pillar)."""
    lib = _lib.load()
    dev = rows.device
    L = len(cfg.units)
    weights = [_f32c(w) for w in weights]
    params = _pfn_struct(cfg, weights, None, None)
    nbytes = ctypes.c_size_t()
    check(lib.mbev_pfn_backward_workspace_bytes(ctypes.byref(params), T, capacity, rows_capacity,
                                                ctypes.byref(nbytes)), "pfn_backward_workspace_bytes")
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
    dws = [torch.empty((cfg.units[l], cfg.in_dims[l]), dtype=torch.float32, device=dev) for l in range(L)]
    dgs = [torch.empty((cfg.units[l],), dtype=torch.float32, device=dev) for l in range(L)]
    dbs = [torch.empty((cfg.units[l],), dtype=torch.float32, device=dev) for l in range(L)]
    dfeats = _f32c(dfeats)
    dw_arr, dg_arr, db_arr = ptr_array(dws), ptr_array(dgs), ptr_array(dbs)
    with torch.cuda.device(dev):
        check(lib.mbev_pfn_backward(ptr(rows), cfg.in_channels, ptr(kept_idx), ptr(num_points), ptr(coors),
                                    ptr(num_pillars_dev), capacity, T, rows_capacity, ctypes.byref(params),
                                    ptr(scale_shift), ptr(batch_stats), cfg.eps, int(train), ptr(dfeats),
                                    ctypes.byref(dw_arr), ctypes.byref(dg_arr), ctypes.byref(db_arr), ptr(ws),
                                    ws.numel(), _stream()), "pfn_backward")
    return dws, dgs, dbs


def pfn_forward_train_rows(rows, kept_idx, num_points, coors, num_pillars_dev, capacity, T, cfg: PfnConfig,
                           weights, gammas, betas, rows_capacity: int = 0):
    """Train-mode forward of a training step in compact row space (mbev_pfn_forward_train_rows): returns
    (feats, scale_shift, batch_stats, saved) — `saved` is the workspace holding every layer's rows, which
    `pfn_backward_rows` consumes instead of recomputing the forward."""
    lib = _lib.load()
    _need_cuda(rows, "features")
    dev = rows.device
    L = len(cfg.units)
    weights = [_f32c(w) for w in weights]
    gammas = [_f32c(g) for g in gammas]
    betas = [_f32c(b) for b in betas]
    params = _pfn_struct(cfg, weights, None, None)
    nbytes = ctypes.c_size_t()
    check(lib.mbev_pfn_backward_workspace_bytes(ctypes.byref(params), T, capacity, rows_capacity,
                                                ctypes.byref(nbytes)), "pfn_backward_workspace_bytes")
    saved = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
    feats = torch.empty((capacity, cfg.units[-1]), dtype=torch.float32, device=dev)
    scale_shift = torch.zeros((L, 2, MAX_UNITS), dtype=torch.float32, device=dev)
    batch_stats = torch.zeros((L, 2, MAX_UNITS), dtype=torch.float32, device=dev)
    g_arr, b_arr = ptr_array(gammas), ptr_array(betas)
    with torch.cuda.device(dev):
        check(lib.mbev_pfn_forward_train_rows(ptr(rows), cfg.in_channels, ptr(kept_idx), ptr(num_points), ptr(coors),
                                              ptr(num_pillars_dev), capacity, T, rows_capacity, ctypes.byref(params),
                                              ctypes.byref(g_arr), ctypes.byref(b_arr), cfg.eps, ptr(feats),
                                              ptr(scale_shift), ptr(batch_stats), ptr(saved), saved.numel(),
                                              _stream()), "pfn_forward_train_rows")
    return feats, scale_shift, batch_stats, saved


def pfn_backward_rows(saved, num_pillars_dev, capacity, T, cfg: PfnConfig, weights, dfeats, rows_capacity: int = 0):
    """Parameter gradients from the rows `pfn_forward_train_rows` kept (same capacity / T / rows_capacity / cfg)."""
    lib = _lib.load()
    dev = saved.device
    L = len(cfg.units)
    weights = [_f32c(w) for w in weights]
    params = _pfn_struct(cfg, weights, None, None)
    dws = [torch.empty((cfg.units[l], cfg.in_dims[l]), dtype=torch.float32, device=dev) for l in range(L)]
    dgs = [torch.empty((cfg.units[l],), dtype=torch.float32, device=dev) for l in range(L)]
    dbs = [torch.empty((cfg.units[l],), dtype=torch.float32, device=dev) for l in range(L)]
    dfeats = _f32c(dfeats)
    dw_arr, dg_arr, db_arr = ptr_array(dws), ptr_array(dgs), ptr_array(dbs)
    with torch.cuda.device(dev):
        check(lib.mbev_pfn_backward_rows(ptr(num_pillars_dev), capacity, cfg.in_channels, T, rows_capacity,
                                         ctypes.byref(params), cfg.eps, ptr(dfeats), ctypes.byref(dw_arr),
                                         ctypes.byref(dg_arr), ctypes.byref(db_arr), ptr(saved), saved.numel(),
                                         _stream()), "pfn_backward_rows")
    return dws, dgs, dbs


# ------------------------------------------------------------------------------------------------------
# scatter
# ------------------------------------------------------------------------------------------------------
def build_cell_table(coors: torch.Tensor, num_pillars_dev: torch.Tensor, capacity: int, batch: int, ny: int,
                     nx: int) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(coors, "coors")
    table = torch.empty((batch, ny * nx), dtype=torch.int32, device=coors.device)
    with torch.cuda.device(coors.device):
        check(lib.mbev_build_cell_table(ptr(coors), ptr(num_pillars_dev), capacity, batch, ny, nx, ptr(table),
                                        _stream()), "build_cell_table")
    return table


def scatter_forward(feats: torch.Tensor, cell_table: torch.Tensor, batch: int, ny: int, nx: int,
                    out: Optional[torch.Tensor] = None, dtype: torch.dtype = torch.float32,
                    stream_ctas_per_sm: int = 0, channels_last: bool = False) -> torch.Tensor:
    """K3. dtype=torch.bfloat16 writes a bf16 canvas (= the fp32 canvas rounded to nearest-even; forward only).
    stream_ctas_per_sm > 0: the TMA-engine form (mbev_scatter_forward_stream), same bits.
    channels_last: the (B, C, ny, nx) result is laid out (B, ny, nx, C) in memory (torch.channels_last)."""
    lib = _lib.load()
    _need_cuda(feats, "voxel_features")
    C = feats.shape[1]
    if dtype not in (torch.float32, torch.bfloat16):
        raise _lib.MbevError(f"canvas dtype {dtype} is not supported (float32 or bfloat16)")
    mf = torch.channels_last if channels_last else torch.contiguous_format
    canvas = out if out is not None else torch.empty((batch, C, ny, nx), dtype=dtype, device=feats.device,
                                                     memory_format=mf)
    if canvas.dtype != dtype or tuple(canvas.shape) != (batch, C, ny, nx) or not canvas.is_contiguous(memory_format=mf):
        raise _lib.MbevError("out must be a (B, C, ny, nx) tensor of the requested dtype and memory format")
    with torch.cuda.device(feats.device):
        if channels_last:
            if dtype != torch.float32:
                raise _lib.MbevError("the channels-last canvas is float32 only")
            check(lib.mbev_scatter_forward_nhwc(ptr(feats), ptr(cell_table), batch, C, ny, nx, ptr(canvas), _stream()),
                  "scatter_forward_nhwc")
        elif dtype == torch.bfloat16:
            check(lib.mbev_scatter_forward_bf16(ptr(feats), ptr(cell_table), batch, C, ny, nx, ptr(canvas), _stream()),
                  "scatter_forward_bf16")
        elif stream_ctas_per_sm > 0:
            check(lib.mbev_scatter_forward_stream(ptr(feats), ptr(cell_table), batch, C, ny, nx, ptr(canvas),
                                                  int(stream_ctas_per_sm), _stream()), "scatter_forward_stream")
        else:
            check(lib.mbev_scatter_forward(ptr(feats), ptr(cell_table), batch, C, ny, nx, ptr(canvas), _stream()),
                  "scatter_forward")
    return canvas


def scatter_layernorm_forward(feats: torch.Tensor, cell_table: torch.Tensor, pillar_base: torch.Tensor, batch: int,
                              ny: int, nx: int, weight: torch.Tensor, bias: torch.Tensor, eps: float,
                              out: Optional[torch.Tensor] = None, walk: Optional[int] = None):
    """K3 + LayerNorm([C, ny, nx]) in one pass (forward only). Returns (out (B, C, ny, nx), stats (B, 2) = mean,
    rstd), or None when the shapes / alignment do not fit the fused kernel (the caller then runs K3 + torch LN).
    walk: _lib.LN_WALK_RUNS / LN_WALK_FRAMES (None: LN_WALK_DEFAULT where it applies)."""
    lib = _lib.load()
    _need_cuda(feats, "voxel_features")
    dev = feats.device
    C = feats.shape[1]
    feats = _f32c(feats)
    weight, bias = _f32c(weight), _f32c(bias)
    if tuple(weight.shape) != (C, ny, nx) or tuple(bias.shape) != (C, ny, nx):
        raise _lib.MbevError(f"LayerNorm weight/bias must be ({C}, {ny}, {nx})")
    out = out if out is not None else torch.empty((batch, C, ny, nx), dtype=torch.float32, device=dev)
    if not lib.mbev_scatter_layernorm_supported(batch, C, ny, nx, ptr(out), ptr(weight), ptr(bias)):
        return None
    stats = torch.empty((batch, 2), dtype=torch.float32, device=dev)
    if walk is None:
        walk = LN_WALK_DEFAULT
        if walk == _lib.LN_WALK_FRAMES and (C % 4 or feats.data_ptr() % 16):
            walk = _lib.LN_WALK_RUNS
    nbytes = ctypes.c_size_t()
    check(lib.mbev_scatter_layernorm_workspace_bytes(batch, ctypes.byref(nbytes)), "scatter_layernorm_workspace_bytes")
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.mbev_scatter_layernorm_forward(ptr(feats), ptr(cell_table), ptr(pillar_base), batch, C, ny, nx,
                                                 ptr(weight), ptr(bias), float(eps), int(walk), ptr(out), ptr(stats),
                                                 ptr(ws), ws.numel(), _stream()), "scatter_layernorm_forward")
    return out, stats


# schedule of the K3+LN forward's streaming pass when the caller does not choose. B200, kitti_b16: frames 1.43 ms,
# runs 1.62 ms (profiles/r2a_bench.json) — the frame walk keeps weight / bias in registers instead of re-reading L2.
LN_WALK_DEFAULT = _lib.LN_WALK_FRAMES


def _aligned16(t: torch.Tensor) -> torch.Tensor:
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def scatter_layernorm_backward_supported(batch: int, C: int, ny: int, nx: int) -> bool:
    return bool(_lib.load().mbev_scatter_layernorm_backward_supported(batch, C, ny, nx))


def scatter_layernorm_backward(dout: torch.Tensor, feats: torch.Tensor, cell_table: torch.Tensor, coors: torch.Tensor,
                               num_pillars_dev: torch.Tensor, weight: torch.Tensor, stats: torch.Tensor):
    """Backward of scatter_layernorm_forward: (dfeats (rows, C), dweight (C, ny, nx), dbias (C, ny, nx)).
    `stats` is the forward's (B, 2) mean / rstd; `feats` the forward's input rows."""
    lib = _lib.load()
    _need_cuda(dout, "grad_output")
    dev = dout.device
    B, C, ny, nx = dout.shape
    dout, feats, weight = _aligned16(_f32c(dout)), _aligned16(_f32c(feats)), _aligned16(_f32c(weight))
    rows = feats.shape[0]
    nbytes = ctypes.c_size_t()
    check(lib.mbev_scatter_layernorm_backward_workspace_bytes(B, C, ny, nx, ctypes.byref(nbytes)),
          "scatter_layernorm_backward_workspace_bytes")
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
    dfeats = torch.empty((rows, C), dtype=torch.float32, device=dev)
    dweight = torch.empty((C, ny, nx), dtype=torch.float32, device=dev)
    dbias = torch.empty((C, ny, nx), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.mbev_scatter_layernorm_backward(ptr(dout), ptr(feats), ptr(cell_table), ptr(coors),
                                                  ptr(num_pillars_dev), rows, B, C, ny, nx, ptr(weight),
                                                  ptr(stats), ptr(dfeats), ptr(dweight), ptr(dbias), ptr(ws),
                                                  ws.numel(), _stream()), "scatter_layernorm_backward")
    return dfeats, dweight, dbias


def scatter_backward(dcanvas: torch.Tensor, cell_table: torch.Tensor, num_rows: int) -> torch.Tensor:
    lib = _lib.load()
    B, C, ny, nx = dcanvas.shape
    dcanvas = _f32c(dcanvas)
    # rows of pillars that never reached the canvas (none in practice) get zero gradient
    dfeats = torch.zeros((num_rows, C), dtype=torch.float32, device=dcanvas.device)
    with torch.cuda.device(dcanvas.device):
        check(lib.mbev_scatter_backward(ptr(dcanvas), ptr(cell_table), B, C, ny, nx, ptr(dfeats), _stream()),
              "scatter_backward")
    return dfeats


def scatter_backward_nhwc(dcanvas: torch.Tensor, cell_table: torch.Tensor, coors: torch.Tensor,
                          num_pillars_dev: torch.Tensor, num_rows: int) -> torch.Tensor:
    """K3' for a channels-last gradient: dcanvas is (B, C, ny, nx) in torch.channels_last memory format."""
    lib = _lib.load()
    B, C, ny, nx = dcanvas.shape
    dcanvas = dcanvas.detach().to(torch.float32).contiguous(memory_format=torch.channels_last)
    dfeats = torch.empty((num_rows, C), dtype=torch.float32, device=dcanvas.device)
    with torch.cuda.device(dcanvas.device):
        check(lib.mbev_scatter_backward_nhwc(ptr(dcanvas), ptr(cell_table), ptr(coors), ptr(num_pillars_dev), num_rows,
                                             B, C, ny, nx, ptr(dfeats), _stream()), "scatter_backward_nhwc")
    return dfeats
