"""`MaskBevEncoder` — host-side mirror of the reference's encoder
(/root/reference/mask_bev/models/encoders/mask_bev_encoders.py:21-123): same constructor arguments, same
sub-module attribute names (``_voxel_layer``, ``_voxel_encoder``, ``_middle_encoder``, ``_layer_norm``) and hence
the same state-dict keys, same ``forward`` / ``voxelize`` / ``encode`` / ``middle_encode`` methods.

``forward`` runs the fused batch path (K1 -> K2 -> K3 on the current stream, no per-frame Python loop, no host
synchronisation) instead of the reference's per-frame loop; ``voxelize`` / ``encode`` / ``middle_encode`` keep the
module-level semantics for callers that use them separately (the reference's tests do).

The reference file itself also imports unchanged against this package through the shims in
``mask_bev_b200/shims`` (``mmcv.ops`` / ``mmdet3d.models``) — see INTEGRATION.md.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Union

import torch
from torch import nn

from . import functional as F_
from ._lib import MbevError
from .collate import PackedFrames
from .pillar_encoder import PillarFeatureNet
from .scatter import PointPillarsScatter, scatter_layernorm_with_table, scatter_with_table
from .voxelize import Voxelization


def _as_points(point_clouds):
    """(points (sum N_i, C) contiguous, frame sizes) of a batch given as the reference's list of per-frame tensors
    (mask_bev_encoders.py:77-81) or as a `PackedFrames` (already one tensor: no concatenation)."""
    if isinstance(point_clouds, PackedFrames):
        return point_clouds.points.contiguous(), [int(s) for s in point_clouds.sizes]
    sizes = [int(pc.shape[0]) for pc in point_clouds]
    pts = point_clouds[0] if len(point_clouds) == 1 else torch.cat(point_clouds, dim=0)
    return pts.contiguous(), sizes


class EncodingType:
    Vanilla = 'vanilla'
    Fourier = 'fourier'
    Cosine = 'cosine'


@dataclass
class EncodeAux:
    """Side outputs of the fused path (all device tensors; `num_pillars` stays on the device)."""
    coors: torch.Tensor        # (cap, 4) int32 (b, z, y, x); rows >= num_pillars undefined
    num_points: torch.Tensor   # (cap,) int32
    kept_idx: torch.Tensor     # (cap, T) int32 rows into the concatenated batch
    pillar_base: torch.Tensor  # (B+1,) int32
    cell_table: torch.Tensor   # (B, ny*nx) int32, pillar id or -1
    feats: torch.Tensor        # (cap, C_out)
    frame_sizes: List[int]

    def occupancy(self, ny: int, nx: int) -> torch.Tensor:
        return (self.cell_table >= 0).view(-1, ny, nx)


class MaskBevEncoder(nn.Module):
    def __init__(self, feat_channels: List[int], x_range, y_range, z_range, voxel_size_x: float,
                 voxel_size_y: float, voxel_size_z: float, max_num_points: int, encoding_type: str,
                 fourier_enc_group: int, max_voxels: Union[tuple, int] = 500 * 500, deterministic: bool = True,
                 encoder_params: Optional[Dict] = None, pc_point_dim: int = 4):
        super().__init__()
        if encoder_params is None:
            encoder_params = {}
        self._feat_channels = feat_channels
        self._out_features = feat_channels[-1]
        self._x_range = x_range
        self._y_range = y_range
        self._z_range = z_range
        if encoding_type == EncodingType.Vanilla:
            self._pos_encoder = None
            pc_in_channels = pc_point_dim
        else:
            # mask_bev_encoders.py:54-61: 'fourier' is never selected by any file in configs/training and is out
            # of scope; anything else raises upstream as well.
            raise NotImplementedError(f'{encoding_type}')
        self._num_voxel_x = int((x_range[1] - x_range[0]) / voxel_size_x)
        self._num_voxel_y = int((y_range[1] - y_range[0]) / voxel_size_y)
        self._num_voxel_z = 1
        point_cloud_range = [x_range[0], y_range[0], z_range[0], x_range[1], y_range[1], z_range[1]]
        voxel_size = [voxel_size_x, voxel_size_y, voxel_size_z]
        self._voxel_layer = Voxelization(voxel_size, point_cloud_range, max_num_points, max_voxels, deterministic)
        self._voxel_encoder = PillarFeatureNet(in_channels=pc_in_channels, feat_channels=self._feat_channels,
                                               voxel_size=voxel_size, point_cloud_range=point_cloud_range,
                                               **encoder_params)
        out_shape = [self._num_voxel_y, self._num_voxel_x]
        self._middle_encoder = PointPillarsScatter(in_channels=self._out_features, output_shape=out_shape)
        self._layer_norm = nn.LayerNorm([self._out_features, *out_shape], eps=1e-3)
        self.apply_layer_norm = True
        # with autograd on: scatter + LayerNorm as the fused forward / backward pair (False: K3, then torch's LayerNorm)
        self.fuse_layer_norm_autograd = True
        gs = self._voxel_layer.grid_size
        if int(gs[0]) != self._num_voxel_x or int(gs[1]) != self._num_voxel_y or int(gs[2]) != 1:
            raise MbevError(f"voxel grid {gs.tolist()} disagrees with the canvas {out_shape} (SURVEY.md a1)")

    # -- fused path -------------------------------------------------------------------------------------
    def encode_batch(self, point_clouds: List[torch.Tensor], return_aux: bool = False,
                     canvas_dtype: torch.dtype = torch.float32, channels_last: bool = False, augment=None):
        """K1 -> K2 -> K3 for the whole batch: (B, C_out, ny, nx) canvas, before the LayerNorm.
        canvas_dtype=torch.bfloat16 (inference only): the canvas is written in bf16; the PFN computes in fp32 unless
        `self._voxel_encoder.gemm_path = 'tcgen05_bf16'` selects the bf16 tensor-core layers as well (BASELINE config 4).
        channels_last=True: the same tensor in torch.channels_last memory format (each pillar's features are one
        contiguous row; north star item 3), forward and backward.
        augment: an ``augment.BatchAugment`` — the reference's point augmentations applied in K1's load stage (f4)."""
        if len(point_clouds) == 0:
            raise MbevError("empty batch")
        if canvas_dtype != torch.float32:
            if torch.is_grad_enabled() and any(p.requires_grad for p in self._voxel_encoder.parameters()):
                raise MbevError("a bfloat16 canvas is forward-only: call under torch.no_grad()")
        pts, sizes = _as_points(point_clouds)
        geo = self._voxel_layer._geometry(pts.shape[1], strict_filter=True)
        vb = F_.voxelize_batch(pts, sizes, geo, augment=augment)
        if vb.points is not None:
            pts = vb.points  # the augmented cloud: what kept_idx indexes
        feats = self._voxel_encoder.apply_rows(pts, vb.kept_idx, vb.num_points, vb.coors, vb.num_pillars_dev,
                                               vb.capacity, geo.max_points)
        if canvas_dtype == torch.float32:
            canvas = scatter_with_table(feats, vb.cell_table, len(sizes), self._num_voxel_y, self._num_voxel_x,
                                        channels_last=channels_last, coors=vb.coors,
                                        num_pillars_dev=vb.num_pillars_dev)
        else:
            canvas = F_.scatter_forward(feats.detach(), vb.cell_table, len(sizes), self._num_voxel_y,
                                        self._num_voxel_x, dtype=canvas_dtype)
        if return_aux:
            return canvas, EncodeAux(vb.coors, vb.num_points, vb.kept_idx, vb.pillar_base, vb.cell_table, feats, sizes)
        return canvas

    def forward(self, point_clouds):
        """list of (N_i, C) float32 CUDA tensors -> (B, C_out, ny, nx)   (mask_bev_encoders.py:77-93)"""
        if self.apply_layer_norm and not torch.is_grad_enabled():
            fused = self._forward_fused_layer_norm(point_clouds)
            if fused is not None:
                return fused
        if self.apply_layer_norm and self.fuse_layer_norm_autograd and len(point_clouds) > 0:
            return self._forward_fused_layer_norm_autograd(point_clouds)
        pseudo_img = self.encode_batch(point_clouds)
        if self.apply_layer_norm:
            pseudo_img = self._layer_norm(pseudo_img)
        return pseudo_img

    def _forward_fused_layer_norm_autograd(self, point_clouds):
        """Training: K1 -> K2 (autograd) -> (K3 + LayerNorm) with the fused backward; the canvas is never written
        un-normalised and torch's LayerNorm (224 ms forward on the kitti_b16 canvas) is not on the path."""
        pts, sizes = _as_points(point_clouds)
        geo = self._voxel_layer._geometry(pts.shape[1], strict_filter=True)
        vb = F_.voxelize_batch(pts, sizes, geo)
        feats = self._voxel_encoder.apply_rows(pts, vb.kept_idx, vb.num_points, vb.coors, vb.num_pillars_dev,
                                               vb.capacity, geo.max_points)
        out = scatter_layernorm_with_table(feats, self._layer_norm, vb.cell_table, vb.pillar_base, vb.coors,
                                           len(sizes), self._num_voxel_y, self._num_voxel_x)
        if out is None:  # shape outside the fused kernels: K3, then torch's LayerNorm
            canvas = scatter_with_table(feats, vb.cell_table, len(sizes), self._num_voxel_y, self._num_voxel_x)
            return self._layer_norm(canvas)
        return out

    def _forward_fused_layer_norm(self, point_clouds):
        """Inference: K1 -> K2 -> (K3 + LayerNorm) — the canvas is written once, already normalised."""
        ln = self._layer_norm
        if len(point_clouds) == 0 or not ln.elementwise_affine or ln.weight is None or ln.bias is None:
            return None
        pts, sizes = _as_points(point_clouds)
        geo = self._voxel_layer._geometry(pts.shape[1], strict_filter=True)
        vb = F_.voxelize_batch(pts, sizes, geo)
        with torch.no_grad():
            feats = self._voxel_encoder.apply_rows(pts, vb.kept_idx, vb.num_points, vb.coors, vb.num_pillars_dev,
                                                   vb.capacity, geo.max_points)
            res = F_.scatter_layernorm_forward(feats, vb.cell_table, vb.pillar_base, len(sizes), self._num_voxel_y,
                                               self._num_voxel_x, ln.weight, ln.bias, ln.eps)
        if res is None:  # shape / alignment outside the fused kernel: K3, then torch's LayerNorm
            canvas = scatter_with_table(feats, vb.cell_table, len(sizes), self._num_voxel_y, self._num_voxel_x)
            return ln(canvas)
        return res[0]

    def forward_patch_tokens(self, point_clouds, patch_embed):
        """K1 -> K2 -> pillar patch embedding: the (B, Hp*Wp, E) tokens that swin.py:745-746 would
        compute from ``self.forward(point_clouds)``, without writing the pseudo image. Inference only.
        Returns (tokens, (Hp, Wp))."""
        if len(point_clouds) == 0:
            raise MbevError("empty batch")
        pts, sizes = _as_points(point_clouds)
        geo = self._voxel_layer._geometry(pts.shape[1], strict_filter=True)
        vb = F_.voxelize_batch(pts, sizes, geo)
        with torch.no_grad():
            feats = self._voxel_encoder.apply_rows(pts, vb.kept_idx, vb.num_points, vb.coors, vb.num_pillars_dev,
                                                   vb.capacity, geo.max_points)
            return patch_embed.forward_pillars(feats, vb.coors, vb.cell_table, vb.pillar_base, len(sizes),
                                               self._num_voxel_y, self._num_voxel_x, self._layer_norm)

    # -- module-level methods with the reference's semantics ---------------------------------------------
    def voxelize(self, point_clouds):
        """mask_bev_encoders.py:95-111: voxels (P,T,C), num_points (P,), coors_batch (P,4) = (b,z,y,x)."""
        pts, sizes = _as_points(point_clouds)
        geo = self._voxel_layer._geometry(pts.shape[1], strict_filter=True)
        vb = F_.voxelize_batch(pts, sizes, geo)
        P = int(vb.pillar_base[-1].item())
        voxels = F_.gather_voxels(pts, vb, P, geo.max_points)
        return voxels, vb.num_points[:P].clone(), vb.coors[:P].clone()

    def _filter_in_range(self, point_cloud):
        """mask_bev_encoders.py:113-117 (kept for API parity; the fused path applies it inside K1)."""
        in_range = (self._x_range[0] < point_cloud[:, 0]) & (point_cloud[:, 0] < self._x_range[1]) & \
                   (self._y_range[0] < point_cloud[:, 1]) & (point_cloud[:, 1] < self._y_range[1]) & \
                   (self._z_range[0] < point_cloud[:, 2]) & (point_cloud[:, 2] < self._z_range[1])
        return point_cloud[in_range]

    def encode(self, voxel, num_points, coords):
        return self._voxel_encoder(voxel, num_points, coords)

    def middle_encode(self, voxel_features, coors, batch_size=None):
        return self._middle_encoder(voxel_features, coors, batch_size)
