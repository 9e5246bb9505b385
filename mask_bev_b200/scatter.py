"""`PointPillarsScatter` — drop-in for mmdet3d==1.1.0's middle encoder as MaskBEV builds and calls it
(/root/reference/mask_bev/models/encoders/mask_bev_encoders.py:74, :122-123). Same constructor and forward;
the canvas is written once by K3 (csrc/scatter.cu) and the backward is the K3' gather. No CPU path.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import functional as F_
from ._lib import MbevError


class _ScatterFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, cell_table, batch, ny, nx):
        ctx.save_for_backward(cell_table)
        ctx.rows = feats.shape[0]
        return F_.scatter_forward(feats.detach().to(torch.float32).contiguous(), cell_table, batch, ny, nx)

    @staticmethod
    def backward(ctx, dcanvas):
        (cell_table,) = ctx.saved_tensors
        return F_.scatter_backward(dcanvas, cell_table, ctx.rows), None, None, None, None


def scatter_with_table(feats: torch.Tensor, cell_table: torch.Tensor, batch: int, ny: int, nx: int) -> torch.Tensor:
    if feats.requires_grad and torch.is_grad_enabled():
        return _ScatterFunction.apply(feats, cell_table, batch, ny, nx)
    return F_.scatter_forward(feats.detach().to(torch.float32).contiguous(), cell_table, batch, ny, nx)


class PointPillarsScatter(nn.Module):
    def __init__(self, in_channels: int, output_shape: List[int]):
        super().__init__()
        self.output_shape = output_shape
        self.ny = output_shape[0]
        self.nx = output_shape[1]
        self.in_channels = in_channels
        self.fp16_enabled = False

    def forward(self, voxel_features: torch.Tensor, coors: torch.Tensor, batch_size: Optional[int] = None):
        """voxel_features (P, C), coors (P, 4) int (b, z, y, x) -> (B, C, ny, nx); batch_size None = one sample
        (upstream forward_single, which ignores the batch column)."""
        if voxel_features.shape[1] != self.in_channels:
            raise MbevError(f"voxel_features have {voxel_features.shape[1]} channels, expected {self.in_channels}")
        coors = coors.to(torch.int32).contiguous()
        if batch_size is None:
            coors = coors.clone()
            coors[:, 0] = 0
            batch_size = 1
        P = voxel_features.shape[0]
        npil = torch.full((1,), P, dtype=torch.int32, device=voxel_features.device)
        table = F_.build_cell_table(coors, npil, P, batch_size, self.ny, self.nx)
        return scatter_with_table(voxel_features, table, batch_size, self.ny, self.nx)
