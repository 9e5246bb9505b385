"""`PointPillarsScatter` — drop-in for mmdet3d==1.1.0's middle encoder as MaskBEV builds and calls it
(/root/reference/mask_bev/models/encoders/mask_bev_encoders.py:74, :122-123). Same constructor and forward;
the canvas is written once by K3 (csrc/scatter.cu) and the backward is the K3' gather; the LayerNorm that follows
the scatter in MaskBevEncoder.forward is fused into both directions (csrc/layernorm.cu). No CPU path.
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


class _ScatterChannelsLastFunction(torch.autograd.Function):
    """Channels-last canvas: forward = one pass over the cell table writing whole 4*C-byte rows, backward = one
    coalesced row read per pillar."""

    @staticmethod
    def forward(ctx, feats, cell_table, coors, num_pillars_dev, batch, ny, nx):
        ctx.save_for_backward(cell_table, coors, num_pillars_dev)
        ctx.rows = feats.shape[0]
        return F_.scatter_forward(feats.detach().to(torch.float32).contiguous(), cell_table, batch, ny, nx,
                                  channels_last=True)

    @staticmethod
    def backward(ctx, dcanvas):
        cell_table, coors, npil = ctx.saved_tensors
        return (F_.scatter_backward_nhwc(dcanvas, cell_table, coors, npil, ctx.rows),) + (None,) * 6


def scatter_with_table(feats: torch.Tensor, cell_table: torch.Tensor, batch: int, ny: int, nx: int,
                       channels_last: bool = False, coors: Optional[torch.Tensor] = None,
                       num_pillars_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    grad = feats.requires_grad and torch.is_grad_enabled()
    if channels_last:
        if feats.shape[1] % 4:
            raise MbevError("the channels-last canvas needs C % 4 == 0")
        if grad:
            if coors is None or num_pillars_dev is None:
                raise MbevError("channels-last scatter under autograd needs coors and the device pillar count")
            return _ScatterChannelsLastFunction.apply(feats, cell_table, coors, num_pillars_dev, batch, ny, nx)
        return F_.scatter_forward(feats.detach().to(torch.float32).contiguous(), cell_table, batch, ny, nx,
                                  channels_last=True)
    if grad:
        return _ScatterFunction.apply(feats, cell_table, batch, ny, nx)
    return F_.scatter_forward(feats.detach().to(torch.float32).contiguous(), cell_table, batch, ny, nx)


class _ScatterLayerNormFunction(torch.autograd.Function):
    """canvas = LayerNorm([C, ny, nx])(scatter(feats)) as ONE pass forward (K3+LN) and one streaming pass over the
    incoming gradient backward (mask_bev_encoders.py:91-92 and its autograd)."""

    @staticmethod
    def forward(ctx, feats, weight, bias, cell_table, pillar_base, coors, batch, ny, nx, eps):
        f = feats.detach().to(torch.float32).contiguous()
        res = F_.scatter_layernorm_forward(f, cell_table, pillar_base, batch, ny, nx, weight.detach(), bias.detach(),
                                           eps)
        if res is None:
            raise MbevError("fused scatter+LayerNorm does not support this shape (check layernorm_autograd_supported)")
        out, stats = res
        ctx.save_for_backward(f, weight, cell_table, pillar_base, coors, stats)
        ctx.batch = batch
        return out

    @staticmethod
    def backward(ctx, dout):
        f, weight, cell_table, pillar_base, coors, stats = ctx.saved_tensors
        dfeats, dweight, dbias = F_.scatter_layernorm_backward(dout, f, cell_table, coors, pillar_base[ctx.batch:],
                                                               weight.detach(), stats)
        return (dfeats, dweight.to(weight.dtype), dbias.to(weight.dtype)) + (None,) * 7


def scatter_layernorm_with_table(feats: torch.Tensor, ln: nn.LayerNorm, cell_table: torch.Tensor,
                                 pillar_base: torch.Tensor, coors: torch.Tensor, batch: int, ny: int, nx: int):
    """Autograd-enabled fused scatter + LayerNorm; None when the shape does not fit the fused kernels (the caller then
    runs K3 followed by torch's LayerNorm)."""
    C = feats.shape[1]
    if not ln.elementwise_affine or ln.weight is None or ln.bias is None:
        return None
    if tuple(ln.weight.shape) != (C, ny, nx) or ln.weight.dtype != torch.float32:
        return None
    if (ny * nx) % 4 or not F_.scatter_layernorm_backward_supported(batch, C, ny, nx):
        return None
    return _ScatterLayerNormFunction.apply(feats, ln.weight, ln.bias, cell_table, pillar_base, coors, batch, ny, nx,
                                           ln.eps)


class PointPillarsScatter(nn.Module):
    def __init__(self, in_channels: int, output_shape: List[int], channels_last: bool = False):
        super().__init__()
        self.output_shape = output_shape
        self.ny = output_shape[0]
        self.nx = output_shape[1]
        self.in_channels = in_channels
        self.fp16_enabled = False
        # additive: return the (B, C, ny, nx) canvas in torch.channels_last memory format
        self.channels_last = channels_last

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
        return scatter_with_table(voxel_features, table, batch_size, self.ny, self.nx,
                                  channels_last=self.channels_last, coors=coors, num_pillars_dev=npil)
