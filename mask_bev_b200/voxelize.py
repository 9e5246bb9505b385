"""`Voxelization` — drop-in for ``mmcv.ops.Voxelization`` as MaskBEV constructs and calls it
(/root/reference/mask_bev/models/encoders/mask_bev_encoders.py:69, :100). Same constructor, attributes,
forward signature and outputs as mmcv==2.0.0's ``mmcv/ops/voxelize.py`` (hard voxelisation only — the
reference never uses ``max_num_points=-1`` dynamic voxelisation); the work runs in K1
(csrc/voxelize.cu) on the tensor's CUDA device. There is no CPU path.
"""
from __future__ import annotations

from typing import List, Tuple, Union

import torch
from torch import nn
from torch.nn.modules.utils import _pair

from . import functional as F_
from ._lib import MbevError


class Voxelization(nn.Module):
    def __init__(self, voxel_size: List, point_cloud_range: List, max_num_points: int,
                 max_voxels: Union[tuple, int] = 20000, deterministic: bool = True):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else _pair(max_voxels)
        # The B200 kernels are always deterministic (ordering comes from row indices, not atomics), so
        # deterministic=False simply gets the deterministic result.
        self.deterministic = deterministic
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        vs = torch.tensor(voxel_size, dtype=torch.float32)
        grid_size = torch.round((pcr[3:] - pcr[:3]) / vs).long()
        self.grid_size = grid_size
        input_feat_shape = grid_size[:2]
        self.pcd_shape = [*input_feat_shape, 1][::-1]

    def _geometry(self, num_feats: int, strict_filter: bool = False):
        if self.max_num_points == -1:
            raise MbevError("dynamic voxelisation (max_num_points=-1) is outside the MaskBEV path and not built")
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return F_.make_geometry(self.voxel_size, self.point_cloud_range, self.max_num_points, max_voxels, num_feats,
                                strict_filter)

    def forward(self, input: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(N, C) float32 CUDA points -> voxels (P, T, C), coors (P, 3) int32 (z, y, x), num_points (P,) int32."""
        geo = self._geometry(input.shape[1])
        vb = F_.voxelize_batch(input, [input.shape[0]], geo)
        P = int(vb.pillar_base[1].item())  # the one host sync of the module-level API (mmcv has the same one)
        voxels = F_.gather_voxels(input.contiguous(), vb, P, self.max_num_points)
        return voxels, vb.coors[:P, 1:].contiguous(), vb.num_points[:P].clone()

    def __repr__(self):
        s = self.__class__.__name__ + '('
        s += 'voxel_size=' + str(self.voxel_size)
        s += ', point_cloud_range=' + str(self.point_cloud_range)
        s += ', max_num_points=' + str(self.max_num_points)
        s += ', max_voxels=' + str(self.max_voxels)
        s += ', deterministic=' + str(self.deterministic)
        s += ')'
        return s
