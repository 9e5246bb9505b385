"""Host side of the input path: the reference collates a batch as a Python LIST of per-frame tensors
(`MaskListCollate` / `MaskListCollateHeight`, mask_bev/datasets/semantic_kitti/semantic_kitti_transforms.py:95-118)
and `MaskBevEncoder.voxelize` then walks that list (mask_bev_encoders.py:98-103) — one host-to-device copy and one
filter + voxelize launch sequence per frame. Here the point clouds of a batch are packed ONCE, in the DataLoader
worker, into a single `(sum N_i, C)` float32 buffer plus the frame sizes: one (pinned, asynchronous) copy moves the
whole batch, and the frames handed to `MaskBevEncoder.forward` are views of that one device tensor, in order — the
layout the fused K1 -> K2 -> K3 entry consumes (`mbev_encode_batch*`: concatenated points + frame offsets).

    loader = DataLoader(dataset, batch_size=4, collate_fn=PackedListCollate(), pin_memory=True)
    for packed, (labels, masks), metadata in loader:
        canvas = encoder(packed.to("cuda", non_blocking=True).frames())
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import torch


@dataclass
class PackedFrames:
    """A batch of point clouds as one tensor: `points` (sum N_i, C) float32 contiguous, `sizes` = N_i per frame."""
    points: torch.Tensor
    sizes: List[int]

    def __post_init__(self):
        if self.points.dim() != 2 or self.points.shape[0] != sum(self.sizes):
            raise ValueError(f"points {tuple(self.points.shape)} do not hold frames of sizes {self.sizes}")

    def __len__(self) -> int:
        return len(self.sizes)

    @property
    def offsets(self) -> List[int]:
        off = [0]
        for s in self.sizes:
            off.append(off[-1] + int(s))
        return off

    def frames(self) -> List[torch.Tensor]:
        """Per-frame views (no copy), in batch order — what `MaskBevEncoder.forward` / the reference's list API take."""
        return list(torch.split(self.points, [int(s) for s in self.sizes], dim=0))

    def pin_memory(self) -> "PackedFrames":
        """DataLoader(pin_memory=True) calls this on custom batch objects."""
        return PackedFrames(self.points.pin_memory(), list(self.sizes))

    def to(self, device, non_blocking: bool = False) -> "PackedFrames":
        return PackedFrames(self.points.to(device, non_blocking=non_blocking), list(self.sizes))


def pack_point_clouds(point_clouds: Sequence[torch.Tensor]) -> PackedFrames:
    """Concatenate per-frame (N_i, C) tensors (any float dtype, same C) into one float32 buffer."""
    if len(point_clouds) == 0:
        raise ValueError("empty batch")
    C = int(point_clouds[0].shape[1])
    for i, pc in enumerate(point_clouds):
        if pc.dim() != 2 or int(pc.shape[1]) != C:
            raise ValueError(f"frame {i} has shape {tuple(pc.shape)}, expected (N, {C})")
    sizes = [int(pc.shape[0]) for pc in point_clouds]
    out = torch.empty((sum(sizes), C), dtype=torch.float32)
    off = 0
    for pc, n in zip(point_clouds, sizes):
        out[off:off + n].copy_(pc)
        off += n
    return PackedFrames(out, sizes)


class PackedListCollate:
    """Drop-in for the reference's `MaskListCollateHeight` (same batch items `(x, (labels, masks), metadata)`, same
    stacked labels / masks, same metadata list) whose first output is a `PackedFrames` instead of a list of tensors;
    with 2-tuples `(x, (labels, masks))` it mirrors `MaskListCollate`."""

    def __call__(self, batch):
        packed = pack_point_clouds([b[0] for b in batch])
        labels = torch.stack([b[1][0] for b in batch])
        masks = torch.stack([b[1][1] for b in batch])
        if len(batch[0]) > 2:
            return packed, (labels, masks), [b[2] for b in batch]
        return packed, (labels, masks)
