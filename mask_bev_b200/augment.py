"""GPU augmentations in K1's load stage.

Host side of csrc/voxelize.cu:k_assign_aug. ``GpuAugment`` takes the constructor arguments of the reference's
augmentation classes (/root/reference/mask_bev/augmentations/semantic_kitti_mask_augmentations.py: ``RandomDropPoints``
:152-162, ``Flip`` :44-56, ``ShufflePoints`` :59-66, ``RandomRotate`` :69-101, ``JitterPoints`` :104-149) in the order the
training configs list them (configs/training/semantic_kitti/01*.yml:34-49) and draws the PER-FRAME decisions on the host
with numpy's global generator in exactly the reference's call order. Two modes for the per-point randomness:

* ``replay=True``: the per-point uniforms (drop) and the standard-normal noise (jitter) are drawn on the host as the
  reference draws them and shipped to the device — bit-exact against the reference run with the same numpy seed
  (tests/golden/augment_reference.npz);
* ``replay=False`` (the B200 way): only ~64 bytes per frame travel; the kernel generates the per-point randomness from
  (seed, point row) with Philox4x32-10. Same distributions, not the same stream.

The label side (``x.mask`` flips / rotation, ``inst_label`` drop) stays with the dataset worker: ``FrameAugment`` carries
every decision it needs (``flip_x``, ``flip_y``, ``theta_deg``, ``keep``)."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import MbevAugment, MbevError, MbevFrameAugment


@dataclass
class FrameAugment:
    flip_x: bool = False
    flip_y: bool = False
    theta_deg: Optional[float] = None           # None: no rotation
    drop_prob: float = 0.0                      # per-point probability, 0: no drop
    keep: Optional[np.ndarray] = None           # replay mode: the keep mask the reference would apply (bool, N)
    drop_u: Optional[np.ndarray] = None         # replay mode: the uniforms behind it (float64, N)
    jitter: bool = False
    noise: Optional[np.ndarray] = None          # replay mode: (N, C) float64, scaled / clipped, in ORIGINAL row indexing
    jitter_std: Sequence[float] = (0.0, 0.0, 0.0, 0.0)
    jitter_max: Sequence[float] = (0.0, 0.0, 0.0, 0.0)


@dataclass
class BatchAugment:
    frames: List[FrameAugment]
    seed: int = 0
    sizes: Sequence[int] = ()   # points per frame (needed to lay the replayed per-point arrays out)

    def to_struct(self, device, batch: int, total: int, C: int):
        """(MbevAugment, tensors to keep alive). Device arrays are built here (tiny unless replaying)."""
        if len(self.frames) != batch:
            raise MbevError(f"{len(self.frames)} frame augmentations for a batch of {batch}")
        arr = (MbevFrameAugment * batch)()
        replay_u = any(f.keep is not None for f in self.frames)
        replay_n = any(f.noise is not None for f in self.frames)
        for i, f in enumerate(self.frames):
            a = arr[i]
            if f.theta_deg is not None:
                a.rotate = 1
                a.cos_t = float(np.cos(np.deg2rad(f.theta_deg)))
                a.sin_t = float(np.sin(np.deg2rad(f.theta_deg)))
            else:
                a.rotate, a.cos_t, a.sin_t = 0, 1.0, 0.0
            a.flip_x, a.flip_y = int(f.flip_x), int(f.flip_y)
            a.drop_prob = float(np.float32(f.drop_prob))
            a.jitter = int(f.jitter)
            if f.jitter and C != 4:
                raise MbevError("JitterPoints is defined on (x, y, z, intensity) clouds: C must be 4")
            for j in range(4):
                a.jitter_std[j] = float(f.jitter_std[j])
                a.jitter_max[j] = float(f.jitter_max[j]) if f.jitter_max[j] else 0.0
        frames_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        keep = [frames_dev]
        out = MbevAugment()
        out.frames = frames_dev.data_ptr()
        out.seed = int(self.seed) & 0xFFFFFFFFFFFFFFFF
        out.drop_u = None
        out.noise = None
        if replay_u:
            # the reference decides in float64 (`keep = u >= p`, :158); the decision itself is replayed: a kept point
            # gets u = 1, a dropped one u = 0, so the kernel's float32 `u < p` cannot disagree
            buf = np.ones(total, np.float32)
            o = 0
            for f, n in zip(self.frames, self.sizes):
                if f.keep is not None:
                    buf[o:o + n] = np.where(f.keep, 1.0, 0.0)
                o += n
            du = torch.from_numpy(buf).to(device)
            keep.append(du)
            out.drop_u = du.data_ptr()
        if replay_n:
            buf = np.zeros((total, C), np.float64)
            o = 0
            for f, n in zip(self.frames, self.sizes):
                if f.noise is not None:
                    buf[o:o + n] = f.noise
                o += n
            dn = torch.from_numpy(buf).to(device)
            keep.append(dn)
            out.noise = dn.data_ptr()
        return out, keep


class GpuAugment:
    """Sampler of the per-frame decisions, mirroring the reference classes' constructor arguments and draw order."""

    def __init__(self, prob_drop: float = 0.0, per_point_drop_prob: float = 0.0, prob_flip_x: float = 0.0,
                 prob_flip_y: float = 0.0, prob_shuffle: float = 0.0, rotate_prob: float = 0.0, rotation_range=0.0,
                 prob_jitter: float = 0.0, jitter_std=0.0, max_delta=None, intensity_std: float = 0.0,
                 intensity_max_delta=None, magnitude: float = 1.0):
        if prob_shuffle:
            raise MbevError("ShufflePoints reorders the cloud on the host (np.random.shuffle); K1 keeps input order — "
                            "shuffle in the dataset worker (semantic_kitti_transforms.py:58-61 does already)")
        self.prob_drop, self.per_point_drop_prob = prob_drop, per_point_drop_prob
        self.prob_flip_x, self.prob_flip_y = prob_flip_x, prob_flip_y
        self.rotate_prob = rotate_prob
        self.rotation_range = (-rotation_range, rotation_range) if np.isscalar(rotation_range) else tuple(rotation_range)
        self.prob_jitter = prob_jitter
        self.jitter_std = (jitter_std,) * 3 if np.isscalar(jitter_std) else tuple(jitter_std)
        self.max_delta = None if max_delta is None else ((max_delta,) * 3 if np.isscalar(max_delta) else tuple(max_delta))
        self.intensity_std, self.intensity_max_delta = intensity_std, intensity_max_delta
        self.magnitude = magnitude

    def sample_frame(self, n_points: int, C: int = 4, replay: bool = False) -> FrameAugment:
        """One frame's decisions; np.random calls in the reference's order (drop, flip, shuffle, rotate, jitter)."""
        m = self.magnitude
        fa = FrameAugment()
        n_after = n_points
        keep = None
        if np.random.uniform(0, 1) < self.prob_drop:                        # RandomDropPoints.__call__ :157
            fa.drop_prob = float(self.per_point_drop_prob * m)
            if replay:
                u = np.random.uniform(0, 1, n_points)                        # :159
                keep = u >= self.per_point_drop_prob * m
                fa.drop_u, fa.keep = u, keep
                n_after = int(keep.sum())
        fa.flip_x = bool(np.random.uniform(0, 1) < self.prob_flip_x * m)    # Flip.__call__ :50
        fa.flip_y = bool(np.random.uniform(0, 1) < self.prob_flip_y * m)    # :53
        np.random.uniform(0, 1)                                              # ShufflePoints.__call__ :64 (prob 0: never taken)
        if np.random.uniform(0, 1) < self.rotate_prob:                       # RandomRotate.__call__ :84
            fa.theta_deg = float(np.random.uniform(self.rotation_range[0] * m, self.rotation_range[1] * m))  # :85-86
        if np.random.uniform(0, 1) < self.prob_jitter:                       # JitterPoints.__call__ :134
            fa.jitter = True
            std = list(self.jitter_std) + [self.intensity_std]
            mx = (list(self.max_delta) if self.max_delta is not None else [0, 0, 0]) + \
                 [self.intensity_max_delta if self.intensity_max_delta is not None else 0]
            fa.jitter_std = [s * m for s in std]
            fa.jitter_max = [d * m for d in mx]
            if replay:
                noise = np.random.standard_normal((n_after, C))              # :135
                for d in range(3):
                    noise[:, d] *= self.jitter_std[d]                        # :136-137
                if self.max_delta is not None:
                    for d in range(3):
                        np.clip(noise[:, d], -self.max_delta[d], self.max_delta[d], noise[:, d])  # :138-140
                noise[:, 3] *= self.intensity_std                            # :141
                if self.intensity_max_delta is not None:
                    np.clip(noise[:, 3], -self.intensity_max_delta, self.intensity_max_delta, noise[:, 3])  # :142-143
                noise = noise * m                                            # :145
                full = np.zeros((n_points, C), np.float64)
                if keep is None:
                    full[:] = noise
                else:
                    full[keep] = noise
                fa.noise = full
        return fa

    def sample(self, frame_sizes: Sequence[int], C: int = 4, replay: bool = False, seed: Optional[int] = None) -> BatchAugment:
        frames = [self.sample_frame(int(n), C, replay) for n in frame_sizes]
        return BatchAugment(frames, seed=int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else seed,
                            sizes=[int(n) for n in frame_sizes])


def augment_mask(mask: np.ndarray, fa: FrameAugment) -> np.ndarray:
    """Label side of one frame, on the host where the reference keeps it (the BEV instance mask never enters K1):
    Flip (:52, :55) then RandomRotate (:99-101, cv2 nearest-neighbour warp about the mask centre) driven by the same
    ``FrameAugment`` the device applied to the points."""
    if fa.flip_x:
        mask = mask[::-1, :].copy()
    if fa.flip_y:
        mask = mask[:, ::-1].copy()
    if fa.theta_deg is not None:
        import cv2  # label plumbing only; the point path never needs it
        sx, sy = mask.shape
        R_2d = cv2.getRotationMatrix2D((sx / 2, sy / 2), fa.theta_deg, 1)
        mask = cv2.warpAffine(mask, R_2d, mask.shape, flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT)
    return mask
