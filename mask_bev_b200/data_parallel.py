"""Data-parallel training step of the front end on the GPUs of one box (BASELINE config 5).

The data path has no collective (frames are independent: `sharding.shard_frames`); the ONLY exchange of a training
step is the gradient allreduce, which the reference gets from Lightning's DDP wrapper around the whole
`MaskBevModule` (`train_mask_bev.py:94-96`: `strategy='ddp'`). This module is that exchange for the encoder's own
parameters, over `torch.distributed` (NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests):

* the PFN parameters (25 792 floats for `[128,128,128]`: Linear weights, BN gamma / beta) travel as ONE flat bucket —
  one collective launch instead of nine latency-bound ones;
* the two LayerNorm gradients (`2*C*ny*nx` floats: 131 MB at 128 x 500 x 500, 656 MB at 128 x 800 x 800) are large
  enough to run at link bandwidth on their own, so they are reduced in place without a staging copy;
* all collectives are issued asynchronously before the first wait, so NCCL pipelines them on its own stream.

Train-mode BatchNorm statistics stay per rank (the reference has no SyncBN), so a sharded step equals the reference run on
each shard, not on the unsharded batch; eval-mode BN makes the sharded gradients' mean equal to the unsharded mean over
frames.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist

from .sharding import shard_frames

SMALL_BUCKET_BYTES = 1 << 20  # gradients below this size share one flat bucket


@dataclass
class AllreduceReport:
    world: int
    collectives: int      # collective launches issued
    bucket_floats: int    # elements that travelled in the flat bucket
    inplace_floats: int   # elements reduced in place (large tensors)


def _avg_op(group, average: bool):
    """(reduce op, scale still to apply): NCCL averages inside the collective (no extra pass over a 328 MB gradient);
    gloo has no AVG, so the sum is scaled afterwards."""
    if average and dist.get_backend(group) == "nccl":
        return dist.ReduceOp.AVG, False
    return dist.ReduceOp.SUM, average


def _world(group) -> int:
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True,
                        small_bucket_bytes: int = SMALL_BUCKET_BYTES) -> AllreduceReport:
    """Sum (or average) `.grad` of every parameter that requires grad over the ranks of `group`.

    Every rank must pass the same parameters in the same order. A rank whose shard produced no gradient for a parameter
    (it owned no frame this step) contributes zeros, so the collective shapes always agree. No-op on a single rank."""
    params = [p for p in params if p.requires_grad]
    world = _world(group)
    if world == 1 or not params:
        return AllreduceReport(world, 0, 0, 0)
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    small = [p for p in params if p.grad.numel() * p.grad.element_size() < small_bucket_bytes]
    large = [p for p in params if p.grad.numel() * p.grad.element_size() >= small_bucket_bytes]
    op, average = _avg_op(group, average)   # from here on `average` = "still to be scaled by 1 / world"
    handles = []
    flat = None
    if small:
        flat = torch.cat([p.grad.reshape(-1).to(torch.float32) for p in small])
        handles.append(dist.all_reduce(flat, op=op, group=group, async_op=True))
    for p in large:
        if not p.grad.is_contiguous():
            p.grad = p.grad.contiguous()
        handles.append(dist.all_reduce(p.grad, op=op, group=group, async_op=True))
    for h in handles:
        h.wait()
    scale = 1.0 / world if average else 1.0
    if flat is not None:
        if average:
            flat.mul_(scale)
        off = 0
        for p in small:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    if average:
        for p in large:
            p.grad.mul_(scale)
    return AllreduceReport(world, len(handles), 0 if flat is None else int(flat.numel()),
                           int(sum(p.grad.numel() for p in large)))


class FrontEndDataParallel:
    """One process per GPU: rank r runs `encoder` on frames r, r+G, ... of the global batch and the gradients of the
    encoder's parameters are averaged over the ranks after `backward`.

        dp = FrontEndDataParallel(encoder)            # after torch.distributed.init_process_group("nccl")
        canvas, owned = dp(frames_of_the_global_batch)  # (B_local, C, ny, nx), global indices of those frames
        loss(canvas).backward()
        dp.reduce_gradients()
        optimizer.step()

    overlap=True: the large gradients (the two LayerNorm tensors, `2*C*ny*nx` floats) are all-reduced from a
    post-accumulate-grad hook, i.e. the moment the fused scatter+LayerNorm backward has produced them — NCCL then
    moves them over NVLink while the PFN backward (K2', which does not depend on them) still computes;
    `reduce_gradients` sends the small bucket and waits for everything. Same result as overlap=False.
    """

    def __init__(self, encoder: torch.nn.Module, group=None, overlap: bool = False,
                 small_bucket_bytes: int = SMALL_BUCKET_BYTES, average: bool = True):
        self.encoder = encoder
        self.group = group
        self.overlap = overlap
        self._average = average
        self.small_bucket_bytes = small_bucket_bytes
        self._pending = []   # (parameter, work handle, scale afterwards?) of hook-issued collectives
        self._hooks = []
        if overlap:
            for p in encoder.parameters():
                if p.requires_grad and p.numel() * p.element_size() >= small_bucket_bytes:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if self.world == 1 or p.grad is None:
            return
        if not p.grad.is_contiguous():
            p.grad = p.grad.contiguous()
        op, post = _avg_op(self.group, self._average)
        self._pending.append((p, dist.all_reduce(p.grad, op=op, group=self.group, async_op=True), post))

    def close(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    @property
    def world(self) -> int:
        return _world(self.group)

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if self.world > 1 else 0

    def owned_frames(self, num_frames: int) -> List[int]:
        return shard_frames(num_frames, self.rank, self.world)

    def __call__(self, frames: Sequence[torch.Tensor]):
        owned = self.owned_frames(len(frames))
        if not owned:  # fewer frames than ranks: this rank idles in the forward and contributes zero gradients
            return None, owned
        return self.encoder([frames[i] for i in owned]), owned

    def reduce_gradients(self, average: bool = True) -> AllreduceReport:
        if not self.overlap:
            return allreduce_gradients(self.encoder.parameters(), group=self.group, average=average,
                                       small_bucket_bytes=self.small_bucket_bytes)
        if average != self._average:
            raise ValueError("overlap=True reduces from hooks with the `average` given at construction")
        done = {id(p) for p, _, _ in self._pending}
        rest = [p for p in self.encoder.parameters() if p.requires_grad and id(p) not in done]
        rep = allreduce_gradients(rest, group=self.group, average=average, small_bucket_bytes=self.small_bucket_bytes)
        world = self.world
        n_inplace = 0
        for p, work, post in self._pending:
            work.wait()
            if post:
                p.grad.mul_(1.0 / world)
            n_inplace += p.grad.numel()
        n_hooked = len(self._pending)
        self._pending = []
        return AllreduceReport(world, rep.collectives + n_hooked, rep.bucket_floats, rep.inplace_floats + n_inplace)


def gradient_bytes(encoder: torch.nn.Module) -> int:
    """Bytes one rank contributes to the step's allreduce (25 792 floats for the PFN + 2*C*ny*nx for the
    LayerNorm)."""
    return int(sum(p.numel() * p.element_size() for p in encoder.parameters() if p.requires_grad))
