"""Frame sharding across the GPUs of one box.

Frames are independent units of the path (voxelize -> PFN -> scatter never mixes frames), so the data path has no
collective: rank r of G processes frames r, r+G, r+2G, ... of the global batch (round-robin keeps per-rank point
counts even when consecutive frames are correlated) and concatenates them into its own `(sum N, C)` tensor.
Eval-mode outputs are shard-invariant; train-mode BatchNorm statistics are per rank, as in the reference (no
SyncBN, mask_bev_module.py has none). The only collective of a training step is the PFN gradient allreduce, which
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) provides.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_frames(num_frames: int, rank: int, world: int) -> List[int]:
    """Global frame indices owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, num_frames, world))


def frame_owner(frame: int, world: int) -> Tuple[int, int]:
    """(rank, position inside that rank's local batch) of a global frame index."""
    return frame % world, frame // world


def gather_order(num_frames: int, world: int) -> List[Tuple[int, int]]:
    """For reassembling a global batch from per-rank outputs: entry f = (rank, local index) of global frame f."""
    return [frame_owner(f, world) for f in range(num_frames)]


def job_throughput(frames_per_rank: Sequence[int], steps: int, max_rank_ms: float) -> float:
    """Whole-job frames/s: all frames every rank processed / the slowest rank's device time."""
    return sum(int(x) for x in frames_per_rank) * steps / (max_rank_ms * 1e-3)
