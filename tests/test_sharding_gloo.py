"""N > 1 host logic on CPU: two gloo processes shard a global batch of frames, every frame is owned exactly once,
the per-rank pillar bookkeeping (oracle voxelizer as a stand-in for K1) reassembles to the unsharded result, the
timing reduction is a MAX over ranks, and a gradient allreduce averages per-rank parameter gradients."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, nframes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mask_bev_b200.sharding import gather_order, job_throughput, shard_frames
        from mask_bev_b200.synthetic import gen_frame
        from oracle import oracle as O
        mine = shard_frames(nframes, rank, world)
        # every rank voxelizes its own frames (oracle = CPU stand-in for K1) and reports per-frame pillar counts
        geo = O.encoder_geometry((-40, 40), (-40, 40), (-20, 20), 0.16, 0.16, 40)
        counts = torch.zeros(nframes, dtype=torch.int64)
        for f in mine:
            pts, _ = O.filter_in_range(gen_frame(3000, 4, 100 + f), (-40, 40), (-40, 40), (-20, 20))
            _, c, _, _ = O.hard_voxelize_np(pts, geo["voxel_size"], geo["point_cloud_range"], 32, 250000)
            counts[f] = len(c)
        owned = torch.zeros(nframes, dtype=torch.int64)
        owned[mine] = 1
        dist.all_reduce(owned)
        dist.all_reduce(counts)
        # timing: max over ranks
        t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # gradient allreduce (mean) of a PFN-sized flat bucket
        g = torch.full((25792,), float(rank + 1))
        dist.all_reduce(g)
        g /= world
        if rank == 0:
            q.put(dict(owned=owned.tolist(), counts=counts.tolist(), t=float(t), g=float(g[0]),
                       order=gather_order(nframes, world),
                       fps=job_throughput([len(shard_frames(nframes, r, world)) for r in range(world)], 4, float(t))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_frame_sharding_gloo():
    from mask_bev_b200.synthetic import gen_frame
    from oracle import oracle as O
    world, nframes = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["owned"] == [1] * nframes                      # every frame owned exactly once
    assert res["t"] == 15.0                                   # MAX over ranks
    assert res["g"] == 1.5                                    # mean of the per-rank gradients
    assert res["order"] == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2)]
    assert abs(res["fps"] - 5 * 4 / 15e-3) < 1e-6
    geo = O.encoder_geometry((-40, 40), (-40, 40), (-20, 20), 0.16, 0.16, 40)
    for f in range(nframes):                                  # sharded bookkeeping == unsharded
        pts, _ = O.filter_in_range(gen_frame(3000, 4, 100 + f), (-40, 40), (-40, 40), (-20, 20))
        _, c, _, _ = O.hard_voxelize_np(pts, geo["voxel_size"], geo["point_cloud_range"], 32, 250000)
        assert res["counts"][f] == len(c)


def test_shard_helpers():
    from mask_bev_b200.sharding import frame_owner, shard_frames
    for world in (1, 2, 4, 8):
        seen = sorted(f for r in range(world) for f in shard_frames(32, r, world))
        assert seen == list(range(32))
        assert all(len(shard_frames(32, r, world)) == 32 // world for r in range(world))
    assert frame_owner(11, 4) == (3, 2)
    with pytest.raises(ValueError):
        shard_frames(4, 4, 4)
