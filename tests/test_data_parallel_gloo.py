"""Gradient exchange of the data-parallel training step on CPU: two gloo ranks, PFN-sized parameters
in one flat bucket, LayerNorm-sized gradients reduced in place, a rank without frames contributing zeros — and the
result equals the mean of the per-rank gradients, as the reference's DDP wrapper produces (train_mask_bev.py:94-96)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _params():
    """Shapes of MaskBevEncoder's trainable parameters for [128,128,128], D=11, on a 40 x 40 canvas."""
    torch.manual_seed(0)
    shapes = [(64, 11), (64, 128), (128, 128), (64,), (64,), (64,), (64,), (128,), (128,), (128, 40, 40), (128, 40, 40)]
    return [torch.nn.Parameter(torch.zeros(s)) for s in shapes]


def _rank_grad(rank, i, shape):
    g = torch.Generator().manual_seed(1000 * rank + i)
    return torch.randn(shape, generator=g)


def _toy_net():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.Tanh(), torch.nn.Linear(4, 2))


def _toy_input(rank):
    return torch.randn((6, 8), generator=torch.Generator().manual_seed(77 + rank))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mask_bev_b200.data_parallel import FrontEndDataParallel, allreduce_gradients
        params = _params()
        frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
        for i, p in enumerate(params):
            if rank == 1 and i in (0, 9):   # this rank produced no gradient for two of them
                continue
            p.grad = _rank_grad(rank, i, p.shape)
        rep = allreduce_gradients(params + [frozen], small_bucket_bytes=128 * 40 * 40 * 4)
        dp = FrontEndDataParallel(torch.nn.Linear(2, 2))
        owned = [dp.owned_frames(n) for n in (5, 1)]
        # overlap=True: large gradients leave from a post-accumulate hook during backward, the rest at the end
        net = _toy_net()
        odp = FrontEndDataParallel(net, overlap=True, small_bucket_bytes=100)   # the (4, 8) weight counts as large
        for step in range(3):   # repeatedly: the pending list must be reset between steps; both zero_grad modes — kept
            net.zero_grad(set_to_none=(step != 0))   # .grad tensors, then torch's default (fresh .grad from the hook on)
            net(_toy_input(rank)).sum().backward()
            orep = odp.reduce_gradients()
        odp.close()
        if rank == 0:
            q.put(dict(grads=[p.grad.numpy().copy() for p in params], rep=dict(rep.__dict__), owned=owned,
                       frozen=frozen.grad, ograds=[p.grad.numpy().copy() for p in net.parameters()],
                       orep=dict(orep.__dict__)))   # numpy: pickled by value (tensors would travel as shared-memory handles)
        else:
            q.put(dict(owned1=owned))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = next(g for g in got if "grads" in g)
    other = next(g for g in got if "owned1" in g)
    params = _params()
    for i, p in enumerate(params):
        g0 = _rank_grad(0, i, p.shape)
        g1 = torch.zeros(p.shape) if i in (0, 9) else _rank_grad(1, i, p.shape)
        assert torch.allclose(torch.from_numpy(res["grads"][i]), (g0 + g1) / 2, rtol=0, atol=1e-6), i
    # nine PFN tensors in ONE flat bucket (25 792 floats), the two LayerNorm-sized tensors in place: 3 collectives
    assert res["rep"] == dict(world=2, collectives=3, bucket_floats=25792, inplace_floats=2 * 128 * 40 * 40)
    assert res["frozen"] is None
    # the overlapped exchange gives the same mean; one hooked collective + one flat bucket
    want = []
    for r in range(world):
        net = _toy_net()
        net(_toy_input(r)).sum().backward()
        want.append([p.grad.clone() for p in net.parameters()])
    for i, g in enumerate(res["ograds"]):
        assert torch.allclose(torch.from_numpy(g), (want[0][i] + want[1][i]) / 2, rtol=0, atol=1e-6), i
    assert res["orep"] == dict(world=2, collectives=2, bucket_floats=4 + 8 + 2, inplace_floats=32)
    assert res["owned"] == [[0, 2, 4], [0]] and other["owned1"] == [[1, 3], []]


def test_single_process_is_a_noop():
    from mask_bev_b200.data_parallel import FrontEndDataParallel, allreduce_gradients, gradient_bytes
    lin = torch.nn.Linear(4, 3)
    lin.weight.grad = torch.ones_like(lin.weight)
    rep = allreduce_gradients(lin.parameters())
    assert rep.world == 1 and rep.collectives == 0
    assert torch.equal(lin.weight.grad, torch.ones_like(lin.weight)) and lin.bias.grad is None
    dp = FrontEndDataParallel(lin)
    assert dp.world == 1 and dp.rank == 0 and dp.owned_frames(3) == [0, 1, 2]
    assert gradient_bytes(lin) == (12 + 3) * 4
