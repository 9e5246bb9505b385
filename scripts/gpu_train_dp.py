"""Data-parallel training step of the front end under torchrun (BASELINE config 5 without the
Swin / mask decoder): per-GPU batch of `--batch` frames, forward + backward of MaskBevEncoder.forward (LayerNorm
included) and the gradient allreduce of mask_bev_b200.data_parallel, timed with CUDA events, MAX over ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \\
      scripts/gpu_train_dp.py --workload semkitti_b1 --batch 4 --steps 10
(written at the end of round 1; not yet run on a multi-GPU box)"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mask_bev_b200 as M  # noqa: E402
from mask_bev_b200.data_parallel import FrontEndDataParallel, gradient_bytes  # noqa: E402
from mask_bev_b200.synthetic import encoder_kwargs, gen_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="semkitti_b1")
    ap.add_argument("--batch", type=int, default=4, help="frames per GPU and step")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.manual_seed(0)                      # identical initial weights on every rank
    enc = M.MaskBevEncoder(**encoder_kwargs(args.workload)).to(dev).train()
    dp = FrontEndDataParallel(enc)
    # the global batch: world * batch frames; every rank generates all of them (cheap) and owns frames r, r+G, ...
    frames = [torch.from_numpy(f).to(dev) for i, f in enumerate(gen_batch(args.workload, batch=world * args.batch))
              if i % world == rank]
    every = [None] * (world * args.batch)
    for j, f in enumerate(frames):
        every[rank + j * world] = f
    opt = torch.optim.AdamW(enc.parameters(), lr=1e-4)
    g = None

    def step():
        nonlocal g
        opt.zero_grad(set_to_none=True)
        out, owned = dp(every)
        if g is None:
            g = torch.randn_like(out)
        out.backward(g)
        rep = dp.reduce_gradients()
        opt.step()
        return rep

    for _ in range(args.warmup):
        rep = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        rep = step()
    e.record()
    e.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = float(ms) / args.steps
        print(json.dumps({"what": "front-end training step incl. LayerNorm, gradient allreduce and AdamW", "n_gpus": world,
                          "workload": args.workload, "frames_per_gpu": args.batch, "ms_per_step": t,
                          "frames_per_s": world * args.batch / (t * 1e-3), "allreduce_bytes": gradient_bytes(enc),
                          "collectives_per_step": rep.collectives}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
