"""BASELINE.json configs[4]: a full MaskBEV-shaped training step, data-parallel over the GPUs of one box.

  front end   = mask_bev_b200.MaskBevEncoder (K1, K2, K3 + LayerNorm, their backward kernels)  <- the product
  back end    = a STAND-IN for the reference's Swin backbone + Mask2Former head (mask_bev_module.py:261-266): Hugging Face
                `Mask2FormerForUniversalSegmentation` over a Swin-T-shaped backbone with a 128-channel input (SURVEY.md
                §7.4: mmdet / mmengine / Lightning are not installable offline, so the reference's own modules cannot be
                imported; this one is architecture-matched — patch 4, window 10, depths 2-2-6-2, heads 3-6-12-24, 6
                pixel-decoder layers, 9 decoder layers, Hungarian matching + CE / dice / mask losses — with random
                weights and synthetic instance masks). It is NOT part of the product and its speed is not a claim.
  exchange    = what Lightning's strategy='ddp' does for the reference (train_mask_bev.py:92-96): the back end under
                torch DDP (NCCL), the front end's gradients through mask_bev_b200.data_parallel (LayerNorm-sized
                tensors leave from a post-accumulate hook and travel under the PFN backward).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 \\
      scripts/full_step_standin.py --batch 4 --steps 5

Prints one JSON line on rank 0: ms per step (CUDA events, MAX over ranks), frames/s of the whole job, the front end's
share of the step (forward + backward timed alone), allreduce bytes."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mask_bev_b200 as M  # noqa: E402
from mask_bev_b200.data_parallel import FrontEndDataParallel, gradient_bytes  # noqa: E402
from mask_bev_b200.synthetic import encoder_kwargs, gen_batch  # noqa: E402


def standin(num_channels):
    from transformers import Mask2FormerConfig, Mask2FormerForUniversalSegmentation, SwinConfig
    bc = SwinConfig(num_channels=num_channels, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=10,
                    out_features=["stage1", "stage2", "stage3", "stage4"])
    cfg = Mask2FormerConfig(backbone_config=bc, num_queries=50, feature_size=256, mask_feature_size=256, hidden_dim=256,
                            encoder_layers=6, decoder_layers=9, num_labels=1, use_pretrained_backbone=False)
    return Mask2FormerForUniversalSegmentation(cfg)


def synthetic_masks(batch, ny, nx, seed, device):
    """A few axis-aligned vehicle-sized boxes per frame as instance masks (the reference rasterises its boxes to BEV
    masks on the CPU, datasets/*: out of scope here)."""
    rng = np.random.default_rng(seed)
    masks, classes = [], []
    for _ in range(batch):
        n = int(rng.integers(3, 9))
        m = torch.zeros((n, ny, nx), dtype=torch.float32)
        for i in range(n):
            y, x = int(rng.integers(0, ny - 30)), int(rng.integers(0, nx - 30))
            m[i, y:y + int(rng.integers(10, 30)), x:x + int(rng.integers(10, 30))] = 1.0
        masks.append(m.to(device))
        classes.append(torch.zeros((n,), dtype=torch.long, device=device))
    return masks, classes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="semkitti_b1")   # 500 x 500 canvas, semantic_kitti/01 geometry
    ap.add_argument("--batch", type=int, default=4, help="frames per GPU and step (semantic_kitti/01:28)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)  # identical initial weights on every rank
    kw = encoder_kwargs(args.workload)
    enc = M.MaskBevEncoder(**kw).to(dev).train()
    head = standin(kw["feat_channels"][-1]).to(dev).train()
    if world > 1:
        head = torch.nn.parallel.DistributedDataParallel(head, device_ids=[local])
    dp = FrontEndDataParallel(enc, overlap=True)
    params = list(enc.parameters()) + list(head.parameters())
    opt = torch.optim.AdamW(params, lr=1e-4)
    frames = [torch.from_numpy(f).to(dev) for f in gen_batch(args.workload, batch=args.batch, first_frame=rank * args.batch)]
    masks, classes = synthetic_masks(args.batch, enc._num_voxel_y, enc._num_voxel_x, 100 + rank, dev)

    def step():
        opt.zero_grad()
        canvas = enc(frames)                                     # K1, K2, K3 + LayerNorm
        out = head(pixel_values=canvas, mask_labels=masks, class_labels=classes)
        out.loss.backward()                                      # ... their backward kernels; DDP reduces the head
        dp.reduce_gradients()                                    # the front end's gradients
        opt.step()
        return out.loss

    def front_end_only():
        enc.zero_grad()
        canvas = enc(frames)
        canvas.backward(torch.ones_like(canvas))

    def timed(fn, n):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            r = fn()
        e.record()
        e.synchronize()
        t = torch.tensor([s.elapsed_time(e) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t), r

    for _ in range(args.warmup):
        loss = step()
    ms, loss = timed(step, args.steps)
    ms_fe, _ = timed(front_end_only, max(2, args.steps))
    if rank == 0:
        print(json.dumps({
            "what": "full training step: B200 front end (incl. LayerNorm) + stand-in Swin-T / Mask2Former (Hugging Face, "
                    "random init, synthetic masks) + AdamW, data-parallel; NOT the reference's mmdet modules",
            "n_gpus": world, "workload": args.workload, "frames_per_gpu": args.batch, "ms_per_step": ms,
            "frames_per_s": world * args.batch / (ms * 1e-3), "front_end_fwd_bwd_ms": ms_fe,
            "front_end_share_of_step": ms_fe / ms, "loss": float(loss),
            "front_end_allreduce_bytes": gradient_bytes(enc),
            "standin_params_m": sum(p.numel() for p in head.parameters()) / 1e6}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
