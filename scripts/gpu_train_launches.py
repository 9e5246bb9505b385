"""Launch list of one encoder training step (forward + backward incl. LayerNorm, train-mode BN) for ncu:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/<tag>_train_launches.csv \\
    python scripts/gpu_train_launches.py [workload] [frames]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mask_bev_b200 as M  # noqa: E402
from mask_bev_b200.synthetic import encoder_kwargs, gen_batch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "kitti_b16"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = M.MaskBevEncoder(**encoder_kwargs(name)).to(dev).train()
frames = [torch.from_numpy(f).to(dev) for f in gen_batch(name, batch=nf)]
g = None
for it in range(3):
    enc.zero_grad()
    y = enc(frames)
    if g is None:
        g = torch.randn_like(y)
    torch.cuda.synchronize()
    if it == 2:
        torch.cuda.nvtx.range_push("step")
    y.backward(g)
    torch.cuda.synchronize()
print("ok")
