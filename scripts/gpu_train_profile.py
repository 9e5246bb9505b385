"""Per-kernel device times of one encoder training step INSIDE the running step (torch.profiler / CUPTI: warm caches,
real clocks, no serialisation) — the cross-check of the cold-cache ncu launch list.
python scripts/gpu_train_profile.py [workload] [frames]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

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


def step():
    global g
    enc.zero_grad()
    y = enc(frames)
    if g is None:
        g = torch.randn_like(y)
    y.backward(g)


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"# torch.profiler (CUPTI) over {N} training steps of {nf} frames of {name}: device time per step {tot / N / 1e3:.3f} ms")
print(f"# {'kernel':72s} {'n/step':>6s} {'us/step':>9s} {'avg_us':>8s} {'share':>6s}")
for k, n, t in rows[:45]:
    k = k.replace("mbev::(anonymous namespace)::", "").replace("void ", "")
    print(f"{k[:72]:72s} {n / N:6.1f} {t / N:9.1f} {t / n:8.2f} {t / tot:6.3f}")
