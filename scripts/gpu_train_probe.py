"""Developer probe: forward+backward time of the front end in train mode (autograd through K2'/K3')."""
import sys, torch, numpy as np
import mask_bev_b200 as M
from mask_bev_b200.synthetic import encoder_kwargs, gen_batch
name = sys.argv[1] if len(sys.argv) > 1 else "kitti_b16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else None
kw = encoder_kwargs(name); frames = gen_batch(name, batch=B)
enc = M.MaskBevEncoder(**kw).to("cuda").train()
enc.apply_layer_norm = False
pcs = [torch.from_numpy(f).cuda() for f in frames]
g = None
def step():
    global g
    for p in enc.parameters(): p.grad = None
    out = enc(pcs)
    if g is None: g = torch.randn_like(out)
    out.backward(g)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize(); return s.elapsed_time(e) / n
print(name, len(frames), "frames: train fwd+bwd %.2f ms" % t(step))
with torch.no_grad():
    print("  train-mode forward only (batch stats): %.2f ms" % t(lambda: enc(pcs)))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
