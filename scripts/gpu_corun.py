"""Developer probe (not a bench value): K2 and the TMA-engine K3 alone and CONCURRENTLY on two streams, each timed
with CUDA events on its own stream — does the 128-thread scatter CTA share the SMs with K2's persistent CTA, and what
does each kernel cost the other?  python scripts/gpu_corun.py [workload] [ctas_per_sm]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mask_bev_b200.runtime import FusedEncoderRunner  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "kitti_b16"
ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
cfg, kwargs, frames = bench.build_workload(name, 0)
enc, _ = bench.make_encoder(kwargs, dev)
r = FusedEncoderRunner(enc, [len(f) for f in frames], dev)
r.set_points(torch.from_numpy(np.concatenate(frames, 0)))
r.run_device()
torch.cuda.synchronize()
s2, s3 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, stream, n=10):
    with torch.cuda.stream(stream):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
    return a, b, n


def report(tag, *runs):
    torch.cuda.synchronize()
    print(tag, "  ".join(f"{nm} {a.elapsed_time(b) / n:.3f} ms" for nm, (a, b, n) in runs))


k3 = lambda: r.run_scatter_stream(ctas)  # noqa: E731
for _ in range(2):
    r.run_pfn(); k3(); r.run_scatter()
torch.cuda.synchronize()
report("alone      :", ("K2", timed(r.run_pfn, s2)))
report("alone      :", (f"K3 stream x{ctas}", timed(k3, s3)))
report("alone      :", ("K3 register", timed(r.run_scatter, s3)))
ta = timed(k3, s3); tb = timed(r.run_pfn, s2)
report("concurrent :", (f"K3 stream x{ctas} (first)", ta), ("K2", tb))
tb = timed(r.run_pfn, s2); ta = timed(k3, s3)
report("concurrent :", ("K2 (first)", tb), (f"K3 stream x{ctas}", ta))
ta = timed(r.run_scatter, s3); tb = timed(r.run_pfn, s2)
report("concurrent :", ("K3 register (first)", ta), ("K2", tb))
# wall time of 10 x (K2 || K3) pairs
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
ta = timed(k3, s3); tb = timed(r.run_pfn, s2)
torch.cuda.current_stream().wait_stream(s2); torch.cuda.current_stream().wait_stream(s3)
b.record()
torch.cuda.synchronize()
print("10 x (K2 || K3 stream): %.3f ms per pair" % (a.elapsed_time(b) / 10))
# zero-only canvas: every run empty (pure TMA zero stream) next to K2
table = r.cell_table.clone()
r.cell_table.fill_(-1)
ta = timed(k3, s3); tb = timed(r.run_pfn, s2)
report("concurrent, empty table:", (f"K3 stream x{ctas} zeros only", ta), ("K2", tb))
report("alone, empty table     :", (f"K3 stream x{ctas} zeros only", timed(k3, s3)))
r.cell_table.copy_(table)

# ---- clocks / power during long loops: is the concurrent slowdown a power cap? ----
def sampled(tag, fns, n=150):
    samp = bench.NvmlSampler(0)
    samp.start()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ts = [timed(fn, st, n) for fn, st in fns]
    for _, st in fns:
        torch.cuda.current_stream().wait_stream(st)
    b.record()
    torch.cuda.synchronize()
    c = samp.stop()
    print(f"{tag}: wall {a.elapsed_time(b) / n:.3f} ms/iter  " + "  ".join(f"{x.elapsed_time(y) / m:.3f}" for x, y, m in ts) +
          f"  sm_mhz median {c.get('sm_mhz')} min {c.get('sm_min_mhz')} power_max {c.get('power_w_max')} reasons {c.get('reasons')}")


sampled("K2 alone        ", [(r.run_pfn, s2)])
sampled("K3 stream alone ", [(k3, s3)])
sampled("K3 register     ", [(r.run_scatter, s3)])
sampled("K2 || K3 stream ", [(k3, s3), (r.run_pfn, s2)])
r.cell_table.fill_(-1)
sampled("K2 || K3 zeros  ", [(k3, s3), (r.run_pfn, s2)])
sampled("K3 zeros alone  ", [(k3, s3)])
r.cell_table.copy_(table)
