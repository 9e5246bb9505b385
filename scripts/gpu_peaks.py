"""Pure-write and copy ceilings of this GPU next to MEASURED_PEAKS.json (developer probe, not a bench value)."""
import torch
x = torch.empty(16 * 128 * 800 * 800, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
tz = t(lambda: x.zero_())
tc = t(lambda: y.copy_(x))
print(f"memset 5.24 GB: {tz:.3f} ms = {x.numel()*4/tz/1e6:.0f} GB/s ; copy: {tc:.3f} ms = {2*x.numel()*4/tc/1e6:.0f} GB/s (r+w)")
