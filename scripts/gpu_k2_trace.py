"""Developer probe: cycle stamps of K2's CTA 0 (library built with MBEV_NVCC_EXTRA=-DMBEV_K2_TRACE).
  python scripts/gpu_k2_trace.py [workload]"""
import ctypes, sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from mask_bev_b200 import _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "kitti_b16"
cfg, kwargs, frames = bench.build_workload(wl, 0)
dev = torch.device("cuda:0")
enc, _ = bench.make_encoder(kwargs, dev)
from mask_bev_b200.runtime import FusedEncoderRunner
r = FusedEncoderRunner(enc, [len(f) for f in frames], dev)
r.set_points(torch.from_numpy(np.concatenate(frames, 0)).pin_memory())
for _ in range(3):
    r.run_device()
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros(18 * 8 * 24, dtype=np.int64)
f = lib.mbev_debug_k2_trace
f.argtypes = [ctypes.c_void_p]; f.restype = ctypes.c_int
assert f(buf.ctypes.data) == 0
t = buf.reshape(18, 8, 24)
base = t[t > 0].min()
np.set_printoptions(linewidth=250)
names_e = ["start", "packed", "x0built", "x0rdv"] + [f"L{l}{n}" for l in range(3) for n in ["_D", "_xst", "_max", "_xc", "_mst"]] + ["?", "end"]
for warp in [0, 4, 1, 8, 16, 17]:
    print("warp", warp)
    for c in range(8):
        row = t[warp, c]
        print("  c", c, " ".join(f"{(v - base) if v > 0 else -1:7d}" for v in row[:21]))
# per-phase mean deltas for an epilogue warp (warp 0, h=0) and (warp 4, h=1)
for warp in [0, 4]:
    d = np.diff(t[warp, 1:7, :21].astype(np.float64), axis=1)
    print("warp", warp, "mean phase deltas:")
    for i, nm in enumerate(names_e[1:21]):
        print(f"   -> {nm:8s} {np.nanmean(d[:, i]):9.0f}")
    print("   chunk period", np.mean(np.diff(t[warp, 1:7, 0])))
