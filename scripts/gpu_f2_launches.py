"""Developer probe: a few pillar-patch-embedding calls on kitti_b16 (run under ncu for the launch list)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
import mask_bev_b200 as M
from mask_bev_b200.runtime import FusedEncoderRunner
wl = sys.argv[1] if len(sys.argv) > 1 else "kitti_b16"
cfg, kwargs, frames = bench.build_workload(wl, 0)
dev = torch.device("cuda:0")
enc, _ = bench.make_encoder(kwargs, dev)
r = FusedEncoderRunner(enc, [len(f) for f in frames], dev)
r.set_points(torch.from_numpy(np.concatenate(frames, 0)).pin_memory())
r.run_device()
pe = M.PillarPatchEmbed(in_channels=r.c_out, embed_dims=192, kernel_size=4, stride=4, norm_cfg=dict(type="LN")).to(dev)
with torch.no_grad():
    for _ in range(3):
        pe.forward_pillars(r.feats, r.coors, r.cell_table, r.pillar_base, r.B, r.ny, r.nx, enc._layer_norm)
torch.cuda.synchronize()
