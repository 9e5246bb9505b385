"""Developer probe: one call of the fused scatter+LayerNorm on kitti_b16 (for ncu)."""
import torch
import mask_bev_b200 as M
from mask_bev_b200.synthetic import encoder_kwargs, gen_batch
from mask_bev_b200 import functional as F_
kw = encoder_kwargs("kitti_b16"); frames = gen_batch("kitti_b16")
enc = M.MaskBevEncoder(**kw).to("cuda").eval()
pcs = [torch.from_numpy(f).cuda() for f in frames]
with torch.no_grad():
    sizes = [len(f) for f in frames]; pts = torch.cat(pcs)
    geo = enc._voxel_layer._geometry(4, strict_filter=True)
    vb = F_.voxelize_batch(pts, sizes, geo)
    feats = enc._voxel_encoder.apply_rows(pts, vb.kept_idx, vb.num_points, vb.coors, vb.num_pillars_dev, vb.capacity, 32)
    out = torch.empty((16, 128, 800, 800), device="cuda")
    ln = enc._layer_norm
    for _ in range(3):
        F_.scatter_layernorm_forward(feats, vb.cell_table, vb.pillar_base, 16, 800, 800, ln.weight, ln.bias, ln.eps, out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        F_.scatter_layernorm_forward(feats, vb.cell_table, vb.pillar_base, 16, 800, 800, ln.weight, ln.bias, ln.eps, out=out)
    e.record(); e.synchronize(); print("K3+LN fused: %.3f ms" % (s.elapsed_time(e) / 5))
