"""Developer probe: the fused scatter+LayerNorm forward / backward on kitti_b16 (for ncu), then the training step of
MaskBevEncoder.forward (LayerNorm included) with the fused pair against K3 + torch's LayerNorm."""
import sys
import torch
import mask_bev_b200 as M
from mask_bev_b200.synthetic import encoder_kwargs, gen_batch
from mask_bev_b200 import functional as F_

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
kw = encoder_kwargs("kitti_b16"); frames = gen_batch("kitti_b16")
enc = M.MaskBevEncoder(**kw).to("cuda").eval()
pcs = [torch.from_numpy(f).cuda() for f in frames]
ln = enc._layer_norm


def t(fn, n):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize(); return s.elapsed_time(e) / n


with torch.no_grad():
    sizes = [len(f) for f in frames]; pts = torch.cat(pcs)
    geo = enc._voxel_layer._geometry(4, strict_filter=True)
    vb = F_.voxelize_batch(pts, sizes, geo)
    feats = enc._voxel_encoder.apply_rows(pts, vb.kept_idx, vb.num_points, vb.coors, vb.num_pillars_dev, vb.capacity, 32)
    out = torch.empty((16, 128, 800, 800), device="cuda")
    _, stats = F_.scatter_layernorm_forward(feats, vb.cell_table, vb.pillar_base, 16, 800, 800, ln.weight, ln.bias, ln.eps, out=out)
    bwd = lambda: F_.scatter_layernorm_backward(out, feats, vb.cell_table, vb.coors, vb.pillar_base[16:], ln.weight, stats)
    print("K3+LN backward: %.3f ms" % t(bwd, 3 if mode == "ncu" else 10), flush=True)
    del out
if mode == "ncu":
    sys.exit(0)

enc.train()
g = None


def step():
    global g
    for p in enc.parameters(): p.grad = None
    o = enc(pcs)
    if g is None: g = torch.randn_like(o)
    o.backward(g)


for fused, n in ((True, 5), (False, 2)):
    enc.fuse_layer_norm_autograd = fused
    print("encoder training step incl. LayerNorm, fused=%s: %.2f ms" % (fused, t(step, n)), flush=True)
