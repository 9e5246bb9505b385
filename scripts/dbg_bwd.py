import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from helpers import *
DEV="cuda:0"
from mask_bev_b200.synthetic import gen_frame
kw = ref_test_kwargs(feat_channels=(128,128,128), T=32)
enc, orc = encoder_pair(kw, seed=3)
enc = enc.to(DEV).train()
frames=[gen_frame(12000,4,s) for s in (1,2)]
voxels, nump, coors, _ = orc.voxelize(frames)
res={}
for path in ("fma","tcgen05"):
    enc._voxel_encoder.gemm_path=path
    enc.zero_grad()
    out = enc.encode(torch.from_numpy(voxels).to(DEV), torch.from_numpy(nump).to(DEV), torch.from_numpy(coors).to(DEV))
    g = torch.Generator(device="cpu").manual_seed(0)
    w = torch.randn(out.shape, generator=g).to(DEV)
    (out*w).sum().backward()
    res[path]=(out.detach().cpu().numpy(), [p.grad.detach().cpu().numpy().copy() for p in enc._voxel_encoder._param_list()])
print("fwd rel", rel_err(res["tcgen05"][0], res["fma"][0]))
for i,(a,b) in enumerate(zip(res["tcgen05"][1], res["fma"][1])):
    print("grad", i, rel_err(a,b))
