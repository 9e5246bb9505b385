"""Whole path (K1 -> K2 -> K3) at BASELINE.json's full sizes, where the oracle is too slow to run: size-independent
properties of the domain checked on the GPU result itself, plus oracle parity on the first frame.
  * the canvas column of every pillar IS its feature row (bit-exact gather), every other cell is exactly zero;
  * checksum of checksums: sum(canvas) == sum(features) in float64;
  * idempotence: a second run over reused buffers gives identical bits;
  * shard invariance (eval mode): encoding the batch in two halves (what two ranks would do) reproduces the
    canvases of the unsharded batch bit for bit.
Through the C ABI (FusedEncoderRunner -> mbev_encode_batch)."""
import numpy as np
import pytest
import torch

from helpers import O, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _encoder(name, seed=0):
    import mask_bev_b200 as M
    from mask_bev_b200.synthetic import encoder_kwargs
    kw = encoder_kwargs(name)
    pfn = O.make_pfn_oracle(in_channels=kw["pc_point_dim"], feat_channels=kw["feat_channels"], with_distance=True,
                            voxel_size=[kw["voxel_size_x"], kw["voxel_size_y"], kw["voxel_size_z"]],
                            point_cloud_range=[kw["x_range"][0], kw["y_range"][0], kw["z_range"][0],
                                               kw["x_range"][1], kw["y_range"][1], kw["z_range"][1]])
    O.randomise_pfn(pfn, seed=seed)
    enc = M.MaskBevEncoder(**kw)
    enc._voxel_encoder.load_state_dict(pfn.state_dict())
    return enc.to(DEV).eval(), pfn.eval(), kw


def _run(enc, frames):
    from mask_bev_b200.runtime import FusedEncoderRunner
    r = FusedEncoderRunner(enc, [len(f) for f in frames], torch.device(DEV))
    pts = torch.from_numpy(np.concatenate(frames, 0)).to(DEV)
    r.canvas.fill_(float("nan"))
    r.run_device(pts)
    torch.cuda.synchronize()
    return r, pts


@pytest.mark.parametrize("name,batch", [("kitti_b16", 16), ("waymo_b32", 8), ("dense_1024", 1), ("semkitti_b1", 1)])
def test_full_size_path_properties(name, batch):
    from mask_bev_b200.synthetic import gen_batch
    enc, pfn, kw = _encoder(name)
    frames = gen_batch(name, batch=batch)
    r, pts = _run(enc, frames)
    canvas = r.canvas
    B, C, ny, nx = canvas.shape
    P = int(r.pillar_base[-1].item())
    assert P > 0 and not torch.isnan(canvas).any(), "every canvas byte must be written"
    coors = r.coors[:P].long()
    feats = r.feats[:P]
    # gather: the column of every pillar is its feature row
    got = canvas[coors[:, 0], :, coors[:, 2], coors[:, 3]]
    assert torch.equal(got, feats)
    # everything else is exactly zero: count of non-zero cells <= P and the table agrees with the canvas support
    occ = r.cell_table.view(B, ny, nx) >= 0
    assert int(occ.sum()) == P
    assert float(canvas.abs().amax(dim=1)[~occ].max() if (~occ).any() else 0.0) == 0.0
    # checksum of checksums
    assert float(canvas.double().sum()) == pytest.approx(float(feats.double().sum()), rel=1e-12, abs=1e-6)
    # idempotence over reused buffers
    first = canvas.clone()
    r.run_device(pts)
    torch.cuda.synchronize()
    assert torch.equal(first, r.canvas)
    del first
    # oracle parity on frame 0 (one frame of the oracle finishes in seconds)
    if name != "dense_1024":
        from oracle import oracle as Or
        orc = Or.MaskBevEncoderOracle(feat_channels=kw["feat_channels"], x_range=kw["x_range"], y_range=kw["y_range"],
                                      z_range=kw["z_range"], voxel_size_x=kw["voxel_size_x"],
                                      voxel_size_y=kw["voxel_size_y"], voxel_size_z=kw["voxel_size_z"],
                                      max_num_points=kw["max_num_points"], pc_point_dim=kw["pc_point_dim"],
                                      with_distance=True)
        orc.pfn.load_state_dict(pfn.state_dict())
        orc.pfn.eval()
        with torch.no_grad():
            ref = orc.forward(frames[:1]).numpy()
        assert_close(r.canvas[0].cpu().numpy(), ref[0], what=f"{name} frame 0 vs oracle")


def test_shard_invariance_eval_mode():
    """Rank r of 2 takes frames r, r+2, ... (mask_bev_b200.sharding): the per-rank canvases are the rows of the
    unsharded canvas, bit for bit."""
    from mask_bev_b200.sharding import shard_frames
    from mask_bev_b200.synthetic import gen_batch
    enc, _, _ = _encoder("kitti_b16", seed=3)
    frames = gen_batch("kitti_b16", batch=6, n=60000)
    r, _ = _run(enc, frames)
    whole = r.canvas.clone()
    for rank in range(2):
        mine = shard_frames(len(frames), rank, 2)
        rr, _ = _run(enc, [frames[i] for i in mine])
        assert torch.equal(rr.canvas, whole[mine]), f"rank {rank} differs from the unsharded batch"
