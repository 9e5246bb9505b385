"""GPU parity of the fused K2+K3 kernel (mbev_pfn_scatter_forward): the PFN walks the pillars in cell order and
writer warps of the same kernel stream the canvas. Bit-identical to mbev_pfn_forward + mbev_scatter_forward (same
arithmetic per pillar, only the walk order changes) and within 1e-5 of the oracle. Through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import O, assert_close, encoder_pair, ref_test_kwargs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _frames(n, C, seeds, kind="lidar"):
    from mask_bev_b200.synthetic import gen_dense_frame, gen_frame
    if kind == "dense":
        return [gen_dense_frame(n, C, s, half=20.0) for s in seeds]
    return [gen_frame(n, C, s) for s in seeds]


def _both_paths(enc, frames):
    """(feats, canvas) of the fused kernel and of the two separate kernels on the same voxelisation."""
    from mask_bev_b200 import functional as F_
    sizes = [len(f) for f in frames]
    pts = torch.from_numpy(np.concatenate(frames, 0)).to(DEV).contiguous()
    geo = enc._voxel_layer._geometry(pts.shape[1], strict_filter=True)
    vb = F_.voxelize_batch(pts, sizes, geo)
    net = enc._voxel_encoder
    ny, nx = enc._num_voxel_y, enc._num_voxel_x
    nan_canvas = torch.full((len(sizes), net.pfn_layers[-1].units, ny, nx), float("nan"), device=DEV)
    with torch.no_grad():
        fused = net.apply_rows_canvas(pts, vb.kept_idx, vb.num_points, vb.coors, vb.capacity, geo.max_points,
                                      vb.cell_table, len(sizes), ny, nx, canvas_out=nan_canvas, force=True)
        feats = net.apply_rows(pts, vb.kept_idx, vb.num_points, vb.coors, vb.num_pillars_dev, vb.capacity,
                               geo.max_points)
        canvas = F_.scatter_forward(feats, vb.cell_table, len(sizes), ny, nx)
    torch.cuda.synchronize()
    P = int(vb.pillar_base[-1].item())
    return fused, (feats, canvas), P


CASES = [
    # chans, C, T, (lo, hi), vs, n, seeds, kind
    ((128, 128, 128), 4, 32, (-40, 40), 0.16, 40000, (1, 2, 3), "lidar"),      # G = 250000: strips straddle frames
    ((128, 128, 128), 4, 32, (-40, 40), 0.1, 60000, (4, 5), "lidar"),           # 800 x 800
    ((128, 64, 128), 5, 32, (-75.2, 75.2), 0.32, 50000, (6, 7, 8), "lidar"),    # 470 x 470: planes 16 B-aligned only
    ((64,), 3, 32, (-40, 40), 0.16, 30000, (9,), "lidar"),                      # one layer, batch 1
    ((128, 128, 128), 4, 8, (-20, 20), 0.16, 200000, (10, 11), "dense"),        # dense: most pillars at T, many strips full
    ((128, 128, 128), 4, 32, (-4, 4), 0.5, 3000, (12, 13, 14, 15, 16), "lidar"),  # 16 x 16 grid: fewer strips than sub-ranges
]


@pytest.mark.parametrize("chans,C,T,rng,vs,n,seeds,kind", CASES)
def test_fused_canvas_bit_identical_to_separate_kernels(chans, C, T, rng, vs, n, seeds, kind):
    kw = ref_test_kwargs(feat_channels=chans, T=T, C=C, x_range=rng, y_range=rng, vs=vs)
    enc, orc = encoder_pair(kw, seed=5)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    frames = _frames(n, C, seeds, kind)
    fused, (feats, canvas), P = _both_paths(enc, frames)
    assert fused is not None, "the fused kernel should take this configuration"
    f_feats, f_canvas = fused
    assert P > 0
    assert torch.equal(f_feats[:P], feats[:P]), "features differ between the fused and the separate kernels"
    assert torch.equal(f_canvas, canvas), "canvas differs between the fused and the separate kernels"
    with torch.no_grad():
        ref = orc.forward(frames).numpy()
    assert_close(f_canvas.cpu().numpy(), ref, what="fused canvas vs oracle")


def test_fused_canvas_empty_and_ragged_frames():
    """A frame with no point in range, a frame with a single point, and a normal one; canvas pre-filled with NaN so
    that any cell the writers miss shows up."""
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32, vs=0.16)
    enc, orc = encoder_pair(kw, seed=6)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    far = np.full((100, 4), 1000.0, np.float32)
    one = np.array([[1.0, 2.0, 0.5, 0.3]], np.float32)
    frames = [far, _frames(20000, 4, (3,))[0], one, np.zeros((0, 4), np.float32)]
    fused, (feats, canvas), P = _both_paths(enc, frames)
    assert fused is not None
    assert not torch.isnan(fused[1]).any()
    assert torch.equal(fused[1], canvas)
    assert float(fused[1][0].abs().max()) == 0.0 and float(fused[1][3].abs().max()) == 0.0
    with torch.no_grad():
        ref = orc.forward(frames).numpy()
    assert_close(fused[1].cpu().numpy(), ref, what="ragged batch")


def test_fused_canvas_all_frames_empty():
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32, vs=0.16)
    enc, _ = encoder_pair(kw, seed=6)
    enc = enc.to(DEV).eval()
    frames = [np.full((50, 4), 1000.0, np.float32), np.full((7, 4), -1000.0, np.float32)]
    fused, (feats, canvas), P = _both_paths(enc, frames)
    assert P == 0 and fused is not None
    assert fused[1].shape == (2, 128, 500, 500)
    assert float(fused[1].abs().max()) == 0.0 and float(canvas.abs().max()) == 0.0


def test_fused_canvas_max_voxels_truncation():
    """Pillars beyond max_voxels are dropped by K1; their cells must stay zero in the fused canvas."""
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32, vs=0.16, max_voxels=3000)
    enc, orc = encoder_pair(kw, seed=7)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    frames = _frames(30000, 4, (1, 2))
    fused, (feats, canvas), P = _both_paths(enc, frames)
    assert fused is not None and P == 6000
    assert torch.equal(fused[1], canvas)
    with torch.no_grad():
        ref = orc.forward(frames).numpy()
    assert_close(fused[1].cpu().numpy(), ref, what="truncated batch")


def test_fused_canvas_run_to_run_identical_and_runner():
    """Repeated calls of the fused kernel on reused buffers give identical bits and match the oracle."""
    from mask_bev_b200.runtime import FusedEncoderRunner
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32, vs=0.1, x_range=(0, 80), y_range=(-40, 40))
    enc, orc = encoder_pair(kw, seed=8)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    frames = _frames(60000, 4, (31, 32, 33))
    pts = torch.from_numpy(np.concatenate(frames, 0)).to(DEV)
    r = FusedEncoderRunner(enc, [len(f) for f in frames], torch.device(DEV))
    outs = []
    r.points_dev.copy_(pts)
    r.run_voxelize()
    for _ in range(3):
        r.canvas.fill_(float("nan"))
        r.run_pfn_scatter()
        outs.append(r.canvas.clone())
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    with torch.no_grad():
        ref = orc.forward(frames).numpy()
    assert_close(outs[0].cpu().numpy(), ref, what="runner canvas (fused kernel)")


def test_fused_canvas_unsupported_shapes_report_so():
    from mask_bev_b200 import functional as F_
    kw = ref_test_kwargs(feat_channels=(16, 32, 64), T=100)  # FMA stack, T > 32
    enc, _ = encoder_pair(kw, seed=1)
    cfg = enc._voxel_encoder._config()
    assert not F_.pfn_scatter_supported(cfg, 100, 2, 500, 500)
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32)
    enc, _ = encoder_pair(kw, seed=1)
    cfg = enc._voxel_encoder._config()
    assert F_.pfn_scatter_supported(cfg, 32, 2, 500, 500)
    assert not F_.pfn_scatter_supported(cfg, 32, 2, 25, 25)   # 625 cells: not a multiple of 4
    assert not F_.pfn_scatter_supported(cfg, 32, 1, 8, 8)     # fewer cells than one strip
