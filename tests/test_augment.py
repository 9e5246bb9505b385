"""F4: the reference's point augmentations in K1's load stage. Pinned to outputs of the reference's OWN
classes (tests/golden/augment_reference.npz, make_golden_augment.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "augment_reference.npz"))
CFG = {
    "all_fire": dict(prob_drop=1.0, per_point_drop_prob=0.05, prob_flip_x=1.0, prob_flip_y=1.0, rotate_prob=1.0,
                     rotation_range=5, prob_jitter=1.0, jitter_std=0.02, intensity_std=0.01),
    "config_01": dict(prob_drop=0.5, per_point_drop_prob=0.05, prob_flip_x=0, prob_flip_y=0.5, rotate_prob=0.5,
                      rotation_range=5, prob_jitter=0.5, jitter_std=0.02, intensity_std=0.01),
    "config_01_b": dict(prob_drop=0.5, per_point_drop_prob=0.05, prob_flip_x=0, prob_flip_y=0.5, rotate_prob=0.5,
                        rotation_range=5, prob_jitter=0.5, jitter_std=0.02, intensity_std=0.01),
    "clipped_jitter": dict(prob_drop=0.0, per_point_drop_prob=0.05, prob_flip_x=0.5, prob_flip_y=0.5, rotate_prob=1.0,
                           rotation_range=(10, 40), prob_jitter=1.0, jitter_std=(0.05, 0.02, 0.01), max_delta=0.03,
                           intensity_std=0.2, intensity_max_delta=0.1),
}


def _replay(name):
    from mask_bev_b200.augment import GpuAugment
    pin = GOLD[f"{name}/points_in"]
    np.random.seed(int(GOLD[f"{name}/seed"]))
    ba = GpuAugment(**CFG[name]).sample([len(pin)], C=4, replay=True, seed=0)
    return pin, ba


@pytest.mark.parametrize("name", sorted(CFG))
def test_sampler_and_oracle_reproduce_the_reference_run_bit_for_bit(name):
    pin, ba = _replay(name)
    fa = ba.frames[0]
    got = O.augment_points_np(pin, keep=fa.keep, flip_x=fa.flip_x, flip_y=fa.flip_y, theta_deg=fa.theta_deg, noise=fa.noise)
    ref = GOLD[f"{name}/points_out"]
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    if fa.keep is not None:  # the label side gets the same decision
        assert np.array_equal(np.nonzero(fa.keep)[0], GOLD[f"{name}/inst_label_out"])


@pytest.mark.parametrize("name", sorted(CFG))
def test_label_side_follows_the_same_decisions(name):
    pytest.importorskip("cv2")
    from mask_bev_b200.augment import augment_mask
    _, ba = _replay(name)
    assert np.array_equal(augment_mask(GOLD[f"{name}/mask_in"], ba.frames[0]), GOLD[f"{name}/mask_out"])


def test_shuffle_is_refused_and_structs_match_the_header():
    import ctypes
    from mask_bev_b200 import MbevError, _lib
    from mask_bev_b200.augment import GpuAugment
    with pytest.raises(MbevError):
        GpuAugment(prob_shuffle=0.5)
    assert ctypes.sizeof(_lib.MbevFrameAugment) == 72 and ctypes.sizeof(_lib.MbevAugment) == 40


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CFG))
def test_k1_load_stage_augmentation_equals_the_reference_run(name):
    """Replay mode on the device: the augmented cloud equals the reference's output bit for bit on every kept row, and
    the pillars are those of voxelising the reference's output (dropped rows removed) with the C oracle."""
    from mask_bev_b200 import functional as F_
    import mask_bev_b200 as M
    dev = torch.device("cuda:0")
    pin, ba = _replay(name)
    fa = ba.frames[0]
    ref = GOLD[f"{name}/points_out"]
    vs, pcr, T, V = [0.5, 0.5, 8.0], [-40.0, -40.0, -4.0, 40.0, 40.0, 4.0], 8, 20000
    layer = M.Voxelization(vs, pcr, T, V)
    geo = layer._geometry(4, strict_filter=True)
    pts = torch.from_numpy(pin).to(dev)
    vb = F_.voxelize_batch(pts, [len(pin)], geo, augment=ba)
    torch.cuda.synchronize()
    keep = fa.keep if fa.keep is not None else np.ones(len(pin), bool)
    aug = vb.points.cpu().numpy()
    assert np.array_equal(aug[keep].view(np.uint32), ref.view(np.uint32))
    # pillars: oracle on the reference's compacted output; our kept_idx rows map through the compaction
    fpts, _ = O.filter_in_range(ref, (pcr[0], pcr[3]), (pcr[1], pcr[4]), (pcr[2], pcr[5]))
    vox, coors, nump = O.hard_voxelize_c(fpts, vs, pcr, T, V)[:3]
    P = int(vb.pillar_base[-1].item())
    assert P == len(coors)
    assert np.array_equal(vb.coors[:P, 1:].cpu().numpy(), coors)
    assert np.array_equal(vb.num_points[:P].cpu().numpy(), nump)
    got_vox = F_.gather_voxels(vb.points, vb, P, T).cpu().numpy()
    assert np.array_equal(got_vox.view(np.uint32), vox.view(np.uint32))


@pytest.mark.gpu
def test_generator_mode_statistics_and_determinism():
    """In-kernel Philox randomness: same distributions as the reference's numpy draws, identical from run to run."""
    from mask_bev_b200 import functional as F_
    import mask_bev_b200 as M
    from mask_bev_b200.augment import BatchAugment, FrameAugment
    dev = torch.device("cuda:0")
    n = 400000
    rng = np.random.default_rng(0)
    pin = np.c_[rng.uniform(-30, 30, (n, 2)), rng.uniform(-2, 1, n), rng.uniform(0.3, 0.7, n)].astype(np.float32)
    fa = FrameAugment(drop_prob=0.05, jitter=True, jitter_std=[0.02, 0.02, 0.02, 0.01], jitter_max=[0.05, 0, 0, 0])
    ba = BatchAugment([fa], seed=1234, sizes=[n])
    geo = M.Voxelization([0.5, 0.5, 8.0], [-40.0, -40.0, -4.0, 40.0, 40.0, 4.0], 8, 40000)._geometry(4, strict_filter=True)
    pts = torch.from_numpy(pin).to(dev)
    vb = F_.voxelize_batch(pts, [n], geo, augment=ba)
    vb2 = F_.voxelize_batch(pts, [n], geo, augment=ba)
    P = int(vb.pillar_base[-1])
    assert torch.equal(vb.points, vb2.points) and P == int(vb2.pillar_base[-1])
    assert torch.equal(vb.num_points[:P], vb2.num_points[:P]) and torch.equal(vb.coors[:P], vb2.coors[:P])
    assert torch.equal(F_.gather_voxels(vb.points, vb, P, 8), F_.gather_voxels(vb2.points, vb2, P, 8))
    d = (vb.points - pts).double().cpu().numpy()
    assert abs(d[:, 1].std() - 0.02) < 2e-4 and abs(d[:, 3].std() - 0.01) < 1e-4 and abs(d.mean()) < 1e-4
    assert np.abs(d[:, 0]).max() <= 0.05 + 1e-6 and np.abs(d[:, 1]).max() > 0.06          # x clipped, y not
    # drop rate: a geometry without truncation (T = 64, ~4 points per cell), every point in range
    geo2 = M.Voxelization([0.25, 0.25, 8.0], [-40.0, -40.0, -4.0, 40.0, 40.0, 4.0], 64, 200000)._geometry(4, strict_filter=True)
    vbd = F_.voxelize_batch(pts, [n], geo2, augment=BatchAugment([FrameAugment(drop_prob=0.05)], seed=7, sizes=[n]))
    kept = int(vbd.num_points[: int(vbd.pillar_base[-1])].sum())
    assert abs(kept / n - 0.95) < 0.003
    assert torch.equal(vbd.points, pts)  # nothing but the drop: the cloud passes through bit for bit
    vb3 = F_.voxelize_batch(pts, [n], geo, augment=BatchAugment([fa], seed=99, sizes=[n]))
    assert not torch.equal(vb.points, vb3.points)


@pytest.mark.gpu
def test_encoder_with_augmentation_equals_encoder_on_the_augmented_cloud():
    import mask_bev_b200 as M
    from mask_bev_b200.synthetic import encoder_kwargs, gen_batch
    dev = torch.device("cuda:0")
    kw = encoder_kwargs("semkitti_b1")
    enc = M.MaskBevEncoder(**kw).to(dev).eval()
    frames = gen_batch("semkitti_b1", batch=2, n=20000)
    np.random.seed(5)
    from mask_bev_b200.augment import GpuAugment
    ba = GpuAugment(prob_drop=1.0, per_point_drop_prob=0.1, prob_flip_y=1.0, rotate_prob=1.0, rotation_range=5,
                    prob_jitter=1.0, jitter_std=0.02, intensity_std=0.01).sample([len(f) for f in frames], replay=True)
    with torch.no_grad():
        a = enc.encode_batch([torch.from_numpy(f).to(dev) for f in frames], augment=ba)
        ref_frames = [O.augment_points_np(f, keep=fa.keep, flip_x=fa.flip_x, flip_y=fa.flip_y, theta_deg=fa.theta_deg,
                                          noise=fa.noise) for f, fa in zip(frames, ba.frames)]
        b = enc.encode_batch([torch.from_numpy(f).to(dev) for f in ref_frames])
    assert torch.equal(a, b)
