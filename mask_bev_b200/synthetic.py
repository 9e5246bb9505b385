"""Deterministic synthetic LiDAR frames. No dataset is available offline, so every
parity test and bench line uses these; points are left in generation (random) order, which is what the
reference's ``ShufflePointCloud`` produces (mask_bev/datasets/semantic_kitti/semantic_kitti_transforms.py:58-61).
"""
from __future__ import annotations

import numpy as np

# name -> (encoder kwargs, frame generator kwargs, batch) for the BASELINE.json configs
CONFIGS = {
    # 1: single SemanticKITTI-shaped frame (configs/training/semantic_kitti/01: +-40, 0.16 -> 500x500)
    "semkitti_b1": dict(x_range=(-40, 40), y_range=(-40, 40), z_range=(-20, 20), voxel_size=0.16, T=32, C=4,
                        feat_channels=(128, 128, 128), n=120_000, batch=1, seed_base=1000, kind="lidar64"),
    # 2: KITTI-shaped frames, batch 16 (configs/training/kitti/01: x 0..80, y +-40, 0.1 -> 800x800)
    "kitti_b16": dict(x_range=(0, 80), y_range=(-40, 40), z_range=(-20, 20), voxel_size=0.1, T=32, C=4,
                      feat_channels=(128, 128, 128), n=120_000, batch=16, seed_base=2000, kind="lidar64"),
    # 3: Waymo-shaped ~180k pts, 5 feats, +-75.2 @0.32 -> 470x470, batch 32
    "waymo_b32": dict(x_range=(-75.2, 75.2), y_range=(-75.2, 75.2), z_range=(-20, 20), voxel_size=0.32, T=32, C=5,
                      feat_channels=(128, 128, 128), n=180_000, batch=32, seed_base=3000, kind="waymo"),
    # 4(i): hi-res 1024x1024 dense pillars (2M uniform points -> >250k cells -> max_voxels truncation)
    "dense_1024": dict(x_range=(-51.2, 51.2), y_range=(-51.2, 51.2), z_range=(-20, 20), voxel_size=0.1, T=32, C=4,
                       feat_channels=(128, 128, 128), n=2_000_000, batch=1, seed_base=4000, kind="dense"),
}


def gen_frame(n: int, feats: int, seed: int, beams: int = 64, elev=(-24.8, 2.0), sensor_h: float = 1.73,
              rmax: float = 120.0) -> np.ndarray:
    """One spinning-LiDAR-shaped frame: (n, feats) float32, xyz + (feats-3) uniform columns."""
    rng = np.random.default_rng(seed)
    az = rng.uniform(0.0, 2.0 * np.pi, n)
    el = np.deg2rad(np.linspace(elev[0], elev[1], beams))[rng.integers(0, beams, n)]
    with np.errstate(divide="ignore"):
        ground = np.where(el < 0, sensor_h / np.tan(-el), np.inf)
    obstacle = np.where(rng.uniform(size=n) < 0.45, rng.uniform(2.0, 80.0, n), np.inf)
    r = np.minimum(np.minimum(ground, obstacle), rmax) * (1.0 + rng.normal(0.0, 0.002, n))
    out = np.empty((n, feats), dtype=np.float32)
    out[:, 0] = r * np.cos(el) * np.cos(az)
    out[:, 1] = r * np.cos(el) * np.sin(az)
    out[:, 2] = r * np.sin(el)
    if feats > 3:
        out[:, 3:] = rng.uniform(0.0, 1.0, (n, feats - 3))
    return out


def gen_dense_frame(n: int, feats: int, seed: int, half: float = 51.2) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.empty((n, feats), dtype=np.float32)
    out[:, 0] = rng.uniform(-half, half, n)
    out[:, 1] = rng.uniform(-half, half, n)
    out[:, 2] = rng.uniform(-3.0, 1.0, n)
    if feats > 3:
        out[:, 3:] = rng.uniform(0.0, 1.0, (n, feats - 3))
    return out


def gen_batch(config: str, batch: int | None = None, n: int | None = None, first_frame: int = 0):
    """Frames ``first_frame .. first_frame+batch-1`` of a named config; frame k uses seed_base + k."""
    cfg = CONFIGS[config]
    batch = cfg["batch"] if batch is None else batch
    n = cfg["n"] if n is None else n
    frames = []
    for k in range(first_frame, first_frame + batch):
        seed = cfg["seed_base"] + k
        if cfg["kind"] == "lidar64":
            frames.append(gen_frame(n, cfg["C"], seed))
        elif cfg["kind"] == "waymo":
            frames.append(gen_frame(n, cfg["C"], seed, elev=(-17.6, 2.4), sensor_h=2.0))
        else:
            frames.append(gen_dense_frame(n, cfg["C"], seed))
    return frames


def encoder_kwargs(config: str) -> dict:
    """Keyword arguments for ``MaskBevEncoder`` exactly as mask_bev_module.py:62,72-75 derives them."""
    cfg = CONFIGS[config]
    return dict(feat_channels=list(cfg["feat_channels"]), x_range=cfg["x_range"], y_range=cfg["y_range"],
                z_range=cfg["z_range"], voxel_size_x=cfg["voxel_size"], voxel_size_y=cfg["voxel_size"],
                voxel_size_z=cfg["z_range"][1] - cfg["z_range"][0], max_num_points=cfg["T"],
                encoding_type="vanilla", fourier_enc_group=1, encoder_params=dict(with_distance=True),
                pc_point_dim=cfg["C"])
