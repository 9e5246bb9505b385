#!/usr/bin/env python
"""Golden vectors for the augmentations, produced by EXECUTING the reference's own classes.

/root/reference/mask_bev/augmentations/semantic_kitti_mask_augmentations.py is imported unchanged from where it lies
(its dataset import is resolved to an empty stand-in class: the augmentations only touch ``x.scan.point_cloud``,
``x.scan.inst_label`` and ``x.mask``), and ``make_semantic_kitti_augmentation_list`` is run on the augmentation list of
configs/training/semantic_kitti/01_point_mask_data_aug_gentle.yml (probabilities raised so that every transform fires in
at least one case) with numpy's global generator seeded. Inputs, seeds and outputs go to augment_reference.npz. Nothing
of the reference is copied into the repository.

    python tests/golden/make_golden_augment.py        # needs /root/reference (this container only)
"""
import os
import sys
import types

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "augment_reference.npz")

CASES = {  # name -> (seed, n points, augmentation list as the yaml would give it)
    "all_fire": (11, 4000, [dict(name='drop', prob_drop=1.0, per_point_drop_prob=0.05),
                            dict(name='flip', prob_flip_x=1.0, prob_flip_y=1.0),
                            dict(name='shuffle', prob_shuffle=0),
                            dict(name='rotate', rotate_prob=1.0, rotation_range=5),
                            dict(name='jitter', prob_jitter=1.0, jitter_std=0.02, intensity_std=0.01)]),
    "config_01": (12, 4000, [dict(name='drop', prob_drop=0.5, per_point_drop_prob=0.05),
                             dict(name='flip', prob_flip_x=0, prob_flip_y=0.5),
                             dict(name='shuffle', prob_shuffle=0),
                             dict(name='rotate', rotate_prob=0.5, rotation_range=5),
                             dict(name='jitter', prob_jitter=0.5, jitter_std=0.02, intensity_std=0.01)]),
    "config_01_b": (13, 3000, [dict(name='drop', prob_drop=0.5, per_point_drop_prob=0.05),
                               dict(name='flip', prob_flip_x=0, prob_flip_y=0.5),
                               dict(name='shuffle', prob_shuffle=0),
                               dict(name='rotate', rotate_prob=0.5, rotation_range=5),
                               dict(name='jitter', prob_jitter=0.5, jitter_std=0.02, intensity_std=0.01)]),
    "clipped_jitter": (14, 2000, [dict(name='drop', prob_drop=0.0, per_point_drop_prob=0.05),
                                  dict(name='flip', prob_flip_x=0.5, prob_flip_y=0.5),
                                  dict(name='shuffle', prob_shuffle=0),
                                  dict(name='rotate', rotate_prob=1.0, rotation_range=(10, 40)),
                                  dict(name='jitter', prob_jitter=1.0, jitter_std=(0.05, 0.02, 0.01), max_delta=0.03,
                                       intensity_std=0.2, intensity_max_delta=0.1)]),
}


def import_reference():
    sys.path.insert(0, "/root/reference")
    for name in ("mask_bev.datasets", "mask_bev.datasets.semantic_kitti",
                 "mask_bev.datasets.semantic_kitti.semantic_kitti_mask_dataset"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["mask_bev.datasets.semantic_kitti.semantic_kitti_mask_dataset"].SemanticKittiMaskScan = type(
        "SemanticKittiMaskScan", (), {})
    import mask_bev.augmentations.semantic_kitti_mask_augmentations as A
    return A


def make_scan(seed, n):
    rng = np.random.default_rng(seed)
    pc = np.empty((n, 4), np.float32)
    pc[:, :2] = rng.uniform(-40, 40, (n, 2))
    pc[:, 2] = rng.uniform(-3, 1, n)
    pc[:, 3] = rng.uniform(0, 1, n)
    pc[:8, :2] = [[0, 0], [40, -40], [39.99999, 0.1], [-0.0, 5], [1e-3, -1e-3], [12.5, 12.5], [-40, 40], [7, -7]]
    x = types.SimpleNamespace()
    x.scan = types.SimpleNamespace(point_cloud=pc, inst_label=np.arange(n, dtype=np.int64))
    m = np.zeros((64, 64), np.uint8)
    m[10:30, 20:50] = 1
    m[40:44, 5:9] = 1
    x.mask = m
    return x


def main():
    A = import_reference()
    out = {}
    for name, (seed, n, cfg) in CASES.items():
        x = make_scan(seed, n)
        out[f"{name}/points_in"] = x.scan.point_cloud.copy()
        out[f"{name}/mask_in"] = x.mask.copy()
        augs = A.make_semantic_kitti_augmentation_list(cfg)
        np.random.seed(seed)
        for aug in augs:  # train_mask_bev.py:71 composes the list in this order
            x = aug(x)
        out[f"{name}/points_out"] = np.ascontiguousarray(x.scan.point_cloud)
        out[f"{name}/inst_label_out"] = np.ascontiguousarray(x.scan.inst_label)
        out[f"{name}/mask_out"] = np.ascontiguousarray(x.mask)
        out[f"{name}/seed"] = np.int64(seed)
        print(name, x.scan.point_cloud.shape, x.scan.point_cloud.dtype)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
