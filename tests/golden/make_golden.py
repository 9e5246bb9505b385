#!/usr/bin/env python
"""Regenerates tests/golden/a6_*.json — the hand-derived golden vectors of SURVEY.md Appendix A.6.

The reference ships no golden vectors for this path (its encoder tests assert shapes only) and mmcv / mmdet3d
cannot be imported here, so these vectors are NOT outputs of the reference: inputs and expected outputs below are
written out by hand from the published semantics of mmcv's deterministic hard voxelization and of
`MaskBevEncoder._filter_in_range` (mask_bev_encoders.py:113-117), and every restatement in oracle/ (and K1 on the
GPU) must reproduce them. Run:  python tests/golden/make_golden.py   (rewrites the two JSON files in place and
checks them against the pure-Python loop restatement).

Geometry: x,y in (-40, 40), z in (-20, 20), voxel 0.16 -> 500 x 500 x 1 cells, T = 2.
  pt0 ( 0,      0,     0)  cell (x 250, y 250)            -> pillar 0, slot 0
  pt1 (39.999996, 0,   0)  passes 39.999996 < 40, but floor(79.999996 / 0.16) = 500 >= grid -> dropped by the voxelizer
  pt2 (-40,     0,     0)  fails the strict -40 < -40                                       -> dropped by the filter
  pt3 ( 0.01,   0.01,  1)  cell (250, 250)                  -> pillar 0, slot 1
  pt4 (-39.99, 39.99,  0)  cell (x 0, y 499)                -> pillar 1, slot 0
  pt5 ( 0.02,   0.02,  2)  cell (250, 250): 3rd point, T = 2 -> dropped (first-T-points)
  pt6 (10,    -10,    25)  fails z < 20                                                      -> dropped by the filter
  pt7 (-39.99, 39.98, -1)  cell (0, 499)                    -> pillar 1, slot 1
  pt8 ( 5,      5,     0)  cell (281, 281)                  -> pillar 2 (max_voxels = 250000) / dropped (max_voxels = 2)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

POINTS = [[0, 0, 0, 0.5], [39.999996, 0, 0, 0.1], [-40, 0, 0, 0.2], [0.01, 0.01, 1, 0.6], [-39.99, 39.99, 0, 0.7],
          [0.02, 0.02, 2, 0.8], [10, -10, 25, 0.9], [-39.99, 39.98, -1, 0.3], [5, 5, 0, 0.4]]
SOURCE = ("SURVEY.md Appendix A.6 (hand-derived: pt1 passes the strict filter but floor(79.999996/0.16)=500>=grid; "
          "pt2 fails -40<-40; pt6 fails z; pt5 is the 3rd point of its pillar)")
BASE = dict(points=POINTS, x_range=[-40, 40], y_range=[-40, 40], z_range=[-20, 20], voxel_size=0.16, max_num_points=2,
            source=SOURCE)
CASES = {
    "a6_v250000": dict(max_voxels=250000, coors_zyx=[[0, 250, 250], [0, 499, 0], [0, 281, 281]], num_points=[2, 2, 1],
                       kept_idx=[[0, 3], [4, 7], [8, -1]]),
    "a6_v2": dict(max_voxels=2, coors_zyx=[[0, 250, 250], [0, 499, 0]], num_points=[2, 2], kept_idx=[[0, 3], [4, 7]]),
}


def main():
    from oracle import oracle as O
    for name, exp in CASES.items():
        g = dict(BASE, **exp)
        pts = np.asarray(g["points"], dtype=np.float32)
        f, src = O.filter_in_range(pts, g["x_range"], g["y_range"], g["z_range"])
        geo = O.encoder_geometry(g["x_range"], g["y_range"], g["z_range"], 0.16, 0.16, 40)
        _, c, n, k = O.hard_voxelize_py(f, geo["voxel_size"], geo["point_cloud_range"], g["max_num_points"], g["max_voxels"])
        k = np.where(k >= 0, src[np.clip(k, 0, None)], -1)
        assert c.tolist() == g["coors_zyx"] and n.tolist() == g["num_points"] and k.tolist() == g["kept_idx"], name
        with open(os.path.join(HERE, name + ".json"), "w") as fh:
            json.dump(g, fh)
        print("wrote", name)


if __name__ == "__main__":
    main()
