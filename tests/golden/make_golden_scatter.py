#!/usr/bin/env python
"""Golden vectors for the scatter's index arithmetic, produced by EXECUTING reference text.

/root/reference/mask_bev/models/encoders/mask_bev_encoders.py keeps, as a commented block, a
`map_voxel_center_to_point(self, pts_coors, voxel_mean, voxel_coors)` that (step 1) scatters per-voxel rows into a
channel-major canvas exactly as upstream's PointPillarsScatter does — `indices = b*ny*nx + y*nx + x` from the
(b, z, y, x) coordinate columns, `canvas[:, indices] = rows.t()`, zeros elsewhere, canvas size
`int((hi - lo) / v)` — and (step 2) gathers a canvas column per point coordinate, which is the scatter's backward
(K3') applied to the canvas. This script reads the file where it lies, strips the comment markers of that one
function, executes it on seeded inputs and stores inputs + output under tests/golden/scatter_fossil.npz. Nothing of
the reference is copied into the repository.

    python tests/golden/make_golden_scatter.py        # needs /root/reference (this container only)
"""
import os
import re
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference/mask_bev/models/encoders/mask_bev_encoders.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scatter_fossil.npz")


def fossil_scatter_gather():
    """The commented `map_voxel_center_to_point` of the reference file, uncommented, whole."""
    lines = open(REF).read().splitlines()
    start = next(i for i, l in enumerate(lines) if re.match(r"#\s+def map_voxel_center_to_point\(", l))
    end = next(i for i in range(start, len(lines)) if re.match(r"#\s+return center_per_point", lines[i]))
    body = [re.sub(r"^# ?", "", l) for l in lines[start:end + 1]]
    src = textwrap.dedent("\n".join(body)) + "\n"
    ns = {"torch": torch}
    exec(compile(src, REF + ":fossil", "exec"), ns)  # noqa: S102 - executing the reference is the point
    return ns["map_voxel_center_to_point"]


def make_inputs():
    rng = np.random.default_rng(20261018)
    B, C, ny, nx = 3, 16, 28, 36          # rectangular canvas: a swapped x / y would show
    vx, vy = 0.5, 0.25
    pcr = (-9.0, -3.5, -3.0, 9.0, 3.5, 1.0)  # (hi - lo) / v = 36 x 28
    counts = (150, 0, 90)                 # frame 1 holds no voxel
    vc = []
    for b, n in enumerate(counts):
        cells = rng.permutation(ny * nx)[:n]
        vc.append(np.stack([np.full(n, b), np.zeros(n, np.int64), cells // nx, cells % nx], 1))
    voxel_coors = np.concatenate(vc, 0).astype(np.int32)
    voxel_mean = rng.normal(size=(len(voxel_coors), C)).astype(np.float32)
    M = 400                               # query coordinates: some on voxels, some on empty cells, last one in frame B-1
    pb = np.sort(rng.integers(0, B, M))
    pb[-1] = B - 1
    pts_coors = np.stack([pb, np.zeros(M, np.int64), rng.integers(0, ny, M), rng.integers(0, nx, M)], 1).astype(np.int32)
    hit = rng.integers(0, len(voxel_coors), M // 2)     # make half of them land on voxels
    pts_coors[: M // 2] = voxel_coors[np.sort(hit)]
    order = np.argsort(pts_coors[:, 0], kind="stable")
    return dict(voxel_coors=voxel_coors, voxel_mean=voxel_mean, pts_coors=pts_coors[order],
                voxel_size=np.array([vx, vy], np.float64), point_cloud_range=np.array(pcr, np.float64),
                shape=np.array([B, C, ny, nx], np.int64))


def run_fossil(fn, d):
    me = types.SimpleNamespace(vx=float(d["voxel_size"][0]), vy=float(d["voxel_size"][1]),
                               point_cloud_range=[float(v) for v in d["point_cloud_range"]])
    out = fn(me, torch.from_numpy(d["pts_coors"]), torch.from_numpy(d["voxel_mean"]), torch.from_numpy(d["voxel_coors"]))
    return out.numpy()


def main():
    d = make_inputs()
    d["center_per_point"] = run_fossil(fossil_scatter_gather(), d)
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, d["center_per_point"].shape, "non-zero rows:",
          int((np.abs(d["center_per_point"]).sum(1) > 0).sum()))


if __name__ == "__main__":
    main()
