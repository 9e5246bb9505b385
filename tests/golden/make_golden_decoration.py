#!/usr/bin/env python
"""Golden vectors for the point decoration, produced by EXECUTING reference text.

The reference keeps a fossil of the pillar encoder's `forward` (the mmdet3d 0.x form: cluster offset, 2-channel pillar
centre offset with the legacy in-place aliasing, distance, padding mask) as a commented block in
/root/reference/mask_bev/models/encoders/mask_bev_encoders.py. This script reads that file where it
lies, strips the comment markers of the block between `def forward(self, features, num_points, coors):` and
`features *= mask`, executes it on seeded inputs and stores inputs + outputs under tests/golden/decoration_vcd2.npz.
Nothing of the reference is copied into the repository; the only restated piece is upstream's four-line
`get_paddings_indicator` helper, which the block calls and the reference does not define.

    python tests/golden/make_golden_decoration.py        # needs /root/reference (this container only)
"""
import os
import re
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference/mask_bev/models/encoders/mask_bev_encoders.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decoration_vcd2.npz")


def get_paddings_indicator(actual_num, max_num, axis=0):
    """mmdet3d.models.voxel_encoders.utils.get_paddings_indicator (upstream, not in the reference tree)."""
    actual_num = torch.unsqueeze(actual_num, axis + 1)
    shape = [1] * len(actual_num.shape)
    shape[axis + 1] = -1
    return actual_num.int() > torch.arange(max_num, dtype=torch.int, device=actual_num.device).view(shape)


def fossil_forward():
    """The commented `forward` of the reference file, uncommented, up to (and including) `features *= mask`."""
    lines = open(REF).read().splitlines()
    start = next(i for i, l in enumerate(lines) if re.match(r"#\s+def forward\(self, features, num_points, coors\):", l))
    end = next(i for i in range(start, len(lines)) if "features *= mask" in lines[i])
    body = [re.sub(r"^# ?", "", l) for l in lines[start:end + 1]]
    src = textwrap.dedent("\n".join(body)) + "\n    return features\n"
    ns = {"torch": torch, "get_paddings_indicator": get_paddings_indicator}
    exec(compile(src, REF + ":fossil", "exec"), ns)  # noqa: S102 - executing the reference is the point
    return ns["forward"]


def main():
    fwd = fossil_forward()
    rng = np.random.default_rng(20261017)
    P, T, C = 48, 8, 4
    vx, vy, vz = 0.16, 0.16, 4.0
    pcr = (-40.0, -40.0, -3.0, 40.0, 40.0, 1.0)  # z centre != 0, so the z channel tells aliasing apart
    nump = rng.integers(1, T + 1, P).astype(np.int32)
    coors = np.stack([rng.integers(0, 3, P), np.zeros(P, np.int64), rng.integers(0, 500, P), rng.integers(0, 500, P)], 1).astype(np.int32)
    vox = np.zeros((P, T, C), np.float32)
    for p in range(P):
        n = nump[p]
        vox[p, :n, 0] = pcr[0] + (coors[p, 3] + rng.uniform(0, 1, n)) * vx
        vox[p, :n, 1] = pcr[1] + (coors[p, 2] + rng.uniform(0, 1, n)) * vy
        vox[p, :n, 2] = rng.uniform(pcr[2], pcr[5], n)
        vox[p, :n, 3] = rng.uniform(0, 1, n)
    out = {}
    for legacy in (True, False):
        for with_distance in (True, False):
            me = types.SimpleNamespace(with_cluster_center=True, with_voxel_center=True, legacy=legacy,
                                       with_distance=with_distance, vx=vx, vy=vy, x_offset=vx / 2 + pcr[0],
                                       y_offset=vy / 2 + pcr[1])
            res = fwd(me, torch.from_numpy(vox.copy()), torch.from_numpy(nump), torch.from_numpy(coors))
            out[f"out_legacy{int(legacy)}_dist{int(with_distance)}"] = res.numpy()
    np.savez_compressed(OUT, voxels=vox, num_points=nump, coors=coors, voxel_size=np.array([vx, vy, vz], np.float64),
                        point_cloud_range=np.array(pcr, np.float64), **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
