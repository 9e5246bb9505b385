#!/usr/bin/env python
"""Collect the encoder-relevant keys of every file under /root/reference/configs/training into
tests/golden/training_configs.json (SURVEY.md §8a "parameter envelope the path must accept"). Defaults for absent keys
are the ones of MaskBevModule.__init__ / train_mask_bev.py (pc_point_dim 4, encoder_encoding_type 'vanilla').

    python tests/golden/make_training_configs.py        # needs /root/reference (this container only)
"""
import glob
import json
import os

import yaml

REF = "/root/reference/configs/training"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "training_configs.json")
KEYS = ("x_range", "y_range", "z_range", "voxel_size", "max_num_points", "encoder_feat_channels", "pc_point_dim",
        "encoder_encoding_type", "encoder_fourier_enc_group", "batch_size")


def collect():
    out = {}
    for f in sorted(glob.glob(os.path.join(REF, "*", "*.yml"))):
        cfg = yaml.safe_load(open(f))
        out[os.path.relpath(f, REF)] = {k: cfg[k] for k in KEYS if k in cfg}
    return out


if __name__ == "__main__":
    d = collect()
    json.dump(d, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, len(d), "configs")
