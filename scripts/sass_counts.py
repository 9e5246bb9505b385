#!/usr/bin/env python
"""Instruction counts per kernel from `cuobjdump -sass` of the built library: the SASS mnemonics that prove the
tcgen05 / TMEM / bulk-copy path (B200_PROFILING.md). python scripts/sass_counts.py > profiles/<tag>_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mask_bev_b200", "_C", "libmask_bev_b200.so")
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "USETMAXREG", "LDGSTS",
        "REDUX", "MATCH", "ELECT", "FFMA", "DFMA", "DADD"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg,")
print("# LDGSTS = cp.async, REDUX / MATCH = redux.sync / match.any, UTMALDG / UTMASTG = tensor-map TMA (none: bulk copies only)")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
total = collections.Counter()
for name, body in zip(names, blocks):
    c = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", body, re.M):
        op = m.group(1)
        for w in WANT:
            if op.startswith(w):
                c[w] += 1
    total.update(c)
    keys = [w for w in WANT if c[w] and w not in ("FFMA",)]
    if any(w in c for w in WANT[:13]):
        print(f"\n{name[:150]}")
        print("    " + "  ".join(f"{w} {c[w]}" for w in WANT if c[w]))
print("\n# whole library: " + "  ".join(f"{w} {total[w]}" for w in WANT if total[w]))
