"""Build the C-ABI shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

``python -m mask_bev_b200.build`` -> mask_bev_b200/_C/libmask_bev_b200.so
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT_DIR = os.path.join(_HERE, "_C")
LIB_PATH = os.path.join(OUT_DIR, "libmask_bev_b200.so")
SOURCES = ["api.cu", "voxelize.cu", "scatter.cu", "layernorm.cu", "pfn.cu", "pfn_bwd.cu", "patch_embed.cu"]
HEADERS = ["common.cuh", "ln_stats.cuh", "tc_ptx.cuh", "rows_gemm_tc.cuh", "pfn_tc.cuh", "pfn_tcw2.cuh", os.path.join("..", "..", "include", "mask_bev_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    # IEEE fp32 everywhere the semantics need it: no fast-math, precise div/sqrt (the defaults, stated)
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; mask_bev_b200 has no CPU fallback and cannot be built without it")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("MBEV_NVCC_EXTRA", "").split()  # developer builds (-DMBEV_K2_TRACE); never set by the product
    tmp = f"{LIB_PATH}.tmp.{os.getpid()}"  # per process: ranks of one torchrun job may build at the same time
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", tmp, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)  # atomic: a concurrent loader sees the old or the new library, never a partial one
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
