#!/bin/bash
# f2 timing: bench line with the patch_embed_f2 block. usage (under gpurun): bash scripts/gpu_f2.sh <tag>
tag=${1:-f2}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("f2", json.dumps(d.get("patch_embed_f2"), indent=1))
print("LN", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (d.get("layernorm_f1") or {}).items() if k.endswith("_ms")})
print("K", {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
PY
