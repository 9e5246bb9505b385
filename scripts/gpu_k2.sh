#!/bin/bash
# K2 iteration loop: PFN parity tests + K2 timing on kitti_b16 / waymo_b32 / dense_1024. usage (under gpurun): bash scripts/gpu_k2.sh <tag>
tag=${1:-k2}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_pfn_scatter.py tests/test_gpu_reference_run.py -m gpu -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
for wl in kitti_b16 waymo_b32 dense_1024 semkitti_b1; do
timeout 120 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-layernorm --no-train > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err || tail -5 gpurun_out/${tag}_bench_$wl.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_$wl.json"))
k = d["kernels"]
print("$wl ms/step %.3f serial %.3f step_frac %.3f | K1 %.4f K2 %.4f K3 %.4f | bf16 K2 %s" % (d["ms_per_step"], d["serial_ms_per_step"], d["roofline"]["step_frac"], k["K1_voxelize"]["ms"], k["K2_pfn"]["ms"], k["K3_scatter"]["ms"], (d.get("bf16_canvas") or {}).get("K2_pfn_bf16_ms")))
PY
done
