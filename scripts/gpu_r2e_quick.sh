#!/bin/bash
# Round-2e probe (1 GPU): the K2' changes (row-space train forward that keeps its rows, element-wise / dW kernels) and the
# bf16 scatter — parity tests that touch them, the training blocks, the launch list of one training step, and the bf16
# scatter's two task orders (second library built with -DMBEV_EXP_BF16_RUNMAJOR). usage: bash scripts/gpu_r2e_quick.sh [tag]
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_pfn_scatter.py tests/test_gpu_smoke.py tests/test_gpu_layernorm_backward.py tests/test_gpu_data_parallel.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -25 > gpurun_out/${tag}_pytest_quick.log
tail -6 gpurun_out/${tag}_pytest_quick.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --train > gpurun_out/${tag}_q_bench.json 2> gpurun_out/${tag}_q_bench.err || tail -5 gpurun_out/${tag}_q_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-layernorm --train --train-batch 16 > gpurun_out/${tag}_q_train16.json 2> gpurun_out/${tag}_q_train16.err || tail -5 gpurun_out/${tag}_q_train16.err
if [ -f _variants/lib_runmajor.so ]; then
  cp _variants/lib_runmajor.so mask_bev_b200/_C/libmask_bev_b200.so
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-layernorm --no-train > gpurun_out/${tag}_q_bf16_runmajor.json 2> gpurun_out/${tag}_q_bf16_runmajor.err || tail -5 gpurun_out/${tag}_q_bf16_runmajor.err
  timeout 300 python bench.py --workload waymo_b32 --steps 10 --warmup 3 --no-cpu-baseline --no-layernorm --no-train > gpurun_out/${tag}_q_bf16_runmajor_waymo.json 2>/dev/null
  cp _variants/lib_default.so mask_bev_b200/_C/libmask_bev_b200.so
  timeout 300 python bench.py --workload waymo_b32 --steps 10 --warmup 3 --no-cpu-baseline --no-layernorm --no-train > gpurun_out/${tag}_q_bf16_default_waymo.json 2>/dev/null
fi
python - <<PY
import json
for f in ("${tag}_q_bench", "${tag}_q_train16", "${tag}_q_bf16_runmajor", "${tag}_q_bf16_runmajor_waymo", "${tag}_q_bf16_default_waymo"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "FAILED", e); continue
    print(f, "ms/step %.3f fps %.0f e2e %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
    if d.get("train"): print("   train", {k: v for k, v in d["train"].items() if k != "what"})
    if d.get("bf16_canvas"): print("   bf16", {k: v for k, v in d["bf16_canvas"].items() if k != "note"})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_train_launches.csv python scripts/gpu_train_launches.py kitti_b16 4 > gpurun_out/${tag}_ncu_train.log 2>&1
tail -2 gpurun_out/${tag}_ncu_train.log
