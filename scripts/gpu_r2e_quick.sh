#!/bin/bash
# Round-2e probe (1 GPU): K2' training-step changes (parity + timing) and the pipelined step with K2 on the caller's
# stream (second library built with -DMBEV_EXP_K2_ON_MAIN). usage: bash scripts/gpu_r2e_quick.sh [tag]
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_smoke.py tests/test_gpu_pfn_scatter.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -25 > gpurun_out/${tag}_pytest_quick.log
tail -6 gpurun_out/${tag}_pytest_quick.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --train > gpurun_out/${tag}_q_bench.json 2> gpurun_out/${tag}_q_bench.err || tail -5 gpurun_out/${tag}_q_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-layernorm --train --train-batch 16 > gpurun_out/${tag}_q_train16.json 2> gpurun_out/${tag}_q_train16.err || tail -5 gpurun_out/${tag}_q_train16.err
if [ -f _variants/lib_k2main.so ]; then
  cp _variants/lib_k2main.so mask_bev_b200/_C/libmask_bev_b200.so
  for wl in kitti_b16 waymo_b32 dense_1024 semkitti_b1; do
    timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-layernorm --no-train > gpurun_out/${tag}_q_k2main_$wl.json 2> gpurun_out/${tag}_q_k2main_$wl.err
  done
  cp _variants/lib_default.so mask_bev_b200/_C/libmask_bev_b200.so
  for wl in waymo_b32 dense_1024 semkitti_b1; do
    timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-layernorm --no-train > gpurun_out/${tag}_q_default_$wl.json 2> gpurun_out/${tag}_q_default_$wl.err
  done
fi
python - <<PY
import json
fs = ["${tag}_q_bench", "${tag}_q_train16"] + ["${tag}_q_%s_%s" % (v, w) for w in ("kitti_b16", "waymo_b32", "dense_1024", "semkitti_b1") for v in ("default", "k2main")]
for f in fs:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "FAILED", e); continue
    print(f, "ms/step %.3f serial %.3f fps %.0f e2e %.0f" % (d["ms_per_step"], d["serial_ms_per_step"], d["value"], d["e2e"]["value"]))
    if d.get("train"): print("   train", {k: v for k, v in d["train"].items() if k != "what"})
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_scatter_run_bf16 -s 2 -c 1 -f -o gpurun_out/${tag}_k_scatter_run_bf16 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-layernorm > gpurun_out/${tag}_ncu_bf16.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_train_launches.csv python scripts/gpu_train_launches.py kitti_b16 4 > gpurun_out/${tag}_ncu_train.log 2>&1
tail -2 gpurun_out/${tag}_ncu_train.log
