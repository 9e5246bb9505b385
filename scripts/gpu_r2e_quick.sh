#!/bin/bash
# Round-2e probe (1 GPU): sparse-run scatter (bf16: in the default library; fp32: second library built with
# -DMBEV_K3_SPARSE_F32) — bit-identity tests and K3 timings of both — and the training step under torch.profiler.
tag=${1:-r2e}
mkdir -p gpurun_out
run_bench() {  # $1 = label
  for wl in kitti_b16 waymo_b32; do
    timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-layernorm --no-train > gpurun_out/${tag}_q_$1_$wl.json 2> gpurun_out/${tag}_q_$1_$wl.err
  done
}
timeout 600 python -m pytest tests/test_gpu_pfn_scatter.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -25 > gpurun_out/${tag}_pytest_default.log
tail -4 gpurun_out/${tag}_pytest_default.log
run_bench default
if [ -f _variants/lib_sparse32.so ]; then
  cp _variants/lib_sparse32.so mask_bev_b200/_C/libmask_bev_b200.so
  timeout 600 python -m pytest tests/test_gpu_pfn_scatter.py tests/test_gpu_full_size.py tests/test_gpu_reference_run.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -25 > gpurun_out/${tag}_pytest_sparse32.log
  tail -4 gpurun_out/${tag}_pytest_sparse32.log
  run_bench sparse32
  cp _variants/lib_default.so mask_bev_b200/_C/libmask_bev_b200.so
fi
python - <<PY
import json
for v in ("default", "sparse32"):
    for w in ("kitti_b16", "waymo_b32"):
        f = "${tag}_q_%s_%s" % (v, w)
        try:
            d = json.load(open(f"gpurun_out/{f}.json"))
        except Exception as e:
            print(f, "FAILED", e); continue
        print(f, "ms/step %.3f serial %.3f fps %.0f e2e %.0f | K3 %.4f frac %.3f of_fill %.3f | bf16 K3 %.4f" % (
            d["ms_per_step"], d["serial_ms_per_step"], d["value"], d["e2e"]["value"], d["kernels"]["K3_scatter"]["ms"],
            d["roofline"]["frac"], d["roofline"].get("frac_of_fill", 0), d["bf16_canvas"]["K3_scatter_bf16_ms"]))
PY
timeout 200 python scripts/gpu_train_profile.py kitti_b16 4 > gpurun_out/${tag}_train_profile.txt 2>&1; head -30 gpurun_out/${tag}_train_profile.txt | cut -c1-110
