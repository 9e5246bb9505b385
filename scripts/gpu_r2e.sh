#!/bin/bash
# Round-2e GPU pass (1 GPU): full parity suite, smoke, bench lines of the four workloads, 16-frame training block,
# ncu launch list + full captures of the hot kernels. usage (under gpurun): bash scripts/gpu_r2e.sh [tag]
tag=${1:-r2e}
mkdir -p gpurun_out
rm -f gpurun_out/parity_elementwise.jsonl
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -8 > gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 --train > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -5 gpurun_out/${tag}_bench.err
for wl in waymo_b32 dense_1024 semkitti_b1; do
  timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err || tail -5 gpurun_out/${tag}_bench_$wl.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-layernorm --train --train-batch 16 > gpurun_out/${tag}_bench_train16.json 2> gpurun_out/${tag}_bench_train16.err || tail -5 gpurun_out/${tag}_bench_train16.err
python - <<PY
import json
for f in ("${tag}_bench", "${tag}_bench_waymo_b32", "${tag}_bench_dense_1024", "${tag}_bench_semkitti_b1", "${tag}_bench_train16"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "FAILED", e); continue
    k = d["kernels"]
    print(f, "ms/step %.3f serial %.3f fps %.0f e2e %.0f step_frac %.3f roof %.3f | K1 %.4f K2 %.4f K3 %.4f" % (
        d["ms_per_step"], d["serial_ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["step_frac"],
        d["roofline"]["frac"], k["K1_voxelize"]["ms"], k["K2_pfn"]["ms"], k["K3_scatter"]["ms"]))
    if d.get("train"): print("   train", d["train"])
    if d.get("patch_embed_f2"): print("   f2", {a: b for a, b in d["patch_embed_f2"].items() if "ms" in a or "err" in a})
    if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_bench.log 2>&1
for k in k_pfn_tcw2 'k_scatter_run$' k_scatter_run_bf16 k_pe_gemm k_pe_tokens k_rank; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${tag}_${k%$} $B > gpurun_out/${tag}_ncu_${k%$}.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_train_launches.csv python scripts/gpu_train_launches.py kitti_b16 4 > gpurun_out/${tag}_ncu_train.log 2>&1
timeout 200 python scripts/gpu_train_profile.py kitti_b16 4 > gpurun_out/${tag}_train_profile.txt 2>&1; head -12 gpurun_out/${tag}_train_profile.txt | cut -c1-110
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dw -s 4 -c 1 -f -o gpurun_out/${tag}_k_gemm_dw python scripts/gpu_train_launches.py kitti_b16 4 > gpurun_out/${tag}_ncu_k_gemm_dw.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
