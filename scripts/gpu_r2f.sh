#!/bin/bash
# Round-2e closing check (1 GPU) of the final tree: full parity suite, smoke, the headline bench line with the training
# block, the 16-frame training block, the training step under torch.profiler. usage: bash scripts/gpu_r2f.sh [tag]
tag=${1:-r2f}
mkdir -p gpurun_out
rm -f gpurun_out/parity_elementwise.jsonl
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -8 > gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 --train > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -5 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-layernorm --train --train-batch 16 > gpurun_out/${tag}_bench_train16.json 2> gpurun_out/${tag}_bench_train16.err || tail -5 gpurun_out/${tag}_bench_train16.err
python - <<PY
import json
for f in ("${tag}_bench", "${tag}_bench_train16"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "FAILED", e); continue
    k = d["kernels"]
    print(f, "ms/step %.3f serial %.3f fps %.0f e2e %.0f step_frac %.3f roof %.3f | K1 %.4f K2 %.4f K3 %.4f" % (
        d["ms_per_step"], d["serial_ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["step_frac"],
        d["roofline"]["frac"], k["K1_voxelize"]["ms"], k["K2_pfn"]["ms"], k["K3_scatter"]["ms"]))
    if d.get("train"): print("   train", {a: b for a, b in d["train"].items() if a != "what"})
    if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
timeout 200 python scripts/gpu_train_profile.py kitti_b16 4 > gpurun_out/${tag}_train_profile.txt 2>&1; head -8 gpurun_out/${tag}_train_profile.txt | cut -c1-110
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_train_launches.csv python scripts/gpu_train_launches.py kitti_b16 4 > gpurun_out/${tag}_ncu_train.log 2>&1
