#!/bin/bash
# Multi-GPU pass (under gpurun --gpus N): headline bench (weak scaling + train block with the NCCL gradient allreduce),
# BASELINE configs[2] (waymo_b32 split over the ranks), configs[4] stand-in full step. usage: bash scripts/gpu_multi.sh N tag
N=${1:-2}; tag=${2:-r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_gpus$N.json 2> gpurun_out/${tag}_bench_gpus$N.err || tail -5 gpurun_out/${tag}_bench_gpus$N.err
$TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --workload waymo_b32 --scaling strong --no-train --no-layernorm > gpurun_out/${tag}_bench_waymo_strong_gpus$N.json 2> gpurun_out/${tag}_bench_waymo_strong_gpus$N.err || tail -5 gpurun_out/${tag}_bench_waymo_strong_gpus$N.err
$TR --master-port 29613 scripts/full_step_standin.py --batch 4 --steps 5 > gpurun_out/${tag}_full_step_gpus$N.json 2> gpurun_out/${tag}_full_step_gpus$N.err || tail -8 gpurun_out/${tag}_full_step_gpus$N.err
python - <<PY
import json
for f in ("${tag}_bench_gpus$N", "${tag}_bench_waymo_strong_gpus$N"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "n_gpus", d["n_gpus"], d["scaling"], "ms/step %.3f fps %.0f e2e %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
        print("   train", d.get("train"))
    except Exception as e:
        print(f, "FAILED", e)
try:
    print(open("gpurun_out/${tag}_full_step_gpus$N.json").read()[-900:])
except Exception as e:
    print("full step FAILED", e)
PY
