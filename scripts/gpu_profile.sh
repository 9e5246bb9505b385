#!/bin/bash
# ncu pass of one round (run under gpurun, 1 GPU): the launch list of a short bench run and one `--set full` capture of
# each hot kernel. Summaries are made on the build box: python scripts/ncu_summary.py {launches,full} ... > profiles/...
# usage: bash scripts/gpu_profile.sh <tag> [workload]
tag=${1:-r2}
wl=${2:-kitti_b16}
B="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-layernorm --no-train"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_bench.log 2>&1
for k in k_scatter_bulk k_scatter_run k_pfn_tcw2 k_rank k_assign k_emit; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${tag}_$k $B > gpurun_out/${tag}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${tag}_*.ncu-rep
