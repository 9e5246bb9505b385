# One GPU call: parity tests, smoke, bench (all workloads), reference arm, ncu launch list, ncu full captures.
set -x
mkdir -p gpurun_out
T=${TAG:-r1b}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 > gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
for w in semkitti_b1 waymo_b32 dense_1024; do
  timeout 300 python bench.py --no-cpu-baseline --workload $w > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-layernorm > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scatter_run -s 3 -c 1 -o gpurun_out/${T}_scatter python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pfn_tcw2 -s 3 -c 1 -o gpurun_out/${T}_pfn python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rank|k_assign|k_emit|k_scan|k_head|k_flag" -s 12 -c 5 -o gpurun_out/${T}_vox python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/
cat gpurun_out/${T}_pytest.log gpurun_out/${T}_smoke.log gpurun_out/${T}_bench.json
