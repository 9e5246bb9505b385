# One GPU call: parity tests, smoke, bench, ncu launch list, ncu full captures of the top kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 > gpurun_out/r1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -s 3 -c 1 -o gpurun_out/r1_scatter python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pfn -s 3 -c 1 -o gpurun_out/r1_pfn python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_rank|k_assign|k_emit|k_scan" -s 12 -c 4 -o gpurun_out/r1_vox python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/
cat gpurun_out/r1_pytest.log gpurun_out/r1_smoke.log gpurun_out/r1_bench.json
