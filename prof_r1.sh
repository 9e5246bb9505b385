set -x
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/r1b_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -s 3 -c 2 -o gpurun_out/r1_scatter python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pfn -s 3 -c 1 -o gpurun_out/r1_pfn python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rank -s 3 -c 1 -o gpurun_out/r1_rank python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/
