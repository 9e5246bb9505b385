# first tcgen05 PFN bring-up: parity of the TC path, then a short bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pfn_scatter.py -m gpu -x -q --timeout 300 -s 2>&1 | tail -40 > gpurun_out/tc1_pytest.log
cat gpurun_out/tc1_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tc1_bench.json 2> gpurun_out/tc1_bench.err
tail -5 gpurun_out/tc1_bench.err
cat gpurun_out/tc1_bench.json
