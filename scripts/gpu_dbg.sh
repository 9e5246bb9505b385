mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pfn_scatter.py -m gpu -q --timeout 300 -k "bf16 or pipelined" > gpurun_out/dbg_pytest.log 2>&1
tail -4 gpurun_out/dbg_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/dbg_bench.json 2>gpurun_out/dbg_bench.err
tail -3 gpurun_out/dbg_bench.err
python -c "
import json;d=json.load(open('gpurun_out/dbg_bench.json'));print(round(d['ms_per_step'],3), round(d['value']), d['bf16_canvas'])"
