mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/dbg_pytest.log 2>&1
tail -2 gpurun_out/dbg_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/final_bench.json 2>gpurun_out/dbg_bench.err
python -c "
import sys,json;d=json.load(open('gpurun_out/final_bench.json'));print('step', round(d['ms_per_step'],3), 'serial', round(d['serial_ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()}, d['gpu_launches'], 'roof', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],2))"
tail -3 gpurun_out/dbg_bench.err
