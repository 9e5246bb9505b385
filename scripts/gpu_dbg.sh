mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/dbg_pytest.log 2>&1
tail -3 gpurun_out/dbg_pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-layernorm > gpurun_out/dbg_bench.json 2>gpurun_out/dbg_bench.err
python -c "
import sys,json;d=json.load(open('gpurun_out/dbg_bench.json'));print('step', round(d['ms_per_step'],3), 'serial', round(d['serial_ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()})"
tail -3 gpurun_out/dbg_bench.err
MBEV_SCATTER=1 timeout 300 python bench.py --no-cpu-baseline --no-layernorm --steps 20 2>/dev/null | python -c "
import sys,json;d=json.loads(sys.stdin.read());print('scatter=1', {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()})"
