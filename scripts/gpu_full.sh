# parity tests + bench (overlap on/off)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 > gpurun_out/full_pytest.log
cat gpurun_out/full_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err
tail -3 gpurun_out/full_bench.err
python -c "
import json;d=json.load(open('gpurun_out/full_bench.json'));print('overlap', {k:round(v['ms'],4) for k,v in d['kernels'].items()}); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --overlap 2>/dev/null | python -c "
import sys,json;d=json.loads(sys.stdin.read());print('overlap-on', d['ms_per_step'], d['value'], d['e2e']['value'])"
