timeout 600 python -m pytest tests/test_gpu_voxelize.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json;d=json.loads(sys.stdin.read());print('step', round(d['ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(v['ms'],4) for k,v in d['kernels'].items()})"
