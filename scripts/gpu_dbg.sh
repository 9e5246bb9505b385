for v in "MBEV_TC_DBG=0" "MBEV_TC_DBG=128" "MBEV_TC_DBG=0"; do
env $v timeout 200 python bench.py --no-cpu-baseline --no-layernorm --steps 30 2>/dev/null | python -c "
import sys,json;d=json.loads(sys.stdin.read());print('$v step', round(d['ms_per_step'],3), {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()})"
done
timeout 400 python -m pytest tests/test_gpu_pfn_scatter.py -m gpu -x -q --timeout 200 2>&1 | tail -2
