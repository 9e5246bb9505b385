for v in "MBEV_SCATTER_SPARSE=0" "MBEV_SCATTER_SPARSE=4" "MBEV_SCATTER_SPARSE=5"; do
env $v timeout 200 python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json;d=json.loads(sys.stdin.read());print('$v step', round(d['ms_per_step'],3), {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()})"
done
