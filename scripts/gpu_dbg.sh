for v in "MBEV_SCATTER_SPARSE=0" "MBEV_SCATTER_SPARSE=0 MBEV_SCATTER_CTAS=3"; do
env $v timeout 200 python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json;d=json.loads(sys.stdin.read());print('$v step', round(d['ms_per_step'],3), {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()})"
done
timeout 300 python -m pytest tests/test_gpu_pfn_scatter.py tests/test_gpu_fused_canvas.py -m gpu -x -q --timeout 120 2>&1 | tail -2
