for cfg in "2 256" "1 128" "4 256"; do
set -- $cfg
MBEV_FILL_CTAS=$1 MBEV_FILL_THREADS=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json;d=json.loads(sys.stdin.read());k=d['kernels']['K3ab_fill_empty+scatter_occupied'];print('fill cfg $cfg: step', round(d['ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'fill', round(k['K3a_fill_empty_ms'],3), 'occ', round(k['K3b_scatter_occupied_ms'],3), 'K2', round(d['kernels']['K2_pfn']['ms'],3))"
done
