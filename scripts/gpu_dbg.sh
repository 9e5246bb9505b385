mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pfn_scatter.py -m gpu -q --timeout 300 -k "pipelin" > gpurun_out/dbg_pytest.log 2>&1
tail -4 gpurun_out/dbg_pytest.log
for w in kitti_b16 semkitti_b1; do
timeout 300 python bench.py --no-cpu-baseline --no-layernorm --workload $w > gpurun_out/dbg_bench_$w.json 2>gpurun_out/dbg_bench.err
python -c "
import sys,json;d=json.load(open('gpurun_out/dbg_bench_$w.json'));print('$w step', round(d['ms_per_step'],3), 'serial', round(d['serial_ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['serial_value']), {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()}, d['gpu_launches'])"
tail -3 gpurun_out/dbg_bench.err
done
