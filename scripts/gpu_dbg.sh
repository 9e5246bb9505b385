mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pfn_scatter.py tests/test_gpu_fused_canvas.py -m gpu -x -q --timeout 100 > gpurun_out/dbg_pytest.log 2>&1
tail -3 gpurun_out/dbg_pytest.log
for w in kitti_b16; do
timeout 200 python bench.py --no-cpu-baseline --no-layernorm --workload $w > gpurun_out/dbg_bench_$w.json 2>gpurun_out/dbg_bench.err
python -c "
import sys,json;d=json.load(open('gpurun_out/dbg_bench_$w.json'));print('$w step', round(d['ms_per_step'],3), 'serial', round(d['serial_ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()})"
tail -3 gpurun_out/dbg_bench.err
done
