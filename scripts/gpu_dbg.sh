mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pfn_scatter.py tests/test_gpu_fused_canvas.py -m gpu -x -q --timeout 200 > gpurun_out/dbg_pytest.log 2>&1
tail -3 gpurun_out/dbg_pytest.log
for w in kitti_b16; do
timeout 200 python bench.py --no-cpu-baseline --workload $w > gpurun_out/dbg_bench_$w.json 2>gpurun_out/dbg_bench.err
python -c "
import sys,json;d=json.load(open('gpurun_out/dbg_bench_$w.json'));print('$w step', round(d['ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['serial_value']), {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()}, 'roof', round(d['roofline']['frac'],3))"
tail -5 gpurun_out/dbg_bench.err
done
MBEV_TC_DBG=8 timeout 100 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>&1 | grep "^chunk" | head -8
