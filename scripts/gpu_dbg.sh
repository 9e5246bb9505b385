mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/dbg_pytest.log 2>&1
tail -3 gpurun_out/dbg_pytest.log
for w in kitti_b16 semkitti_b1 dense_1024 waymo_b32; do
timeout 300 python bench.py --no-cpu-baseline --workload $w > gpurun_out/dbg_bench_$w.json 2>gpurun_out/dbg_bench.err
python -c "
import sys,json;d=json.load(open('gpurun_out/dbg_bench_$w.json'));print('$w step', round(d['ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), {k[:5]:round(v['ms'],3) for k,v in d['kernels'].items()}, 'roof', round(d['roofline']['frac'],3), 'bf16', round(d['bf16_canvas']['K3_scatter_bf16_ms'],3), round(d['bf16_canvas']['ms_per_step'],3), 'LN', d['layernorm_f1'] and round(d['layernorm_f1']['K3+LN_fused_ms'],3))"
tail -3 gpurun_out/dbg_bench.err
done
