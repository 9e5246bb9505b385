# quick GPU check of the fused canvas kernel
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_fused_canvas.py -m gpu -x -q --timeout 60 > gpurun_out/dbg_pytest.log 2>&1
tail -3 gpurun_out/dbg_pytest.log
run() {
env $1 timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/dbg_bench_$2.json 2>gpurun_out/dbg_bench.err
python -c "
import sys,json;d=json.load(open('gpurun_out/dbg_bench_$2.json'));print('$1 step', round(d['ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), {k[:3]:round(v['ms'],3) for k,v in d['kernels'].items()})"
tail -5 gpurun_out/dbg_bench.err
}
run MBEV_TC_DBG=0 d0
run MBEV_TC_DBG=16 d16
run MBEV_TC_DBG=32 d32
run MBEV_TC_DBG=80 d80
run MBEV_CANVAS_KAPPA=2 k2
run MBEV_CANVAS_KAPPA=4 k4
run MBEV_CANVAS_KAPPA=8 k8
run MBEV_CANVAS_KAPPA=32 k32
