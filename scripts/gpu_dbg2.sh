mkdir -p gpurun_out
run() {
env $1 timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $3 > gpurun_out/dbg_bench_$2.json 2>gpurun_out/dbg_bench.err
python -c "
import sys,json;d=json.load(open('gpurun_out/dbg_bench_$2.json'));print('$1 $3 step', round(d['ms_per_step'],3), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), {k[:3]:round(v['ms'],3) for k,v in d['kernels'].items()})"
tail -5 gpurun_out/dbg_bench.err
}
run MBEV_NO_FUSED_CANVAS=1 nofuse
run MBEV_NO_FUSED_CANVAS=1 nofuse_ov --overlap
run MBEV_NO_FUSED_CANVAS=1 nofuse2
run MBEV_NO_FUSED_CANVAS=1 nofuse_ov2 --overlap
