# sweep: K2 register cap x K3p build x scatter CTAs per SM of the pipelined step (developer probe)
run() { python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-layernorm --no-train "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels']
print('   ms/step %.3f serial %.3f e2e_ms %.3f | K2 %.3f K3run %.3f K3s1 %.3f K3s2 %.3f K3s4 %.3f' % (d['ms_per_step'], d['serial_ms_per_step'], 16e3/d['e2e']['value'], k['K2_pfn']['ms'], k['K3_scatter']['ms'], k['K3_scatter_stream_1cta']['ms'], k['K3_scatter_stream_2cta']['ms'], k['K3_scatter_stream_4cta']['ms']))"; }
variant() {  # $1 = K2 maxnreg, $2 = nvcc extra
  sed -i "s/^__global__ void __maxnreg__([0-9]*)\$/__global__ void __maxnreg__($1)/" mask_bev_b200/csrc/pfn_tcw2.cuh
  MBEV_NVCC_EXTRA="$2" python -m mask_bev_b200.build --force > /dev/null 2>&1
  echo "=== K2 maxnreg $1, K3p [$2]"
  for c in 0 1 2 4; do echo " ctas $c"; run --scatter-ctas $c; done
}
variant 88 "-DMBEV_EXP_K3_MINBLOCKS=7"
variant 80 "-DMBEV_EXP_K3_MINBLOCKS=5 -DMBEV_EXP_K3_DOUBLE"
variant 96 "-DMBEV_EXP_K3_MINBLOCKS=4 -DMBEV_EXP_K3_DOUBLE"
