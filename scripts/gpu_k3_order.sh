# developer probe: K3 task order (frame, channel chunk, run) x channel split
for v in "-DMBEV_SCATTER_PLANE_ORDER -DMBEV_SCATTER_CS=16" "-DMBEV_SCATTER_PLANE_ORDER -DMBEV_SCATTER_CS=32" "-DMBEV_SCATTER_PLANE_ORDER -DMBEV_SCATTER_CS=64" "-DMBEV_SCATTER_PLANE_ORDER -DMBEV_SCATTER_CS=128" "-DMBEV_SCATTER_PLANE_ORDER -DMBEV_SCATTER_CS=32 -DMBEV_SCATTER_V8"; do
  MBEV_NVCC_EXTRA="$v" python -m mask_bev_b200.build --force > /dev/null 2>&1
  for wl in kitti_b16 waymo_b32; do
  timeout 120 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-layernorm --no-train > gpurun_out/k3o.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/k3o.json')); k=d['kernels']; print('[$v] $wl ms/step %.3f serial %.3f K3 %.4f frac %.3f'%(d['ms_per_step'],d['serial_ms_per_step'],k['K3_scatter']['ms'],k['K3_scatter']['frac_hbm']))"
  done
done
