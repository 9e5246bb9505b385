for v in "96 112" "88 104"; do
  set -- $v
  MBEV_NVCC_EXTRA="-DMBEV_W2_REGS_LAUNCH=$1 -DMBEV_W2_REGS_EPI=$2" python -m mask_bev_b200.build --force > /dev/null 2>&1
  echo "=== K2 regs launch $1 epi $2"
  for c in 1 2; do timeout 120 python scripts/gpu_corun.py kitti_b16 $c 2>&1 | grep -v "^$" | head -12; done
done
