mkdir -p gpurun_out
export PYTHONPATH=$PWD
( timeout 70 python -m pytest tests/test_gpu_layernorm_backward.py -q 2>&1 | tail -15
  echo "--- MBEV_LN_BWD=0"
  MBEV_LN_BWD=0 timeout 40 python -m pytest tests/test_gpu_layernorm_backward.py -q 2>&1 | tail -5
  echo "--- probe default"
  timeout 30 python scripts/gpu_ln_bwd_probe.py ncu 2>&1 | tail -2
  echo "--- probe MBEV_LN_BWD=0"
  MBEV_LN_BWD=0 timeout 30 python scripts/gpu_ln_bwd_probe.py ncu 2>&1 | tail -2 ) > gpurun_out/r1g3.log 2>&1
cat gpurun_out/r1g3.log
timeout 40 ncu --set full --clock-control none --import-source on -k regex:k_ln_bwd_dense -c 1 -f -o gpurun_out/r1g_lnbwd_async python scripts/gpu_ln_bwd_probe.py ncu > gpurun_out/r1g3_ncu.log 2>&1
tail -2 gpurun_out/r1g3_ncu.log
