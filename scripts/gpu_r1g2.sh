mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 45 python scripts/gpu_ln_bwd_probe.py > gpurun_out/r1g_ln_train.log 2>&1
cat gpurun_out/r1g_ln_train.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_ln_bwd_dense -c 1 -f -o gpurun_out/r1g_lnbwd python scripts/gpu_ln_bwd_probe.py ncu > gpurun_out/r1g_ncu.log 2>&1
tail -2 gpurun_out/r1g_ncu.log
