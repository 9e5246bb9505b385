# final-tree run of round 1: bench line, full GPU parity suite, LN backward probe + ncu capture
mkdir -p gpurun_out
timeout 120 python bench.py --steps 50 --warmup 10 > gpurun_out/r1g_bench.json 2> gpurun_out/r1g_bench.err
tail -2 gpurun_out/r1g_bench.err
timeout 140 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -12 > gpurun_out/r1g_pytest.log
cat gpurun_out/r1g_pytest.log
timeout 45 python scripts/gpu_ln_bwd_probe.py > gpurun_out/r1g_ln_train.log 2>&1
cat gpurun_out/r1g_ln_train.log
timeout 50 ncu --set full --clock-control none --import-source on -k regex:k_ln_bwd_dense -c 1 -f -o gpurun_out/r1g_lnbwd python scripts/gpu_ln_bwd_probe.py ncu > gpurun_out/r1g_ncu.log 2>&1
tail -2 gpurun_out/r1g_ncu.log
