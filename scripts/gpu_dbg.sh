mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/dbg_pytest.log 2>&1
tail -3 gpurun_out/dbg_pytest.log
PYTHONPATH=. timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_scatter_ln -s 2 -c 1 -f -o gpurun_out/r1f_ln python scripts/gpu_ln_probe.py 2>&1 | tail -1
