mkdir -p gpurun_out
PYTHONPATH=. timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scatter_ln -s 2 -c 1 -f -o gpurun_out/r1f_ln python scripts/gpu_ln_probe.py 2>&1 | tail -2
