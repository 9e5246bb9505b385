PYTHONPATH=. timeout 300 python scripts/gpu_ln_probe.py 2>&1 | tail -3
