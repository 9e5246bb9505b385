mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layernorm.py -m gpu -q --timeout 200 > gpurun_out/dbg_pytest.log 2>&1
tail -25 gpurun_out/dbg_pytest.log
