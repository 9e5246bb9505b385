mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q --timeout 300 > gpurun_out/dbg_pytest.log 2>&1
tail -3 gpurun_out/dbg_pytest.log
PYTHONPATH=. timeout 300 python scripts/gpu_train_probe.py kitti_b16 2>&1 | grep -v Warn | cut -c1-60,150-215 | head -14
