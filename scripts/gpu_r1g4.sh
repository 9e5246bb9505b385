mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 50 python -m pytest tests/test_gpu_layernorm_backward.py -q 2>&1 | tail -8 > gpurun_out/r1g4_pytest.log
cat gpurun_out/r1g4_pytest.log
timeout 60 python bench.py --steps 50 --warmup 10 > gpurun_out/r1g4_bench.json 2> gpurun_out/r1g4_bench.err
tail -2 gpurun_out/r1g4_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r1g4_bench.json'));print(d['value'], d['e2e']['value'], d['layernorm_f1'])"
