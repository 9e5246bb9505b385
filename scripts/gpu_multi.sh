mkdir -p gpurun_out
N=${N:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/multi_$N.json 2> gpurun_out/multi_$N.err
tail -3 gpurun_out/multi_$N.err
python -c "
import json;d=json.load(open('gpurun_out/multi_$N.json'));print('N', d['n_gpus'], 'fps', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'serial', round(d['e2e']['serial_value']), d['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/multi_ref_$N.json 2> gpurun_out/multi_ref_$N.err
tail -2 gpurun_out/multi_ref_$N.err; cut -c1-300 gpurun_out/multi_ref_$N.json
