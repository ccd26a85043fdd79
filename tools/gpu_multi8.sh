#!/bin/bash
# scaling check on one N-GPU box: bash tools/gpu_multi8.sh N
N=$1; mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_${N}_gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --extra x25519,p256_mul --extra-steps 3 > gpurun_out/multi_${N}_bench.json 2> gpurun_out/multi_${N}_bench.err; echo "rc=$?" >> gpurun_out/multi_${N}_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu --extra x25519,p256_mul --extra-steps 3 > gpurun_out/multi_${N}_same_box_1gpu.json 2> gpurun_out/multi_1_bench.err
timeout 300 python -m pytest tests -m gpu -x -q -k "all_visible_devices" 2>&1 | tail -2
tail -3 gpurun_out/multi_${N}_bench.err; python -c "
import json
for f in ('gpurun_out/multi_${N}_bench.json','gpurun_out/multi_${N}_same_box_1gpu.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['value'], d['e2e']['value'], d['clocks'], d.get('workloads') and {k:(v.get('value'), v.get('e2e')) for k,v in d['workloads'].items()})
    except Exception as e: print(f, 'ERR', e)
"
