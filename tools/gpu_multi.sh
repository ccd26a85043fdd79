#!/bin/bash
# bench.py under torchrun at N ranks (one process per GPU), both arms
N=$1; mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_${N}_gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/multi_${N}_bench.json 2> gpurun_out/multi_${N}_bench.err; echo "rc=$?" >> gpurun_out/multi_${N}_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/multi_${N}_ref.json 2> gpurun_out/multi_${N}_ref.err; echo "rc=$?" >> gpurun_out/multi_${N}_ref.err
python bench.py --steps 10 --warmup 3 --extra '' --no-cpu > gpurun_out/multi_1_bench.json 2> gpurun_out/multi_1_bench.err
tail -3 gpurun_out/multi_${N}_bench.err; python -c "
import json
for f in ('gpurun_out/multi_${N}_bench.json','gpurun_out/multi_1_bench.json','gpurun_out/multi_${N}_ref.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['value'], d['e2e']['value'], d.get('workloads') and {k:v.get('value') for k,v in d['workloads'].items()})
    except Exception as e: print(f, 'ERR', e)
"
