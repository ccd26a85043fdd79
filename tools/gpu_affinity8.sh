#!/bin/bash
# 8-GPU box: does binding each rank to its GPU's CPUs (bench.py bind_to_gpu_numa_node) change the per-rank copy rates?
N=$1; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
nvidia-smi topo -m > gpurun_out/r02_t_topo.txt 2>&1
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-check --extra ed25519_mul_base > gpurun_out/r02_t_weak_${N}_bound.json 2> gpurun_out/r02_t_bound.err; echo "bound rc=$?"
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-check --no-bind --extra ed25519_mul_base > gpurun_out/r02_t_weak_${N}_unbound.json 2> gpurun_out/r02_t_unbound.err; echo "unbound rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r02_t_weak_${N}_bound.json","gpurun_out/r02_t_weak_${N}_unbound.json"):
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, round(d['value']/1e6), round(d['e2e']['value']/1e6), d['cpu_affinity'], d['e2e']['pcie'].get('per_rank_simultaneous_h2d_d2h_duplex_gbs'), round(d['workloads']['ed25519_mul_base']['e2e']['value']/1e6))
PY
head -14 gpurun_out/r02_t_topo.txt
