#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01k_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01k_pytest_gpu.log
for c in 65536 131072 262144 524288; do timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-check --extra '' --opt chunk=$c 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk=$c e2e', d['e2e']['value'], 'dev', d['value'])"; done > gpurun_out/r01k_chunk.log 2>&1
for p in 8 16 32 64; do timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-check --extra '' --opt inv_per_thread=$p 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inv_per_thread=$p', d['value'], d['roofline']['kernels_ms'])"; done > gpurun_out/r01k_inv.log 2>&1
timeout 900 python tools/sweep.py > gpurun_out/r01k_sweep.jsonl 2> gpurun_out/r01k_sweep.err
tail -3 gpurun_out/r01k_pytest_gpu.log; cat gpurun_out/r01k_chunk.log gpurun_out/r01k_inv.log; python -c "
import json
for l in open('gpurun_out/r01k_sweep.jsonl'):
    d=json.loads(l); print(d.get('workload'), d.get('log2_n'), d.get('value'), d.get('ms_per_batch'), d.get('error'))"
