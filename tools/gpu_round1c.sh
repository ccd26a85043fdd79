#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01c_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --extra x25519,p256_mul,p256_ecdsa_verify,bls12_381_g1_mul,p384_mul > gpurun_out/r01c_bench.json 2> gpurun_out/r01c_bench.err; echo "bench rc=$?" >> gpurun_out/r01c_bench.err
CMD="python bench.py --workload p256_mul --steps 3 --warmup 3 --extra x25519 --extra-steps 2 --no-cpu --no-check"
eval $CMD > gpurun_out/r01c_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_wei_mul|k_x25519' -s 2 -c 4 -o gpurun_out/r01c_prof bash -c "$CMD" > gpurun_out/r01c_ncu.log 2>&1
tail -4 gpurun_out/r01c_pytest_gpu.log; python -c "
import json; d=json.loads(open('gpurun_out/r01c_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['parity_check']); [print(k, v) for k,v in d['workloads'].items()]"
