#!/bin/bash
# second GPU pass: Jacobian/P-256 special reduction; parity + bench + ncu of k_wei_mul (P-256)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01b_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --extra x25519,p256_mul,p256_ecdsa_verify,bls12_381_g1_mul,p384_mul,x448,ed25519_mul > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err; echo "bench rc=$?" >> gpurun_out/r01b_bench.err
for w in 10 12 13 14 16; do timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-check --extra '' --comb-w $w 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('w=$w', d['value'], d['ms_per_step'], d['roofline']['kernels_ms'])" ; done > gpurun_out/r01b_combw.log 2>&1
CMD="python bench.py --workload p256_mul --steps 3 --warmup 3 --extra '' --no-cpu --no-check"
eval $CMD > gpurun_out/r01b_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_wei_mul' -s 2 -c 1 -o gpurun_out/r01b_prof_p256 bash -c "$CMD" > gpurun_out/r01b_ncu.log 2>&1
tail -4 gpurun_out/r01b_pytest_gpu.log; cat gpurun_out/r01b_combw.log; python -c "
import json; d=json.loads(open('gpurun_out/r01b_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['parity_check']); [print(k, v) for k,v in d['workloads'].items()]"
