#!/bin/bash
# round 2, step aa: persistent kernels fetch their work by warp from a counter (no tail round)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "wei or ecdsa or p256 or p384 or kats or k256 or bls or ed25519_mul or verify or ristretto or cpp or fullsize" > gpurun_out/r2aa_pytest.log 2>&1; tail -3 gpurun_out/r2aa_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --workload p256_mul --extra p256_ecdsa_verify,p384_mul,bls12_381_g1_mul,bls12_381_g1_mul_glv,p256k1_mul,ed25519_mul,ed25519_verify > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; tail -3 gpurun_out/r2aa_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r2aa_bench.json').read().strip().splitlines()[-1])
print(d['config']['workload'], round(d['value'] / 1e6, 2), d.get('parity_check'))
for k, v in d['workloads'].items():
    print(k, round(v['value'] / 1e6, 2), v.get('parity_check'))
PY
