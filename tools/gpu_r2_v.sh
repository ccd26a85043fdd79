#!/bin/bash
# round 2, step v: P-256 / P-384 field layer with merged small multiples / double subtractions and the rare refold as a branch
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "wei or ecdsa or p256 or p384 or kats or k256 or bls or cpp" > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --workload p256_mul --extra p256_ecdsa_verify,p384_mul,p256_mul_base,bls12_381_g1_mul,p256_ecdsa_sign > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print(d['config']['workload'], round(d['value'] / 1e6, 2), d['roofline'].get('kernels_ms'), d.get('parity_check'))
for k, v in d['workloads'].items():
    print(k, round(v['value'] / 1e6, 2), v['roofline'].get('kernels_ms'), v.get('parity_check'))
PY
