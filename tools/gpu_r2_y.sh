#!/bin/bash
# round 2, step y: BLS12-381 G1 Point::mul through the endomorphism (option bls12_381_g1_glv) beside the plain kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "endomorphism or wei_mul or bls" > gpurun_out/r2y_pytest.log 2>&1; tail -4 gpurun_out/r2y_pytest.log
for o in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload bls12_381_g1_mul --extra '' --opt bls12_381_g1_glv=$o > gpurun_out/r2y_bench_glv$o.json 2> gpurun_out/r2y_bench_glv$o.err
python -c "
import json
d = json.loads(open('gpurun_out/r2y_bench_glv$o.json').read().strip().splitlines()[-1]); print('glv=$o', d['config']['workload'], round(d['value'] / 1e6, 2), 'e2e', round(d['e2e']['value'] / 1e6, 2), d['roofline'].get('kernels_ms', {}).get('scalar_mult'), d.get('parity_check'))"
done
