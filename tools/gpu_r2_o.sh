#!/bin/bash
# round 2, step o: constant-time key generation / signing — whole GPU suite, CT vs vartime throughput
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; tail -6 gpurun_out/r2o_pytest.log
timeout 900 python bench.py --no-cpu --workload ed25519_keygen --extra ed25519_keygen_vartime,ed25519_sign,ed25519_sign_vartime,p256_ecdsa_sign,p256_ecdsa_sign_vartime > gpurun_out/r2o_bench_ct.json 2> gpurun_out/r2o_bench_ct.err; tail -3 gpurun_out/r2o_bench_ct.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2o_bench_ct.json') if l.startswith('{')][-1])
print(d['config']['workload'], round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), d['parity_check'], d['roofline']['frac_executed'])
for k,v in d['workloads'].items(): print(k, v.get('error') or (round(v['value']/1e6,1), round(v['e2e']['value']/1e6,1), v['parity_check'], v['roofline']['frac_executed']))
PY
