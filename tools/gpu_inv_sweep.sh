#!/bin/bash
# batch-inversion launch-shape sweep: bash tools/gpu_inv_sweep.sh TAG
TAG=$1; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for cfg in "inv_block=0" "inv_block=1,inv_per_thread=32" "inv_block=1,inv_per_thread=16" "inv_block=1,inv_per_thread=8" "inv_block=1,inv_per_thread=4" "inv_block=1,inv_per_thread=8,inv_fill_per_sm=512" "inv_block=1,inv_per_thread=8,inv_fill_per_sm=1024" "inv_block=1,inv_per_thread=16,inv_fill_per_sm=512" "inv_block=1,inv_per_thread=12" ; do
  OPTS=""; for kv in ${cfg//,/ }; do OPTS="$OPTS --opt $kv"; done
  echo "== $cfg"
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-check $OPTS --extra ed25519_mul_base_2p16,p256_mul_base,bls12_381_g1_mul_base 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' headline', round(d['value']/1e6,1), d['roofline']['kernels_ms'], 'e2e', round(d['e2e']['value']/1e6,1)); [print(' ', k, round(v.get('value',0)/1e6,1), v.get('kernels_ms')) for k,v in d['workloads'].items()]"
done
