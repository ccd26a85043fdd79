#!/bin/bash
# pipeline chunk-schedule comparison: bash tools/gpu_ramp.sh
timeout 900 python -m pytest tests -m gpu -x -q -k "chunks or config1 or ragged or shards or device_resident" 2>&1 | tail -3
for cfg in "ramp=0" "ramp=1" "ramp=2" "ramp=3" "ramp=2,chunk=151552" "ramp=3,chunk=236800"; do
  OPTS=""; for kv in ${cfg//,/ }; do OPTS="$OPTS --opt $kv"; done
  echo "== $cfg"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-check $OPTS --extra-steps 3 --extra ed25519_mul_base_2p16,p256_mul_base,x25519_base 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' headline', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), d['e2e']['pcie']); [print(' ', k, round(v.get('value',0)/1e6,1), 'e2e', round(v.get('e2e',0)/1e6,1), v.get('error')) for k,v in d['workloads'].items()]"
done
