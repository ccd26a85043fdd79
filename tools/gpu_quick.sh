#!/bin/bash
# quick perf check of selected workloads: bash tools/gpu_quick.sh TAG "wl1,wl2" [ncu-kernel-regex wl]
TAG=$1; WLS=$2; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "wei or ecdsa or bls or p256 or kats" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-check --extra "$WLS" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
if [ -n "$3" ]; then
CMD="python bench.py --workload $4 --steps 3 --warmup 3 --extra '' --no-cpu --no-check"
eval $CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$3" -s 2 -c 1 -o gpurun_out/${TAG}_prof bash -c "$CMD" > gpurun_out/${TAG}_ncu.log 2>&1
fi
tail -3 gpurun_out/${TAG}_pytest.log; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value']); [print(k, v.get('value'), v.get('roofline_frac'), v.get('kernels_ms')) for k,v in d['workloads'].items()]"
