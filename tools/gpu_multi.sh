#!/bin/bash
# Everything measured on an N-GPU box (bash tools/gpu_multi.sh N TAG): weak scaling (the driver's contract), strong scaling
# of one 2^20 batch, the library's own single-process multi-device path, the config-5 sweep sliced over N ranks,
# and the multi-device parity test.  One GPU of the same box is measured beside each.
N=$1; TAG=$2; mkdir -p gpurun_out; O=gpurun_out/${TAG}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L > ${O}_gpus.txt; nproc >> ${O}_gpus.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "all_visible_devices" > ${O}_pytest_multidev.txt 2>&1; tail -2 ${O}_pytest_multidev.txt
$TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --extra ed25519_mul_base,p256_mul,x25519 > ${O}_weak_${N}.json 2> ${O}_weak_${N}.err; echo "weak rc=$?"
python bench.py --steps 20 --warmup 3 --no-cpu --extra ed25519_mul_base,p256_mul,x25519 > ${O}_weak_1.json 2> ${O}_weak_1.err
$TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --scaling strong --workload ed25519_mul_base --extra p256_mul,x25519 > ${O}_strong_${N}.json 2> ${O}_strong_${N}.err; echo "strong rc=$?"
for g in 1 2 4 8; do
  [ $g -le $N ] || continue
  timeout 600 python bench.py --single-process --gpus $g --steps 10 --workload ed25519_mul_base --extra p256_mul,x25519,ed25519_keygen > ${O}_single_process_${g}.json 2> ${O}_single_process_${g}.err; echo "single-process $g rc=$?"
done
timeout 900 $TR tools/sweep.py --workloads ed25519_mul_base,p256_mul,p384_mul,x448 > ${O}_sweep_${N}.jsonl 2> ${O}_sweep_${N}.err; echo "sweep rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("${O}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['scaling'], round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), {k:(round(v['value']/1e6,1), round((v.get('e2e') or {}).get('value',0)/1e6,1) if isinstance(v.get('e2e'),dict) else None) for k,v in (d.get('workloads') or {}).items() if 'value' in v})
    except Exception as e: print(f,'ERR',e)
PY
