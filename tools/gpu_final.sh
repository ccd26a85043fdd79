#!/bin/bash
# end-of-round record: tests, both bench arms (headline + x25519 + p256), launch list + full capture of the headline kernels
TAG=$1; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt; nproc >> gpurun_out/${TAG}_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
python bench.py --extra ed25519_mul_base_2p16,x25519,x25519_base,p256_mul,p256_ecdsa_verify,bls12_381_g1_mul,p384_mul,x448,ed25519_mul,ed25519_verify,p256_mul_base,bls12_381_g1_mul_base,p256_decompress,bls12_381_g1_from_compressed,ed25519_keygen,ed25519_sign,p256_ecdsa_sign,p256k1_mul,bls12_381_g1_mul_glv > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
for w in x25519 p256_mul; do python bench.py --workload $w --steps 5 --extra "" > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; python bench.py --impl reference --workload $w --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_$w.json 2>&1; done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-check --extra x25519,p256_mul --extra-steps 1"
eval $CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv bash -c "$CMD" > gpurun_out/${TAG}_list.log 2>&1
CMD2="python bench.py --profile-run --steps 1 --warmup 1 --no-cpu --no-check --extra ''"
eval $CMD2 > gpurun_out/${TAG}_plain2.log 2>&1 && ncu --set full --clock-control none -k regex:'k_ed25519_mul_base|k_batch_inv' -s 2 -c 2 --import-source on -o /tmp/${TAG}_prof bash -c "$CMD2" > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_summary.txt 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[2:]:
    print(r[ix['Kernel Name']][:60], 'grid', r[ix['launch__grid_size']], 'dram_read', r[ix['dram__bytes_read.sum']], rows[1][ix['dram__bytes_read.sum']], 'dram_write', r[ix['dram__bytes_write.sum']], rows[1][ix['dram__bytes_write.sum']], 'dur', r[ix['gpu__time_duration.sum']])
" > gpurun_out/${TAG}_traffic.txt 2>&1
tail -2 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_traffic.txt; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'], d['clocks'], d['parity_check']); [print(k, v.get('value'), v.get('roofline_frac'), v.get('e2e'), v.get('kernels_ms'), v.get('error')) for k,v in d['workloads'].items()]
for w in ('x25519','p256_mul'):
    d=json.loads(open('gpurun_out/${TAG}_bench_%s.json'%w).read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/${TAG}_bench_ref_%s.json'%w).read().strip().splitlines()[-1]); print(w, d['value'], d['e2e']['value'], d['cpu_baseline']['value'], r['value'], d['parity_check'])
"
