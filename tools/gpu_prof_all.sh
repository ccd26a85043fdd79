#!/bin/bash
# one ncu --set full capture of the kernels of every workload (one step each); the summary is made
# on the box so that only text has to come back if the report is large
TAG=$1; mkdir -p gpurun_out
CMD="python bench.py --profile-run --steps 1 --warmup 1 --no-cpu --no-check --extra-steps 1 --extra x25519,p256_mul,p256_ecdsa_verify,bls12_381_g1_mul,p384_mul,x448,ed25519_mul,p256_mul_base"
eval $CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none -k regex:'k_ed25519_mul_base|k_batch_inv|k_x25519|k_x448|k_wei_mul|k_ecdsa_main|k_ed25519_mul' -s 4 -c 48 -o /tmp/${TAG}_prof bash -c "$CMD" > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_summary.txt 2>&1
ls -la /tmp/${TAG}_prof.ncu-rep; tail -2 gpurun_out/${TAG}_ncu.log
