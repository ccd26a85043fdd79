#!/bin/bash
# first full GPU pass of round 1: parity tests, bench (both arms), ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r01_gpu.txt
nproc >> gpurun_out/r01_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r01_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r01_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err; echo "bench rc=$?" >> gpurun_out/r01_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_ref.json 2> gpurun_out/r01_bench_ref.err
CMD="python bench.py --steps 3 --warmup 3 --extra x25519,p256_mul --extra-steps 1 --no-cpu --no-check"
$CMD > gpurun_out/r01_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/r01_ncu_list.log 2>&1
$CMD > gpurun_out/r01_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_ed25519_mul_base|k_batch_inv|k_x25519|k_wei_mul' -s 8 -c 8 -o gpurun_out/r01_prof $CMD > gpurun_out/r01_ncu_full.log 2>&1
tail -5 gpurun_out/r01_pytest_gpu.log; cat gpurun_out/r01_bench.json | head -c 3000; tail -3 gpurun_out/r01_bench.err
