#!/bin/bash
# round 2, step l: whole GPU suite, bench (both arms), entry-stride comparison (tune + ncu DRAM bytes)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; tail -8 gpurun_out/r2l_pytest.log
timeout 900 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -3 gpurun_out/r2l_bench.err; cut -c1-400 gpurun_out/r2l_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err; cut -c1-300 gpurun_out/r2l_bench_ref.json
timeout 900 python tools/tune_ed25519.py --w 24 --stride 24,32 --logs 16,20 > gpurun_out/r2l_tune_stride.jsonl 2> gpurun_out/r2l_tune.err; cat gpurun_out/r2l_tune_stride.jsonl | cut -c1-200
CMD="python bench.py --workload ed25519_mul_base --profile-run --opt ed25519_entry_stride=32"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_ed25519_mul_base -s 2 -c 1 --csv --log-file gpurun_out/r2l_stride32_dram.csv bash -c "$CMD" > /dev/null 2>&1
grep -v "^==" gpurun_out/r2l_stride32_dram.csv | cut -d, -f5,13,14,15 | tail -4
