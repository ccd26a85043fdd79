#!/bin/bash
# round 2, step j: whole GPU suite + the restructured bench (headline = configs[0], 2^16) + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; tail -5 gpurun_out/r2j_pytest.log
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -3 gpurun_out/r2j_bench.err; cut -c1-1500 gpurun_out/r2j_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; cat gpurun_out/r2j_bench_ref.json
