#!/bin/bash
# round 2, step x: MSM with balanced bucket sums (segments of the sorted index array) and running-sum chunks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "msm" > gpurun_out/r2x_pytest.log 2>&1; tail -8 gpurun_out/r2x_pytest.log
timeout 900 python tools/msm_bench.py > gpurun_out/r2x_msm.jsonl 2> gpurun_out/r2x_msm.err; tail -3 gpurun_out/r2x_msm.err; cat gpurun_out/r2x_msm.jsonl
