#!/bin/bash
# round 2, step u: multi-scalar multiplication, ristretto255 on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "msm or ristretto" > gpurun_out/r2u_pytest.log 2>&1; tail -8 gpurun_out/r2u_pytest.log
timeout 600 python tools/msm_bench.py > gpurun_out/r2u_msm.jsonl 2> gpurun_out/r2u_msm.err; tail -3 gpurun_out/r2u_msm.err; cat gpurun_out/r2u_msm.jsonl
