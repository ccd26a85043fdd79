#!/bin/bash
# round 2, step z: secp256k1 on its pseudo-Mersenne field (mont.cuh kind 3); whole GPU suite; MSM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; tail -5 gpurun_out/r2z_pytest.log
timeout 300 python tools/tune_wei_lib.py --curve p256k1 --log 20 --steps 10 > gpurun_out/r2z_tune_k256.jsonl 2> gpurun_out/r2z_tune.err; cat gpurun_out/r2z_tune_k256.jsonl
timeout 900 python tools/msm_bench.py > gpurun_out/r2z_msm.jsonl 2> gpurun_out/r2z_msm.err; tail -3 gpurun_out/r2z_msm.err; cut -c1-200 gpurun_out/r2z_msm.jsonl
