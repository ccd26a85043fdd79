#!/bin/bash
# round 2, call A: fused small-batch kernel — parity first, then the launch-shape sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or stride or ed25519_mul_base or x25519_base or native" > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python tools/tune_ed25519.py --w 24,26 --stride 24 > gpurun_out/r2a_tune.jsonl 2> gpurun_out/r2a_tune.err
timeout 600 python tools/tune_ed25519.py --w 20,24 --stride 32 --logs 12,16,20 >> gpurun_out/r2a_tune.jsonl 2>> gpurun_out/r2a_tune.err
tail -3 gpurun_out/r2a_tune.err
cat gpurun_out/r2a_tune.jsonl | cut -c1-200
