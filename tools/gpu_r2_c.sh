#!/bin/bash
mkdir -p gpurun_out
python tools/latency_probe.py > gpurun_out/r2c_latency.jsonl 2>&1; cat gpurun_out/r2c_latency.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or stride or ed25519_mul_base or x25519_base" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
timeout 900 python tools/tune_ed25519.py --w 24,26 --stride 24 --logs 10,12,14,16,20 > gpurun_out/r2c_tune.jsonl 2> gpurun_out/r2c_tune.err
tail -3 gpurun_out/r2c_tune.err
cat gpurun_out/r2c_tune.jsonl | cut -c1-200
