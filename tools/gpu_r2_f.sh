#!/bin/bash
mkdir -p gpurun_out
python tools/fused_trace.py > gpurun_out/r2f_trace.jsonl 2>&1; cat gpurun_out/r2f_trace.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused" > gpurun_out/r2f_pytest.log 2>&1; tail -2 gpurun_out/r2f_pytest.log
timeout 900 python tools/tune_ed25519.py --w 26 --stride 24 --logs 10,12,14,15,16 > gpurun_out/r2f_tune.jsonl 2> gpurun_out/r2f_tune.err
tail -3 gpurun_out/r2f_tune.err
cat gpurun_out/r2f_tune.jsonl | cut -c1-200
