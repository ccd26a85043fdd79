#!/bin/bash
# round 2, step q: jump-table divsteps in the warp-level inversion — latency, parity, phase trace, launch-shape sweep
mkdir -p gpurun_out
python tools/latency_probe.py > gpurun_out/r2q_latency.jsonl 2>&1; cat gpurun_out/r2q_latency.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or batch_inversion or ed25519_mul_base or x25519_base or hostsim" > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
python tools/fused_trace.py > gpurun_out/r2q_trace.jsonl 2>&1; cat gpurun_out/r2q_trace.jsonl
timeout 900 python tools/tune_ed25519.py --w 26 --stride 24 --logs 10,12,14,15,16,20 > gpurun_out/r2q_tune.jsonl 2> gpurun_out/r2q_tune.err
tail -3 gpurun_out/r2q_tune.err
cat gpurun_out/r2q_tune.jsonl | cut -c1-200
