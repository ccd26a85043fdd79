#!/bin/bash
# round 2, step m: the new large-size parity tests, the config-5 sweep on one GPU, single-process bench on one GPU
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_bench_contract.py -m gpu -q -k "2p18 or 2p22 or above_2p20 or never_allocate or counts_launches or contract_line or all_visible" -rs > gpurun_out/r2m_pytest.log 2>&1; tail -8 gpurun_out/r2m_pytest.log
timeout 900 python tools/sweep.py > gpurun_out/r2m_sweep_1.jsonl 2> gpurun_out/r2m_sweep.err; tail -2 gpurun_out/r2m_sweep.err; cut -c1-160 gpurun_out/r2m_sweep_1.jsonl
timeout 600 python bench.py --single-process --gpus 1 --steps 10 --workload ed25519_mul_base --extra p256_mul,ed25519_keygen > gpurun_out/r2m_single_process_1.json 2> gpurun_out/r2m_sp.err; tail -2 gpurun_out/r2m_sp.err; cut -c1-600 gpurun_out/r2m_single_process_1.json
