#!/bin/bash
# round 2, step p: secp256k1, BLS small-order inputs, Ed25519 decompress, BLS uncompressed — new GPU tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "p256k1 or small_order or decompress or uncompressed or wei_mul or batch_inversion or standard_encodings" > gpurun_out/r2p_pytest.log 2>&1; tail -6 gpurun_out/r2p_pytest.log
