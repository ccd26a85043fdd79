#!/bin/bash
# round 2, step k: whole GPU suite, ncu of the headline kernels at the new state, e2e experiments at 2^16
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; tail -5 gpurun_out/r2k_pytest.log
bash tools/gpu_prof_static.sh r02_k "ed25519_mul_base_2p16:k_ed25519_mul_base_fused:n_=_2^16,_fused_small-batch_kernel,_comb_W_=_24 ed25519_mul_base:k_ed25519_mul_base:n_=_2^20,_comb_W_=_24_(11_windows,_8.9_GB_table)"
for o in "ramp=2" "ramp=0"; do
  echo "== e2e 2^16 with $o"; timeout 300 python bench.py --no-cpu --extra '' --steps 10 --opt $o 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['ms_per_batch'])"
done
