#!/bin/bash
# round 2, step s: state after the jump-table inversion — whole GPU suite, bench (both arms), ncu of the headline kernels and
# of the constant-time comb, block-level inversion for the wider fields, FP64 probe under ncu, Ed25519 sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2s_pytest.log 2>&1; tail -6 gpurun_out/r2s_pytest.log
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -3 gpurun_out/r2s_bench.err; cut -c1-300 gpurun_out/r2s_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2s_bench_ref.json 2> gpurun_out/r2s_bench_ref.err; cut -c1-200 gpurun_out/r2s_bench_ref.json
bash tools/gpu_prof_static.sh r02_s "ed25519_mul_base_2p16:k_ed25519_mul_base_fused:n_=_2^16,_fused_small-batch_kernel,_comb_W_=_24,_aligned_entries ed25519_mul_base:k_ed25519_mul_base:n_=_2^20,_comb_W_=_24_(11_windows,_11.8_GB_aligned_table) ed25519_keygen:k_ed25519_mul_base_ct:n_=_2^20,_constant-time_comb_(W_=_4,_48_KB_in_shared_memory)"
for o in 1 2; do
  echo "== inv_block=$o"; timeout 300 python bench.py --no-cpu --no-check --workload p256_mul_base --extra bls12_381_g1_mul_base,x25519_base --steps 10 --opt inv_block=$o 2>/dev/null | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['config']['workload'], round(d['value']/1e6,1), d['roofline']['kernels_ms']['scalar_mult'], d['roofline']['kernels_ms']['batch_inversion_encode']); [print(k, round(v['value']/1e6,1), v['roofline']['kernels_ms']['scalar_mult'], v['roofline']['kernels_ms']['batch_inversion_encode']) for k,v in d['workloads'].items()]"
done
ncu --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:k_fieldmul -c 12 --csv --log-file gpurun_out/r2s_fp64_pipes.csv python tools/fieldmul_probe.py > gpurun_out/r2s_fp64_probe_under_ncu.log 2>&1
timeout 600 python tools/sweep.py --workloads ed25519_mul_base > gpurun_out/r2s_sweep_ed25519.jsonl 2> /dev/null; cut -c1-140 gpurun_out/r2s_sweep_ed25519.jsonl
