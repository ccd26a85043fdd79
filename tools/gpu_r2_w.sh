#!/bin/bash
# round 2, step w: A/B of builds of the Weierstrass variable-base kernels (blocks per SM, window width); the variants/ libraries
# come from tools/build_variant.sh (e.g. `build_variant.sh p256_mb6 tu_wei_p256 -DECB_WEI_MINBLOCKS=6`); see tools/tune_wei_lib.py
mkdir -p gpurun_out
D=eccoxide_b200/libeccbatch.so
O=gpurun_out/r2w2_tune_wei_lib.jsonl; : > $O
timeout 300 python tools/tune_wei_lib.py --libs $D,variants/libeccbatch_p256_mb6.so --steps 10 >> $O 2> gpurun_out/r2w2_tune.err
timeout 300 python tools/tune_wei_lib.py --libs $D,variants/libeccbatch_others_win5.so --curve p384r1 --log 18 --steps 10 >> $O 2>> gpurun_out/r2w2_tune.err
timeout 300 python tools/tune_wei_lib.py --libs $D,variants/libeccbatch_others_win5.so --curve bls12_381_g1 --log 18 --steps 10 >> $O 2>> gpurun_out/r2w2_tune.err
timeout 300 python tools/tune_wei_lib.py --libs $D,variants/libeccbatch_others_win5.so --curve p256k1 --log 20 --steps 10 >> $O 2>> gpurun_out/r2w2_tune.err
tail -2 gpurun_out/r2w2_tune.err; cat $O
for l in $D variants/libeccbatch_ecdsa_mb5.so; do
timeout 300 python bench.py --lib $l --steps 10 --warmup 3 --no-cpu --workload p256_ecdsa_verify --extra '' 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$l', d['config']['workload'], round(d['value'] / 1e6, 2), d.get('parity_check'))"
done
