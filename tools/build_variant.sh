#!/bin/bash
# Build variants/libeccbatch_NAME.so: the in-tree objects with ONE translation unit recompiled under extra flags —
# the A/B builds timed by tools/tune_wei_lib.py and bench.py --lib (window width, blocks per SM, ...).
#   bash tools/build_variant.sh mb6 tu_wei_p256 -DECB_WEI_MINBLOCKS=6
#   bash tools/build_variant.sh win4 tu_wei_p256 -DECB_P256_WIN=4
# variants/ is git-ignored (it still travels to the GPU box with gpurun).  Run `make -C eccoxide_b200/csrc` first.
set -e
NAME=$1; TU=$2; shift 2
cd "$(dirname "$0")/../eccoxide_b200/csrc"
mkdir -p ../../variants build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 128 "$@" -c $TU.cu -o build/variant_${NAME}_$TU.o
OBJS=""
for t in eccbatch tu_ed25519 tu_x25519 tu_x448 tu_wei_p256 tu_wei_p384 tu_wei_bls tu_wei_k256 tu_ecdsa_p256 tu_ecdsa_p384 tu_probe; do
  if [ "$t" = "$TU" ]; then OBJS="$OBJS build/variant_${NAME}_$TU.o"; else OBJS="$OBJS build/$t.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libeccbatch_$NAME.so $OBJS
ls -la ../../variants/libeccbatch_$NAME.so
