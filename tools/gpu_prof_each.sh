#!/bin/bash
# one ncu --set full capture per workload (second launch of its main kernel), summaries only come back:
#   bash tools/gpu_prof_each.sh TAG "workload:kernel-regex workload:kernel-regex ..."
TAG=$1; shift; mkdir -p gpurun_out; : > gpurun_out/${TAG}_ncu_each_summary.txt
for pair in $1; do
  WL=${pair%%:*}; RX=${pair##*:}
  CMD="python bench.py --workload $WL --profile-run --steps 1 --warmup 1 --no-cpu --no-check --extra ''"
  eval $CMD > gpurun_out/${TAG}_plain_$WL.log 2>&1 || { echo "$WL: plain run failed" >> gpurun_out/${TAG}_ncu_each_summary.txt; continue; }
  ncu --set full --clock-control none -k regex:"$RX" -s 1 -c 1 -o /tmp/${TAG}_$WL bash -c "$CMD" > gpurun_out/${TAG}_ncu_$WL.log 2>&1
  echo "#### workload $WL" >> gpurun_out/${TAG}_ncu_each_summary.txt
  python tools/ncu_summary.py /tmp/${TAG}_$WL.ncu-rep >> gpurun_out/${TAG}_ncu_each_summary.txt 2>&1
done
grep -E "^####|^==|duration|fmaheavy|alu pipe|regs/thread|stall reasons" gpurun_out/${TAG}_ncu_each_summary.txt
