#!/bin/bash
# One `ncu --set full` capture per workload (third launch of its main kernel) -> text summary + a JSON fragment for
# profiles/ncu_static.json, and the launch list of the same command (gpu__time_duration only).
#   bash tools/gpu_prof_static.sh TAG "workload:kernel-regex:launch-description ..."     (description: no spaces, _ shown as space)
TAG=$1; shift; mkdir -p gpurun_out; : > gpurun_out/${TAG}_ncu_summary.txt; : > gpurun_out/${TAG}_static.jsonl
for triple in $1; do
  WL=${triple%%:*}; REST=${triple#*:}; RX=${REST%%:*}; DESC=${REST#*:}; DESC=${DESC//_/ }
  CMD="python bench.py --workload $WL --profile-run"
  eval $CMD > gpurun_out/${TAG}_plain_$WL.log 2>&1 || { echo "$WL: plain run failed" >> gpurun_out/${TAG}_ncu_summary.txt; tail -3 gpurun_out/${TAG}_plain_$WL.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 2 -c 1 -o /tmp/${TAG}_$WL bash -c "$CMD" > gpurun_out/${TAG}_ncu_$WL.log 2>&1
  echo "#### workload $WL" >> gpurun_out/${TAG}_ncu_summary.txt
  python tools/ncu_summary.py /tmp/${TAG}_$WL.ncu-rep >> gpurun_out/${TAG}_ncu_summary.txt 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_$WL.ncu-rep --json $WL "$DESC" "profiles/${TAG}_ncu_summary.txt" >> gpurun_out/${TAG}_static.jsonl 2>> gpurun_out/${TAG}_ncu_summary.txt
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_$WL.csv bash -c "$CMD" > /dev/null 2>&1
done
grep -E "^####|^==|duration|fmaheavy|alu pipe|regs/thread|stall reasons|DRAM read|DRAM write|occupancy %" gpurun_out/${TAG}_ncu_summary.txt
