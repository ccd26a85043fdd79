#!/bin/bash
# full capture of the P-256 variable-base kernel: bash tools/gpu_profile_p256.sh TAG [workload] [kernel-regex]
TAG=$1; WL=${2:-p256_mul}; RX=${3:-k_wei_mul}; mkdir -p gpurun_out
CMD2="python bench.py --workload $WL --profile-run --steps 1 --warmup 1 --no-cpu --no-check --extra ''"
eval $CMD2 > gpurun_out/${TAG}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 2 -c 1 -o /tmp/${TAG}_prof bash -c "$CMD2" > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_summary.txt 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[2:]:
    print(r[ix['Kernel Name']][:60], 'grid', r[ix['launch__grid_size']], 'dram_read', r[ix['dram__bytes_read.sum']], rows[1][ix['dram__bytes_read.sum']], 'dram_write', r[ix['dram__bytes_write.sum']], rows[1][ix['dram__bytes_write.sum']], 'dur', r[ix['gpu__time_duration.sum']])
" > gpurun_out/${TAG}_traffic.txt 2>&1
ls -la /tmp/${TAG}_prof.ncu-rep gpurun_out/${TAG}_source.csv
cat gpurun_out/${TAG}_ncu_summary.txt gpurun_out/${TAG}_traffic.txt
