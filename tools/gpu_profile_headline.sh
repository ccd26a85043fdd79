#!/bin/bash
# launch list of the default bench command + full capture of the headline kernels (small, fast)
TAG=$1; mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-check --extra x25519,p256_mul --extra-steps 1"
eval $CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv bash -c "$CMD" > gpurun_out/${TAG}_list.log 2>&1
CMD2="python bench.py --profile-run --steps 1 --warmup 1 --no-cpu --no-check --extra ''"
eval $CMD2 > gpurun_out/${TAG}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_ed25519_mul_base|k_batch_inv' -s 2 -c 4 -o /tmp/${TAG}_prof bash -c "$CMD2" > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_summary.txt 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[2:]:
    print(r[ix['Kernel Name']][:60], 'grid', r[ix['launch__grid_size']], 'dram_read', r[ix['dram__bytes_read.sum']], rows[1][ix['dram__bytes_read.sum']], 'dram_write', r[ix['dram__bytes_write.sum']], rows[1][ix['dram__bytes_write.sum']], 'dur', r[ix['gpu__time_duration.sum']])
" > gpurun_out/${TAG}_traffic.txt 2>&1
cp /tmp/${TAG}_prof.ncu-rep gpurun_out/ 2>/dev/null; ls -la gpurun_out/${TAG}_prof.ncu-rep
cat gpurun_out/${TAG}_traffic.txt
