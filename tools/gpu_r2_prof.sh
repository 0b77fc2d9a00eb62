#!/bin/bash
# ncu: launch list of one bench run + full captures of the pipeline kernels named in $1 (regex list)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
grep -c . gpurun_out/launches.csv
for k in ${1:-fk_cull_kernel mid_kernel narrow_kernel}; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_$k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
done
ls -la gpurun_out | tail -6
