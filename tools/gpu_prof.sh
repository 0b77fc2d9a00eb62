#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:validity_kernel -s 3 -c 1 -o gpurun_out/prof_validity python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -3
