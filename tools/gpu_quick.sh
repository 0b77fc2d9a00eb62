#!/bin/bash
# quick GPU check: parity tests + bench (no cpu baseline) + launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -3
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
grep -E "validity|recheck" gpurun_out/launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -8
