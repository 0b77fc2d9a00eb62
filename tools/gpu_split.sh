#!/bin/bash
# two-kernel pipeline (MJB_SPLIT=1): parity suites, then timing against the single kernel on the same box
MJB_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_planning.py -x -q 2>&1 | tail -15
for r in 1 2; do
echo "single: $(python tools/ab_time.py | tail -1)"
echo "split:  $(MJB_SPLIT=1 python tools/ab_time.py | tail -1)"
done
MJB_SPLIT=1 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -s 30 -c 3 --csv --log-file gpurun_out/split_metrics.csv python tools/ab_time.py > /dev/null 2>&1
