#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planning.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 900 python tools/bench_extra.py plans --queries 4096 2>&1 | tail -2 | tee gpurun_out/bench_plans.jsonl
MAX_ACTIVE=4096 timeout 900 python tools/bench_extra.py plans --queries 32768 2>&1 | tail -1 | tee -a gpurun_out/bench_plans.jsonl
