#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_extra.py plans --queries 4096 2>&1 | tail -3 | tee gpurun_out/bench_plans.jsonl
MAX_ACTIVE=1024 timeout 900 python tools/bench_extra.py plans --queries 4096 2>&1 | tail -1 | tee -a gpurun_out/bench_plans.jsonl
