#!/bin/bash
# IK kernel: gpu tests + the move-to-pose bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ik.py -x -q -s 2>&1 | tail -25 | tee gpurun_out/pytest_ik.log
timeout 600 python tools/bench_extra.py poses --queries 4096 2>&1 | tail -3 | tee gpurun_out/bench_poses.json
