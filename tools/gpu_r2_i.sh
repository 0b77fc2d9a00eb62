#!/bin/bash
python tools/rowk_crossover.py 2>&1 | tail -11
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_planning.py tests/test_gpu_pose.py -x -q 2>&1 | tail -3
