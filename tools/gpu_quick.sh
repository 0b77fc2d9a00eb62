#!/bin/bash
# quick GPU check: parity tests + raw timing of the committed library
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_planning.py -x -q 2>&1 | tail -4
python tools/ab_time.py | tail -1
python tools/ab_time.py | tail -1
