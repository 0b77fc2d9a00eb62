#!/bin/bash
# same-box A/B: smallest adaptive tile (rows per warp) for small batches, single kernel
build() { nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error; }
for k in 8 4; do build -DVK_MIN_TILE_ROWS=$k -o /tmp/v_m$k.so & done
wait
for k in 8 4; do echo "min_tile_rows=$k"; MJPL_B200_LIB=/tmp/v_m$k.so timeout 300 python tools/split_crossover.py 2>&1 | grep -E "rows +(4096|8192|16384)"; done
for k in 8 4; do MJPL_B200_LIB=/tmp/v_m$k.so timeout 300 python tools/bench_extra.py plans --queries 4096 2>&1 | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('min_tile_rows=$k plans/s', round(d['plans_per_s']), 'solved', d['solved'], 'replay failures', d['replay_failures'])"; done
MJPL_B200_LIB=/tmp/v_m4.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_planning.py -x -q 2>&1 | tail -1
