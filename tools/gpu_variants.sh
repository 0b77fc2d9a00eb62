#!/bin/bash
# same-box A/B of compile-time variants (tools/ab_time.py times the raw call, 1M Franka rows)
build() { nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error; }
for k in 0 1; do build -DVK_A_PAIRS2=$k -o /tmp/v_a$k.so & done
wait
for r in 1 2; do for k in 0 1; do echo "single kernel, sphere pairs2=$k: $(MJB_SPLIT=0 MJPL_B200_LIB=/tmp/v_a$k.so python tools/ab_time.py | tail -1)"; done; done
echo "pipeline (committed lib): $(python tools/ab_time.py | tail -1)"
MJB_SPLIT=0 MJPL_B200_LIB=/tmp/v_a1.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_planning.py -x -q 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
