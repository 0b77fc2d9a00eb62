#!/bin/bash
# same-box A/B of compile-time variants (tools/ab_time.py times the raw call, 1M Franka rows)
build() { nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error; }
for k in 1 8 12 16 20 24; do build -DVK_REFILL_MIN=$k -o /tmp/v_r$k.so & done
wait
for r in 1 2; do for k in 1 8 12 16 20 24; do echo "refill_min=$k: $(MJPL_B200_LIB=/tmp/v_r$k.so python tools/ab_time.py | tail -1)"; done; done
