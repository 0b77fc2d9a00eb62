#!/bin/bash
build() { nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error; }
build -o /tmp/v_base.so
build -DVK_SCAN_UNROLL=4 -o /tmp/v_scan4.so
build -DVK_SCAN_UNROLL=8 -o /tmp/v_scan8.so
build -DVK_A_UNROLL=2 -o /tmp/v_a2.so
build -DVK_A_UNROLL=4 -o /tmp/v_a4.so
build -DVK_GRP=1 -o /tmp/v_g1.so
build -DVK_GRP=1 -DVK_SCAN_UNROLL=4 -o /tmp/v_g1s4.so
build -DVK_SOLVE_INLINE -o /tmp/v_inl.so
for r in 1 2; do for v in base scan4 scan8 a2 a4 g1 g1s4 inl; do
echo "$v: $(MJPL_B200_LIB=/tmp/v_$v.so python tools/ab_time.py | tail -1)"
done; done
