#!/bin/bash
# same-box A/B of compile-time variants (tools/ab_time.py times the raw call, 1M Franka rows)
build() { nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error; }
build -o /tmp/v_base.so &
build -DVK_NARROW_THREADS=128 -DVK_NARROW_CTAS=6 -o /tmp/v_128x6.so &
build -DVK_NARROW_THREADS=128 -DVK_NARROW_CTAS=7 -o /tmp/v_128x7.so &
build -DVK_NARROW_THREADS=128 -DVK_NARROW_CTAS=5 -o /tmp/v_128x5.so &
build -DVK_NARROW_THREADS=64 -DVK_NARROW_CTAS=13 -o /tmp/v_64x13.so &
build -DVK_NARROW_THREADS=512 -DVK_NARROW_CTAS=1 -o /tmp/v_512x1.so &
wait
for r in 1 2; do for v in base 128x5 128x6 128x7 64x13 512x1; do echo "narrow $v: $(MJPL_B200_LIB=/tmp/v_$v.so python tools/ab_time.py | tail -1)"; done; done
