#!/bin/bash
# same-box A/B of variants of the validity kernel (tools/ab_time.py times the raw call)
build() { nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error; }
for k in 0 4 8 12 16 24; do build -DVK_SUSPEND=$k -o /tmp/v_s$k.so & done
build -DVK_SUSPEND=16 -DVK_FLUSH_FILL=64 -o /tmp/v_s16f64.so &
build -DVK_SUSPEND=16 -DVK_FLUSH_FILL=112 -o /tmp/v_s16f112.so &
wait
t() { echo "$1: $(env $2 python tools/ab_time.py | tail -1)"; }
for r in 1 2; do
for k in 0 4 8 12 16 24; do t s$k MJPL_B200_LIB=/tmp/v_s$k.so; done
t s16f64 MJPL_B200_LIB=/tmp/v_s16f64.so
t s16f112 MJPL_B200_LIB=/tmp/v_s16f112.so
t s16_items4 "MJPL_B200_LIB=/tmp/v_s16.so MJB_ROUND_ITEMS=4 MJB_ROUND_SPH=8"
t s16_items1.5 "MJPL_B200_LIB=/tmp/v_s16.so MJB_ROUND_ITEMS=1.5 MJB_ROUND_SPH=3"
done
