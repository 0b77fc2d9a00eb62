#!/bin/bash
# same-box A/B: a commit's kernel sources (tools/ab_prev, made by tools/make_ab_prev.sh) against the working tree
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC"
nvcc $F -o /tmp/prev.so tools/ab_prev/mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error &
nvcc $F -o /tmp/cur.so mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error &
wait
for r in 1 2 3; do
echo "prev: $(MJPL_B200_LIB=/tmp/prev.so python tools/ab_time.py 11 | tail -1)"
echo "cur:  $(MJPL_B200_LIB=/tmp/cur.so python tools/ab_time.py 11 | tail -1)"
done
