#!/bin/bash
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC -o /tmp/lib_prev.so tools/ab_prev/mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error
for i in 1 2 3; do
echo "prev: $(MJPL_B200_LIB=/tmp/lib_prev.so python tools/ab_time.py | tail -1)"
echo "cur : $(python tools/ab_time.py | tail -1)"
done
