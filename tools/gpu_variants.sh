#!/bin/bash
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC -o /tmp/lib_prev.so tools/ab_prev/mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error
for i in 1 2; do
echo "prev: $(MJPL_B200_LIB=/tmp/lib_prev.so timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])')"
echo "cur : $(timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])')"
done
