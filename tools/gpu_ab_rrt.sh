#!/bin/bash
# same-box A/B of compile-time variants on the planner: tools/gpu_ab_rrt.sh "name1:-DFLAG" "name2:"
build() { nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep -E "error" ; }
for v in "$@"; do name="${v%%:*}"; flags="${v#*:}"; build $flags -o /tmp/v_$name.so & done
wait
for rep in 1 2; do for v in "$@"; do name="${v%%:*}"; echo "== $name"; MJPL_B200_LIB=/tmp/v_$name.so timeout 300 python tools/rrt_time.py 2>&1 | tail -3 | cut -c1-60; done; done
