#!/bin/bash
# same-box A/B of compile-time variants: tools/gpu_ab2.sh "name1:-DFLAG=1 -DX=2" "name2:..."
build() { nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep -E "error" ; }
for v in "$@"; do name="${v%%:*}"; flags="${v#*:}"; build $flags -o /tmp/v_$name.so & done
wait
for v in "$@"; do name="${v%%:*}"; echo "== $name (${v#*:})"; MJPL_B200_LIB=/tmp/v_$name.so timeout 300 python tools/ab_kernels.py 2>&1 | tail -1; done
