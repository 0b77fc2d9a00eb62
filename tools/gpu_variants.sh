#!/bin/bash
# NB: gpurun_out/ is not shipped to the box; variants are rebuilt there
mkdir -p /tmp/variants
for v in "4 8 4" "2 8 4" "8 8 4" "4 8 8" "4 16 8" "2 8 8"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DVK_GRP=$1 -DVK_Q1=$2 -DVK_Q2=$3 -o /tmp/variants/lib_g$1_q$2_$3.so mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error
  echo "GRP=$1 Q1=$2 Q2=$3: $(MJPL_B200_LIB=/tmp/variants/lib_g$1_q$2_$3.so timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["stats"])')"
done
