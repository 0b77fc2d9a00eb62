#!/bin/bash
# ncu --set full of the pipeline kernels at a small batch (AB_ROWS), to see what the fixed cost per launch is made of
mkdir -p gpurun_out
for k in ${KERNELS:-narrow_kernel}; do
  AB_ROWS=${AB_ROWS:-250000} timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -o gpurun_out/small_$k -f python tools/ab_kernels.py > gpurun_out/ncu_small_$k.log 2>&1
done
ls -la gpurun_out | grep small
