#!/bin/bash
# round 2, call A: every GPU test (new parity tests included), smoke, baseline bench of the round-1 kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench.err
cut -c1-1500 gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench.err
nproc; lscpu | grep -E "Model name|Socket|NUMA" | head -8
