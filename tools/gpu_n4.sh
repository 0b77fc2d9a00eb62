#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n4.json | cut -c1-300
