#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/bench_multi.py --rows 250000000 --queries 2048 2>&1 | tail -2 | tee gpurun_out/bench_multi_n2.json
