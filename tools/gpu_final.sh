#!/bin/bash
# Final numbers of the round: bench (both arms), launch list, ncu capture of the top kernel, workload benches.
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:validity_kernel -s 3 -c 1 -o gpurun_out/prof_validity python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 python tools/bench_extra.py edges plans poses --queries 4096 2>&1 | grep '^{' > gpurun_out/bench_extra.jsonl
cat gpurun_out/bench_extra.jsonl | cut -c1-400
PYTHONPATH=. timeout 600 python examples/benchmark.py 2>&1 | tail -1 > gpurun_out/bench_config1.json
cat gpurun_out/bench_config1.json
