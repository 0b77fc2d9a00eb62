#!/bin/bash
# round 2, call B: parity of the new multi-kernel pipeline + bench with per-kernel times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/pytest_gpu_b.log
tail -30 gpurun_out/pytest_gpu_b.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['kernel_ms'], d['stats'])
PY
tail -3 gpurun_out/bench_b.err
MJB_SPLIT=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2b_single.json 2>> gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b_single.json').read())
print("single kernel:", {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms'])
PY
