#!/bin/bash
mkdir -p gpurun_out
python tools/split_crossover.py 2>&1 | tail -9
timeout 900 python bench.py --steps 30 --warmup 5 --no-plans --no-sweep --no-cpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_d.err
tail -5 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2d.json').read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
print('kernel_ms', d['roofline']['kernel_ms'])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
