#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 30 --warmup 5 --no-plans --no-sweep --no-cpu-baseline > gpurun_out/bench_r2e.json 2> gpurun_out/bench_e.err
tail -5 gpurun_out/bench_e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2e.json').read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
print('kernel_ms', d['roofline']['kernel_ms'])
PY
for sp in 1 0; do for n in 4096 32768; do echo "== split=$sp rows=$n"; MJB_SPLIT=$sp ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 6 --csv python tools/dbg/small_batch.py $n 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
if rows:
  h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
  for r in rows[1:]: print('  ', r[ki][:50], r[vi])
"; done; done
