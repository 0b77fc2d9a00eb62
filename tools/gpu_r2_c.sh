#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_c.err
tail -5 gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2c.json').read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('peak_source','flop_source')})
print('plans', d.get('plans'))
print('sweep', d.get('sweep'))
print('cpu', d.get('cpu_baseline'), d['stats'])
PY
for sp in 1 0; do for n in 4096; do echo "== split=$sp rows=$n"; MJB_SPLIT=$sp ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 8 --csv python tools/dbg/small_batch.py $n 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]: print('  ', r[ki][:50], r[vi])
"; done; done
