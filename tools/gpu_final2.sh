#!/bin/bash
# round 2, final session: every GPU test, smoke, bench (both arms); the ncu captures come from tools/gpu_r2_prof.sh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -3 gpurun_out/bench_r2.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2>> gpurun_out/bench_r2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2.json').read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('bound','achieved','peak','frac','traffic','kernel','kernel_ms','step','issue')})
print('plans', {k:v for k,v in d['plans'].items() if k not in ('what',)})
print('sweep', {k:v for k,v in d['sweep'].items() if k not in ('what',)})
print('cpu', d.get('cpu_baseline'))
r=json.loads(open('gpurun_out/bench_r2_ref.json').read()); print('ref arm', r['value'], r['config']==d['config'])
PY
