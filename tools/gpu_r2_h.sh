#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "min_distance" 2>&1 | tail -6
timeout 600 python - <<'PY'
import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import mjpl_b200 as mj
from tests.test_gpu_pose import _constrained_problem
for nqs in (1024, 4096):
    model, allowed, joints, q_init, ref, lim, cons, goals = _constrained_problem(nqs)
    for kw in ({"sync_every": 8}, {"sync_every": 32}):
        pl = mj.BatchedRRT(model, joints, cons, max_planning_time=120, epsilon=0.05, seed=17, goal_biasing_probability=0.1, **kw)
        pl.plan(np.tile(q_init, (4, 1)), goals[:4]); torch.cuda.synchronize()
        t0 = time.perf_counter(); paths = pl.plan(np.tile(q_init, (len(goals), 1)), goals); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ok = sum(1 for p in paths if p)
        print(nqs, kw, "plans/s", ok / dt, "seconds", dt, {k: pl.stats[k] for k in ("ticks", "iterations", "solved", "gave_up", "host_syncs", "configs_checked")})
PY
timeout 900 python bench.py --steps 20 --warmup 5 --no-plans --no-cpu-baseline > gpurun_out/bench_r2h.json 2> gpurun_out/bench_h.err; tail -3 gpurun_out/bench_h.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2h.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value']); print('sweep', d.get('sweep'))
PY
