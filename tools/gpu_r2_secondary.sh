#!/bin/bash
# round 2, secondary numbers of the final build: constrained planner (configs[3]), UR5e edges (configs[2]), small launches
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r2_constrained.txt
import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import mjpl_b200 as mj
from tests.test_gpu_pose import _constrained_problem
for nqs in (1024, 4096):
    model, allowed, joints, q_init, ref, lim, cons, goals = _constrained_problem(nqs)
    pl = mj.BatchedRRT(model, joints, cons, max_planning_time=120, epsilon=0.05, seed=17, goal_biasing_probability=0.1, sync_every=32)
    pl.plan(np.tile(q_init, (4, 1)), goals[:4]); torch.cuda.synchronize()
    t0 = time.perf_counter(); paths = pl.plan(np.tile(q_init, (len(goals), 1)), goals); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ok = sum(1 for p in paths if p)
    print("constrained", nqs, "plans/s", round(ok / dt, 1), "seconds", round(dt, 3), {k: pl.stats[k] for k in ("ticks", "iterations", "solved", "gave_up", "host_syncs", "configs_checked")})
PY
timeout 600 python tools/bench_extra.py edges 2>&1 | tail -1 | tee gpurun_out/r2_edges.jsonl
timeout 600 python tools/rowk_crossover.py 2>&1 | tail -10 | tee gpurun_out/r2_small_launches.txt
