#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planning.py -x -q -s 2>&1 | tail -15
timeout 600 python - <<'PY'
import sys, time, json; sys.path.insert(0,'.')
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
import bench
model = models.load(bench.MODEL); eng = mj.get_engine(model, bench.ALLOWED)
for kw in ({}, {"use_cuda_graph": False}, {"sync_every": 16}, {"sync_every": 32}):
    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, bench.ALLOWED)]
    q_init = model.keyframe("home").qpos.copy()
    rows = eng.sweep_rows(7, 0, 8 * 4096).double().cpu().numpy(); rows[:, 7:] = q_init[7:]
    goals = rows[np.asarray(mj.obeys_constraints_batch(rows, c))][:4096]
    pl = mj.BatchedRRT(model, bench.PLAN_JOINTS, c, max_planning_time=60.0, epsilon=0.05, seed=0, goal_biasing_probability=0.1, max_active=4096, max_iterations_per_query=2000, **kw)
    pl.plan(np.tile(q_init, (8, 1)), goals[:8]); torch.cuda.synchronize()
    t0 = time.perf_counter(); paths = pl.plan(np.tile(q_init, (len(goals), 1)), goals); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(kw, "plans/s", sum(1 for p in paths if p) / dt, "seconds", dt, pl.stats)
PY
