#!/bin/bash
timeout 600 python - <<'PY'
import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
import bench
from tests.test_gpu_pose import _constrained_problem
model = models.load(bench.MODEL); eng = mj.get_engine(model, bench.ALLOWED)
c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, bench.ALLOWED)]
q_init = model.keyframe("home").qpos.copy()
rows = eng.sweep_rows(7, 0, 8 * 4096).double().cpu().numpy(); rows[:, 7:] = q_init[7:]
goals = rows[np.asarray(mj.obeys_constraints_batch(rows, c))][:4096]
for nqs in (512, 4096):
    pl = mj.BatchedRRT(model, bench.PLAN_JOINTS, c, max_planning_time=60.0, epsilon=0.05, seed=0, goal_biasing_probability=0.1, max_active=4096, max_iterations_per_query=2000, sync_every=32)
    pl.plan(np.tile(q_init, (8, 1)), goals[:8]); torch.cuda.synchronize()
    t0 = time.perf_counter(); paths = pl.plan(np.tile(q_init, (nqs, 1)), goals[:nqs]); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("unconstrained", nqs, "plans/s", sum(1 for p in paths if p) / dt, "seconds", dt, pl.stats["solved"], pl.stats["iterations"])
for nqs in (1024, 4096):
    model2, allowed, joints, qi, ref, lim, cons, g2 = _constrained_problem(nqs)
    pl = mj.BatchedRRT(model2, joints, cons, max_planning_time=120, epsilon=0.05, seed=17, goal_biasing_probability=0.1, sync_every=32)
    pl.plan(np.tile(qi, (4, 1)), g2[:4]); torch.cuda.synchronize()
    t0 = time.perf_counter(); paths = pl.plan(np.tile(qi, (len(g2), 1)), g2); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("constrained", nqs, "plans/s", sum(1 for p in paths if p) / dt, "seconds", dt, {k: pl.stats[k] for k in ("ticks", "iterations", "solved", "gave_up")})
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
