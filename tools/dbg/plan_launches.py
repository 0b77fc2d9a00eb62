"""A few planner iterations (eager, no graph) for an ncu launch list: which kernels make up an iteration?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
import bench
model = models.load(bench.MODEL); eng = mj.get_engine(model, bench.ALLOWED)
c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, bench.ALLOWED)]
q_init = model.keyframe("home").qpos.copy()
rows = eng.sweep_rows(7, 0, 8 * 4096).double().cpu().numpy(); rows[:, 7:] = q_init[7:]
goals = rows[np.asarray(mj.obeys_constraints_batch(rows, c))][:4096]
pl = mj.BatchedRRT(model, bench.PLAN_JOINTS, c, max_planning_time=60.0, epsilon=0.05, seed=0, goal_biasing_probability=0.1,
                   max_active=4096, max_iterations_per_query=int(sys.argv[1]) if len(sys.argv) > 1 else 40, use_cuda_graph=False)
pl.plan(np.tile(q_init, (len(goals), 1)), goals)
print(pl.stats)
