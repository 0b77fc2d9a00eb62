import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
from torch.profiler import profile, ProfilerActivity
model = models.load("franka_scene_with_obstacles"); allowed = [("left_finger", "right_finger")]
joints = [f"joint{i}" for i in range(1, 8)]
c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, allowed)]
eng = c[1].engine
q_init = model.keyframe("home").qpos.copy()
import os
NQ = int(os.environ.get('RRT_QUERIES', '256'))
rows = eng.sweep_rows(7, 0, 8 * NQ).double().cpu().numpy(); rows[:, 7:] = q_init[7:]
goals = rows[np.asarray(mj.obeys_constraints_batch(rows, c))][:NQ]
pl = mj.BatchedRRT(model, joints, c, max_planning_time=30, epsilon=0.05, seed=0, goal_biasing_probability=0.1, max_active=4096, max_iterations_per_query=int(os.environ.get('RRT_ITERS', '300')), sync_every=32)
pl.plan(np.tile(q_init, (8, 1)), goals[:8])
torch.cuda.synchronize()
t0 = time.perf_counter(); pl.plan(np.tile(q_init, (len(goals), 1)), goals); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("plain:", pl.stats, "ms/iter", dt / pl.stats["iterations"] * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    pl.plan(np.tile(q_init, (len(goals), 1)), goals); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
