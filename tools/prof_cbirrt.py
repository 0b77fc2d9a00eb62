"""torch-profiler breakdown of the constrained planner's ticks (configs[3])."""
import sys, time, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, torch
import mjpl_b200 as mj
from tests.test_gpu_pose import _constrained_problem
from torch.profiler import profile, ProfilerActivity
nqs = int(os.environ.get("RRT_QUERIES", "4096"))
model, allowed, joints, q_init, ref, lim, cons, goals = _constrained_problem(nqs)
pl = mj.BatchedRRT(model, joints, cons, max_planning_time=120, epsilon=0.05, seed=17, goal_biasing_probability=0.1, sync_every=32)
pl.plan(np.tile(q_init, (4, 1)), goals[:4]); torch.cuda.synchronize()
t0 = time.perf_counter(); paths = pl.plan(np.tile(q_init, (len(goals), 1)), goals); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("plain", dt, pl.stats)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    pl.plan(np.tile(q_init, (len(goals), 1)), goals); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70))
