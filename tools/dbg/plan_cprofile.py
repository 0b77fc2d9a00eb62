import sys, time, cProfile, pstats
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
model = models.load("franka_scene_with_obstacles"); allowed = [("left_finger", "right_finger")]
joints = [f"joint{i}" for i in range(1, 8)]
c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, allowed)]
eng = c[1].engine
q_init = model.keyframe("home").qpos.copy()
rows = eng.sweep_rows(7, 0, 8 * 4096).double().cpu().numpy(); rows[:, 7:] = q_init[7:]
goals = rows[np.asarray(mj.obeys_constraints_batch(rows, c))][:4096]
pl = mj.BatchedRRT(model, joints, c, max_planning_time=60, epsilon=0.05, seed=0, goal_biasing_probability=0.1,
                   max_active=4096, max_iterations_per_query=2000, sync_every=32)
QI = np.tile(q_init, (4096, 1))
pl.plan(QI, goals); pl.plan(QI, goals)
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter(); pl.plan(QI, goals); torch.cuda.synchronize(); dt = time.perf_counter() - t0; pr.disable()
print("plan", dt)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
