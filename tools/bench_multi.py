"""BASELINE configs[4]: sharded validity sweep (rows generated on the device from the global row
id, no PCIe traffic) + NCCL gather of the per-rank masks, and independent planning queries sharded
across ranks.  Launch with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/bench_multi.py [--rows 1000000000] [--queries 4096]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

import mjpl_b200 as mj
from mjpl_b200 import models
from mjpl_b200.parallel import gather_masks, reduce_counts, shard_range

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000_000)
ap.add_argument("--queries", type=int, default=4096)
ap.add_argument("--chunk", type=int, default=25_000_000)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
model = models.load("franka_scene_with_obstacles")
allowed = [("left_finger", "right_finger")]
eng = mj.get_engine(model, allowed)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

# ---- sweep: rank r owns global rows [lo, hi); masks stay on the device, bit counts are reduced
lo, hi = shard_range(a.rows, rank, world)
mask = torch.empty(hi - lo, dtype=torch.uint8, device="cuda")
eng.sweep(0, lo, min(a.chunk, hi - lo), out=mask[: min(a.chunk, hi - lo)])  # warm-up
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for c0 in range(lo, hi, a.chunk):
    n = min(a.chunk, hi - c0)
    eng.sweep(0, c0, n, out=mask[c0 - lo : c0 - lo + n])
e1.record()
barrier()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
valid_total = int(reduce_counts([int(mask.sum())], device="cuda")[0]) if world > 1 else int(mask.sum())
# gather of one logical batch that was split: the first 8M global rows' masks, every rank gets all
g0 = time.perf_counter()
nb = min(8_000_000, a.rows)
blo, bhi = shard_range(nb, rank, world)
part = eng.sweep(0, blo, bhi - blo)
full = gather_masks(part, nb) if world > 1 else part
torch.cuda.synchronize()
gather_s = time.perf_counter() - g0
check = eng.sweep(0, 0, min(nb, 1_000_000))
assert bool((full[: len(check)] == check).all())

# ---- planning queries: rank r plans its own contiguous slice
joints = [f"joint{i}" for i in range(1, 8)]
cons = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, allowed)]
q_init = model.keyframe("home").qpos.copy()
rows = eng.sweep_rows(11, 0, 8 * a.queries).double().cpu().numpy()
rows[:, 7:] = q_init[7:]
goals = rows[np.asarray(mj.obeys_constraints_batch(rows, cons))][: a.queries]
qlo, qhi = shard_range(len(goals), rank, world)
pl = mj.BatchedRRT(model, joints, cons, max_planning_time=120, epsilon=0.05, seed=rank, goal_biasing_probability=0.1,
                   max_iterations_per_query=2000)
pl.plan(np.tile(q_init, (4, 1)), goals[:4])
barrier()
t0 = time.perf_counter()
paths = pl.plan(np.tile(q_init, (qhi - qlo, 1)), goals[qlo:qhi])
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
solved = torch.tensor([sum(1 for p in paths if p)], dtype=torch.int64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(solved, op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps({"case": "sharded sweep + sharded planning (BASELINE configs[4])", "n_gpus": world, "rows": a.rows,
                      "sweep_ms": float(ms[0]), "configs_per_s": a.rows / float(ms[0]) * 1e3, "valid_fraction": valid_total / a.rows,
                      "mask_gather": {"rows": nb, "seconds_incl_compute": gather_s, "backend": "nccl all_gather" if world > 1 else "none"},
                      "queries": len(goals), "solved": int(solved[0]), "plan_seconds": float(dt[0]),
                      "plans_per_s": int(solved[0]) / float(dt[0])}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
