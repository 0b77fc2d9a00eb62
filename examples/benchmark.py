"""BASELINE config #1 on the B200 engine: Franka Panda joint-space bi-RRT move-to-pose.

Same experiment as the reference's ``examples/benchmark.py`` (scene.xml, joints 1-7, epsilon 0.05,
goal bias 0.1, seed 42, 15 attempts, success rate + median planning time), run twice:
  1. one query at a time through ``RRT.plan_to_pose`` (the reference's call, block-extend inside);
  2. the same family of queries (seed 42 + i) as ONE batch through ``BatchedRRT.plan_to_poses``.
Needs a GPU: the package has no CPU path.

    python examples/benchmark.py [--attempts 15] [--batch 1024]
"""

import argparse
import json
import time

import numpy as np

import mjpl_b200 as mjpl
from mjpl_b200 import models

SITE = "ee_site"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--attempts", type=int, default=15)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()

    model = models.load("franka_scene")
    joints = [f"joint{i}" for i in range(1, 8)]
    constraints = [mjpl.JointLimitConstraint(model), mjpl.CollisionConstraint(model)]
    q_init = model.keyframe("home").qpos.copy()

    # ---- 1. sequential queries, the reference's loop
    times = []
    for i in range(args.attempts):
        q_goal = mjpl.random_config(model, q_init, joints, args.seed, constraints)
        goal_pose = mjpl.site_pose(model, q_goal, SITE)
        planner = mjpl.RRT(model, joints, constraints, max_planning_time=10, epsilon=0.05, seed=args.seed,
                           goal_biasing_probability=0.1)
        t0 = time.time()
        path = planner.plan_to_pose(q_init, goal_pose, SITE)
        if path:
            times.append(time.time() - t0)
    seq = {"attempts": args.attempts, "succeeded": len(times),
           "median_planning_time_s": float(np.median(times)) if times else None}
    print(f"sequential: {seq['succeeded']}/{args.attempts} plans succeeded, median {seq['median_planning_time_s']} s")

    # ---- 2. a batch of queries of the same family
    B = args.batch
    goals = np.stack([mjpl.random_config(model, q_init, joints, args.seed + i, constraints) for i in range(B)])
    poses = [mjpl.site_pose(model, q, SITE) for q in goals]
    planner = mjpl.BatchedRRT(model, joints, constraints, max_planning_time=60, epsilon=0.05, seed=args.seed,
                              goal_biasing_probability=0.1)
    inits = np.tile(q_init, (B, 1))
    planner.plan_to_poses(inits[:4], poses[:4], SITE)  # warm-up
    t0 = time.time()
    paths = planner.plan_to_poses(inits, poses, SITE)
    dt = time.time() - t0
    solved = sum(1 for p in paths if p)
    bat = {"queries": B, "solved": solved, "seconds": dt, "queries_per_s": solved / dt}
    print(f"batched: {solved}/{B} move-to-pose queries in {dt:.2f} s = {solved / dt:.0f} queries/s")
    print(json.dumps({"sequential": seq, "batched": bat}))


if __name__ == "__main__":
    main()
