"""Secondary measurements (BASELINE configs 3 and 5): UR5e edge validation and batched bi-RRT
planning queries per second, each next to its CPU restatement.  Prints one JSON line per case.

    python tools/bench_extra.py [edges] [plans] [--queries N]
"""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import mjpl_b200 as mj
import oracle
from mjpl_b200 import models


def bench_edges(ne=100_000, step=0.05):
    model = models.load("ur5e_scene")
    eng = mj.get_engine(model, [])
    rng = np.random.default_rng(0)
    q0 = rng.uniform(-3.1415, 3.1415, size=(ne, 6)).astype(np.float32)
    q1 = rng.uniform(-3.1415, 3.1415, size=(ne, 6)).astype(np.float32)
    d0, d1 = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
    for _ in range(3):
        v, fb = eng.valid_edges(d0, d1, step, want_first_bad=True)
    torch.cuda.synchronize()
    eng.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        v, fb = eng.valid_edges(d0, d1, step, want_first_bad=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = eng.stats()
    # CPU restatement on a subsample
    orc = oracle.Oracle(model)
    oracle.Oracle.set_threads(len(os.sched_getaffinity(0)))
    ns = 300
    t0 = time.perf_counter()
    want = [oracle.valid_collision_interval(orc, q0[i].astype(float), q1[i].astype(float), step) for i in range(ns)]
    cpu_s = time.perf_counter() - t0
    agree = np.mean([bool(v[i]) == w[0] and int(fb[i]) == w[1] for i, w in enumerate(want)])
    print(json.dumps({"case": "UR5e edge validation (BASELINE configs[2])", "edges": ne, "step": step,
                      "edges_per_s": ne / ms * 1e3, "waypoints_total": st["rows"] // reps,
                      "configs_per_s": st["rows"] / reps / ms * 1e3, "ms": ms, "valid_fraction": float(v.float().mean()),
                      "cpu_port_edges_per_s_one_core_python_loop": ns / cpu_s, "agreement_on_subsample": float(agree)}))


def bench_plans(nq_queries=1024, time_limit=60.0):
    model = models.load("franka_scene_with_obstacles")
    allowed = [("left_finger", "right_finger")]
    joints = [f"joint{i}" for i in range(1, 8)]
    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, allowed)]
    eng = c[1].engine
    q_init = model.keyframe("home").qpos.copy()
    # goals: device-generated uniform rows with the fingers pinned to the home value, valid ones kept
    rows = eng.sweep_rows(7, 0, 8 * nq_queries).double().cpu().numpy()
    rows[:, 7:] = q_init[7:]
    ok = np.asarray(mj.obeys_constraints_batch(rows, c))
    goals = rows[ok][:nq_queries]
    B = len(goals)
    planner = mj.BatchedRRT(model, joints, c, max_planning_time=time_limit, epsilon=0.05, seed=0, goal_biasing_probability=0.1,
                            max_active=int(os.environ.get("MAX_ACTIVE", "4096")), max_iterations_per_query=int(os.environ.get("MAX_ITERS", "2000")))
    planner.plan(np.tile(q_init, (8, 1)), goals[:8])  # warm-up (allocator, kernels)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    paths = planner.plan(np.tile(q_init, (B, 1)), goals)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    solved = [i for i, p in enumerate(paths) if p]
    # replay a sample under the fp64 oracle
    orc = oracle.Oracle(model, allowed)
    oracle.Oracle.set_threads(len(os.sched_getaffinity(0)))
    bad = 0
    for i in solved[:200]:
        P = np.array(paths[i])
        okp = orc.check(P, 3).all() and (np.linalg.norm(np.diff(P, axis=0), axis=1) <= 0.05 + 1e-9).all()
        okp = okp and np.array_equal(P[0], q_init) and np.array_equal(P[-1], goals[i])
        bad += 0 if okp else 1
    # CPU restatement: the reference's sequential algorithm on the oracle, a few queries, one core
    from tests.doubles import OracleCollisionConstraint, OracleJointLimitConstraint
    from mjpl_b200.planning.rrt import RRT

    oracle.Oracle.set_threads(1)
    cc = [OracleJointLimitConstraint(model), OracleCollisionConstraint(model, allowed)]
    ncpu = 6
    t1 = time.perf_counter()
    ok_cpu = 0
    for i in range(ncpu):
        r = RRT(model, joints, cc, max_planning_time=20, epsilon=0.05, seed=i, goal_biasing_probability=0.1)
        ok_cpu += bool(r.plan_to_config(q_init, goals[i]))
    cpu_dt = time.perf_counter() - t1
    cores = len(os.sched_getaffinity(0))
    print(json.dumps({"case": "batched bi-RRT, Franka scene_with_obstacles, home -> random valid goal", "queries": B,
                      "solved": len(solved), "seconds": dt, "plans_per_s": len(solved) / dt, "stats": planner.stats,
                      "replayed_under_oracle": min(200, len(solved)), "replay_failures": bad,
                      "cpu_port": {"queries": ncpu, "solved": ok_cpu, "seconds": cpu_dt, "plans_per_s_one_core": ok_cpu / cpu_dt,
                                   "plans_per_s_all_cores_extrapolated": ok_cpu / cpu_dt * cores, "cores": cores,
                                   "what": "block-extend RRT host logic on the fp64 C oracle (not MuJoCo), sequential queries"}}))


def bench_poses(nq_queries=1024, time_limit=60.0):
    """BASELINE config #1 as the reference runs it (examples/benchmark.py): the goal is a POSE (the
    end-effector pose of a random valid configuration); each query is IK + bi-RRT."""
    model = models.load("franka_scene_with_obstacles")
    allowed = [("left_finger", "right_finger")]
    joints = [f"joint{i}" for i in range(1, 8)]
    site = "ee_site"
    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, allowed)]
    eng = c[1].engine
    q_init = model.keyframe("home").qpos.copy()
    rows = eng.sweep_rows(11, 0, 8 * nq_queries).double().cpu().numpy()
    rows[:, 7:] = q_init[7:]
    ok = np.asarray(mj.obeys_constraints_batch(rows, c))
    qg = rows[ok][:nq_queries]
    B = len(qg)
    # goal poses through the engine's fp64 site kinematics, one block
    solver = mj.DLSIKSolver(model=model, joints=joints, constraints=c, seed=0, max_attempts=8)
    po = oracle.PoseOracle(model, site, [0, 0, 0], [1, 0, 0, 0], [(-np.inf, np.inf)] * 6)
    poses = []
    for q in qg:
        p, r = po.site_pose(q)
        poses.append(mj.SE3(mj.SO3(r), p))
    inits = np.tile(q_init, (B, 1))
    solver.solve_ik_batch(poses[:8], site, inits[:8])  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Q, solved = solver.solve_ik_batch(poses, site, inits)
    torch.cuda.synchronize()
    ik_dt = time.perf_counter() - t0
    # kernel-only rate of the IK launch (rows = queries x attempts)
    tp = np.repeat(np.array([p.translation() for p in poses]), 8, axis=0)
    tq = np.repeat(np.array([p.rotation().wxyz for p in poses]), 8, axis=0)
    G = solver._guess_block(inits).reshape(-1, model.nq)
    t1 = time.perf_counter()
    solver.solve_rows(tp, tq, G, site)
    rows_dt = time.perf_counter() - t1
    bad = 0
    for i in np.flatnonzero(solved)[:200]:
        p, r = po.site_pose(Q[i])
        e = poses[i].minus(mj.SE3(mj.SO3(r), p))
        bad += not (np.linalg.norm(e[:3]) <= 1e-3 and np.linalg.norm(e[3:]) <= 1e-3)
    planner = mj.BatchedRRT(model, joints, c, max_planning_time=time_limit, epsilon=0.05, seed=0, goal_biasing_probability=0.1,
                            max_active=int(os.environ.get("MAX_ACTIVE", "4096")), max_iterations_per_query=int(os.environ.get("MAX_ITERS", "2000")))
    planner.plan_to_poses(inits[:8], poses[:8], site, solver)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    paths = planner.plan_to_poses(inits, poses, site, solver)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t2
    nsolved = sum(1 for p in paths if p)
    print(json.dumps({"case": "move-to-pose queries (IK + batched bi-RRT), Franka scene_with_obstacles", "queries": B,
                      "ik": {"solved": int(solved.sum()), "attempts_per_query": 8, "seconds": ik_dt, "poses_per_s": B / ik_dt,
                             "kernel_call_rows": len(G), "kernel_call_seconds_incl_copies": rows_dt,
                             "pose_check_failures_of_200": int(bad)},
                      "solved": nsolved, "seconds": dt, "queries_per_s": nsolved / dt, "stats": planner.stats}))


def bench_constrained(nq_queries=256, time_limit=120.0):
    """BASELINE config #4 as a batch: PoseConstraint (roll, pitch within +-0.1) + limits + collision,
    Franka obstacle scene, lock-step projected extends; the sequential planner timed on a few of
    the same queries beside it."""
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
    from tests.test_gpu_pose import _constrained_problem

    model, allowed, joints, q_init, ref, lim, cons, goals = _constrained_problem(nq_queries)
    B = len(goals)
    planner = mj.BatchedRRT(model, joints, cons, max_planning_time=time_limit, epsilon=0.05, seed=17, goal_biasing_probability=0.1)
    planner.plan(np.tile(q_init, (4, 1)), goals[:4])
    t0 = time.perf_counter()
    paths = planner.plan(np.tile(q_init, (B, 1)), goals)
    dt = time.perf_counter() - t0
    solved = [b for b in range(B) if paths[b]]
    ok = all(np.asarray(mj.obeys_constraints_batch(np.asarray(paths[b]), cons)).all() for b in solved[:64])
    nseq = 4
    t1 = time.perf_counter()
    seq_ok = 0
    for b in range(nseq):
        r = mj.RRT(model, joints, cons, max_planning_time=60, epsilon=0.05, seed=17 + b, goal_biasing_probability=0.1)
        seq_ok += bool(r.plan_to_config(q_init, goals[b]))
    seq_dt = time.perf_counter() - t1
    print(json.dumps({"case": "constrained bi-RRT (PoseConstraint + limits + collision), Franka scene_with_obstacles", "queries": B,
                      "solved": len(solved), "seconds": dt, "plans_per_s": len(solved) / dt, "paths_valid": bool(ok),
                      "stats": planner.stats,
                      "sequential_same_engine": {"queries": nseq, "solved": seq_ok, "seconds": seq_dt, "plans_per_s": seq_ok / seq_dt}}))


if __name__ == "__main__":
    args = sys.argv[1:]
    nqq = int(args[args.index("--queries") + 1]) if "--queries" in args else 1024
    which = [a for a in args if a in ("edges", "plans", "poses", "constrained")] or ["edges", "plans"]
    if "edges" in which:
        bench_edges()
    if "plans" in which:
        bench_plans(nqq)
    if "poses" in which:
        bench_poses(nqq)
    if "constrained" in which:
        bench_constrained(nqq)
