"""GPU: PoseConstraint (pose_kernel through the C ABI) against the numpy restatement."""

import numpy as np
import pytest

import oracle
import mjpl_b200 as mj
from mjpl_b200 import models
from mjpl_b200.lie import SE3, SO3

pytestmark = pytest.mark.gpu
INF = (-np.inf, np.inf)


def test_reference_translation_case():
    # reference test/test_pose_constraint.py:17-52
    model = models.load("two_dof_ball")
    home = mj.site_pose(model, np.zeros(2), "ball_site")
    np.testing.assert_allclose(home.translation(), [0, 0, 1.0], atol=1e-15)
    pc = mj.PoseConstraint(model, "ball_site", home, x_translation=(-0.1, 0.1), q_step=np.inf)
    q = np.array([0.2, 0.0])
    assert not pc.valid_config(q)
    qc = pc.apply(np.array([0.0, 0.0]), q)
    assert qc is not None
    np.testing.assert_allclose(qc, [0.1, 0.0], rtol=0, atol=1e-12)
    assert pc.valid_config(qc)
    pc.q_step = 1e-5
    assert pc.apply(np.array([0.0, 0.0]), q) is None
    with pytest.raises(ValueError, match="tolerance"):
        mj.PoseConstraint(model, "ball_site", home, tolerance=-1.0)
    with pytest.raises(ValueError, match="q_step"):
        mj.PoseConstraint(model, "ball_site", home, q_step=0.0)
    with pytest.raises(KeyError):
        mj.PoseConstraint(model, "no_such_site", home)


@pytest.mark.parametrize("mname,site", [("ur5e_scene", "attachment_site"), ("franka_scene", "ee_site")])
def test_projection_matches_restatement(mname, site):
    model = models.load(mname)
    q0 = model.keyframe("home").qpos.copy()
    ref = mj.site_pose(model, q0, site)
    free = oracle.PoseOracle(model, site, [0, 0, 0], [1, 0, 0, 0], [INF] * 6)
    p, r = free.site_pose(q0)
    np.testing.assert_allclose(ref.translation(), p, atol=1e-12)
    assert min(np.abs(ref.rotation().wxyz - r).max(), np.abs(ref.rotation().wxyz + r).max()) < 1e-12
    lim = (-0.1, 0.1)
    pc = mj.PoseConstraint(model, site, ref, z_translation=(-0.05, 0.2), roll=lim, pitch=lim, q_step=0.5)
    po = oracle.PoseOracle(model, site, p, r, [INF, INF, (-0.05, 0.2), lim, lim, INF], q_step=0.5)
    rng = np.random.default_rng(1)
    n = 300
    Q = q0[None, :] + rng.uniform(-0.35, 0.35, size=(n, model.nq))
    Q[:, 7:] = q0[7:] if model.nq > 7 else Q[:, 7:]
    Q = np.clip(Q, model.jnt_range[:, 0], model.jnt_range[:, 1])
    want_valid = np.array([po.valid_config(q) for q in Q])
    np.testing.assert_array_equal(pc.valid_configs(Q), want_valid)
    out, ok, iters = pc.apply_batch(np.tile(q0, (n, 1)), Q, want_iterations=True)
    want = [po.apply(q0, q) for q in Q]
    want_ok = np.array([w is not None for w in want])
    assert (ok == want_ok).mean() > 0.99  # borderline aborts (|q - q_old| == 2*q_step) may flip
    both = ok & want_ok
    assert both.sum() > 15
    err = max(np.abs(out[i] - want[i]).max() for i in np.flatnonzero(both))
    print(f"{mname}: projected {both.sum()}/{n}, max |dq| vs restatement {err:.2e}, mean iterations {iters[ok].mean():.1f}")
    assert err < 1e-8
    assert pc.valid_configs(out[ok]).all()
    # scalar API = block of one
    i = int(np.flatnonzero(both)[0])
    np.testing.assert_allclose(pc.apply(q0, Q[i]), want[i], atol=1e-8)
    j = np.flatnonzero(~want_ok)
    if len(j):
        assert pc.apply(q0, Q[int(j[0])]) is None


def test_constrained_rrt_with_pose_constraint():
    """BASELINE configs[3] in miniature: bi-RRT with PoseConstraint + limits + collision
    (examples/franka_constrained_move_to_pose.py:51-95 parameters), sequential projected extends."""
    model = models.load("franka_scene_with_obstacles")
    allowed = [("left_finger", "right_finger")]
    joints = [f"joint{i}" for i in range(1, 8)]
    q_init = model.keyframe("home").qpos.copy()
    ref = mj.site_pose(model, q_init, "ee_site")
    lim = (-0.1, 0.1)
    pose = mj.PoseConstraint(model, "ee_site", ref, roll=lim, pitch=lim, q_step=0.05)
    cons = [mj.JointLimitConstraint(model), pose, mj.CollisionConstraint(model, allowed)]
    assert mj.obeys_constraints(q_init, cons)
    # a goal that satisfies all constraints: project random configurations until one sticks
    rng = np.random.default_rng(17)
    q_goal = None
    for _ in range(400):
        cand = q_init.copy()
        cand[:7] = q_init[:7] + rng.uniform(-0.8, 0.8, 7)
        pose.q_step = np.inf
        proj = mj.apply_constraints(q_init, cand, cons)
        pose.q_step = 0.05
        if proj is not None and np.linalg.norm(proj - q_init) > 0.5:
            q_goal = proj
            break
    assert q_goal is not None
    planner = mj.RRT(model, joints, cons, max_planning_time=60, epsilon=0.05, seed=17, goal_biasing_probability=0.1)
    path = planner.plan_to_config(q_init, q_goal)
    assert path, "planner timed out"
    np.testing.assert_equal(path[0], q_init)
    np.testing.assert_equal(path[-1], q_goal)
    P = np.array(path)
    assert np.asarray(mj.obeys_constraints_batch(P, cons)).all()
    po = oracle.PoseOracle(model, "ee_site", ref.translation(), ref.rotation().wxyz, [INF] * 3 + [lim, lim, INF])
    assert all(po.valid_config(q) for q in P)
    assert oracle.Oracle(model, allowed).check(P, 3).all()


def _constrained_problem(n_goals, seed=17):
    """Franka obstacle scene, PoseConstraint(roll, pitch in +-0.1) + limits + collision, home as the
    start and ``n_goals`` goals obtained by projecting random configurations onto the constraints
    (examples/franka_constrained_move_to_pose.py:51-95 parameters)."""
    model = models.load("franka_scene_with_obstacles")
    allowed = [("left_finger", "right_finger")]
    joints = [f"joint{i}" for i in range(1, 8)]
    q_init = model.keyframe("home").qpos.copy()
    ref = mj.site_pose(model, q_init, "ee_site")
    lim = (-0.1, 0.1)
    pose = mj.PoseConstraint(model, "ee_site", ref, roll=lim, pitch=lim, q_step=0.05)
    cons = [mj.JointLimitConstraint(model), pose, mj.CollisionConstraint(model, allowed)]
    rng = np.random.default_rng(seed)
    goals = np.empty((0, model.nq))
    while len(goals) < n_goals:
        cand = np.tile(q_init, (4 * n_goals, 1))
        cand[:, :7] += rng.uniform(-0.8, 0.8, size=(len(cand), 7))
        pose.q_step = np.inf
        proj, ok = mj.apply_constraints_batch(np.tile(q_init, (len(cand), 1)), cand, cons)
        pose.q_step = 0.05
        ok &= np.linalg.norm(proj - q_init, axis=1) > 0.5
        goals = np.concatenate([goals, proj[ok]])
    return model, allowed, joints, q_init, ref, lim, cons, goals[:n_goals]


def _check_constrained_paths(model, allowed, ref, lim, cons, q_init, goals, paths, solved):
    po = oracle.PoseOracle(model, "ee_site", ref.translation(), ref.rotation().wxyz, [INF] * 3 + [lim, lim, INF])
    orc = oracle.Oracle(model, allowed)
    for b in solved:
        P = np.asarray(paths[b])
        np.testing.assert_array_equal(P[0], q_init)
        np.testing.assert_array_equal(P[-1], goals[b])
        assert np.asarray(mj.obeys_constraints_batch(P, cons)).all()
        assert (np.linalg.norm(np.diff(P, axis=0), axis=1) <= 2 * 0.05 + 1e-9).all()   # a projected step moves at most 2 q_step
    for b in solved[:8]:
        P = np.asarray(paths[b])
        assert all(po.valid_config(q) for q in P)
        assert orc.check(P, 3).all()


def test_batched_constrained_rrt():
    """BASELINE configs[3] as a batch, lock-step host driver: every query takes each projected extend
    step together with the others.  Paths start and end where they should, every waypoint satisfies all
    constraints (pose checked with the numpy restatement, collisions with the CPU oracle), and a query
    gives the same path as it does alone through the sequential planner."""
    model, allowed, joints, q_init, ref, lim, cons, goals = _constrained_problem(48)
    B = len(goals)
    planner = mj.BatchedRRT(model, joints, cons, max_planning_time=120, epsilon=0.05, seed=17, goal_biasing_probability=0.1,
                            device_projection=False)
    paths = planner.plan(np.tile(q_init, (B, 1)), goals)
    solved = [b for b in range(B) if paths[b]]
    assert len(solved) >= 0.9 * B, planner.stats
    _check_constrained_paths(model, allowed, ref, lim, cons, q_init, goals, paths, solved)
    b = solved[0]
    want = mj.RRT(model, joints, cons, max_planning_time=120, epsilon=0.05, seed=17 + b, goal_biasing_probability=0.1).plan_to_config(q_init, goals[b])
    assert len(want) == len(paths[b]) and all(np.array_equal(x, y) for x, y in zip(want, paths[b]))
    print(f"batched constrained rrt (host lock-step): {len(solved)}/{B} solved in {planner.stats['seconds']:.2f} s, {planner.stats}")


def test_constrained_rrt_device_ticks():
    """BASELINE configs[3] with all planner state on the device (mjb_cbirrt_tick): asynchronous slots,
    one projected step per tick, ticks replayed as a CUDA graph.  Same acceptance test as the host
    driver: start / goal, every waypoint valid under all constraints (numpy pose restatement + CPU
    oracle), steps bounded; also without a graph and without the collision constraint."""
    model, allowed, joints, q_init, ref, lim, cons, goals = _constrained_problem(256)
    B = len(goals)
    planner = mj.BatchedRRT(model, joints, cons, max_planning_time=120, epsilon=0.05, seed=17, goal_biasing_probability=0.1)
    paths = planner.plan(np.tile(q_init, (B, 1)), goals)
    assert planner.stats["driver"] == "device ticks"
    solved = [b for b in range(B) if paths[b]]
    print(f"constrained rrt (device ticks): {len(solved)}/{B} solved in {planner.stats['seconds']:.2f} s, {planner.stats}")
    assert len(solved) >= 0.9 * B, planner.stats
    assert planner.stats["appends_refused_at_capacity"] == 0
    _check_constrained_paths(model, allowed, ref, lim, cons, q_init, goals, paths, solved)
    eager = mj.BatchedRRT(model, joints, cons, max_planning_time=120, epsilon=0.05, seed=17, goal_biasing_probability=0.1,
                          use_cuda_graph=False)
    p2 = eager.plan(np.tile(q_init, (32, 1)), goals[:32])
    for a, b in zip(p2, paths[:32]):     # same counter-based random stream, same arithmetic: same paths
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
    no_coll = [cons[0], cons[1]]
    p3 = mj.BatchedRRT(model, joints, no_coll, max_planning_time=60, epsilon=0.05, seed=3, goal_biasing_probability=0.1).plan(
        np.tile(q_init, (16, 1)), goals[:16])
    assert sum(1 for p in p3 if p) >= 14
    for p in p3:
        if p:
            assert np.asarray(mj.obeys_constraints_batch(np.asarray(p), no_coll)).all()
