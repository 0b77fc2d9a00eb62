"""GPU: the IK kernel through the C ABI (mjb_ik_solve) and the solver / planner entry points on top.

The acceptance criteria are the reference's (test/test_mink_ik_solver.py, test/test_rrt.py
plan_to_pose cases): solutions reproduce the target pose within tolerance (checked with the CPU
oracle's kinematics, not the engine's), obey the constraints, and keep fixed joints fixed."""

import numpy as np
import pytest

import oracle
import mjpl_b200 as mj
from mjpl_b200 import models
from mjpl_b200.lie import SE3, SO3
from tests.hostsim import HostSim

pytestmark = pytest.mark.gpu
INF = (-np.inf, np.inf)


def _free(model, site):
    return oracle.PoseOracle(model, site, [0, 0, 0], [1, 0, 0, 0], [INF] * 6)


def _err(po, target, q):
    p, r = po.site_pose(q)
    return target.minus(SE3(SO3(r), p))


@pytest.mark.parametrize("mname,site", [("ur5e_scene", "attachment_site"), ("franka_scene", "ee_site")])
def test_kernel_matches_host_execution(mname, site):
    """Same fp64 iteration on the device and on the CPU (tests/hostsim): same verdicts, same rows."""
    model = models.load(mname)
    po = _free(model, site)
    rng = np.random.default_rng(11)
    lo, hi = model.jnt_range.T
    n = 256
    poses = [po.site_pose(q) for q in rng.uniform(lo, hi, size=(n, model.nq))]
    tp = np.array([p for p, _ in poses])
    tq = np.array([r for _, r in poses])
    q0 = rng.uniform(lo, hi, size=(n, model.nq))
    # one iteration at a time from identical inputs: the iterates agree to rounding (rows resting
    # on a joint limit exercise the active-set re-solve from the second step on)
    hs = HostSim(model)
    one = mj.DLSIKSolver(model, mj.all_joints(model), iterations=1)
    q, at_limit = q0, 0
    for _ in range(5):
        Qd = one.solve_rows(tp, tq, q, site)[0]
        Qh = hs.ik(one._spec(site), tp, tq, q)[0]
        assert np.abs(Qd - Qh).max() < 1e-9
        q = Qh
        at_limit += int(((q <= lo) | (q >= hi)).any(axis=1).sum())
    assert at_limit > 50
    # full runs: fp64 on both sides, but fma contraction / libm differ and the iteration is
    # chaotic near singular configurations, so a few rows may end in different basins
    solver = mj.DLSIKSolver(model, mj.all_joints(model), iterations=200)
    Q, ok, iters, errs = solver.solve_rows(tp, tq, q0, site)
    Qh, okh, itersh, errsh = hs.ik(solver._spec(site), tp, tq, q0)
    same = ok == okh
    assert same.mean() > 0.9, same.mean()
    quick = okh & ok & (itersh <= 15)
    close = np.abs(Q[quick] - Qh[quick]).max(axis=1) < 1e-6
    assert quick.sum() > n // 16 and close.mean() > 0.9, (quick.sum(), close.mean())
    for i in np.flatnonzero(ok):
        e = _err(po, SE3(SO3(tq[i]), tp[i]), Q[i])
        assert np.linalg.norm(e[:3]) <= 1e-3 and np.linalg.norm(e[3:]) <= 1e-3
    assert (Q >= lo - 1e-12).all() and (Q <= hi + 1e-12).all()
    print(f"{mname}: device ok {ok.sum()}/{n}, host ok {okh.sum()}/{n}, verdicts equal {same.mean():.3f}, "
          f"quick rows equal {close.mean():.3f}")


def test_ik_reference_case():
    # reference test/test_mink_ik_solver.py:12-66
    model = models.load("ur5e_scene")
    site = "attachment_site"
    cons = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model)]
    q_init = model.keyframe("home").qpos.copy()
    rng = np.random.default_rng(seed=12345)
    target = mj.site_pose(model, rng.uniform(*model.jnt_range.T), site)
    solver = mj.DLSIKSolver(model=model, joints=mj.all_joints(model), constraints=cons, pos_tolerance=1e-3,
                            ori_tolerance=1e-3, seed=12345, max_attempts=5)
    sols = solver.solve_ik(target, site, q_init) + solver.solve_ik(target, site, None)
    assert len(sols) == 2
    po = _free(model, site)
    for q in sols:
        assert mj.obeys_constraints(q, cons)
        e = _err(po, target, q)
        assert np.linalg.norm(e[:3]) <= 1e-3 and np.linalg.norm(e[3:]) <= 1e-3


def test_ik_subset_joints():
    # reference test/test_mink_ik_solver.py:68-113
    model = models.load("ur5e_scene")
    site = "attachment_site"
    cons = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model)]
    q_init = model.keyframe("home").qpos.copy()
    joints = ["shoulder_pan_joint", "elbow_joint"]
    q_rand = mj.random_config(model, q_init, joints, seed=12345, constraints=cons)
    target = mj.site_pose(model, q_rand, site)
    solver = mj.DLSIKSolver(model=model, joints=joints, constraints=cons, seed=12345, max_attempts=5)
    sols = solver.solve_ik(target, site, q_init)
    assert len(sols) == 1 and mj.obeys_constraints(sols[0], cons)
    fixed = [i for i in range(model.nq) if i not in mj.qpos_idx(model, joints)]
    np.testing.assert_allclose(sols[0][fixed], q_init[fixed], rtol=0, atol=1e-12)


def test_ik_batch_throughput_shape():
    """Thousands of (target, attempt) rows in one launch; every solved row is checked."""
    model = models.load("franka_scene")
    site = "ee_site"
    cons = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model)]
    arm = mj.all_joints(model)[:7]
    q_init = model.keyframe("home").qpos.copy()
    n = 512
    targets = [mj.site_pose(model, mj.random_config(model, q_init, arm, seed=1000 + i, constraints=cons), site)
               for i in range(n)]
    solver = mj.DLSIKSolver(model=model, joints=arm, constraints=cons, seed=3, max_attempts=8)
    Q, solved = solver.solve_ik_batch(targets, site, q_init)
    assert solved.mean() > 0.95, solved.mean()
    assert mj.obeys_constraints_batch(Q[solved], cons).all()
    po = _free(model, site)
    for i in np.flatnonzero(solved)[:128]:
        e = _err(po, targets[i], Q[i])
        assert np.linalg.norm(e[:3]) <= 1e-3 and np.linalg.norm(e[3:]) <= 1e-3
    np.testing.assert_array_equal(Q[solved][:, 7:], np.tile(q_init[7:], (solved.sum(), 1)))
    print(f"franka: {solved.sum()}/{n} pose goals solved with 8 attempts each")


def test_plan_to_pose():
    """examples/benchmark.py's query: goal pose from a random valid configuration, plan_to_pose."""
    model = models.load("franka_scene")
    site = "ee_site"
    arm = mj.all_joints(model)[:7]
    cons = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model)]
    q_init = model.keyframe("home").qpos.copy()
    q_goal = mj.random_config(model, q_init, arm, seed=24, constraints=cons)
    goal_pose = mj.site_pose(model, q_goal, site)
    planner = mj.RRT(model, arm, cons, seed=24, max_planning_time=20.0)
    path = planner.plan_to_pose(q_init, goal_pose, site)
    assert len(path) >= 2
    np.testing.assert_array_equal(path[0], q_init)
    e = _err(_free(model, site), goal_pose, path[-1])
    assert np.linalg.norm(e[:3]) <= 1e-3 and np.linalg.norm(e[3:]) <= 1e-3
    assert mj.obeys_constraints_batch(np.asarray(path), cons).all()
    # an unreachable pose gives no path (reference rrt.py:139)
    far = SE3(SO3([1, 0, 0, 0]), [4.0, 4.0, 4.0])
    assert planner.plan_to_pose(q_init, far, site) == []


def test_ik_argument_errors():
    model = models.load("ur5e_scene")
    s = mj.DLSIKSolver(model, mj.all_joints(model))
    with pytest.raises(ValueError):
        s.solve_rows(np.zeros((2, 3)), np.zeros((1, 4)), np.zeros((2, 6)), "attachment_site")
    with pytest.raises(KeyError):
        s.solve_rows(np.zeros((1, 3)), np.array([[1.0, 0, 0, 0]]), np.zeros((1, 6)), "nope")


def test_cartesian_path():
    # reference test/test_cartesian_planner.py:114-200, on the engine
    model = models.load("ur5e_scene")
    site = "attachment_site"
    cons = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model)]
    q_init = model.keyframe("home").qpos.copy()
    current = mj.site_pose(model, q_init, site)
    nxt = current.multiply(SE3.from_translation(np.array([0.02, 0.0, 0.0])))
    final = nxt.multiply(SE3.from_translation(np.array([0.0, 0.02, 0.0])))
    mid = current.multiply(SE3.from_translation(np.array([0.02, 0.01, 0.0])))
    solver = mj.DLSIKSolver(model=model, joints=mj.all_joints(model), constraints=[cons[1]], pos_tolerance=1e-3,
                            ori_tolerance=1e-3, seed=12345, max_attempts=5)
    wps = mj.cartesian_plan(q_init, [nxt, final], site, solver, cons, lin_threshold=0.01, ori_threshold=0.1,
                            collision_interval_check=(0.01, cons[1]))
    assert len(wps) == 4
    assert mj.obeys_constraints_batch(np.asarray(wps), cons).all()
    np.testing.assert_equal(wps[0], q_init)
    po = _free(model, site)
    for w, want in zip(wps[1:], (nxt, mid, final)):
        e = _err(po, want, w)
        assert np.linalg.norm(e[:3]) <= 1e-3 and np.linalg.norm(e[3:]) <= 1e-3
    with pytest.raises(ValueError, match="site"):
        mj.cartesian_plan(q_init, [], "", solver, [])
