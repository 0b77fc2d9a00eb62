"""The IK iteration (vk_core.cuh ik_row) executed on the CPU through tests/hostsim, checked
against the oracle's kinematics: every row reported as converged must reproduce the target pose
within tolerance (reference acceptance test: test/test_mink_ik_solver.py:59-66), stay inside the
joint limits and leave joints outside the movable set untouched (:68-104)."""

import numpy as np
import pytest

import oracle
from mjpl_b200 import _abi, models
from mjpl_b200.lie import SE3, SO3
from tests.hostsim import HostSim

SITES = {"ur5e_scene": "attachment_site", "franka_scene": "ee_site"}


def _spec(m, site, mask, pos_tol=1e-3, ori_tol=1e-3, iterations=500):
    s = m.site(site).id
    sp = _abi.IkSpec()
    sp.site_bodyid = int(m.site_bodyid[s])
    sp.site_pos[:] = [float(x) for x in m.site_pos[s]]
    sp.site_quat[:] = [float(x) for x in m.site_quat[s]]
    sp.movable_mask = mask
    sp.pos_tolerance, sp.ori_tolerance = pos_tol, ori_tol
    sp.lm_damping, sp.damping, sp.max_step = -1.0, 0.0, 0.0   # defaults
    sp.iterations = iterations
    return sp


def _site_name(m, name):
    want = SITES[name]
    try:
        m.site(want)
        return want
    except Exception:
        return m.site(0).name


@pytest.mark.parametrize("name", ["ur5e_scene", "franka_scene"])
def test_converged_rows_reach_the_target(name):
    m = models.load(name)
    site = _site_name(m, name)
    po = oracle.PoseOracle(m, site, [0, 0, 0], [1, 0, 0, 0], [(-np.inf, np.inf)] * 6)
    rng = np.random.default_rng(12345)
    lo, hi = m.jnt_range.T
    n = 64
    Qt = rng.uniform(lo, hi, size=(n, m.nq))
    poses = [po.site_pose(q) for q in Qt]
    tp = np.array([p for p, _ in poses])
    tq = np.array([r for _, r in poses])
    q0 = np.tile(m.keyframe("home").qpos, (n, 1))
    hs = HostSim(m)
    Q, ok, iters, errs = hs.ik(_spec(m, site, (1 << m.njnt) - 1), tp, tq, q0)
    # One guess alone solves a good share of random reachable targets; the rest end in a
    # constrained local minimum (a joint resting on a limit), which is what the caller's random
    # restarts are for (reference mink_ik_solver.py:110-116).
    assert ok.mean() > {"ur5e_scene": 0.25, "franka_scene": 0.7}[name], ok.mean()
    assert (Q >= lo - 1e-12).all() and (Q <= hi + 1e-12).all()
    for i in np.flatnonzero(ok):
        p, r = po.site_pose(Q[i])
        err = SE3(SO3(tq[i]), tp[i]).minus(SE3(SO3(r), p))
        assert np.linalg.norm(err[:3]) <= 1e-3 * (1 + 1e-6)
        assert np.linalg.norm(err[3:]) <= 1e-3 * (1 + 1e-6)
        assert errs[i, 0] <= 1e-3 and errs[i, 1] <= 1e-3
        assert iters[i] < 500
    print(f"{name}: {ok.sum()}/{n} converged from one guess, median iterations {np.median(iters[ok]):.0f}")


def test_fixed_joints_do_not_move():
    # reference test/test_mink_ik_solver.py:68-104: joints outside `joints` keep their value
    m = models.load("ur5e_scene")
    site = _site_name(m, "ur5e_scene")
    po = oracle.PoseOracle(m, site, [0, 0, 0], [1, 0, 0, 0], [(-np.inf, np.inf)] * 6)
    q0 = m.keyframe("home").qpos.copy()
    rng = np.random.default_rng(3)
    n = 32
    Qt = np.tile(q0, (n, 1))
    Qt[:, 1:] += rng.uniform(-0.4, 0.4, size=(n, m.nq - 1))      # joint 0 identical in target and guess
    poses = [po.site_pose(q) for q in Qt]
    tp = np.array([p for p, _ in poses])
    tq = np.array([r for _, r in poses])
    mask = ((1 << m.njnt) - 1) & ~1
    Q, ok, _, _ = HostSim(m).ik(_spec(m, site, mask), tp, tq, np.tile(q0, (n, 1)))
    assert ok.sum() >= n // 2
    assert np.all(Q[:, 0] == q0[0])


def test_argument_errors():
    m = models.load("ur5e_scene")
    site = _site_name(m, "ur5e_scene")
    hs = HostSim(m)
    z3, z4, q = np.zeros((1, 3)), np.array([[1.0, 0, 0, 0]]), np.zeros((1, m.nq))
    with pytest.raises(ValueError, match="joints"):
        hs.ik(_spec(m, site, 0), z3, z4, q)
    with pytest.raises(ValueError, match="iterations"):
        hs.ik(_spec(m, site, 1, iterations=0), z3, z4, q)


def test_unreachable_target_reports_failure():
    m = models.load("ur5e_scene")
    site = _site_name(m, "ur5e_scene")
    Q, ok, iters, errs = HostSim(m).ik(_spec(m, site, (1 << m.njnt) - 1, iterations=100), [[5.0, 5.0, 5.0]],
                                       [[1.0, 0, 0, 0]], m.keyframe("home").qpos[None, :])
    assert not ok[0] and iters[0] == 100 and errs[0, 0] > 1.0


# ---- the solver's host logic (attempt schedule, constraint filter) with the kernel core on the CPU ------
from mjpl_b200 import DLSIKSolver, all_joints, qpos_idx, random_config
from tests.doubles import OracleCollisionConstraint, OracleJointLimitConstraint


class HostDLSIKSolver(DLSIKSolver):
    """DLSIKSolver whose row solver is the same C++ core run by tests/hostsim (no GPU)."""

    def solve_rows(self, target_pos, target_quat, q_init, site):
        return HostSim(self.model).ik(self._spec(site), target_pos, target_quat, q_init)


def _ur5e():
    m = models.load("ur5e_scene")
    cons = [OracleJointLimitConstraint(m), OracleCollisionConstraint(m)]
    po = oracle.PoseOracle(m, "attachment_site", [0, 0, 0], [1, 0, 0, 0], [(-np.inf, np.inf)] * 6)
    return m, cons, po


def _pose_err(po, target, q):
    p, r = po.site_pose(q)
    return target.minus(SE3(SO3(r), p))


def test_solver_ik_like_the_reference_test():
    # reference test/test_mink_ik_solver.py:12-66
    m, cons, po = _ur5e()
    q_init = m.keyframe("home").qpos.copy()
    rng = np.random.default_rng(seed=12345)
    p, r = po.site_pose(rng.uniform(*m.jnt_range.T))
    target = SE3(SO3(r), p)
    solver = HostDLSIKSolver(model=m, joints=all_joints(m), constraints=cons, pos_tolerance=1e-3, ori_tolerance=1e-3,
                             seed=12345, max_attempts=5)
    sols = solver.solve_ik(target, "attachment_site", q_init) + solver.solve_ik(target, "attachment_site", None)
    assert len(sols) == 2
    for q in sols:
        assert all(c.valid_config(q) for c in cons)
        err = _pose_err(po, target, q)
        assert np.linalg.norm(err[:3]) <= 1e-3 and np.linalg.norm(err[3:]) <= 1e-3


def test_solver_subset_joints():
    # reference test/test_mink_ik_solver.py:68-113
    m, cons, po = _ur5e()
    q_init = m.keyframe("home").qpos.copy()
    joints = ["shoulder_pan_joint", "elbow_joint"]
    q_rand = random_config(m, q_init, joints, seed=12345, constraints=cons)
    p, r = po.site_pose(q_rand)
    solver = HostDLSIKSolver(model=m, joints=joints, constraints=cons, seed=12345, max_attempts=5)
    sols = solver.solve_ik(SE3(SO3(r), p), "attachment_site", q_init)
    assert len(sols) == 1 and all(c.valid_config(sols[0]) for c in cons)
    fixed = [i for i in range(m.nq) if i not in qpos_idx(m, joints)]
    np.testing.assert_allclose(sols[0][fixed], q_init[fixed], rtol=0, atol=1e-12)


def test_solver_invalid_args():
    # reference test/test_mink_ik_solver.py:115-147
    m = models.load("ur5e_scene")
    for bad in (-2, 0):
        with pytest.raises(ValueError, match="`max_attempts` must be > 0"):
            DLSIKSolver(model=m, joints=all_joints(m), max_attempts=bad)
        with pytest.raises(ValueError, match="`iterations` must be > 0"):
            DLSIKSolver(model=m, joints=all_joints(m), iterations=bad)
    with pytest.raises(ValueError, match="cannot be empty"):
        DLSIKSolver(model=m, joints=[])


def test_solver_batch_returns_first_passing_attempt():
    m, cons, po = _ur5e()
    rng = np.random.default_rng(7)
    lo, hi = m.jnt_range.T
    targets = []
    while len(targets) < 12:
        q = rng.uniform(lo, hi)
        if all(c.valid_config(q) for c in cons):
            p, r = po.site_pose(q)
            targets.append(SE3(SO3(r), p))
    solver = HostDLSIKSolver(model=m, joints=all_joints(m), constraints=cons, seed=5, max_attempts=6)
    Q, solved = solver.solve_ik_batch(targets, "attachment_site", m.keyframe("home").qpos)
    assert solved.sum() >= 8
    for i in np.flatnonzero(solved):
        assert all(c.valid_config(Q[i]) for c in cons)
        err = _pose_err(po, targets[i], Q[i])
        assert np.linalg.norm(err[:3]) <= 1e-3 and np.linalg.norm(err[3:]) <= 1e-3
    # an unreachable target is reported unsolved, reachable ones keep their rows
    far = SE3(SO3([1, 0, 0, 0]), [4.0, 4.0, 4.0])
    Q2, s2 = solver.solve_ik_batch([targets[0], far], "attachment_site", m.keyframe("home").qpos)
    assert s2.tolist() == [bool(solved[0]), False]
    assert solver.solve_ik(far, "attachment_site", None) == []


def test_solver_batch_many_distinct_inits():
    """A block of queries with distinct q_init rows takes the block sampler for the restarts."""
    m, cons, po = _ur5e()
    rng = np.random.default_rng(9)
    lo, hi = m.jnt_range.T
    inits, targets = [], []
    while len(inits) < 12:
        q = rng.uniform(lo, hi)
        if all(c.valid_config(q) for c in cons):
            inits.append(q)
    while len(targets) < 12:
        q = rng.uniform(lo, hi)
        if all(c.valid_config(q) for c in cons):
            p, r = po.site_pose(q)
            targets.append(SE3(SO3(r), p))
    solver = HostDLSIKSolver(model=m, joints=all_joints(m), constraints=cons, seed=5, max_attempts=6)
    G = solver._guess_block(np.asarray(inits))
    assert G.shape == (12, 6, m.nq)
    np.testing.assert_array_equal(G[:, 0], np.asarray(inits))
    assert all(c.valid_configs(G.reshape(-1, m.nq)).all() for c in cons)
    Q, solved = solver.solve_ik_batch(targets, "attachment_site", np.asarray(inits))
    assert solved.sum() >= 8
    for i in np.flatnonzero(solved):
        err = _pose_err(po, targets[i], Q[i])
        assert np.linalg.norm(err[:3]) <= 1e-3 and np.linalg.norm(err[3:]) <= 1e-3
