"""Cartesian planner (reference: test/test_cartesian_planner.py, same cases).  The planner's host
logic runs here on the oracle-backed constraint doubles and the IK core executed on the CPU."""

import numpy as np
import pytest

import oracle
from mjpl_b200 import all_joints, cartesian_plan, models
from mjpl_b200.lie import SE3, SO3
from mjpl_b200.planning.cartesian_planner import _interpolate_poses
from tests.doubles import OracleCollisionConstraint, OracleJointLimitConstraint
from tests.test_ik_host import HostDLSIKSolver


def approx(p1, p2, tol=1e-9):
    np.testing.assert_allclose(p1.translation(), p2.translation(), rtol=0, atol=tol)
    a, b = p1.rotation().parameters(), p2.rotation().parameters()
    assert min(np.abs(a - b).max(), np.abs(a + b).max()) <= tol


def test_interpolate_pose():
    # reference test/test_cartesian_planner.py:27-112
    start = SE3.from_rotation_and_translation(SO3.from_x_radians(0), np.array([0, 0, 0]))
    end = SE3.from_rotation_and_translation(SO3.from_x_radians(np.pi), np.array([1, 0, 0]))
    poses = _interpolate_poses(start, end, np.inf, np.inf)
    assert len(poses) == 2 and poses[0] == start and poses[1] == end
    poses = _interpolate_poses(start, end, 0.65, np.inf)
    assert len(poses) == 3 and poses[0] == start and poses[2] == end
    approx(poses[1], SE3.from_rotation_and_translation(SO3.from_x_radians(np.pi / 2), np.array([0.5, 0.0, 0.0])))
    quarter = [SE3.from_rotation_and_translation(SO3.from_x_radians(np.pi * f), np.array([f, 0.0, 0.0])) for f in (0.25, 0.5, 0.75)]
    for lin in (np.inf, 0.65):   # the axis that needs more steps wins
        poses = _interpolate_poses(start, end, lin, np.pi * 0.3)
        assert len(poses) == 5 and poses[0] == start and poses[4] == end
        for got, want in zip(poses[1:4], quarter):
            approx(got, want)
    for bad in (0.0, -1.0):
        with pytest.raises(ValueError, match="`lin_threshold` must be > 0"):
            _interpolate_poses(start, end, bad, np.inf)
        with pytest.raises(ValueError, match="`ori_threshold` must be > 0"):
            _interpolate_poses(start, end, np.inf, bad)


def test_se3_exp_log_round_trip():
    rng = np.random.default_rng(0)
    for _ in range(50):
        tg = rng.normal(size=6) * np.array([1, 1, 1, 0.8, 0.8, 0.8])
        np.testing.assert_allclose(SE3.exp(tg).log(), tg, atol=1e-12)
    a = SE3(SO3.from_rpy_radians(0.3, -0.2, 1.0), [0.1, 0.2, 0.3])
    b = SE3(SO3.from_rpy_radians(-0.5, 0.4, 0.2), [-0.3, 0.0, 0.5])
    np.testing.assert_allclose(a.interpolate(b, 0.5).minus(a), 0.5 * b.minus(a), atol=1e-12)
    assert a.interpolate(b, 0.0) == a and a.interpolate(b, 1.0) == b and not (a == b)


def test_cartesian_path():
    # reference test/test_cartesian_planner.py:114-200
    m = models.load("ur5e_scene")
    site = "attachment_site"
    cons = [OracleJointLimitConstraint(m), OracleCollisionConstraint(m)]
    po = oracle.PoseOracle(m, site, [0, 0, 0], [1, 0, 0, 0], [(-np.inf, np.inf)] * 6)
    q_init = m.keyframe("home").qpos.copy()
    p, r = po.site_pose(q_init)
    current = SE3(SO3(r), p)
    nxt = current.multiply(SE3.from_translation(np.array([0.02, 0.0, 0.0])))
    final = nxt.multiply(SE3.from_translation(np.array([0.0, 0.02, 0.0])))
    mid = current.multiply(SE3.from_translation(np.array([0.02, 0.01, 0.0])))
    solver = HostDLSIKSolver(model=m, joints=all_joints(m), constraints=[cons[1]], pos_tolerance=1e-3, ori_tolerance=1e-3,
                             seed=12345, max_attempts=5)
    wps = cartesian_plan(q_init, [nxt, final], site, solver, cons, lin_threshold=0.01, ori_threshold=0.1)
    assert len(wps) == 4
    assert all(c.valid_config(w) for w in wps for c in cons)
    np.testing.assert_equal(wps[0], q_init)
    for w, want in zip(wps[1:], (nxt, mid, final)):
        pw, rw = po.site_pose(w)
        err = want.minus(SE3(SO3(rw), pw))
        assert np.linalg.norm(err[:3]) <= 1e-3 and np.linalg.norm(err[3:]) <= 1e-3
    # with the interval check switched on the same path is found (tiny joint motions)
    wps2 = cartesian_plan(q_init, [nxt, final], site, solver, cons, collision_interval_check=(0.01, cons[1]))
    assert len(wps2) == 4
    # an unreachable pose ends the plan with an empty list
    far = SE3(SO3([1, 0, 0, 0]), [3.0, 3.0, 3.0])
    assert cartesian_plan(q_init, [nxt, far], site, solver, cons, lin_threshold=np.inf, ori_threshold=np.inf) == []


def test_invalid_args():
    m = models.load("ur5e_scene")
    solver = HostDLSIKSolver(m, all_joints(m))
    with pytest.raises(ValueError, match="site"):
        cartesian_plan(m.keyframe("home").qpos, [], "", solver, [])
