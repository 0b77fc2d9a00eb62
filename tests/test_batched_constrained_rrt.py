"""Lock-step bi-RRT with a projecting constraint (BASELINE config #4 in miniature, on the CPU
doubles): every query of the batch must get exactly the path the reference's sequential algorithm
(``RRT`` with the same per-query seed) produces, and every waypoint must satisfy all constraints."""

import numpy as np
import pytest

from mjpl_b200 import all_joints, apply_constraints, apply_constraints_batch, models
from mjpl_b200.planning import RRT, BatchedRRT
from tests.doubles import OracleCollisionConstraint, OracleJointLimitConstraint, OraclePoseConstraint

INF = (-np.inf, np.inf)


def _ur5e_constraints():
    m = models.load("ur5e_scene")
    q0 = m.keyframe("home").qpos.copy()
    probe = OraclePoseConstraint(m, "attachment_site", [0, 0, 0], [1, 0, 0, 0], [INF] * 6)
    p, r = probe.po.site_pose(q0)
    pose = OraclePoseConstraint(m, "attachment_site", p, r, [INF, INF, (-0.05, 0.05), (-0.1, 0.1), (-0.1, 0.1), INF], q_step=0.2)
    return m, q0, [OracleJointLimitConstraint(m), OracleCollisionConstraint(m), pose]


def test_apply_constraints_batch_matches_scalar():
    m, q0, cons = _ur5e_constraints()
    rng = np.random.default_rng(0)
    Q = q0 + rng.uniform(-0.3, 0.3, size=(60, m.nq))
    out, ok = apply_constraints_batch(np.tile(q0, (60, 1)), Q, cons)
    n_ok = 0
    for i in range(60):
        want = apply_constraints(q0, Q[i], cons)
        assert (want is not None) == bool(ok[i])
        if want is not None:
            np.testing.assert_array_equal(out[i], want)
            n_ok += 1
    assert 5 < n_ok < 60
    # no projecting constraint: rows are only accepted or rejected
    out2, ok2 = apply_constraints_batch(Q, Q, cons[:2])
    np.testing.assert_array_equal(out2, Q)
    assert ok2.tolist() == [all(c.valid_config(q) for c in cons[:2]) for q in Q]


def test_batched_constrained_rrt_equals_sequential():
    m, q0, cons = _ur5e_constraints()
    joints = all_joints(m)
    # goals: project random perturbations of home onto the constraint manifold
    rng = np.random.default_rng(3)
    goals = []
    while len(goals) < 4:
        q = apply_constraints(q0, q0 + rng.uniform(-0.5, 0.5, m.nq), cons)
        if q is not None and np.linalg.norm(q - q0) > 0.3:
            goals.append(q)
    goals = np.asarray(goals)
    B = len(goals)
    kw = dict(max_planning_time=120.0, epsilon=0.1, goal_biasing_probability=0.2)
    batched = BatchedRRT(m, joints, cons, seed=10, **kw)
    paths = batched.plan(np.tile(q0, (B, 1)), goals)
    assert batched.stats["solved"] == B
    for b in range(B):
        want = RRT(m, joints, cons, seed=10 + b, **kw).plan_to_config(q0, goals[b])
        assert len(want) >= 2 and len(paths[b]) == len(want)
        for x, y in zip(paths[b], want):
            np.testing.assert_array_equal(x, y)
        P = np.asarray(paths[b])
        np.testing.assert_array_equal(P[0], q0)
        np.testing.assert_array_equal(P[-1], goals[b])
        assert all(c.valid_configs(P).all() for c in cons)
