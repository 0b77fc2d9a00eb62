"""Trajectory re-validation helpers (reference: test/test_trajectory_utils.py, same cases) and the
batched ``generate_constrained_trajectory`` on oracle-backed constraints."""

import numpy as np
import pytest
from scipy.interpolate import make_interp_spline

from mjpl_b200 import models
from mjpl_b200.trajectory import Trajectory, TrajectoryGenerator, first_invalid_position, generate_constrained_trajectory
from mjpl_b200.trajectory.utils import _add_intermediate_waypoint, _waypoint_timing
from tests.doubles import OracleCollisionConstraint, OracleJointLimitConstraint

WAYPOINTS = [np.array([0, 0]), np.array([1, 1]), np.array([2, 1]), np.array([2, 0])]


class SplineGenerator(TrajectoryGenerator):
    """Interpolating spline through the waypoints sampled at n points (what the reference's test builds by hand)."""

    def __init__(self, n=100, k=3):
        self.n, self.k, self.calls = n, k, 0

    def generate_trajectory(self, waypoints):
        self.calls += 1
        x = np.linspace(0, 1, len(waypoints))
        spl = make_interp_spline(x, np.stack(waypoints).astype(float), k=min(self.k, len(waypoints) - 1))
        pos = spl(np.linspace(0, 1, self.n))
        return Trajectory(dt=1 / self.n, q_init=pos[0], positions=[r for r in pos], velocities=[], accelerations=[])


def _traj():
    return SplineGenerator().generate_trajectory(WAYPOINTS)


def test_waypoint_timing():
    # reference test/test_trajectory_utils.py:31-41
    traj = _traj()
    splx = np.linspace(0, 1, len(WAYPOINTS))
    times = _waypoint_timing(WAYPOINTS, traj)
    assert len(times) == len(splx)
    assert all(abs(a - b) <= traj.dt for a, b in zip(splx, times))
    assert all(times[i] < times[i + 1] for i in range(len(times) - 1))
    with pytest.raises(ValueError, match="at least two waypoints"):
        _waypoint_timing(WAYPOINTS[:1], traj)


def test_add_intermediate_waypoint():
    # reference test/test_trajectory_utils.py:43-131
    traj = _traj()
    splx = np.linspace(0, 1, len(WAYPOINTS))
    times = _waypoint_timing(WAYPOINTS, traj)
    for outside in (-1, 2):
        wp = list(WAYPOINTS)
        assert not _add_intermediate_waypoint(wp, times, outside)
        assert len(wp) == len(WAYPOINTS)
    cases = [(splx[1] + (splx[2] - splx[1]) / 3, 2, (1, 2)), (times[0], 1, (0, 1)), (times[-1], 3, (2, 3)), (times[1], 1, (0, 1))]
    for stamp, where, (a, b) in cases:
        wp = list(WAYPOINTS)
        assert _add_intermediate_waypoint(wp, times, stamp)
        assert len(wp) == len(WAYPOINTS) + 1
        np.testing.assert_allclose(wp[where], (WAYPOINTS[a] + WAYPOINTS[b]) / 2, rtol=0, atol=1e-8)
        rest = wp[:where] + wp[where + 1:]
        assert all(np.array_equal(x, y) for x, y in zip(rest, WAYPOINTS))
    with pytest.raises(ValueError, match="must be the same length"):
        _add_intermediate_waypoint([], [0.0], 0.0)


def test_generate_constrained_trajectory_inserts_waypoints_until_valid():
    """two_dof_ball: a spline through three valid waypoints overshoots into the wall region; the
    loop must add midpoints until every sample is valid, and every sample is checked as one block."""
    m = models.load("two_dof_ball")
    cons = [OracleJointLimitConstraint(m), OracleCollisionConstraint(m)]
    ok = lambda q: all(c.valid_config(np.asarray(q, float)) for c in cons)
    # find the edge of the valid region along x at y = 0
    xs = np.linspace(0.0, 1.0, 201)
    edge = next(x for x in xs if not ok([x, 0.0]))
    assert 0.1 < edge < 1.0
    a, b, c = np.array([0.0, 0.0]), np.array([edge - 0.03, 0.25]), np.array([0.0, 0.5])
    assert ok(a) and ok(b) and ok(c)
    waypoints = [a, b, c]
    gen = SplineGenerator(n=400, k=2)
    first = gen.generate_trajectory(list(waypoints))
    traj = generate_constrained_trajectory(waypoints, gen, cons)
    assert traj is not None
    assert first_invalid_position(traj, cons) == -1
    assert all(ok(p) for p in traj.positions[::7])
    if first_invalid_position(first, cons) >= 0:       # the unconstrained spline did overshoot
        assert len(waypoints) > 3 and gen.calls > 2
    # a generator that fails, and constraints nothing can satisfy
    class Never(TrajectoryGenerator):
        def generate_trajectory(self, waypoints):
            return None
    assert generate_constrained_trajectory([a, c], Never(), cons) is None
    bad = [np.array([0.0, 0.0]), np.array([edge + 0.2, 0.0])]   # the far end is in collision
    assert generate_constrained_trajectory(bad, SplineGenerator(n=50, k=1), cons) is None
    assert first_invalid_position(Trajectory(0.1, a, [], [], []), cons) == -1
