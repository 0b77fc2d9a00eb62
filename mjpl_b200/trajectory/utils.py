"""Constraint re-validation of generated trajectories.

Reference: ``src/mjpl/trajectory/utils.py``.  ``generate_constrained_trajectory`` (:8-56) checks
every trajectory sample with ``obeys_constraints`` one at a time -- thousands of configurations
at dt = 2 ms.  Here all samples of a trajectory are one block through the validity engine
(``first_invalid_position``), and the first failing index -- the only thing the reference's loop
uses (:40-47) -- is read off the mask.  Waypoint timing (:59-84) and waypoint insertion (:87-118)
are the reference's algorithms.
"""

from __future__ import annotations

import numpy as np

from ..constraint.constraint_interface import Constraint
from ..constraint.utils import apply_constraints, obeys_constraints_batch
from .trajectory_interface import Trajectory, TrajectoryGenerator


def first_invalid_position(trajectory: Trajectory, constraints: list[Constraint]) -> int:
    """Index of the first trajectory position that violates ``constraints``, or -1."""
    if not trajectory.positions or not constraints:
        return -1
    ok = np.asarray(obeys_constraints_batch(np.stack(trajectory.positions).astype(np.float64), constraints), dtype=bool)
    bad = np.flatnonzero(~ok)
    return int(bad[0]) if len(bad) else -1


def generate_constrained_trajectory(waypoints: list[np.ndarray], generator: TrajectoryGenerator,
                                    constraints: list[Constraint]) -> Trajectory | None:
    """Generate a trajectory that follows ``waypoints`` and obeys ``constraints``.

    Generate; if a sample violates the constraints, insert a constrained waypoint in the middle of
    the waypoint segment that sample belongs to; repeat (section 3.5 of Richter et al., ISRR 2013).
    ``waypoints`` is extended in place, as in the reference.
    """
    while True:
        traj = generator.generate_trajectory(waypoints)
        if traj is None:
            return None
        i = first_invalid_position(traj, constraints)
        if i < 0:
            return traj
        path_timestamps = _waypoint_timing(waypoints, traj)
        trajectory_timestamp = (i + 1) * traj.dt
        if not _add_intermediate_waypoint(waypoints, path_timestamps, trajectory_timestamp, constraints):
            # the intermediate waypoint cannot obey the constraints
            return None


def _waypoint_timing(waypoints: list[np.ndarray], trajectory: Trajectory) -> list[float]:
    """Timestamps of ``waypoints`` along ``trajectory``: 0 for the first, the duration for the
    last, the time of the closest trajectory position for the ones in between."""
    if len(waypoints) < 2:
        raise ValueError("There must be at least two waypoints defined.")
    timestamps = [0.0]
    if len(waypoints) > 2:
        positions = np.stack(trajectory.positions)
        inner = np.stack(waypoints[1:-1])
        d2 = ((positions[None, :, :] - inner[:, None, :]) ** 2).sum(axis=2)
        timestamps.extend(((np.argmin(d2, axis=1) + 1) * trajectory.dt).tolist())
    timestamps.append(len(trajectory.positions) * trajectory.dt)
    return timestamps


def _add_intermediate_waypoint(waypoints: list[np.ndarray], timing: list[float], timestamp: float,
                               constraints: list[Constraint] = []) -> bool:
    """Insert a constrained midpoint into the waypoint segment that contains ``timestamp``."""
    if len(waypoints) != len(timing):
        raise ValueError("`waypoints` and `timing` must be the same length.")
    for i in range(len(waypoints) - 1):
        if timing[i] <= timestamp <= timing[i + 1]:
            mid = (waypoints[i] + waypoints[i + 1]) / 2
            constrained = apply_constraints(mid, mid, constraints)
            if constrained is None:
                return False
            waypoints.insert(i + 1, constrained)
            return True
    return False
