"""Cartesian path following (reference: ``src/mjpl/planning/cartesian_planner.py``).

Same algorithm and argument meaning: interpolate the pose list so that adjacent poses are within
``lin_threshold`` / ``ori_threshold`` (:11-43), solve IK for each pose seeded with the previous
waypoint, keep the solutions that obey the constraints and (optionally) whose connecting interval
is collision free (:88-100), take the one closest to the previous waypoint (:103-104).  The poses
are inherently sequential (each IK starts from the last waypoint); what is batched is each pose's
work: all IK attempts are one launch, and the candidates' validity and interval checks are blocks.
"""

from __future__ import annotations

import numpy as np

from ..constraint.collision_constraint import CollisionConstraint
from ..constraint.constraint_interface import Constraint
from ..constraint.utils import obeys_constraints_batch
from ..inverse_kinematics.ik_solver_interface import IKSolver
from ..lie import SE3
from .utils import _valid_collision_interval


def _interpolate_poses(pose_from: SE3, pose_to: SE3, lin_threshold: float, ori_threshold: float) -> list[SE3]:
    """Poses from ``pose_from`` to ``pose_to`` no further than the thresholds apart."""
    if lin_threshold <= 0.0:
        raise ValueError("`lin_threshold` must be > 0.0")
    if ori_threshold <= 0.0:
        raise ValueError("`ori_threshold` must be > 0.0")
    pose_diff = pose_to.minus(pose_from)
    lin_dist = np.linalg.norm(pose_diff[:3])
    ori_dist = np.linalg.norm(pose_diff[3:])
    # a relative 1e-9 guard keeps a distance that is a whole number of thresholds up to rounding
    # (0.02 m at 0.01 m) from costing an extra step; the reference relies on mink's rounding here
    lin_steps = int(np.ceil(lin_dist / lin_threshold * (1.0 - 1e-9)))
    ori_steps = int(np.ceil(ori_dist / ori_threshold * (1.0 - 1e-9)))
    num_steps = max(lin_steps, ori_steps, 1)
    return [pose_from.interpolate(pose_to, alpha) for alpha in np.linspace(0, 1, num_steps + 1)]


def cartesian_plan(q_init: np.ndarray, poses: list[SE3], site: str, solver: IKSolver, constraints: list[Constraint],
                   collision_interval_check: tuple[float, CollisionConstraint] | None = None,
                   lin_threshold: float = 0.01, ori_threshold: float = 0.1) -> list[np.ndarray]:
    """Joint configurations that follow the Cartesian path ``poses`` (world frame) with ``site``,
    starting from ``q_init``; an empty list if some pose cannot be reached."""
    if not site:
        raise ValueError("`site` must be defined.")
    if not poses:
        return [q_init]
    interpolated = [poses[0]]
    for i in range(len(poses) - 1):
        interpolated.extend(_interpolate_poses(poses[i], poses[i + 1], lin_threshold, ori_threshold)[1:])
    waypoints = [q_init]
    for p in interpolated:
        cands = solver.solve_ik(p, site, q_init_guess=waypoints[-1])
        if cands and constraints:
            ok = np.asarray(obeys_constraints_batch(np.asarray(cands, dtype=np.float64), constraints), dtype=bool)
            cands = [q for q, k in zip(cands, ok) if k]
        if cands and collision_interval_check:
            cands = [q for q in cands if _valid_collision_interval(waypoints[-1], q, *collision_interval_check)]
        if not cands:
            return []
        waypoints.append(min(cands, key=lambda q: np.linalg.norm(q - waypoints[-1])))
    return waypoints
