"""Pluggable inverse kinematics (the ``solver=`` argument of ``RRT.plan_to_pose(s)`` and
``cartesian_plan``).  Mirrors the reference's interface, ``src/mjpl/inverse_kinematics/
ik_solver_interface.py:7-28``: one method, ``solve_ik``."""

from __future__ import annotations

import abc

import numpy as np

from ..lie import SE3


class IKSolver(abc.ABC):
    """Anything that can turn a site pose into joint configurations."""

    @abc.abstractmethod
    def solve_ik(self, pose: SE3, site: str, q_init_guess: np.ndarray | None) -> list[np.ndarray]:
        """Configurations that put ``site`` at ``pose``.

        ``pose`` is expressed in the world frame; ``q_init_guess`` seeds the search (``None``
        means the model's default configuration).  The result is empty when no solution was
        found; every returned configuration is a full ``qpos`` vector.
        """
