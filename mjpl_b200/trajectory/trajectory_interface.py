"""Trajectory container and generator interface (reference:
``src/mjpl/trajectory/trajectory_interface.py``).  The reference's generators wrap Ruckig and
TOPP-RA, which are not part of the validity path; any object with ``generate_trajectory`` fits."""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Trajectory:
    """Trajectory data: ``n`` states at increments of ``dt`` over ``t = [dt, n*dt]``."""

    dt: float
    q_init: np.ndarray
    positions: list[np.ndarray]
    velocities: list[np.ndarray]
    accelerations: list[np.ndarray]


class TrajectoryGenerator(ABC):
    """Abstract base class for generating trajectories."""

    @abstractmethod
    def generate_trajectory(self, waypoints: list[np.ndarray]) -> Trajectory | None:
        """A trajectory that follows ``waypoints``, or None if one cannot be generated."""
