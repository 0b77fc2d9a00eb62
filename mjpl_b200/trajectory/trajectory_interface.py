"""What a trajectory is and what produces one.

Same shape as the reference's ``src/mjpl/trajectory/trajectory_interface.py`` so that generators
written against it plug in unchanged; the reference's own generators wrap Ruckig and TOPP-RA,
which are third-party libraries outside the validity path.
"""

from __future__ import annotations

import abc
import dataclasses

import numpy as np


@dataclasses.dataclass(frozen=True)
class Trajectory:
    """Sampled joint-space motion.

    ``positions[k]``, ``velocities[k]`` and ``accelerations[k]`` describe the state at time
    ``(k + 1) * dt``; ``q_init`` is the state at time zero, so ``n`` samples span ``n * dt`` seconds.
    """

    dt: float
    q_init: np.ndarray
    positions: list[np.ndarray]
    velocities: list[np.ndarray]
    accelerations: list[np.ndarray]


class TrajectoryGenerator(abc.ABC):
    """Turns a list of waypoints into a :class:`Trajectory`."""

    @abc.abstractmethod
    def generate_trajectory(self, waypoints: list[np.ndarray]) -> Trajectory | None:
        """The trajectory through ``waypoints``; ``None`` when none can be produced."""
