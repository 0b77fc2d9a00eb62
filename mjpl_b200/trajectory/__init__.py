from .trajectory_interface import Trajectory, TrajectoryGenerator
from .utils import first_invalid_position, generate_constrained_trajectory

__all__ = ("Trajectory", "TrajectoryGenerator", "first_invalid_position", "generate_constrained_trajectory")
