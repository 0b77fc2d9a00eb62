from .batched_rrt import BatchedRRT
from .cartesian_planner import cartesian_plan
from .rrt import RRT
from .utils import path_length, smooth_path

__all__ = ("BatchedRRT", "RRT", "cartesian_plan", "path_length", "smooth_path")
