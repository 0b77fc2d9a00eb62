"""Planners on the batched validity engine: sequential and lock-step bi-RRT, Cartesian paths,
shortcut smoothing."""

from .utils import path_length, smooth_path  # noqa: I001
from .rrt import RRT
from .batched_rrt import BatchedRRT
from .cartesian_planner import cartesian_plan

__all__ = ["RRT", "BatchedRRT", "cartesian_plan", "smooth_path", "path_length"]
