from .batched_rrt import BatchedRRT
from .rrt import RRT
from .utils import path_length, smooth_path

__all__ = ("BatchedRRT", "RRT", "path_length", "smooth_path")
