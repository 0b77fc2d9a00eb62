"""IK for pose goals: the solver interface and the batched damped-least-squares solver."""

from .ik_solver_interface import IKSolver  # noqa: I001  (interface first: the solver imports it)
from .dls_ik_solver import DLSIKSolver

__all__ = ["IKSolver", "DLSIKSolver"]
