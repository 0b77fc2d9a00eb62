from .dls_ik_solver import DLSIKSolver
from .ik_solver_interface import IKSolver

__all__ = ("DLSIKSolver", "IKSolver")
