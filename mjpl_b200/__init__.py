"""mjpl_b200: B200-native configuration-validity engine behind mjpl's Constraint API.

Drop-in surface (same names as the reference's ``mjpl`` package for the validity path):
``CollisionConstraint``, ``JointLimitConstraint``, ``obeys_constraints``, ``apply_constraints``,
``RRT``, ``smooth_path``, ``path_length``, ``random_config``, ``all_joints``, ``qpos_idx``,
``qvel_idx`` -- plus the batched entry points ``obeys_constraints_batch``,
``Constraint.valid_configs``, ``CollisionConstraint.valid_edges`` and ``BatchedRRT``.
"""

from . import mjcf, models
from .constraint import (
    CollisionConstraint,
    CollisionRuleset,
    Constraint,
    JointLimitConstraint,
    PoseConstraint,
    apply_constraints,
    apply_constraints_batch,
    obeys_constraints,
    obeys_constraints_batch,
)
from .inverse_kinematics import DLSIKSolver, IKSolver
from .lie import SE3, SO3
from .engine import EngineUnavailable, ValidityEngine, get_engine
from .model import Model
from .planning import RRT, BatchedRRT, cartesian_plan, path_length, smooth_path
from .trajectory import Trajectory, TrajectoryGenerator, generate_constrained_trajectory
from .utils import all_joints, qpos_idx, qvel_idx, random_config, site_pose

__all__ = (
    "BatchedRRT",
    "CollisionConstraint",
    "CollisionRuleset",
    "DLSIKSolver",
    "IKSolver",
    "Constraint",
    "EngineUnavailable",
    "JointLimitConstraint",
    "Model",
    "PoseConstraint",
    "SE3",
    "SO3",
    "Trajectory",
    "TrajectoryGenerator",
    "RRT",
    "ValidityEngine",
    "all_joints",
    "apply_constraints",
    "apply_constraints_batch",
    "cartesian_plan",
    "generate_constrained_trajectory",
    "get_engine",
    "mjcf",
    "models",
    "obeys_constraints",
    "obeys_constraints_batch",
    "path_length",
    "qpos_idx",
    "qvel_idx",
    "random_config",
    "site_pose",
    "smooth_path",
)
