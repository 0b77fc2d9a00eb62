from .collision_constraint import CollisionConstraint, CollisionRuleset
from .constraint_interface import Constraint
from .joint_limit_constraint import JointLimitConstraint
from .pose_constraint import PoseConstraint
from .utils import apply_constraints, apply_constraints_batch, obeys_constraints, obeys_constraints_batch

__all__ = (
    "Constraint",
    "CollisionConstraint",
    "CollisionRuleset",
    "JointLimitConstraint",
    "PoseConstraint",
    "apply_constraints",
    "apply_constraints_batch",
    "obeys_constraints",
    "obeys_constraints_batch",
)
