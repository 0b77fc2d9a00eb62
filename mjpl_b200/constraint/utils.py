"""Constraint dispatch: the reference's two helpers plus their batched twins.

Reference: ``src/mjpl/constraint/utils.py:6-43``.
"""

from __future__ import annotations

import numpy as np

from .. import engine as _engine
from .collision_constraint import CollisionConstraint
from .constraint_interface import Constraint
from .joint_limit_constraint import JointLimitConstraint


def obeys_constraints(q: np.ndarray, constraints: list[Constraint]) -> bool:
    """True if ``q`` obeys each constraint (short-circuit AND, reference :6-19)."""
    for c in constraints:
        if not c.valid_config(q):
            return False
    return True


def apply_constraints(q_old: np.ndarray, q: np.ndarray, constraints: list[Constraint]) -> np.ndarray | None:
    """Apply constraints in order, then re-validate the result (reference :22-43)."""
    q_constrained = q
    for c in constraints:
        q_constrained = c.apply(q_old, q_constrained)
        if q_constrained is None:
            return None
    return q_constrained if obeys_constraints(q_constrained, constraints) else None


def _fusable(constraints):
    """``[JointLimitConstraint, CollisionConstraint]`` (any order, at most one of each) on ONE model
    -> ``(engine, flags)`` for a single fused kernel launch (limits + FK + collision); else None."""
    if not constraints:
        return None
    flags, eng, model = 0, None, None
    for c in constraints:
        if type(c) is JointLimitConstraint and not flags & _engine.CHECK_LIMITS:
            flags |= _engine.CHECK_LIMITS
        elif type(c) is CollisionConstraint and not flags & _engine.CHECK_COLLISION:
            flags |= _engine.CHECK_COLLISION
            eng = c.engine
        else:
            return None
        if model is not None and c.model is not model:
            return None
        model = c.model
    return (eng if eng is not None else constraints[0].engine), flags


def obeys_constraints_batch(Q, constraints: list[Constraint]):
    """Batched :func:`obeys_constraints`: ``(n,nq) -> (n,) bool`` (numpy in -> numpy out,
    tensor in -> tensor out).  Non-projecting built-in constraints on one model are fused into
    one kernel launch; anything else is AND-ed constraint by constraint."""
    fused = _fusable(constraints)
    if fused is not None:
        eng, flags = fused
        return eng.valid_configs(Q, flags)
    out = None
    for c in constraints:
        v = c.valid_configs(Q)
        out = v if out is None else (out & v)
    if out is None:
        n = len(Q)
        return np.ones(n, dtype=bool)
    return out
