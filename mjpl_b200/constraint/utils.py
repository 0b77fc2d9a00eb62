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
    tensor in -> tensor out).  The built-in non-projecting constraints on one model are fused into
    one kernel launch (also when other constraints are in the list); anything else is AND-ed
    constraint by constraint."""
    fused = _fusable(constraints)
    if fused is not None:
        eng, flags = fused
        return eng.valid_configs(Q, flags)
    builtin = [c for c in constraints if type(c) in (JointLimitConstraint, CollisionConstraint)]
    rest = [c for c in constraints if type(c) not in (JointLimitConstraint, CollisionConstraint)]
    out = None
    if len(builtin) > 1 and _fusable(builtin) is not None:
        eng, flags = _fusable(builtin)
        out = eng.valid_configs(Q, flags)
    else:
        rest = list(constraints)
    for c in rest:
        v = c.valid_configs(Q)
        out = v if out is None else (out & v)
    if out is None:
        n = len(Q)
        return np.ones(n, dtype=bool)
    return out


def apply_constraints_batch(Q_old: np.ndarray, Q: np.ndarray, constraints: list[Constraint]):
    """Batched :func:`apply_constraints` -> ``(Q_constrained, ok)``: every constraint is applied in
    order to the rows that are still alive (a projecting constraint moves them, the others only
    accept or reject), then the survivors are re-validated against all constraints as one block
    (reference ``constraint/utils.py:38-43``).  Rows with ``ok == False`` are the ones for which the
    reference returns ``None``; their output row is unspecified."""
    Q_old = np.asarray(Q_old, dtype=np.float64)
    out = np.array(Q, dtype=np.float64, copy=True)
    ok = np.ones(len(out), dtype=bool)
    for c in constraints:
        idx = np.flatnonzero(ok)
        if not len(idx):
            break
        if getattr(c, "projects", True):
            if hasattr(c, "apply_batch"):
                moved, good = c.apply_batch(Q_old[idx], out[idx])
                out[idx] = np.asarray(moved, dtype=np.float64)
                ok[idx] = np.asarray(good, dtype=bool)
            else:   # a constraint written against the reference interface only: one apply() per row
                for i in idx:
                    r = c.apply(Q_old[i], out[i])
                    ok[i] = r is not None
                    if r is not None:
                        out[i] = r
        else:
            ok[idx] = np.asarray(c.valid_configs(out[idx]), dtype=bool)
    # Re-validation: a non-projecting constraint that ran after the last projection has already
    # seen the final rows, so only the constraints up to that projection are checked again (the
    # answer is the reference's; it re-checks everything, :43).
    proj = [i for i, c in enumerate(constraints) if getattr(c, "projects", True)]
    again = constraints[: proj[-1] + 1] if proj else []
    idx = np.flatnonzero(ok)
    if len(idx) and again:
        ok[idx] = np.asarray(obeys_constraints_batch(out[idx], again), dtype=bool)
    return out, ok
