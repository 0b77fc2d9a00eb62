"""``CollisionConstraint`` / ``CollisionRuleset`` on the B200 engine.

Reference: ``src/mjpl/constraint/collision_constraint.py``.  There, every call copies ``q``
into an ``MjData``, runs ``mj_kinematics`` + ``mj_collision`` and tests the contact list against
an allow-list of body pairs (:26-30, :66-95).  Here the allow-list is folded into the static
geom-pair table once (a contact between allowed bodies can never invalidate a configuration,
so those pairs are simply not tested) and FK + broad phase + narrow phase run on the GPU for a
whole block of rows; ``valid_config(q)`` is a block of one row through the same kernels.
"""

from __future__ import annotations

import numpy as np

from .. import engine as _engine
from .constraint_interface import Constraint


class CollisionRuleset:
    """Which body pairs may be in collision (reference :36-95, same semantics).

    ``obeys_ruleset`` works on an explicit ``(n,2)`` matrix of colliding geom ids, exactly like
    the reference; the engine never materialises such a matrix, it drops allowed pairs up front.
    """

    def __init__(self, model, allowed_collision_bodies: list[tuple[str, str]] = []) -> None:
        self.model = model
        self.allowed_collisions: np.ndarray | None = None
        if allowed_collision_bodies:
            ids = _engine.allowed_body_ids(model, allowed_collision_bodies)
            self.allowed_collisions = np.sort(np.asarray(ids, dtype=np.int64), axis=1)

    def obeys_ruleset(self, collision_geometries: np.ndarray) -> bool:
        g = np.asarray(collision_geometries)
        if g.ndim != 2 or g.shape[1] != 2:
            raise ValueError("`collision_geometries` must be a nx2 matrix.")
        if g.shape[0] == 0:
            return True
        if self.allowed_collisions is None:
            return False
        bodies = np.sort(np.asarray(self.model.geom_bodyid)[g.astype(np.int64)], axis=1)
        allowed = {tuple(p) for p in self.allowed_collisions.tolist()}
        return all(tuple(b) in allowed for b in bodies.tolist())


class CollisionConstraint(Constraint):
    """Constraint that enforces collision rules on a configuration."""

    projects = False   # apply() never moves q: chains of configurations can be validated as one block

    def __init__(self, model, allowed_collision_bodies: list[tuple[str, str]] = []) -> None:
        self.model = model
        self.cr = CollisionRuleset(model, allowed_collision_bodies)
        self.allowed_collision_bodies = list(allowed_collision_bodies)
        self.engine = _engine.get_engine(model, self.allowed_collision_bodies)

    def valid_config(self, q: np.ndarray) -> bool:
        return bool(self.valid_configs(np.asarray(q, dtype=np.float64)[None, :])[0])

    def valid_configs(self, Q):
        return self.engine.valid_configs(Q, _engine.CHECK_COLLISION)

    def valid_edges(self, Q0, Q1, step_dist: float, want_first_bad: bool = False):
        """Batched ``_valid_collision_interval`` (reference ``planning/utils.py:188-216``)."""
        return self.engine.valid_edges(Q0, Q1, step_dist, _engine.CHECK_COLLISION, want_first_bad)

    def apply(self, q_old: np.ndarray, q: np.ndarray) -> np.ndarray | None:
        return q if self.valid_config(q) else None
