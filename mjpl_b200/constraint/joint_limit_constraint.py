"""``JointLimitConstraint`` on the B200 engine.

Reference: ``src/mjpl/constraint/joint_limit_constraint.py:7-23`` --
``np.all((q >= lower) & (q <= upper))`` with ``lower/upper = model.jnt_range`` columns, closed
interval, every joint.  The comparison runs in the fused validity kernel (flag
``MJB_CHECK_LIMITS``), in fp64 on the fp32 row, so a row that is representable in fp32 gets
exactly the reference's answer.  The scalar call is a batch of one through the same kernel.
"""

from __future__ import annotations

import numpy as np

from .. import engine as _engine
from .constraint_interface import Constraint


class JointLimitConstraint(Constraint):
    """Constraint that enforces joint limits on a configuration."""

    projects = False   # apply() never moves q: chains of configurations can be validated as one block

    def __init__(self, model) -> None:
        self.model = model
        self.lower = model.jnt_range[:, 0]
        self.upper = model.jnt_range[:, 1]
        self.engine = _engine.get_engine(model, ())

    def valid_config(self, q: np.ndarray) -> bool:
        return bool(self.valid_configs(np.asarray(q, dtype=np.float64)[None, :])[0])

    def valid_configs(self, Q):
        return self.engine.valid_configs(Q, _engine.CHECK_LIMITS)

    def apply(self, q_old: np.ndarray, q: np.ndarray) -> np.ndarray | None:
        return q if self.valid_config(q) else None
