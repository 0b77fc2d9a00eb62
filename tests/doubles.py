"""Test doubles: Constraint implementations backed by the fp64 CPU oracle.

They let the host-side planner logic (chain building, stop rules, tree bookkeeping, batched
lock-step RRT) run in the ``-m "not gpu"`` suite.  Test infrastructure only -- the product's
constraints call the CUDA engine and have no CPU path.
"""

from __future__ import annotations

import numpy as np

import oracle
from mjpl_b200.constraint.constraint_interface import Constraint


from oracle.constraints import OracleCollisionConstraint, OracleJointLimitConstraint  # noqa: E402,F401


class OraclePoseConstraint(Constraint):
    """``PoseConstraint`` double on the numpy restatement (``oracle.PoseOracle``); projects."""

    projects = True

    def __init__(self, model, site, ref_pos, ref_quat, box, tolerance=0.001, q_step=0.05):
        self.model = model
        self.po = oracle.PoseOracle(model, site, ref_pos, ref_quat, box, tolerance=tolerance, q_step=q_step)

    def valid_config(self, q):
        return bool(self.po.valid_config(np.asarray(q, float)))

    def valid_configs(self, Q):
        return np.array([self.valid_config(q) for q in np.asarray(Q, float)], dtype=bool)

    def apply(self, q_old, q):
        return self.po.apply(np.asarray(q_old, float), np.asarray(q, float))

    def apply_batch(self, Q_old, Q):
        Q_old, Q = np.asarray(Q_old, float), np.asarray(Q, float)
        out, ok = Q.copy(), np.zeros(len(Q), dtype=bool)
        for i in range(len(Q)):
            r = self.apply(Q_old[i], Q[i])
            if r is not None:
                out[i], ok[i] = r, True
        return out, ok
