"""``Constraint`` implementations backed by the fp64 CPU oracle.

TEST INFRASTRUCTURE / CPU BASELINE ONLY: used by ``tests/`` (host-side planner logic without a GPU)
and by ``bench.py``'s CPU planner baseline.  The product's constraints (``mjpl_b200.constraint``) call
the CUDA engine and have no CPU path.

The reference's constraints (``src/mjpl/constraint/joint_limit_constraint.py:7-23``,
``collision_constraint.py:7-33``) answer one configuration per call and its planner calls them step by
step (``planning/utils.py:139-164``); the ``Reference*`` classes below keep exactly that calling
pattern (``projects`` left at the interface's default, so no whole-chain block evaluation), which is
what the CPU baseline times.  The ``Oracle*`` classes also answer blocks, for the host-logic tests.
"""

from __future__ import annotations

import numpy as np

import oracle
from mjpl_b200.constraint.constraint_interface import Constraint


class OracleJointLimitConstraint(Constraint):
    projects = False

    def __init__(self, model):
        self.model = model
        self.orc = oracle.Oracle(model)

    def valid_config(self, q):
        return bool(self.orc.check(np.asarray(q, float), oracle.CHECK_LIMITS)[0])

    def valid_configs(self, Q):
        Q = np.asarray(Q, float)
        return self.orc.check(Q, oracle.CHECK_LIMITS) if len(Q) else np.zeros(0, bool)

    def apply(self, q_old, q):
        return q if self.valid_config(q) else None


class OracleCollisionConstraint(Constraint):
    projects = False

    def __init__(self, model, allowed_collision_bodies=()):
        self.model = model
        self.orc = oracle.Oracle(model, allowed_collision_bodies)

    def valid_config(self, q):
        return bool(self.orc.check(np.asarray(q, float), oracle.CHECK_COLLISION)[0])

    def valid_configs(self, Q):
        Q = np.asarray(Q, float)
        return self.orc.check(Q, oracle.CHECK_COLLISION) if len(Q) else np.zeros(0, bool)

    def valid_edges(self, Q0, Q1, step_dist, want_first_bad=False):
        res = [oracle.valid_collision_interval(self.orc, a, b, step_dist) for a, b in zip(np.asarray(Q0, float), np.asarray(Q1, float))]
        v = np.array([r[0] for r in res], dtype=bool)
        fb = np.array([r[1] for r in res], dtype=np.int32)
        return (v, fb) if want_first_bad else v

    def apply(self, q_old, q):
        return q if self.valid_config(q) else None


class ReferenceJointLimitConstraint(OracleJointLimitConstraint):
    """one configuration per call, the reference's calling pattern (CPU baseline)"""

    projects = True


class ReferenceCollisionConstraint(OracleCollisionConstraint):
    """one configuration per call, the reference's calling pattern (CPU baseline)"""

    projects = True
