"""The ``Constraint`` boundary of the validity path.

Same abstract interface as the reference (``src/mjpl/constraint/constraint_interface.py:6-33``:
``valid_config(q) -> bool`` and ``apply(q_old, q) -> q | None``) plus one optional batched
method, ``valid_configs(Q) -> (n,) bool``, that the batched planners call with whole blocks of
configurations.  A constraint that does not override it is evaluated row by row.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np


class Constraint(ABC):
    """Abstract base class for a constraint."""

    @abstractmethod
    def valid_config(self, q: np.ndarray) -> bool:
        """True if configuration ``q`` (full ``(nq,)`` vector) obeys the constraint."""

    @abstractmethod
    def apply(self, q_old: np.ndarray, q: np.ndarray) -> np.ndarray | None:
        """A configuration derived from ``q`` that obeys the constraint, or ``None``."""

    #: constraints that never change ``q`` in ``apply`` (``q if valid else None``) can be
    #: evaluated for a whole extend chain at once; projecting constraints set this to False.
    projects: bool = False

    def valid_configs(self, Q) -> np.ndarray:
        """Batched twin of :meth:`valid_config`; default = one scalar call per row."""
        Q = np.asarray(Q)
        return np.fromiter((bool(self.valid_config(q)) for q in Q), dtype=bool, count=len(Q))
