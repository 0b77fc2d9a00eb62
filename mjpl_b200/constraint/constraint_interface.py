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

    #: True (the safe default, and what the reference assumes of every constraint): ``apply`` may
    #: move ``q``, so planners call it step by step.  A constraint whose ``apply`` is exactly
    #: ``q if valid_config(q) else None`` sets this to False; whole extend chains can then be
    #: validated as one block through ``valid_configs``.
    projects: bool = True

    def valid_configs(self, Q) -> np.ndarray:
        """Batched twin of :meth:`valid_config`; default = one scalar call per row."""
        Q = np.asarray(Q)
        return np.fromiter((bool(self.valid_config(q)) for q in Q), dtype=bool, count=len(Q))
