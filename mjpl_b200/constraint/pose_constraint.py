"""``PoseConstraint`` on the B200 engine (CBiRRT projection).

Reference: ``src/mjpl/constraint/pose_constraint.py`` -- same constructor, same error texts,
same algorithm (``valid_config`` :72-76, ``apply`` :78-91, displacement :93-123, RPY Jacobian
:125-171).  The per-row loop (site FK, geometric Jacobian, ``E_rpy``, 6x6 pseudo-inverse step,
limit / ``2*q_step`` abort) runs in fp64 in ``pose_kernel`` for a whole block of rows
(``apply_batch`` / ``valid_configs``); the scalar calls are blocks of one row.
This constraint *projects* (``projects = True``), so planners step it sequentially per query
and batch across queries.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _abi
from .. import engine as _engine
from ..lie import SE3
from .constraint_interface import Constraint
from .joint_limit_constraint import JointLimitConstraint


class PoseConstraint(Constraint):
    """Constraint that enforces pose constraints on a site."""

    projects = True

    def __init__(self, model, site: str, reference_frame: SE3,
                 x_translation=(-np.inf, np.inf), y_translation=(-np.inf, np.inf), z_translation=(-np.inf, np.inf),
                 roll=(-np.inf, np.inf), pitch=(-np.inf, np.inf), yaw=(-np.inf, np.inf),
                 tolerance: float = 0.001, q_step: float = 0.05, max_iterations: int = 1000) -> None:
        if tolerance < 0.0:
            raise ValueError("`tolerance` must be >= 0.")
        if q_step <= 0.0:
            raise ValueError("`q_step` must be > 0.")
        self.model = model
        self.C = np.array([x_translation, y_translation, z_translation, roll, pitch, yaw], dtype=np.float64)
        self.reference_frame = reference_frame
        self.C_T_world = reference_frame.inverse()
        self.site = site
        self.tolerance = tolerance
        self.q_step = q_step
        self.max_iterations = max_iterations  # the reference loops forever; a cap turns that into None
        self.site_id = model.site(site).id
        self.joint_limit_constraint = JointLimitConstraint(model)
        self.engine = _engine.get_engine(model, ())

    def _spec(self) -> _abi.PoseSpec:
        sp = _abi.PoseSpec()
        m, s = self.model, self.site_id
        sp.site_bodyid = int(m.site_bodyid[s])
        sp.site_pos[:] = [float(x) for x in m.site_pos[s]]
        sp.site_quat[:] = [float(x) for x in m.site_quat[s]]
        sp.ref_pos[:] = [float(x) for x in self.reference_frame.translation()]
        sp.ref_quat[:] = [float(x) for x in self.reference_frame.rotation().wxyz]
        sp.lower[:] = [float(x) for x in self.C[:, 0]]
        sp.upper[:] = [float(x) for x in self.C[:, 1]]
        sp.tolerance, sp.q_step = float(self.tolerance), float(self.q_step)  # read at call time (tests mutate q_step)
        return sp

    def _dev(self, Q):
        import torch

        e = self.engine
        if torch.is_tensor(Q):
            return Q.to(device=e.torch_device, dtype=torch.float64).contiguous(), ("cuda" if Q.is_cuda else "cpu")
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        if Q.ndim != 2 or Q.shape[1] != e.nq:
            raise ValueError(f"expected an (n, {e.nq}) array of configurations")
        return torch.from_numpy(Q).to(e.torch_device), "numpy"

    def valid_configs(self, Q):
        import torch

        e = self.engine
        with torch.cuda.device(e.device):
            q, kind = self._dev(Q)
            out = torch.empty(len(q), dtype=torch.uint8, device=e.torch_device)
            sp = self._spec()
            _abi.check(e._L.mjb_pose_valid(e._h, C.byref(sp), q.data_ptr(), len(q), out.data_ptr(), e._stream()))
            return e._back(out.bool(), kind)

    def apply_batch(self, Q_old, Q, want_iterations: bool = False):
        """Projection of many rows at once -> ``(Q_projected, ok)``; rows with ``ok == False`` are
        the ones for which the reference's ``apply`` returns ``None`` (their output row is a copy
        of the input)."""
        import torch

        e = self.engine
        with torch.cuda.device(e.device):
            q, kind = self._dev(Q)
            q_old, _ = self._dev(Q_old)
            out = q.clone()
            ok = torch.empty(len(q), dtype=torch.uint8, device=e.torch_device)
            iters = torch.zeros(len(q), dtype=torch.int32, device=e.torch_device)
            sp = self._spec()
            _abi.check(e._L.mjb_pose_project(e._h, C.byref(sp), q_old.data_ptr(), q.data_ptr(), len(q), int(self.max_iterations),
                                             out.data_ptr(), ok.data_ptr(), iters.data_ptr(), e._stream()))
            res = (e._back(out, kind), e._back(ok.bool(), kind))
            return res + (e._back(iters, kind),) if want_iterations else res

    def valid_config(self, q: np.ndarray) -> bool:
        return bool(self.valid_configs(np.asarray(q, dtype=np.float64)[None, :])[0])

    def apply(self, q_old: np.ndarray, q: np.ndarray) -> np.ndarray | None:
        q_old = np.asarray(q_old, dtype=np.float64)
        if q_old.shape != np.shape(q):  # the reference's tests pass np.array([]) when q_old is unused
            q_old = np.asarray(q, dtype=np.float64)
        out, ok = self.apply_batch(q_old[None, :], np.asarray(q, dtype=np.float64)[None, :])
        return out[0] if ok[0] else None
