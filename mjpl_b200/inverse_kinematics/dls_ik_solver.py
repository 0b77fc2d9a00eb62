"""Batched damped-least-squares IK on the B200 engine.

Takes the place of the reference's ``MinkIKSolver`` (``src/mjpl/inverse_kinematics/
mink_ik_solver.py``): same constructor arguments (minus the QP solver name and mink tasks,
which have no meaning here), same ``solve_ik`` contract -- a list with one configuration whose
site pose is within ``pos_tolerance`` / ``ori_tolerance`` of the target and which obeys the
constraints, or ``[]``.

What differs is the schedule.  The reference runs its attempts one after the other
(:93-116: iterate from the current guess; on convergence return if the constraints hold,
otherwise draw a random valid configuration and start over).  Here every attempt of every
target is one row of a single ``mjb_ik_solve`` launch (attempt 0 from ``q_init_guess``, attempt
``a > 0`` from ``random_config(..., seed + a - 1)``, the reference's seeds), the constraints are
checked on all converged rows as one block, and the lowest-numbered attempt that passed is
returned -- which is the row the sequential loop would have returned had it used the same
iteration.  The iteration itself is Levenberg-Marquardt on the geometric Jacobian rather than
mink's QP; solutions are therefore not the same configurations as the reference's, only
solutions of the same problem (IK is not on the parity path: SURVEY.md section 8(f)).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _abi
from .. import engine as _engine
from .. import utils
from ..constraint.constraint_interface import Constraint
from ..constraint.utils import obeys_constraints_batch
from ..lie import SE3
from .ik_solver_interface import IKSolver


class DLSIKSolver(IKSolver):
    """Damped-least-squares implementation of IKSolver (all attempts in one launch)."""

    def __init__(self, model, joints: list[str], constraints: list[Constraint] = [],
                 pos_tolerance: float = 1e-3, ori_tolerance: float = 1e-3, seed: int | None = None,
                 max_attempts: int = 1, iterations: int = 500, lm_damping: float = 0.1,
                 damping: float = 1e-9, max_step: float = 0.5):
        if not joints:
            raise ValueError("`joints` cannot be empty.")
        if max_attempts < 1:
            raise ValueError("`max_attempts` must be > 0.")
        if iterations < 1:
            raise ValueError("`iterations` must be > 0.")
        self.model = model
        self.joints = joints
        self.constraints = constraints
        self.pos_tolerance = pos_tolerance
        self.ori_tolerance = ori_tolerance
        self.seed = seed
        self.max_attempts = max_attempts
        self.iterations = iterations
        self.lm_damping = lm_damping
        self.damping = damping
        self.max_step = max_step
        self._engine = None
        self._mask = 0
        for name in joints:
            self._mask |= 1 << int(model.joint(name).id)

    # -- engine call ---------------------------------------------------------------------------
    @property
    def engine(self):
        if self._engine is None:
            self._engine = _engine.get_engine(self.model, ())
        return self._engine

    def _spec(self, site: str) -> _abi.IkSpec:
        m = self.model
        s = m.site(site).id
        sp = _abi.IkSpec()
        sp.site_bodyid = int(m.site_bodyid[s])
        sp.site_pos[:] = [float(x) for x in m.site_pos[s]]
        sp.site_quat[:] = [float(x) for x in m.site_quat[s]]
        sp.movable_mask = self._mask
        # the kernel's tolerance test uses |p_t - p| and the rotation angle; the reference measures
        # the se(3) log of the relative pose, which differs from those by O(angle^2): leave a margin
        sp.pos_tolerance = float(self.pos_tolerance) * 0.99
        sp.ori_tolerance = float(self.ori_tolerance) * 0.99
        sp.lm_damping, sp.damping, sp.max_step = float(self.lm_damping), float(self.damping), float(self.max_step)
        sp.iterations = int(self.iterations)
        return sp

    def solve_rows(self, target_pos, target_quat, q_init, site: str):
        """Raw kernel call: one IK iteration loop per row -> ``(Q, converged, iterations, errors)``."""
        import torch

        e = self.engine
        tp = np.ascontiguousarray(target_pos, dtype=np.float64).reshape(-1, 3)
        tq = np.ascontiguousarray(target_quat, dtype=np.float64).reshape(-1, 4)
        q0 = np.ascontiguousarray(q_init, dtype=np.float64).reshape(-1, e.nq)
        if not (len(tp) == len(tq) == len(q0)):
            raise ValueError("targets and initial guesses must have the same number of rows")
        n = len(q0)
        with torch.cuda.device(e.device):
            dev = e.torch_device
            d_tp, d_tq, d_q0 = (torch.from_numpy(a).to(dev) for a in (tp, tq, q0))
            out = torch.empty_like(d_q0)
            ok = torch.zeros(n, dtype=torch.uint8, device=dev)
            iters = torch.zeros(n, dtype=torch.int32, device=dev)
            errs = torch.zeros((n, 2), dtype=torch.float64, device=dev)
            sp = self._spec(site)
            _abi.check(e._L.mjb_ik_solve(e._h, C.byref(sp), d_tp.data_ptr(), d_tq.data_ptr(), d_q0.data_ptr(), n,
                                         out.data_ptr(), ok.data_ptr(), iters.data_ptr(), errs.data_ptr(), e._stream()))
            return out.cpu().numpy(), ok.bool().cpu().numpy(), iters.cpu().numpy(), errs.cpu().numpy()

    # -- reference interface -------------------------------------------------------------------
    def _guesses(self, q0: np.ndarray) -> np.ndarray:
        """Initial guesses of all attempts: q0, then the reference's seeded random configurations."""
        rows = [q0]
        for attempt in range(self.max_attempts - 1):
            _seed = self.seed + attempt if self.seed is not None else self.seed
            rows.append(utils.random_config(self.model, q0, self.joints, _seed, self.constraints))
        return np.asarray(rows, dtype=np.float64)

    def _guess_block(self, q0: np.ndarray) -> np.ndarray:
        """(n, attempts, nq) initial guesses.  A few distinct ``q0`` rows use the reference's seeded
        sequence per row; a large block draws the restart configurations of all rows together
        (one validity launch per round of rejection sampling instead of one per row)."""
        n, nq = q0.shape
        A = self.max_attempts
        G = np.empty((n, A, nq))
        uniq, inv = np.unique(q0, axis=0, return_inverse=True)
        if len(uniq) <= 8 or A == 1:
            per = np.stack([self._guesses(u) for u in uniq])
            return per[inv.reshape(-1)]
        G[:, 0] = q0
        q_idx = utils.qpos_idx(self.model, self.joints)
        lo, hi = self.model.jnt_range.T
        for a in range(1, A):
            rng = np.random.default_rng(None if self.seed is None else self.seed + a - 1)
            cand = q0.copy()
            todo = np.arange(n)
            while len(todo):
                cand[np.ix_(todo, q_idx)] = rng.uniform(lo, hi, size=(len(todo), len(lo)))[:, q_idx]
                ok = (np.asarray(obeys_constraints_batch(cand[todo], self.constraints), dtype=bool)
                      if self.constraints else np.ones(len(todo), bool))
                todo = todo[~ok]
            G[:, a] = cand
        return G

    def solve_ik_batch(self, poses: list[SE3], site: str, q_init_guesses=None):
        """IK for many targets at once -> ``(Q (n,nq), solved (n,) bool)``.

        ``q_init_guesses``: None (``qpos0`` for every target, what the reference's
        ``q_init_guess=None`` amounts to), one configuration, or one per target.
        """
        n = len(poses)
        nq = int(self.model.nq)
        if q_init_guesses is None:
            q0 = np.tile(np.asarray(self.model.qpos0, dtype=np.float64), (n, 1))
        else:
            q0 = np.asarray(q_init_guesses, dtype=np.float64)
            q0 = np.tile(q0, (n, 1)) if q0.ndim == 1 else q0
        if q0.shape != (n, nq):
            raise ValueError(f"expected {n} initial guesses of {nq} values")
        A = self.max_attempts
        G = self._guess_block(q0)
        tp = np.repeat(np.asarray([p.translation() for p in poses], dtype=np.float64).reshape(n, 3), A, axis=0)
        tq = np.repeat(np.asarray([p.rotation().wxyz for p in poses], dtype=np.float64).reshape(n, 4), A, axis=0)
        Q, conv, _, _ = self.solve_rows(tp, tq, G.reshape(n * A, nq), site)
        good = conv.copy()
        if self.constraints and conv.any():
            good[conv] = np.asarray(obeys_constraints_batch(Q[conv], self.constraints), dtype=bool)
        good = good.reshape(n, A)
        first = np.argmax(good, axis=1)
        solved = good.any(axis=1)
        return Q.reshape(n, A, nq)[np.arange(n), first], solved

    def solve_ik(self, pose: SE3, site: str, q_init_guess: np.ndarray | None) -> list[np.ndarray]:
        Q, solved = self.solve_ik_batch([pose], site, None if q_init_guess is None else np.asarray(q_init_guess))
        return [Q[0]] if solved[0] else []
