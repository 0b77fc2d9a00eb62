"""Index helpers and the rejection sampler on the validity path.

Reference: ``src/mjpl/utils.py`` (``all_joints`` :10-19, ``qpos_idx`` :22-38, ``qvel_idx``
:41-57, ``site_pose`` :60-75, ``random_config`` :78-107).
"""

from __future__ import annotations

import numpy as np

from .constraint.constraint_interface import Constraint
from .constraint.utils import apply_constraints, obeys_constraints_batch
from .model import JNT_DOF_WIDTH, JNT_QPOS_WIDTH


def all_joints(model) -> list[str]:
    """All joint names of the model, in joint-id order."""
    return [model.joint(j).name for j in range(model.njnt)]


def qpos_idx(model, joints: list[str]) -> list[int]:
    """Indices into qpos of the given joints (query order preserved)."""
    idx: list[int] = []
    for name in joints:
        j = model.joint(name).id
        a = int(model.jnt_qposadr[j])
        idx.extend(range(a, a + JNT_QPOS_WIDTH[int(model.jnt_type[j])]))
    return idx


def qvel_idx(model, joints: list[str]) -> list[int]:
    """Indices into qvel of the given joints (query order preserved)."""
    idx: list[int] = []
    for name in joints:
        j = model.joint(name).id
        a = int(model.jnt_dofadr[j])
        idx.extend(range(a, a + JNT_DOF_WIDTH[int(model.jnt_type[j])]))
    return idx


def random_config(model, q_init: np.ndarray, joints: list[str], seed: int | None = None,
                  constraints: list[Constraint] = [], batch: int = 64) -> np.ndarray:
    """Random configuration that obeys ``constraints`` (rejection sampling).

    Same candidate sequence as the reference for a given seed: each candidate consumes one
    ``rng.uniform(*model.jnt_range.T)`` draw over ALL joints, of which only ``joints`` are kept
    (reference :103-105).  Candidates are drawn ``batch`` at a time and validated as one block;
    the first valid one is returned, which is what the reference's loop would return.  With a
    projecting constraint in the list the reference's one-at-a-time loop is used.
    """
    q_idx = qpos_idx(model, joints)
    rng = np.random.default_rng(seed=seed)
    lo, hi = model.jnt_range.T
    if any(getattr(c, "projects", True) for c in constraints):
        q = q_init.copy()
        while True:
            q[q_idx] = rng.uniform(lo, hi)[q_idx]
            qc = apply_constraints(q_init, q, constraints)
            if qc is not None:
                return qc
    while True:
        draws = rng.uniform(lo, hi, size=(batch, len(lo)))  # row i == the i-th sequential draw
        Q = np.tile(np.asarray(q_init, dtype=np.float64), (batch, 1))
        Q[:, q_idx] = draws[:, q_idx]
        ok = np.asarray(obeys_constraints_batch(Q, constraints)) if constraints else np.ones(batch, bool)
        hit = np.flatnonzero(ok)
        if len(hit):
            # NB: a block draw advances the generator past the accepted candidate; the
            # generator is local to this call, so nothing observable depends on that.
            return Q[hit[0]].copy()


def site_pose(model, q: np.ndarray, site_name: str):
    """Pose of a site in the world frame at configuration ``q`` (fp64, ``mjb_site_pose``).

    The reference reads it from an ``MjData`` after ``mj_kinematics`` (``utils.py:60-75``); there is
    no ``MjData`` here, so this takes the configuration itself.
    """
    import ctypes as C

    import torch

    from . import _abi
    from .engine import get_engine
    from .lie import SE3, SO3

    s = model.site(site_name).id
    eng = get_engine(model, ())
    with torch.cuda.device(eng.device):
        qd = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float64).reshape(1, -1)).to(eng.torch_device)
        pos = torch.empty((1, 3), dtype=torch.float64, device=eng.torch_device)
        quat = torch.empty((1, 4), dtype=torch.float64, device=eng.torch_device)
        sp = (C.c_double * 3)(*[float(x) for x in model.site_pos[s]])
        sq = (C.c_double * 4)(*[float(x) for x in model.site_quat[s]])
        _abi.check(eng._L.mjb_site_pose(eng._h, int(model.site_bodyid[s]), sp, sq, qd.data_ptr(), 1, pos.data_ptr(),
                                        quat.data_ptr(), eng._stream()))
        return SE3(SO3(quat[0].cpu().numpy()), pos[0].cpu().numpy())
