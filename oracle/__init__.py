"""ctypes front-end of the fp64 CPU oracle (``oracle.c``) -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; nothing under ``mjpl_b200/`` does.

"parity unpinned" against a real MuJoCo (see ``oracle.h``): MuJoCo cannot be installed in
this image, so the oracle is pinned by the reference's own known-answer tests and by
closed-form cases only.

Besides the C restatement of ``mj_kinematics`` / ``mj_collision`` + ``CollisionRuleset``
this module restates, in plain numpy, the host-side reference algorithms that sit on the
validity path (``_step``, ``_valid_collision_interval``, ``_constrained_extend``), each citing
the reference lines it follows, so the batched product code can be compared with the
sequential original.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"

CHECK_LIMITS = 1
CHECK_COLLISION = 2
DEPTH_CAP = 1e-3

_I32P = C.POINTER(C.c_int32)
_F64P = C.POINTER(C.c_double)
_I64P = C.POINTER(C.c_int64)


class _Desc(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in
         "nq nbody njnt ngeom nmesh nmeshvert nexclude nallowed disable_contact disable_filterparent".split()]
        + [(n, _I32P) for n in "body_parentid body_weldid body_jntadr body_jntnum".split()]
        + [(n, _F64P) for n in "body_pos body_quat".split()]
        + [(n, _I32P) for n in "jnt_type jnt_qposadr jnt_bodyid jnt_limited".split()]
        + [(n, _F64P) for n in "jnt_pos jnt_axis jnt_range qpos0".split()]
        + [(n, _I32P) for n in "geom_type geom_bodyid geom_contype geom_conaffinity geom_dataid".split()]
        + [(n, _F64P) for n in "geom_size geom_pos geom_quat geom_margin geom_gap".split()]
        + [(n, _I32P) for n in "mesh_vertadr mesh_vertnum".split()]
        + [("mesh_vert", _F64P), ("exclude_signature", _I64P), ("allowed_body_pairs", _I32P)]
    )


def build(force: bool = False) -> Path:
    """Compile ``liboracle.so`` with the committed Makefile (gcc only)."""
    src_m = max((_HERE / f).stat().st_mtime for f in ("oracle.c", "oracle.h", "Makefile"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.orc_model_create.argtypes = [C.POINTER(_Desc), C.POINTER(C.c_void_p)]
        L.orc_model_create.restype = C.c_int
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_npair.argtypes = [C.c_void_p]
        L.orc_npair.restype = C.c_int32
        L.orc_pairs.argtypes = [C.c_void_p, _I32P, _I32P]
        L.orc_fk.argtypes = [C.c_void_p, _F64P, C.c_int64, _F64P, _F64P]
        L.orc_geom_poses.argtypes = [C.c_void_p, _F64P, C.c_int64, _F64P, _F64P]
        L.orc_check.argtypes = [C.c_void_p, _F64P, C.c_int64, C.c_uint32,
                                C.POINTER(C.c_uint8), _F64P, _I32P]
        L.orc_check.restype = C.c_int
        L.orc_pair_distance.argtypes = [C.c_void_p, _F64P, C.c_int32, _F64P]
        L.orc_pair_distance.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_get_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


class Oracle:
    """fp64 reference checker for one model + one ``CollisionRuleset`` allow-list.

    ``model`` is duck-typed on MuJoCo's ``MjModel`` field names; ``allowed_collision_bodies``
    has the meaning of the reference's ``CollisionConstraint`` argument
    (``collision_constraint.py:10-24``).
    """

    def __init__(self, model, allowed_collision_bodies=()):
        self.model = model
        L = lib()
        keep = []

        def i32(x):
            a = np.ascontiguousarray(np.asarray(x), dtype=np.int32).reshape(-1)
            keep.append(a)
            return _p(a, _I32P)

        def f64(x):
            a = np.ascontiguousarray(np.asarray(x), dtype=np.float64).reshape(-1)
            keep.append(a)
            return _p(a, _F64P)

        allowed = [(model.body(a).id, model.body(b).id) for a, b in allowed_collision_bodies]
        excl = np.ascontiguousarray(np.asarray(getattr(model, "exclude_signature", [])), dtype=np.int64)
        keep.append(excl)
        d = _Desc()
        d.nq, d.nbody, d.njnt, d.ngeom, d.nmesh = (
            int(model.nq), int(model.nbody), int(model.njnt), int(model.ngeom), int(model.nmesh))
        d.nmeshvert = int(len(model.mesh_vert))
        d.nexclude = int(len(excl))
        d.nallowed = len(allowed)
        flags = int(model.opt.disableflags)
        d.disable_contact = int(bool(flags & 16))
        d.disable_filterparent = int(bool(flags & 512))
        for n in "body_parentid body_weldid body_jntadr body_jntnum jnt_type jnt_qposadr jnt_bodyid jnt_limited geom_type geom_bodyid geom_contype geom_conaffinity geom_dataid mesh_vertadr mesh_vertnum".split():
            setattr(d, n, i32(getattr(model, n)))
        for n in "body_pos body_quat jnt_pos jnt_axis jnt_range qpos0 geom_size geom_pos geom_quat geom_margin geom_gap mesh_vert".split():
            setattr(d, n, f64(getattr(model, n)))
        d.exclude_signature = _p(excl, _I64P)
        d.allowed_body_pairs = i32(np.array(allowed, dtype=np.int32).reshape(-1, 2))
        h = C.c_void_p()
        if L.orc_model_create(C.byref(d), C.byref(h)) != 0:
            raise ValueError(L.orc_last_error().decode())
        self._h = h
        self._keep = keep
        self.nq, self.nbody, self.ngeom = d.nq, d.nbody, d.ngeom

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().orc_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def set_threads(n: int) -> None:
        lib().orc_set_threads(int(n))

    def _q(self, q):
        q = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
        if q.ndim == 1:
            q = q[None]
        if q.ndim != 2 or q.shape[1] != self.nq:
            raise ValueError(f"q must be (n,{self.nq})")
        return q

    def pairs(self) -> np.ndarray:
        n = lib().orc_npair(self._h)
        g1, g2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        lib().orc_pairs(self._h, _p(g1, _I32P), _p(g2, _I32P))
        return np.stack([g1, g2], axis=1)

    def fk(self, q):
        q = self._q(q)
        n = len(q)
        xpos = np.zeros((n, self.nbody, 3))
        xquat = np.zeros((n, self.nbody, 4))
        lib().orc_fk(self._h, _p(q, _F64P), n, _p(xpos, _F64P), _p(xquat, _F64P))
        return xpos, xquat

    def geom_poses(self, q):
        q = self._q(q)
        n = len(q)
        gp = np.zeros((n, self.ngeom, 3))
        gm = np.zeros((n, self.ngeom, 9))
        lib().orc_geom_poses(self._h, _p(q, _F64P), n, _p(gp, _F64P), _p(gm, _F64P))
        return gp, gm.reshape(n, self.ngeom, 3, 3)

    def check(self, q, flags=CHECK_COLLISION, want_dist=False):
        """-> valid (n,) bool [, min signed distance (n,), arg-min pair (n,)]"""
        q = self._q(q)
        n = len(q)
        valid = np.zeros(n, np.uint8)
        if want_dist:
            dist = np.zeros(n)
            pair = np.zeros(n, np.int32)
            rc = lib().orc_check(self._h, _p(q, _F64P), n, flags, _p(valid, C.POINTER(C.c_uint8)),
                                 _p(dist, _F64P), _p(pair, _I32P))
        else:
            rc = lib().orc_check(self._h, _p(q, _F64P), n, flags, _p(valid, C.POINTER(C.c_uint8)),
                                 None, None)
        if rc != 0:
            raise ValueError(lib().orc_last_error().decode())
        return (valid.astype(bool), dist, pair) if want_dist else valid.astype(bool)

    def pair_distance(self, q, pair: int) -> float:
        q = self._q(q)
        out = C.c_double()
        if lib().orc_pair_distance(self._h, _p(q, _F64P), pair, C.byref(out)) != 0:
            raise ValueError(lib().orc_last_error().decode())
        return out.value

    # ---- reference Constraint-style scalar calls -----------------------------------------
    def collision_valid_config(self, q) -> bool:
        """CollisionConstraint.valid_config (collision_constraint.py:26-30)."""
        return bool(self.check(q, CHECK_COLLISION)[0])

    def limits_valid_config(self, q) -> bool:
        """JointLimitConstraint.valid_config (joint_limit_constraint.py:19-20)."""
        return bool(self.check(q, CHECK_LIMITS)[0])


# ------------------------------------------------------------------------------------------
# numpy restatements of the host-side reference algorithms on the path
# ------------------------------------------------------------------------------------------
def step(start: np.ndarray, target: np.ndarray, max_step_dist: float) -> np.ndarray:
    """``_step`` (reference: src/mjpl/planning/utils.py:167-185)."""
    if max_step_dist <= 0.0:
        raise ValueError("`max_step_dist` must be > 0.0")
    if np.array_equal(start, target):
        return start.copy()
    direction = target - start
    magnitude = np.linalg.norm(direction)
    unit_vec = direction / magnitude
    return start + (unit_vec * min(max_step_dist, magnitude))


def interval_waypoints(start, end, step_dist) -> list[np.ndarray]:
    """Interior waypoints of ``_valid_collision_interval`` (planning/utils.py:206-214)."""
    if step_dist <= 0.0:
        raise ValueError("`step_dist` must be > 0")
    wps = [start]
    while not np.array_equal(wps[-1], end):
        wps.append(step(wps[-1], end, step_dist))
    return wps[1:-1]


def valid_collision_interval(oracle: Oracle, start, end, step_dist):
    """``_valid_collision_interval`` (planning/utils.py:188-216) -> (ok, first_bad_index)."""
    wps = interval_waypoints(np.asarray(start, float), np.asarray(end, float), step_dist)
    if not wps:
        return True, -1
    v = oracle.check(np.array(wps), CHECK_COLLISION)
    bad = np.flatnonzero(~v)
    return (len(bad) == 0), (int(bad[0]) if len(bad) else -1)


def constrained_extend_chain(oracle: Oracle, q_near, q_target, eps, flags=CHECK_LIMITS | CHECK_COLLISION,
                             interval=None, equality_threshold=1e-8, max_steps=100000):
    """``_constrained_extend`` for non-projecting constraints (planning/utils.py:105-164).

    Returns the list of configurations the reference would add to the tree (in order) and
    the configuration it would return.
    """
    q = np.asarray(q_near, float)
    q_old = q
    q_target = np.asarray(q_target, float)
    added = []
    for _ in range(max_steps):
        if np.array_equal(q_target, q):
            return added, q
        q = step(q, q_target, eps)
        ok = bool(oracle.check(q, flags)[0])
        if (
            not ok
            or np.linalg.norm(q - q_old) < equality_threshold
            or np.linalg.norm(q_target - q) > np.linalg.norm(q_target - q_old)
            or (interval is not None and not valid_collision_interval(oracle, q_old, q, interval)[0])
        ):
            return added, q_old
        added.append(q)
        q_old = q
    return added, q_old


# ------------------------------------------------------------------------------------------
# PoseConstraint restated (reference: src/mjpl/constraint/pose_constraint.py)
# ------------------------------------------------------------------------------------------
def _q2mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def _rpy(q):
    """mink / jaxlie SO3.as_rpy_radians for a wxyz quaternion."""
    w, x, y, z = q
    return np.array([np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y)),
                     np.arcsin(np.clip(2 * (w * y - z * x), -1, 1)),
                     np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))])


class PoseOracle:
    """fp64 numpy restatement of ``PoseConstraint`` on top of the C oracle's ``mj_kinematics``.

    ``ref_pos`` / ``ref_quat`` (wxyz) are the reference frame ``world_T_C``; ``C`` is the (6,2) box
    of allowed x, y, z, roll, pitch, yaw.
    """

    def __init__(self, model, site, ref_pos, ref_quat, C, tolerance=0.001, q_step=0.05):
        self.model, self.orc = model, Oracle(model)
        self.sid = model.site(site).id
        self.C = np.asarray(C, dtype=np.float64)
        rq = np.asarray(ref_quat, float) / np.linalg.norm(ref_quat)
        self.cw_quat = rq * np.array([1, -1, -1, -1.0])
        self.cw_pos = -_q2mat(self.cw_quat) @ np.asarray(ref_pos, float)
        self.tolerance, self.q_step = tolerance, q_step

    def site_pose(self, q):
        xpos, xquat = self.orc.fk(q)
        b = int(self.model.site_bodyid[self.sid])
        p = xpos[0, b] + _q2mat(xquat[0, b]) @ self.model.site_pos[self.sid]
        r = _qmul(xquat[0, b], self.model.site_quat[self.sid])
        return p, r / np.linalg.norm(r)

    def displacement(self, q):
        """``_displacement_from_constraint`` (pose_constraint.py:93-123)."""
        p, r = self.site_pose(q)
        d = np.concatenate([self.cw_pos + _q2mat(self.cw_quat) @ p, _rpy(_qmul(self.cw_quat, r))])
        over, under = d > self.C[:, 1], d < self.C[:, 0]
        dx = np.zeros(6)
        dx[over] = d[over] - self.C[over, 1]
        dx[under] = d[under] - self.C[under, 0]
        return dx

    def jacobian(self, q):
        """``_get_jacobian`` + ``_e_rpy`` (pose_constraint.py:125-171): mj_jacSite restated
        (hinge: jacr = axis, jacp = axis x (site - anchor); slide: jacp = axis)."""
        m = self.model
        xpos, xquat = self.orc.fk(q)
        p, r = self.site_pose(q)
        J = np.zeros((6, m.nv))
        b = int(m.site_bodyid[self.sid])
        while b > 0:
            for k in range(int(m.body_jntnum[b])):
                j = int(m.body_jntadr[b]) + k
                # the joint's own motion does not move its anchor / axis: the body pose after the
                # joint transform gives the same world axis, and the anchor from jnt_pos
                R = _q2mat(xquat[0, b])
                axis = R @ m.jnt_axis[j]
                anchor = xpos[0, b] + R @ m.jnt_pos[j]
                c = int(m.jnt_dofadr[j])
                if int(m.jnt_type[j]) == 2:
                    J[:3, c] = axis
                else:
                    J[:3, c] = np.cross(axis, p - anchor)
                    J[3:, c] = axis
            b = int(m.body_parentid[b])
        _, pitch, yaw = _rpy(r)
        cp, sp, cy, sy = np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
        E = np.eye(6)
        E[3:6, 3:5] = np.array([[cy / cp, sy / cp], [-sy, cp], [cy * (sp / cp), sy * (sp / cp)]])
        return E @ J

    def limits_ok(self, q):
        return bool(np.all((q >= self.model.jnt_range[:, 0]) & (q <= self.model.jnt_range[:, 1])))

    def valid_config(self, q):
        q = np.asarray(q, float)
        return self.limits_ok(q) and np.linalg.norm(self.displacement(q)) <= self.tolerance

    def apply(self, q_old, q, max_iters=1000):
        """``apply`` (pose_constraint.py:78-91) -> projected q or None."""
        q_old, qp = np.asarray(q_old, float), np.asarray(q, float).copy()
        for _ in range(max_iters):
            dx = self.displacement(qp)
            if np.linalg.norm(dx) <= self.tolerance:
                return qp
            J = self.jacobian(qp)
            qp = qp - J.T @ np.linalg.pinv(J @ J.T) @ dx
            if not self.limits_ok(qp) or np.linalg.norm(qp - q_old) > 2 * self.q_step:
                return None
        return None
