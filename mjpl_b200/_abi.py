"""ctypes binding of the C ABI in ``include/mjpl_b200.h`` (``libmjpl_b200.so``).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable,
every compute call raises :class:`EngineUnavailable` -- loudly.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libmjpl_b200.so"

MJB_OK, MJB_ERR_ARG, MJB_ERR_MODEL, MJB_ERR_CUDA = 0, 1, 2, 3
CHECK_LIMITS, CHECK_COLLISION, NO_OBB_CULL, NO_FP64_RECHECK, LIMITS_OUTWARD = 1, 2, 4, 8, 16

_I32P = C.POINTER(C.c_int32)
_F64P = C.POINTER(C.c_double)
_I64P = C.POINTER(C.c_int64)

_INT_FIELDS = "nq nbody njnt ngeom nmesh nmeshvert nexclude nallowed disable_contact disable_filterparent".split()
_I32_ARRAYS_A = "body_parentid body_weldid body_jntadr body_jntnum".split()
_F64_ARRAYS_A = "body_pos body_quat".split()
_I32_ARRAYS_B = "jnt_type jnt_qposadr jnt_bodyid jnt_limited".split()
_F64_ARRAYS_B = "jnt_pos jnt_axis jnt_range qpos0".split()
_I32_ARRAYS_C = "geom_type geom_bodyid geom_contype geom_conaffinity geom_dataid".split()
_F64_ARRAYS_C = "geom_size geom_pos geom_quat geom_margin geom_gap".split()
_I32_ARRAYS_D = "mesh_vertadr mesh_vertnum".split()


class ModelDesc(C.Structure):
    """``mjb_model_desc``: MjModel-named constant tables."""

    _fields_ = (
        [(n, C.c_int32) for n in _INT_FIELDS]
        + [(n, _I32P) for n in _I32_ARRAYS_A]
        + [(n, _F64P) for n in _F64_ARRAYS_A]
        + [(n, _I32P) for n in _I32_ARRAYS_B]
        + [(n, _F64P) for n in _F64_ARRAYS_B]
        + [(n, _I32P) for n in _I32_ARRAYS_C]
        + [(n, _F64P) for n in _F64_ARRAYS_C]
        + [(n, _I32P) for n in _I32_ARRAYS_D]
        + [("mesh_vert", _F64P), ("exclude_signature", _I64P), ("allowed_body_pairs", _I32P)]
        + [("mesh_graphadr", _I32P), ("mesh_graph", _I32P), ("nmeshgraph", C.c_int32)]
    )


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in "rows narrow_items uncertain_rows queue_overflow launches".split()]


class PoseSpec(C.Structure):
    """``mjb_pose_spec``"""

    _fields_ = [("site_bodyid", C.c_int32), ("site_pos", C.c_double * 3), ("site_quat", C.c_double * 4),
                ("ref_pos", C.c_double * 3), ("ref_quat", C.c_double * 4), ("lower", C.c_double * 6),
                ("upper", C.c_double * 6), ("tolerance", C.c_double), ("q_step", C.c_double)]


class IkSpec(C.Structure):
    """``mjb_ik_spec``"""

    _fields_ = [("site_bodyid", C.c_int32), ("site_pos", C.c_double * 3), ("site_quat", C.c_double * 4),
                ("movable_mask", C.c_uint32), ("pos_tolerance", C.c_double), ("ori_tolerance", C.c_double),
                ("lm_damping", C.c_double), ("damping", C.c_double), ("max_step", C.c_double),
                ("iterations", C.c_int32)]


class CbirrtState(C.Structure):
    """``mjb_cbirrt_state``: device pointers of the tick-based constrained planner"""

    _fields_ = [("nslots", C.c_int64), ("cap", C.c_int64), ("nq", C.c_int32), ("check_limits_before", C.c_int32),
                ("eps", C.c_double), ("goal_bias", C.c_double), ("seed", C.c_uint64), ("max_age", C.c_int64),
                ("q_init", C.c_void_p), ("q_goal", C.c_void_p), ("plan_mask", C.c_void_p), ("lo", C.c_void_p), ("hi", C.c_void_p),
                ("nodes", C.c_void_p * 2), ("parent", C.c_void_p * 2), ("count", C.c_void_p * 2),
                ("phase", C.c_void_p), ("swapped", C.c_void_p), ("age", C.c_void_p),
                ("target", C.c_void_p), ("tip", C.c_void_p), ("qa", C.c_void_p), ("last", C.c_void_p), ("ia", C.c_void_p),
                ("cand", C.c_void_p), ("cand32", C.c_void_p), ("proj", C.c_void_p),
                ("proj_ok", C.c_void_p), ("valid", C.c_void_p), ("stepping", C.c_void_p),
                ("res_start", C.c_void_p), ("res_goal", C.c_void_p), ("counters", C.c_void_p)]


class EngineUnavailable(RuntimeError):
    """The CUDA engine cannot run here (library not built, or no GPU).  Never caught internally."""


def make_desc(model, allowed_body_ids):
    """Marshal a model (any object with MjModel field names) into ``mjb_model_desc``.

    Returns ``(desc, keepalive)``; ``keepalive`` owns the numpy buffers the struct points to.
    """
    keep = []

    def arr(x, dtype):
        a = np.ascontiguousarray(np.asarray(x), dtype=dtype).reshape(-1)
        keep.append(a)
        return a

    d = ModelDesc()
    d.nq, d.nbody, d.njnt, d.ngeom, d.nmesh = (
        int(model.nq), int(model.nbody), int(model.njnt), int(model.ngeom), int(model.nmesh))
    mesh_vert = arr(model.mesh_vert, np.float64)
    d.nmeshvert = len(mesh_vert) // 3
    excl = arr(getattr(model, "exclude_signature", np.zeros(0)), np.int64)
    d.nexclude = len(excl)
    allowed = arr(np.asarray(allowed_body_ids, dtype=np.int32).reshape(-1, 2), np.int32)
    d.nallowed = len(allowed) // 2
    flags = int(model.opt.disableflags)
    d.disable_contact = int(bool(flags & 16))        # mjDSBL_CONTACT
    d.disable_filterparent = int(bool(flags & 512))  # mjDSBL_FILTERPARENT
    for n in _I32_ARRAYS_A + _I32_ARRAYS_B + _I32_ARRAYS_C + _I32_ARRAYS_D:
        setattr(d, n, arr(getattr(model, n), np.int32).ctypes.data_as(_I32P))
    for n in _F64_ARRAYS_A + _F64_ARRAYS_B + _F64_ARRAYS_C:
        setattr(d, n, arr(getattr(model, n), np.float64).ctypes.data_as(_F64P))
    d.mesh_vert = mesh_vert.ctypes.data_as(_F64P)
    d.exclude_signature = excl.ctypes.data_as(_I64P)
    d.allowed_body_pairs = allowed.ctypes.data_as(_I32P)
    graph = getattr(model, "mesh_graph", None)
    gadr = getattr(model, "mesh_graphadr", None)
    if graph is not None and gadr is not None and len(graph):
        d.mesh_graphadr = arr(gadr, np.int32).ctypes.data_as(_I32P)
        d.mesh_graph = arr(graph, np.int32).ctypes.data_as(_I32P)
        d.nmeshgraph = len(graph)
    return d, keep


_lib = None

EXPORTS = (
    "mjb_last_error mjb_device_count mjb_model_create mjb_model_destroy mjb_model_npair "
    "mjb_model_pairs mjb_check_configs mjb_check_configs_host mjb_fk mjb_check_edges "
    "mjb_check_sweep mjb_sweep_rows mjb_get_stats mjb_reset_stats mjb_nearest_batch mjb_rrt_extend mjb_pose_valid mjb_pose_project mjb_site_pose "
    "mjb_ik_solve mjb_kernel_timing mjb_fma_peak mjb_rrt_extend_masked mjb_rrt_sample mjb_rrt_meet mjb_cbirrt_tick mjb_min_distance mjb_tree_paths mjb_set_chain_hint"
).split()


def lib():
    """Load ``libmjpl_b200.so`` (built by ``__graft_entry__.build()`` / ``mjpl_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("MJPL_B200_LIB", LIB_PATH))
    if not path.exists():
        raise EngineUnavailable(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  mjpl_b200 has no CPU fallback."
        )
    L = C.CDLL(str(path))
    vp, u8p, f32p = C.c_void_p, C.c_void_p, C.c_void_p
    L.mjb_last_error.restype = C.c_char_p
    L.mjb_device_count.restype = C.c_int
    L.mjb_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
    L.mjb_model_destroy.argtypes = [vp]
    L.mjb_model_destroy.restype = None
    L.mjb_model_npair.argtypes = [vp]
    L.mjb_model_npair.restype = C.c_int32
    L.mjb_model_pairs.argtypes = [vp, _I32P, _I32P]
    L.mjb_check_configs.argtypes = [vp, f32p, C.c_int64, C.c_int32, u8p, C.c_uint32, vp]
    L.mjb_check_configs_host.argtypes = [vp, f32p, C.c_int64, u8p, C.c_uint32]
    L.mjb_fk.argtypes = [vp, f32p, C.c_int64, C.c_int32, f32p, f32p, vp]
    L.mjb_check_edges.argtypes = [vp, f32p, f32p, C.c_int64, C.c_int32, C.c_float, u8p, vp, C.c_uint32, vp]
    L.mjb_check_sweep.argtypes = [vp, C.c_uint64, C.c_int64, C.c_int64, u8p, C.c_uint32, vp]
    L.mjb_sweep_rows.argtypes = [vp, C.c_uint64, C.c_int64, C.c_int64, f32p, vp]
    L.mjb_nearest_batch.argtypes = [vp, C.c_int64, C.c_int32, vp, vp, vp, C.c_int64, vp, vp]
    L.mjb_rrt_extend.argtypes = [vp, vp, vp, vp, C.c_int64, vp, vp, C.c_int64, C.c_double, C.c_int32, C.c_uint32, vp, vp, vp]
    L.mjb_rrt_extend_masked.argtypes = [vp, vp, vp, vp, C.c_int64, vp, vp, vp, C.c_int64, C.c_double, C.c_int32, C.c_uint32, vp, vp, vp]
    L.mjb_rrt_sample.argtypes = [C.c_uint64, vp, C.c_int64, C.c_int32, vp, vp, vp, vp, vp, C.c_double, vp, vp, vp]
    L.mjb_set_chain_hint.argtypes = [vp, C.c_int64]
    L.mjb_tree_paths.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, C.c_int64, vp, vp, vp]
    L.mjb_rrt_meet.argtypes = [C.c_int64, C.c_int32, vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp]
    L.mjb_cbirrt_tick.argtypes = [vp, C.POINTER(CbirrtState), C.POINTER(PoseSpec), C.c_int32, C.c_uint32, vp]
    L.mjb_min_distance.argtypes = [vp, f32p, C.c_int64, C.c_int32, C.c_double, vp, vp, vp]
    L.mjb_site_pose.argtypes = [vp, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), vp, C.c_int64, vp, vp, vp]
    L.mjb_pose_valid.argtypes = [vp, C.POINTER(PoseSpec), vp, C.c_int64, vp, vp]
    L.mjb_pose_project.argtypes = [vp, C.POINTER(PoseSpec), vp, vp, C.c_int64, C.c_int32, vp, vp, vp, vp]
    L.mjb_ik_solve.argtypes = [vp, C.POINTER(IkSpec), vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp]
    L.mjb_kernel_timing.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.mjb_fma_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mjb_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mjb_reset_stats.argtypes = [vp]
    for n in EXPORTS:
        if n not in ("mjb_last_error", "mjb_device_count", "mjb_model_destroy", "mjb_model_npair"):
            getattr(L, n).restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    """Map a C status to the reference's error convention (bad argument -> ValueError)."""
    if rc == MJB_OK:
        return
    msg = lib().mjb_last_error().decode()
    if rc == MJB_ERR_CUDA:
        raise EngineUnavailable(msg)
    raise ValueError(msg)
