"""TEST TOOL: CPU execution of the CUDA kernels' templated core (see hostsim.cpp).

Built on demand with g++; used only by the ``-m "not gpu"`` tests to compare the geometric
core (FK, culls, GJK classifier, fp64 re-evaluation) with the fp64 oracle where no GPU exists.
Not a product path.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from mjpl_b200 import _abi

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libhostsim.so"
_ROOT = _HERE.parent.parent


def build(force=False):
    srcs = [_HERE / "hostsim.cpp", _ROOT / "mjpl_b200/csrc/vk_core.cuh", _ROOT / "mjpl_b200/csrc/vk_build.h"]
    newest = max(p.stat().st_mtime for p in srcs)
    if force or not _SO.exists() or _SO.stat().st_mtime < newest:
        subprocess.run(
            ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DVK_HILL=1",   # hull graphs stay testable on the host
             *os.environ.get("HOSTSIM_CXXFLAGS", "").split(), "-x", "c++",   # e.g. -DVK_OBB_EDGE_AXES=1 for experiments
             str(_HERE / "hostsim.cpp"), "-o", str(_SO)],
            check=True, capture_output=True, text=True,
        )
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        L.hs_create.argtypes = [C.POINTER(_abi.ModelDesc), C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        L.hs_destroy.argtypes = [C.c_void_p]
        L.hs_npair.argtypes = [C.c_void_p]
        L.hs_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_fk.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.hs_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.hs_support_check.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_pose.argtypes = [C.c_void_p, C.POINTER(_abi.PoseSpec), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.hs_ik.argtypes = [C.c_void_p, C.POINTER(_abi.IkSpec), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.hs_pair_census.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.hs_bins.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_pair_verdict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.hs_check_pipe.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p]
        L.hs_min_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
        L.hs_smap_check.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.hs_bounds_check.argtypes = [C.c_void_p]
        L.hs_bounds_check.restype = C.c_double
        L.hs_inner_check.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
        L.hs_inner_check.restype = C.c_double
        L.hs_l0_sq_err.argtypes = [C.c_void_p]
        L.hs_l0_sq_err.restype = C.c_double
        L.hs_point_segment_check.argtypes = [C.c_int, C.c_uint64]
        L.hs_point_segment_check.restype = C.c_double
        L.hs_segseg_check.argtypes = [C.c_int, C.c_uint64]
        L.hs_segseg_check.restype = C.c_double
        _lib = L
    return _lib


class HostSim:
    def __init__(self, model, allowed_collision_bodies=()):
        allowed = [(model.body(a).id, model.body(b).id) for a, b in allowed_collision_bodies]
        desc, keep = _abi.make_desc(model, allowed)
        h = C.c_void_p()
        err = C.create_string_buffer(256)
        if lib().hs_create(C.byref(desc), C.byref(h), err, 256) != 0:
            raise ValueError(err.value.decode())
        self._h = h
        self.model = model

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hs_destroy(self._h)
            self._h = None

    def pairs(self):
        n = lib().hs_npair(self._h)
        g1, g2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        lib().hs_pairs(self._h, g1.ctypes.data, g2.ctypes.data)
        return np.stack([g1, g2], 1)

    def fk(self, q):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.model.nq)
        n = len(q)
        xpos = np.zeros((n, self.model.nbody, 3), np.float32)
        xquat = np.zeros((n, self.model.nbody, 4), np.float32)
        lib().hs_fk(self._h, q.ctypes.data, n, xpos.ctypes.data, xquat.ctypes.data)
        return xpos, xquat

    def check(self, q, flags=3, obb=True, recheck=True):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.model.nq)
        n = len(q)
        valid = np.zeros(n, np.uint8)
        stats = np.zeros(48, np.int64)
        lib().hs_check(self._h, q.ctypes.data, n, flags, int(obb), int(recheck), valid.ctypes.data, stats.ctypes.data)
        names = "items gjk_iters uncertain_rows rows sphere_survivors gjk_calls max_gjk_iters uncertain_items".split()
        d = dict(zip(names, stats[:8].tolist()))
        d['gjk_iter_hist'] = stats[8:40].tolist()
        d['gjk_verdicts_sep_pen_unc'] = stats[40:43].tolist()
        d['gjk_iters_by_verdict'] = stats[44:47].tolist()
        d['vertex_evals'] = int(stats[47])
        return valid, d

    def check_pipe(self, q, flags=3):
        """validity through the pipeline's culling order (group level -> capsule -> OBB -> narrow), no
        early exit -> (valid, {level0, expanded, capsule, items, contacts})"""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.model.nq)
        valid = np.zeros(len(q), np.uint8)
        stats = np.zeros(48, np.int64)
        lib().hs_check_pipe(self._h, q.ctypes.data, len(q), flags, valid.ctypes.data, stats.ctypes.data)
        d = dict(zip("level0 expanded capsule items contacts closed_form_items".split(), stats[:6].tolist()))
        d["gjk_iter_hist"] = stats[8:24].tolist()
        d["gjk_verdicts_sep_pen_unc"] = stats[24:27].tolist()
        d["gjk_one_iteration_sep_pen_unc"] = stats[27:30].tolist()
        d["fp64_iters_of_uncertain_le_4_8_12_16_24_32_48_more"] = stats[40:48].tolist()
        d["inner"] = dict(zip("invalid_rows caught_level0 caught_any level0_of_caught0 expanded_of_caught0 items_of_caught0 items_of_caught false_positives".split(),
                              stats[32:40].tolist()))
        return valid, d

    def min_distance(self, q, far_cap=0.01):
        """signed distance to contact per row through the fp64 core -> (dist, pair index)"""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.model.nq)
        dist = np.zeros(len(q)); pair = np.zeros(len(q), np.int32)
        lib().hs_min_distance(self._h, q.ctypes.data, len(q), far_cap, dist.ctypes.data, pair.ctypes.data)
        return dist, pair

    def smap_check(self, ndir=4000, seed=2):
        """support maps vs full scans -> (mapped shapes, largest shortfall of the mapped support value, mean candidates per query)"""
        worst, avg = C.c_double(0), C.c_double(0)
        n = lib().hs_smap_check(self._h, ndir, seed, C.byref(worst), C.byref(avg))
        return n, worst.value, avg.value

    def bounds_check(self):
        """largest distance by which a vertex sticks out of its bounding capsule / group sphere (<= 0: contained)"""
        return lib().hs_bounds_check(self._h)

    def l0_sq_err(self):
        """what the squared limits of level 0 carry for fp32 rounding (m^2)"""
        return lib().hs_l0_sq_err(self._h)

    def inner_check(self, nsample=400, seed=5):
        """inner capsules against their shapes -> (shapes with an inner capsule, largest distance by which a
        sampled surface point of an inner capsule sticks out of its shape; <= 0: none does)"""
        n = C.c_int(0)
        worst = lib().hs_inner_check(self._h, nsample, seed, C.byref(n))
        return n.value, worst

    def pair_verdict(self, q, g1, g2, fp64=False):
        """(verdict, iterations) of one geom pair: 0 separated, 1 contact, 2 uncertain."""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(self.model.nq)
        it = C.c_int(0)
        v = lib().hs_pair_verdict(self._h, q.ctypes.data, int(g1), int(g2), int(fp64), C.byref(it))
        return v, it.value

    def support_check(self, ndir=2000, seed=1):
        """hill-climbing support vs full scan -> (shapes with a graph, max value shortfall, evals hill, evals scan)"""
        gap = C.c_double(0)
        eh, es = C.c_int64(0), C.c_int64(0)
        n = lib().hs_support_check(self._h, ndir, seed, C.byref(gap), C.byref(eh), C.byref(es))
        return n, gap.value, eh.value, es.value

    def pose(self, spec, q_old, q, project=True, max_iters=1000):
        """PoseConstraint rows through the kernel core on the CPU -> (q_out, ok, iters)"""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.model.nq)
        q_old = np.ascontiguousarray(q_old, dtype=np.float64).reshape(-1, self.model.nq)
        n = len(q)
        out = q.copy()
        ok = np.zeros(n, np.uint8)
        iters = np.zeros(n, np.int32)
        err = C.create_string_buffer(256)
        rc = lib().hs_pose(self._h, C.byref(spec), q_old.ctypes.data, q.ctypes.data, n, int(project), max_iters,
                           out.ctypes.data, ok.ctypes.data, iters.ctypes.data, err, 256)
        if rc:
            raise ValueError(err.value.decode())
        return out, ok.astype(bool), iters

    def ik(self, spec, target_pos, target_quat, q_init):
        """IK rows through the kernel core on the CPU -> (q_out, ok, iters, errs)"""
        q = np.ascontiguousarray(q_init, dtype=np.float64).reshape(-1, self.model.nq)
        tp = np.ascontiguousarray(target_pos, dtype=np.float64).reshape(-1, 3)
        tq = np.ascontiguousarray(target_quat, dtype=np.float64).reshape(-1, 4)
        n = len(q)
        out = q.copy()
        ok = np.zeros(n, np.uint8)
        iters = np.zeros(n, np.int32)
        errs = np.zeros((n, 2))
        err = C.create_string_buffer(256)
        rc = lib().hs_ik(self._h, C.byref(spec), tp.ctypes.data, tq.ctypes.data, q.ctypes.data, n, out.ctypes.data,
                         ok.ctypes.data, iters.ctypes.data, errs.ctypes.data, err, 256)
        if rc:
            raise ValueError(err.value.decode())
        return out, ok.astype(bool), iters, errs

    def pair_census(self, q):
        """per pair (in the engine's pair order, see ``pairs()``): sphere survivors, narrow items,
        vertex evaluations, contacts over the rows of ``q`` (no early exit)"""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.model.nq)
        out = np.zeros((lib().hs_npair(self._h), 4), np.int64)
        lib().hs_pair_census(self._h, q.ctypes.data, len(q), out.ctypes.data)
        return out

    def bins(self):
        """(bin_expect[8], calibrated items per row, bin of every pair)"""
        be = np.zeros(8)
        tot = C.c_double(0)
        pb = np.zeros(lib().hs_npair(self._h), np.int32)
        lib().hs_bins(self._h, be.ctypes.data, C.byref(tot), pb.ctypes.data)
        return be, tot.value, pb


def point_segment_check(ncase=200000, seed=4):
    """largest error of level 0's expanded squared point-segment distance (fp32) against fp64, relative to |e|_max^2"""
    return lib().hs_point_segment_check(ncase, seed)


def segseg_check(ncase=20000, seed=3):
    """largest overestimate of the fp32 segment-segment distance against an fp64 search"""
    return lib().hs_segseg_check(ncase, seed)
