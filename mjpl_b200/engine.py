"""Host-side handle of the B200 validity engine: one ``mjb_model`` per (model, allow-list, GPU).

PyTorch is used here only for device memory and streams; all arithmetic happens in the CUDA
kernels behind ``include/mjpl_b200.h``.  There is no CPU fallback.
"""

from __future__ import annotations

import atexit
import ctypes as C
import threading

import numpy as np

from . import _abi
from ._abi import CHECK_COLLISION, CHECK_LIMITS, EngineUnavailable  # noqa: F401

_lock = threading.Lock()
_cache: dict = {}


def _torch():
    import torch

    return torch


def allowed_body_ids(model, allowed_collision_bodies) -> list[tuple[int, int]]:
    """Body-name pairs -> id pairs; unknown names raise ``KeyError`` like ``model.body(name)``
    does in the reference (``collision_constraint.py:59-63``)."""
    return [(model.body(a).id, model.body(b).id) for a, b in allowed_collision_bodies]


class ValidityEngine:
    """Device tables + kernels for one model and one ``CollisionRuleset`` allow-list."""

    def __init__(self, model, allowed_collision_bodies=(), device: int | None = None):
        torch = _torch()
        L = _abi.lib()
        if not torch.cuda.is_available() or L.mjb_device_count() < 1:
            raise EngineUnavailable(
                "no CUDA device: mjpl_b200 runs its validity checks on the GPU only (no CPU fallback)"
            )
        self.model = model
        self.nq = int(model.nq)
        self.nbody = int(model.nbody)
        self.device = torch.cuda.current_device() if device is None else int(device)
        ids = allowed_body_ids(model, allowed_collision_bodies)
        desc, keep = _abi.make_desc(model, ids)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _abi.check(L.mjb_model_create(C.byref(desc), C.byref(h)))
        self._h = h
        self._L = L
        self._host_mask = None   # pinned result buffer of the host entry point (grown on demand)
        # a handle is not re-entrant (its scratch is per handle): host threads are serialised here,
        # streams are ordered inside the library (csrc/mjpl_b200.cu: enter_stream / leave_stream)
        self._call_lock = threading.RLock()
        self._lo = np.ascontiguousarray(model.jnt_range[:, 0], dtype=np.float64)
        self._hi = np.ascontiguousarray(model.jnt_range[:, 1], dtype=np.float64)
        self._lim_dev = None
        del keep

    def close(self):
        if getattr(self, "_h", None):
            self._L.mjb_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    @property
    def torch_device(self):
        return _torch().device("cuda", self.device)

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _rows(self, Q):
        """-> (float32 contiguous CUDA tensor (n,nq), kind) where kind tells how to hand results back."""
        torch = _torch()
        if isinstance(Q, np.ndarray):
            if Q.ndim != 2 or Q.shape[1] != self.nq:
                raise ValueError(f"expected an (n, {self.nq}) array of configurations")
            t = torch.from_numpy(np.ascontiguousarray(Q, dtype=np.float32))
            return t.to(self.torch_device, non_blocking=False), "numpy"
        if not torch.is_tensor(Q):
            raise TypeError("configurations must be a numpy array or a torch tensor")
        if Q.ndim != 2 or Q.shape[1] != self.nq:
            raise ValueError(f"expected an (n, {self.nq}) tensor of configurations")
        kind = "cuda" if Q.is_cuda else "cpu"
        t = Q.to(device=self.torch_device, dtype=torch.float32, non_blocking=True).contiguous()
        return t, kind

    @staticmethod
    def _back(t, kind):
        if kind == "numpy":
            return t.cpu().numpy()
        if kind == "cpu":
            return t.cpu()
        return t

    def pairs(self) -> np.ndarray:
        n = self._L.mjb_model_npair(self._h)
        g1, g2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        _abi.check(self._L.mjb_model_pairs(self._h, g1.ctypes.data_as(_abi._I32P), g2.ctypes.data_as(_abi._I32P)))
        return np.stack([g1, g2], axis=1)

    # ------------------------------------------------------------------ entry points
    def _exact_limits(self, Q):
        """``JointLimitConstraint.valid_config`` for rows that are wider than fp32, in the rows' own
        precision and container (reference ``joint_limit_constraint.py:19-20``: a closed-interval fp64
        compare).  The device kernels see the fp32 rounding of such rows, so they are asked for
        outward-rounded limits (``MJB_LIMITS_OUTWARD``) and this mask settles the rows within one
        fp32 ulp of a limit.  Returns None for rows that are fp32 (or narrower) already."""
        torch = _torch()
        if isinstance(Q, np.ndarray):
            if Q.dtype != np.float64 or Q.ndim != 2 or Q.shape[1] != self.nq:
                return None
            return ((Q >= self._lo) & (Q <= self._hi)).all(axis=1)
        if torch.is_tensor(Q) and Q.dtype == torch.float64 and Q.ndim == 2 and Q.shape[1] == self.nq:
            if Q.is_cuda:
                if self._lim_dev is None or self._lim_dev[0].device != Q.device:
                    self._lim_dev = (torch.from_numpy(self._lo).to(Q.device), torch.from_numpy(self._hi).to(Q.device))
                lo, hi = self._lim_dev
            else:
                lo, hi = torch.from_numpy(self._lo), torch.from_numpy(self._hi)
            return ((Q >= lo) & (Q <= hi)).all(dim=1)
        return None

    def valid_configs(self, Q, flags: int = CHECK_LIMITS | CHECK_COLLISION):
        """(n,nq) -> (n,) bool, same container kind as the input (numpy / CPU tensor / CUDA tensor).

        Host containers go through ``mjb_check_configs_host``: the rows are copied in chunks while the
        (single) validity launch is already consuming them, so copy and compute overlap.
        Joint limits of fp64 rows are decided on the fp64 values (see ``_exact_limits``).
        """
        torch = _torch()
        exact = self._exact_limits(Q) if (flags & CHECK_LIMITS) and not (flags & _abi.NO_FP64_RECHECK) else None
        if exact is not None:
            flags |= _abi.LIMITS_OUTWARD
        with self._call_lock:
            if isinstance(Q, np.ndarray) or (torch.is_tensor(Q) and not Q.is_cuda):
                res = self._valid_host(Q, flags)
            else:
                with torch.cuda.device(self.device):
                    q, kind = self._rows(Q)
                    n = q.shape[0]
                    out = torch.empty(n, dtype=torch.uint8, device=self.torch_device)
                    if n:
                        _abi.check(self._L.mjb_check_configs(self._h, q.data_ptr(), n, q.stride(0), out.data_ptr(), flags, self._stream()))
                    res = self._back(out.bool() if not (flags & _abi.NO_FP64_RECHECK) else out, kind)
        return res if exact is None else (res & exact)

    def _valid_host(self, Q, flags: int):
        torch = _torch()
        is_np = isinstance(Q, np.ndarray)
        if Q.ndim != 2 or Q.shape[1] != self.nq:
            raise ValueError(f"expected an (n, {self.nq}) {'array' if is_np else 'tensor'} of configurations")
        if is_np:
            rows = np.ascontiguousarray(Q, dtype=np.float32)
            ptr = rows.ctypes.data
        else:
            rows = Q.to(dtype=torch.float32).contiguous()
            ptr = rows.data_ptr()
        n = int(rows.shape[0])
        if self._host_mask is None or self._host_mask.numel() < n:
            self._host_mask = torch.empty(max(n, 4096), dtype=torch.uint8, pin_memory=True)
        out = self._host_mask[:n]
        if n:
            with torch.cuda.device(self.device):
                _abi.check(self._L.mjb_check_configs_host(self._h, ptr, n, out.data_ptr(), flags))
        res = out.clone() if (flags & _abi.NO_FP64_RECHECK) else out.bool()
        return res.numpy() if is_np else res

    def valid_configs_host(self, Q: np.ndarray, flags: int = CHECK_LIMITS | CHECK_COLLISION) -> np.ndarray:
        """Host buffers straight through ``mjb_check_configs_host`` (copies inside the call)."""
        return np.asarray(self._valid_host(np.asarray(Q), flags)).astype(bool)

    def fk(self, Q):
        """mj_kinematics for a block: -> xpos (n,nbody,3), xquat (n,nbody,4), fp32."""
        torch = _torch()
        with torch.cuda.device(self.device):
            q, kind = self._rows(Q)
            n = q.shape[0]
            xpos = torch.empty((n, self.nbody, 3), dtype=torch.float32, device=self.torch_device)
            xquat = torch.empty((n, self.nbody, 4), dtype=torch.float32, device=self.torch_device)
            if n:
                _abi.check(self._L.mjb_fk(self._h, q.data_ptr(), n, q.stride(0), xpos.data_ptr(), xquat.data_ptr(), self._stream()))
            return self._back(xpos, kind), self._back(xquat, kind)

    def min_distance(self, Q, far_cap: float = 0.0):
        """Signed distance to contact of every row (``mjb_min_distance``, fp64) -> ``(dist (n,) float64,
        pair (n,) int32)``, same container kind as the input.  ``dist <= 0`` iff the row is in contact;
        values are capped at ``far_cap`` (default 0.01 m, pair -1) and at -0.001 m."""
        torch = _torch()
        with self._call_lock, torch.cuda.device(self.device):
            q, kind = self._rows(Q)
            n = q.shape[0]
            dist = torch.empty(n, dtype=torch.float64, device=self.torch_device)
            pair = torch.empty(n, dtype=torch.int32, device=self.torch_device)
            if n:
                _abi.check(self._L.mjb_min_distance(self._h, q.data_ptr(), n, q.stride(0), float(far_cap), dist.data_ptr(),
                                                    pair.data_ptr(), self._stream()))
            return self._back(dist, kind), self._back(pair, kind)

    def valid_edges(self, Q0, Q1, step: float, flags: int = CHECK_COLLISION, want_first_bad: bool = False):
        """``_valid_collision_interval`` for many edges: interior waypoints only, early exit."""
        # reference: raise ValueError("`step_dist` must be > 0") (planning/utils.py:203-204)
        if not step > 0.0:
            raise ValueError("`step_dist` must be > 0")
        torch = _torch()
        with self._call_lock, torch.cuda.device(self.device):
            q0, kind = self._rows(Q0)
            q1, _ = self._rows(Q1)
            if q0.shape != q1.shape:
                raise ValueError("Q0 and Q1 must have the same shape")
            ne = q0.shape[0]
            valid = torch.empty(ne, dtype=torch.uint8, device=self.torch_device)
            fb = torch.empty(ne, dtype=torch.int32, device=self.torch_device)
            if ne:
                _abi.check(self._L.mjb_check_edges(self._h, q0.data_ptr(), q1.data_ptr(), ne, q0.stride(0), float(step),
                                                   valid.data_ptr(), fb.data_ptr(), flags, self._stream()))
            v = self._back(valid.bool(), kind)
            return (v, self._back(fb, kind)) if want_first_bad else v

    def sweep(self, seed: int, row0: int, n: int, flags: int = CHECK_LIMITS | CHECK_COLLISION, out=None):
        """Validity of device-generated uniform rows [row0, row0+n) -> uint8 CUDA tensor."""
        torch = _torch()
        with self._call_lock, torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(n, dtype=torch.uint8, device=self.torch_device)
            if n:
                _abi.check(self._L.mjb_check_sweep(self._h, seed, row0, n, out.data_ptr(), flags, self._stream()))
            return out

    def sweep_rows(self, seed: int, row0: int, n: int):
        torch = _torch()
        with torch.cuda.device(self.device):
            q = torch.empty((n, self.nq), dtype=torch.float32, device=self.torch_device)
            if n:
                _abi.check(self._L.mjb_sweep_rows(self._h, seed, row0, n, q.data_ptr(), self._stream()))
            return q

    def kernel_timing(self, enable: bool = True, read: bool = False):
        """Switch per-kernel CUDA-event timing of the validity launches on/off; with ``read`` return
        ``{"first_ms", "mid_ms", "narrow_ms", "fp64_ms", "launches", "pipeline"}`` summed since the last
        read (``first`` = ``validity_kernel``, or ``fk_cull_kernel`` of the multi-kernel pipeline)."""
        with _torch().cuda.device(self.device):
            if not read:
                _abi.check(self._L.mjb_kernel_timing(self._h, int(enable), None, None))
                return None
            ms = (C.c_double * 4)()
            n = C.c_int64(0)
            _abi.check(self._L.mjb_kernel_timing(self._h, int(enable), ms, C.byref(n)))
            return {"first_ms": ms[0], "mid_ms": ms[1], "narrow_ms": ms[2], "fp64_ms": ms[3], "launches": int(n.value),
                    "pipeline": "fk_cull+mid+narrow" if ms[2] > 0.0 else "single"}

    def stats(self) -> dict:
        st = _abi.Stats()
        _abi.check(self._L.mjb_get_stats(self._h, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in _abi.Stats._fields_}

    def reset_stats(self) -> None:
        _abi.check(self._L.mjb_reset_stats(self._h))


def _close_cached_engines():
    """Release the cached engines while the CUDA context is still alive (registered with atexit:
    finalising them during interpreter teardown would race the runtime's own shutdown)."""
    with _lock:
        for e in list(_cache.values()):
            try:
                e.close()
            except Exception:
                pass
        _cache.clear()


atexit.register(_close_cached_engines)


def get_engine(model, allowed_collision_bodies=(), device: int | None = None) -> ValidityEngine:
    """Engines are cached per (model object, allow-list, device): constraints on the same model
    share device tables, which is also what lets ``obeys_constraints_batch`` fuse them."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise EngineUnavailable(
            "no CUDA device: mjpl_b200 runs its validity checks on the GPU only (no CPU fallback)"
        )
    dev = torch.cuda.current_device() if device is None else int(device)
    key = (id(model), tuple(sorted(tuple(sorted(p)) for p in allowed_collision_bodies)), dev)
    with _lock:
        e = _cache.get(key)
        if e is None or e.model is not model:
            e = ValidityEngine(model, allowed_collision_bodies, dev)
            _cache[key] = e
        return e


def fma_peak(device: int | None = None) -> dict:
    """Measured FP32 FMA throughput of a GPU (``mjb_fma_peak``): ``{"tflops", "ms"}``."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise EngineUnavailable("no CUDA device")
    t, ms = C.c_double(0), C.c_double(0)
    with torch.cuda.device(torch.cuda.current_device() if device is None else device):
        _abi.check(_abi.lib().mjb_fma_peak(C.byref(t), C.byref(ms)))
    return {"tflops": t.value, "ms": ms.value}


def sweep_rows_host(model, seed: int, row0: int, n: int) -> np.ndarray:
    """numpy mirror of the device row generator (``vk::sweep_value``), bit-identical in fp32."""
    lo = model.jnt_range[:, 0].astype(np.float32)
    hi = model.jnt_range[:, 1].astype(np.float32)
    nq = int(model.nq)
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(row0))[:, None]
    j = np.arange(nq, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (rows * np.uint64(64) + j + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (lo[None, :] + (u * (hi - lo)[None, :]).astype(np.float32)).astype(np.float32)
