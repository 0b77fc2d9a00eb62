"""Minimal SE3 / SO3 types with the mink API subset mjpl's validity path uses.

The reference takes poses as ``mink.SE3`` (``pose_constraint.py:3-4,21``, ``utils.py:60-75``,
``rrt.py:5``).  mink is not installable here; this module provides the few operations those
call sites need under the same names: ``SE3.from_rotation_and_translation``, ``.inverse()``,
``.multiply()`` / ``@``, ``.rotation()``, ``.translation()``, ``SO3.from_matrix``,
``SO3.from_rpy_radians``, ``SO3.as_rpy_radians()`` (jaxlie convention: R = Rz(yaw) Ry(pitch)
Rx(roll)), ``SO3.wxyz``.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class RollPitchYaw:
    roll: float
    pitch: float
    yaw: float


def _qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


class SO3:
    def __init__(self, wxyz):
        q = np.asarray(wxyz, dtype=np.float64)
        self.wxyz = q / np.linalg.norm(q)

    @staticmethod
    def identity():
        return SO3([1.0, 0.0, 0.0, 0.0])

    @staticmethod
    def from_matrix(R):
        R = np.asarray(R, dtype=np.float64).reshape(3, 3)
        t = np.trace(R)
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
        else:
            i = int(np.argmax(np.diag(R)))
            j, k = (i + 1) % 3, (i + 2) % 3
            s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
            q = [0.0] * 4
            q[0] = (R[k, j] - R[j, k]) / s
            q[1 + i] = 0.25 * s
            q[1 + j] = (R[j, i] + R[i, j]) / s
            q[1 + k] = (R[k, i] + R[i, k]) / s
        return SO3(q)

    @staticmethod
    def from_rpy_radians(roll, pitch, yaw):
        cr, sr, cp, sp, cy, sy = np.cos(roll / 2), np.sin(roll / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(yaw / 2), np.sin(yaw / 2)
        return SO3([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])

    def as_matrix(self):
        w, x, y, z = self.wxyz
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    def as_rpy_radians(self) -> RollPitchYaw:
        w, x, y, z = self.wxyz
        return RollPitchYaw(
            roll=float(np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))),
            pitch=float(np.arcsin(np.clip(2 * (w * y - z * x), -1.0, 1.0))),
            yaw=float(np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))),
        )

    @staticmethod
    def exp(omega):
        """Rotation of angle |omega| about omega."""
        om = np.asarray(omega, dtype=np.float64).reshape(3)
        th = np.linalg.norm(om)
        if th < 1e-12:
            return SO3(np.concatenate([[1.0], 0.5 * om]))
        return SO3(np.concatenate([[np.cos(0.5 * th)], np.sin(0.5 * th) / th * om]))

    @staticmethod
    def from_x_radians(theta):
        return SO3.exp([theta, 0.0, 0.0])

    @staticmethod
    def from_y_radians(theta):
        return SO3.exp([0.0, theta, 0.0])

    @staticmethod
    def from_z_radians(theta):
        return SO3.exp([0.0, 0.0, theta])

    def parameters(self):
        return self.wxyz

    def inverse(self):
        w, x, y, z = self.wxyz
        return SO3([w, -x, -y, -z])

    def multiply(self, other):
        return SO3(_qmul(self.wxyz, other.wxyz))

    __matmul__ = multiply

    def apply(self, v):
        return self.as_matrix() @ np.asarray(v, dtype=np.float64)

    def log(self) -> np.ndarray:
        """Rotation vector (axis * angle), angle in [0, pi]."""
        w, v = self.wxyz[0], self.wxyz[1:]
        if w < 0:
            w, v = -w, -v
        n = np.linalg.norm(v)
        if n < 1e-12:
            return 2.0 * v
        return (2.0 * np.arctan2(n, w) / n) * v


class SE3:
    def __init__(self, rotation: SO3, translation):
        self._r = rotation
        self._t = np.asarray(translation, dtype=np.float64).reshape(3)

    @staticmethod
    def identity():
        return SE3(SO3.identity(), np.zeros(3))

    @staticmethod
    def from_rotation_and_translation(rotation: SO3, translation):
        return SE3(rotation, translation)

    @staticmethod
    def from_translation(translation):
        return SE3(SO3.identity(), translation)

    def rotation(self) -> SO3:
        return self._r

    def translation(self) -> np.ndarray:
        return self._t

    def inverse(self):
        ri = self._r.inverse()
        return SE3(ri, -ri.apply(self._t))

    def multiply(self, other):
        return SE3(self._r.multiply(other._r), self._t + self._r.apply(other._t))

    __matmul__ = multiply

    def log(self) -> np.ndarray:
        """se(3) tangent ``[v, omega]`` (translation part first, as in mink / jaxlie)."""
        om = self._r.log()
        th = np.linalg.norm(om)
        K = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
        if th < 1e-6:
            Vinv = np.eye(3) - 0.5 * K + (K @ K) / 12.0
        else:
            half = 0.5 * th
            Vinv = np.eye(3) - 0.5 * K + (1.0 - half * np.cos(half) / np.sin(half)) / (th * th) * (K @ K)
        return np.concatenate([Vinv @ self._t, om])

    @staticmethod
    def exp(tangent):
        """Inverse of ``log``: tangent ``[v, omega]`` -> SE3."""
        tg = np.asarray(tangent, dtype=np.float64).reshape(6)
        v, om = tg[:3], tg[3:]
        th = np.linalg.norm(om)
        K = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
        if th < 1e-6:
            V = np.eye(3) + 0.5 * K + (K @ K) / 6.0
        else:
            V = np.eye(3) + (1.0 - np.cos(th)) / (th * th) * K + (th - np.sin(th)) / (th ** 3) * (K @ K)
        return SE3(SO3.exp(om), V @ v)

    def interpolate(self, other, alpha: float = 0.5):
        """Geodesic interpolation: ``self @ exp(alpha * log(self^-1 @ other))``; the end points are
        returned as they are."""
        if alpha <= 0.0:
            return self
        if alpha >= 1.0:
            return other
        return self.multiply(SE3.exp(alpha * self.inverse().multiply(other).log()))

    def __eq__(self, other):
        if not isinstance(other, SE3):
            return NotImplemented
        same_rot = min(np.abs(self._r.wxyz - other._r.wxyz).max(), np.abs(self._r.wxyz + other._r.wxyz).max()) < 1e-12
        return bool(same_rot and np.abs(self._t - other._t).max() < 1e-12)

    __hash__ = None

    def rminus(self, other) -> np.ndarray:
        return other.inverse().multiply(self).log()

    minus = rminus  # mink's SE3.minus is the right-minus

    @property
    def wxyz_xyz(self):
        return np.concatenate([self._r.wxyz, self._t])
