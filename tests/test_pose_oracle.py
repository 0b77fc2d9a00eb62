"""Pins the numpy restatement of PoseConstraint (oracle.PoseOracle) with the reference's own
test cases (test/test_pose_constraint.py) and with finite differences."""

import numpy as np
import pytest

import oracle
from mjpl_b200 import models
from mjpl_b200.lie import SE3, SO3

INF = (-np.inf, np.inf)


def test_translation_limit_known_answer():
    # reference test/test_pose_constraint.py:17-52 (two_dof_ball, ball_site at (0,0,1) at q=0)
    m = models.load("two_dof_ball")
    po = oracle.PoseOracle(m, "ball_site", [0, 0, 1.0], [1, 0, 0, 0], [(-0.1, 0.1)] + [INF] * 5, q_step=np.inf)
    q = np.array([0.2, 0.0])
    assert not po.valid_config(q)
    qc = po.apply(np.zeros(2), q)
    np.testing.assert_allclose(qc, [0.1, 0.0], rtol=0, atol=1e-12)
    assert po.valid_config(qc)
    po.q_step = 1e-5
    assert po.apply(np.zeros(2), q) is None


def test_rotation_limit_properties():
    # reference test/test_pose_constraint.py:54-119 on the local ur5e.xml
    u = models.load("ur5e_scene")
    q0 = u.keyframe("home").qpos
    free = oracle.PoseOracle(u, "attachment_site", [0, 0, 0], [1, 0, 0, 0], [INF] * 6)
    p, r = free.site_pose(q0)
    lim = (-0.1, 0.1)
    pc = oracle.PoseOracle(u, "attachment_site", p, r, [INF] * 3 + [lim, lim, INF], q_step=np.inf)
    assert pc.valid_config(q0)
    rng = np.random.default_rng(5)
    done = 0
    for _ in range(200):
        qr = q0 + rng.uniform(-0.6, 0.6, 6)
        if pc.valid_config(qr):
            continue
        qa = pc.apply(q0, qr)
        if qa is None:
            continue
        done += 1
        assert pc.valid_config(qa)
        ps, rs = pc.site_pose(qa)
        rel = SE3(SO3(r), p).inverse().multiply(SE3(SO3(rs), ps)).rotation().as_rpy_radians()
        assert lim[0] - 1e-3 <= rel.roll <= lim[1] + 1e-3 and lim[0] - 1e-3 <= rel.pitch <= lim[1] + 1e-3
        pc.q_step = 1e-5
        assert pc.apply(q0, qr) is None
        pc.q_step = np.inf
    assert done >= 20


def test_jacobian_against_finite_differences():
    """mj_jacSite restated: translational rows and the roll / yaw rows of the RPY Jacobian equal
    finite differences of (site position, rpy).  The pitch row does NOT: the reference's _e_rpy
    (pose_constraint.py:165) uses cos(pitch) where the exact inverse has cos(yaw); that quirk is
    restated faithfully, so the row is compared with the reference's own formula instead."""
    u = models.load("ur5e_scene")
    po = oracle.PoseOracle(u, "attachment_site", [0, 0, 0], [1, 0, 0, 0], [INF] * 6)

    def d_c(q):
        p, r = po.site_pose(q)
        return np.concatenate([p, oracle._rpy(r)])

    rng = np.random.default_rng(0)
    for _ in range(5):
        q = rng.uniform(-2, 2, 6)
        J, eps = po.jacobian(q), 1e-7
        Jn = np.stack([(d_c(q + eps * e) - d_c(q - eps * e)) / (2 * eps) for e in np.eye(6)], axis=1)
        np.testing.assert_allclose(J[[0, 1, 2, 3, 5]], Jn[[0, 1, 2, 3, 5]], atol=2e-6)
        _, pitch, yaw = oracle._rpy(po.site_pose(q)[1])
        # angular velocity rows recovered from the exact rows, then the reference's pitch row
        cp, sp, cy, sy = np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
        Einv = np.array([[cy / cp, sy / cp, 0], [-sy, cy, 0], [cy * sp / cp, sy * sp / cp, 1]])
        w = np.linalg.solve(Einv, Jn[3:])
        np.testing.assert_allclose(J[4], -sy * w[0] + cp * w[1], atol=2e-6)


def test_lie_types():
    a = SE3(SO3.from_rpy_radians(0.3, -0.4, 1.1), [0.1, 0.2, 0.3])
    b = SE3(SO3.from_rpy_radians(-1.0, 0.2, 0.5), [-0.4, 0.0, 0.9])
    i = a.multiply(a.inverse())
    np.testing.assert_allclose(i.translation(), 0, atol=1e-15)
    np.testing.assert_allclose(np.abs(i.rotation().wxyz), [1, 0, 0, 0], atol=1e-15)
    rpy = a.rotation().as_rpy_radians()
    np.testing.assert_allclose([rpy.roll, rpy.pitch, rpy.yaw], [0.3, -0.4, 1.1], atol=1e-14)
    np.testing.assert_allclose(SO3.from_matrix(a.rotation().as_matrix()).wxyz, a.rotation().wxyz, atol=1e-14)
    c = a @ b
    np.testing.assert_allclose(c.translation(), a.translation() + a.rotation().as_matrix() @ b.translation(), atol=1e-15)
    np.testing.assert_allclose(c.rotation().as_matrix(), a.rotation().as_matrix() @ b.rotation().as_matrix(), atol=1e-15)


def test_kernel_pose_core_matches_restatement_on_cpu():
    """The CUDA kernel's row functions (pose_valid_row / pose_project_row in vk_core.cuh),
    compiled for the host, against the numpy restatement of the reference."""
    from mjpl_b200 import _abi

    from .hostsim import HostSim

    for mname, site in (("ur5e_scene", "attachment_site"), ("franka_scene", "ee_site")):
        model = models.load(mname)
        q0 = model.keyframe("home").qpos.copy()
        free = oracle.PoseOracle(model, site, [0, 0, 0], [1, 0, 0, 0], [INF] * 6)
        p, r = free.site_pose(q0)
        lim = (-0.1, 0.1)
        box = [INF, INF, (-0.05, 0.2), lim, lim, INF]
        po = oracle.PoseOracle(model, site, p, r, box, q_step=0.5)
        sid = model.site(site).id
        sp = _abi.PoseSpec()
        sp.site_bodyid = int(model.site_bodyid[sid])
        sp.site_pos[:] = list(model.site_pos[sid])
        sp.site_quat[:] = list(model.site_quat[sid])
        sp.ref_pos[:] = list(p)
        sp.ref_quat[:] = list(r)
        sp.lower[:] = [b[0] for b in box]
        sp.upper[:] = [b[1] for b in box]
        sp.tolerance, sp.q_step = 0.001, 0.5
        rng = np.random.default_rng(1)
        n = 200
        Q = q0[None, :] + rng.uniform(-0.35, 0.35, size=(n, model.nq))
        Q[:, 7:] = q0[7:]
        Q = np.clip(Q, model.jnt_range[:, 0], model.jnt_range[:, 1])
        h = HostSim(model)
        _, valid, _ = h.pose(sp, Q, Q, project=False)
        np.testing.assert_array_equal(valid, [po.valid_config(q) for q in Q])
        out, ok, _ = h.pose(sp, np.tile(q0, (n, 1)), Q)
        want = [po.apply(q0, q) for q in Q]
        np.testing.assert_array_equal(ok, [w is not None for w in want])
        assert ok.sum() > 10
        for i in np.flatnonzero(ok):
            np.testing.assert_allclose(out[i], want[i], atol=1e-9)
