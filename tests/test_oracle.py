"""Pins the fp64 CPU oracle: the reference's known-answer tests, closed-form geometry, the
model census, the committed golden vectors, and an independent second implementation.

"parity unpinned" against a real MuJoCo (not installable here); everything the reference's own
tests assert about this path that does not need the network is restated below with its
reference file:line.
"""

import numpy as np
import pytest

import oracle
from mjpl_b200 import mjcf, models

from . import toy_models as toys
from .hostsim import HostSim

GOLDEN = __import__("pathlib").Path(__file__).parent / "golden"


# ---------------------------------------------------------------- reference known answers
def test_collision_constraint_known_answers():
    # reference test/test_collision_constraint.py:17-33 on two_dof_ball.xml
    o = oracle.Oracle(models.load("two_dof_ball"))
    assert o.collision_valid_config(np.array([0.0, 0.0]))
    assert not o.collision_valid_config(np.array([0.6, 0.0]))


def test_joint_limit_known_answers():
    # reference test/test_joint_limit_constraint.py:15-31 (range -2..2)
    o = oracle.Oracle(models.load("two_dof_ball"))
    assert o.limits_valid_config(np.array([0.0, 0.0]))
    assert not o.limits_valid_config(np.array([2.5, 0.0]))
    assert o.limits_valid_config(np.array([2.0, -2.0]))  # closed interval


def test_valid_collision_interval_known_answers():
    # reference test/test_planning_utils.py:321-344 on one_dof_ball.xml
    o = oracle.Oracle(models.load("one_dof_ball"))
    assert not oracle.valid_collision_interval(o, [0.8], [1.5], 0.1)[0]
    assert oracle.valid_collision_interval(o, [0.8], [1.5], 0.2)[0]
    assert oracle.valid_collision_interval(o, [0.0], [0.2], 0.01)[0]
    with pytest.raises(ValueError, match="step_dist"):
        oracle.valid_collision_interval(o, [0.0], [0.2], 0.0)


def test_constrained_extend_known_answers():
    # reference test/test_planning_utils.py:207-256
    o = oracle.Oracle(models.load("one_dof_ball"))
    added, reached = oracle.constrained_extend_chain(o, [-0.1], [0.15], 0.1)
    np.testing.assert_allclose(np.array(added).ravel(), [0.0, 0.1, 0.15], atol=1e-9)
    np.testing.assert_equal(reached, [0.15])
    added, reached = oracle.constrained_extend_chain(o, [0.0], [1.0], 0.1)
    assert 0.0 < reached[0] < 0.85 and reached[0] < 1.0
    # :258-285 interval check with eps = inf
    _, r = oracle.constrained_extend_chain(o, [0.8], [1.8], np.inf, interval=0.3)
    np.testing.assert_equal(r, [1.8])
    _, r = oracle.constrained_extend_chain(o, [0.8], [1.8], np.inf, interval=0.1)
    np.testing.assert_equal(r, [0.8])


def test_home_keyframes_are_collision_free():
    # implicit known answers: rrt.py:154-155 raises unless q_init is valid, and the reference's CI
    # runs these scenes from the home keyframe (examples/benchmark.py:47,54-55; ci.yml:41-57)
    f = models.load("franka_scene")
    assert oracle.Oracle(f).check(f.keyframe("home").qpos, 3)[0]  # no allowed pairs (benchmark.py:47)
    fo = models.load("franka_scene_with_obstacles")
    assert oracle.Oracle(fo, [("left_finger", "right_finger")]).check(fo.keyframe("home").qpos, 3)[0]
    u = models.load("ur5e_scene")
    assert oracle.Oracle(u).check(u.keyframe("home").qpos, 3)[0]


def test_pair_census():
    # SURVEY.md Appendix B (static pair lists after MuJoCo's filters and mjpl's allow-list)
    f = models.load("franka_scene")
    assert len(oracle.Oracle(f).pairs()) == 206
    assert len(oracle.Oracle(f, [("left_finger", "right_finger")]).pairs()) == 170
    fo = models.load("franka_scene_with_obstacles")
    assert len(oracle.Oracle(fo, [("left_finger", "right_finger")]).pairs()) == 324
    assert len(oracle.Oracle(models.load("ur5e_scene")).pairs()) == 30
    # link0-link1 ARE tested (world exemption of the parent filter), fingers vs hand/link7 are not
    o = oracle.Oracle(f)
    bodies = {tuple(sorted((f.geom_bodyid[a], f.geom_bodyid[b]))) for a, b in o.pairs()}
    b = {n: f.body(n).id for n in f.body_names}
    assert (b["link0"], b["link1"]) in bodies
    assert (b["link1"], b["link2"]) not in bodies
    assert (b["hand"], b["left_finger"]) not in bodies and (b["link7"], b["left_finger"]) not in bodies
    assert (b["left_finger"], b["right_finger"]) in bodies
    assert (b["world"], b["link0"]) not in bodies
    # allowed pair order does not matter (collision_constraint.py:55-64 sorts)
    assert len(oracle.Oracle(f, [("right_finger", "left_finger")]).pairs()) == 170
    with pytest.raises(KeyError):
        oracle.Oracle(f, [("left_finger", "no_such_body")])


def test_link0_link1_standing_gap():
    # two parallel discs 0.993 mm apart at every q
    f = models.load("franka_scene")
    o = oracle.Oracle(f)
    pr = o.pairs()
    b0, b1 = f.body("link0").id, f.body("link1").id
    k = [i for i, (a, b) in enumerate(pr) if {f.geom_bodyid[a], f.geom_bodyid[b]} == {b0, b1}]
    assert len(k) == 1
    q = f.keyframe("home").qpos.copy()
    for j1 in (-2.0, 0.0, 1.3):
        q[0] = j1
        assert abs(o.pair_distance(q, k[0]) - 0.993e-3) < 2e-6


# ---------------------------------------------------------------- closed-form geometry
def _two_geom_model(g1: str, g2: str, extra=""):
    xml = f"""<mujoco><worldbody>
      <body name="a" pos="0 0 0"><joint type="slide" axis="1 0 0" range="-10 10"/>
        <joint type="slide" axis="0 1 0" range="-10 10"/><joint type="slide" axis="0 0 1" range="-10 10"/>
        <joint type="hinge" axis="0 0 1" range="-10 10"/>{g1}</body>
      {g2}{extra}</worldbody></mujoco>"""
    return mjcf.from_xml_string(xml)


@pytest.mark.parametrize("g1,g2,q,expect", [
    ('<geom type="sphere" size="0.1"/>', '<geom type="sphere" size="0.2" pos="1 0 0"/>', [0, 0, 0, 0], 0.7),
    ('<geom type="sphere" size="0.1"/>', '<geom type="plane" size="1 1 1"/>', [0.3, 0.2, 0.5, 0], 0.4),
    ('<geom type="box" size="0.1 0.2 0.3"/>', '<geom type="plane" size="1 1 1"/>', [0, 0, 0.5, 0.7], 0.2),
    ('<geom type="capsule" size="0.1 0.3"/>', '<geom type="plane" size="1 1 1"/>', [0, 0, 0.5, 0], 0.1),
    ('<geom type="cylinder" size="0.1 0.3"/>', '<geom type="plane" size="1 1 1"/>', [0, 0, 0.5, 0], 0.2),
    ('<geom type="capsule" size="0.1 0.3"/>', '<geom type="capsule" size="0.05 0.2" pos="1 0 0"/>', [0, 0, 0, 0], 0.85),
    ('<geom type="sphere" size="0.1"/>', '<geom type="box" size="0.2 0.2 0.2" pos="1 0 0"/>', [0, 0, 0, 0], 0.7),
    ('<geom type="sphere" size="0.1"/>', '<geom type="box" size="0.2 0.2 0.2" pos="1 1 0"/>', [0, 0, 0, 0], np.hypot(0.8, 0.8) - 0.1),
    ('<geom type="box" size="0.1 0.1 0.1"/>', '<geom type="box" size="0.2 0.2 0.2" pos="1 0 0"/>', [0, 0, 0, 0], 0.7),
    ('<geom type="box" size="0.1 0.1 0.1"/>', '<geom type="box" size="0.2 0.2 0.2" pos="1 0 0"/>', [0, 0, 0, np.pi / 4], 0.8 - 0.1 * np.sqrt(2)),
    ('<geom type="capsule" size="0.1 0.3"/>', '<geom type="box" size="0.2 0.2 0.2" pos="1 0 0"/>', [0, 0, 0, 0], 0.7),
    ('<geom type="cylinder" size="0.1 0.3"/>', '<geom type="cylinder" size="0.2 0.1" pos="1 0 0"/>', [0, 0, 0, 0], 0.7),
    ('<geom type="cylinder" size="0.1 0.3"/>', '<geom type="capsule" size="0.2 0.1" pos="0 0 1"/>', [0, 0, 0, 0], 0.4),
    ('<geom type="sphere" size="0.1"/>', '<geom type="cylinder" size="0.2 0.1" pos="0 0 1"/>', [0, 0, 0, 0], 0.8),
])
def test_closed_form_distances(g1, g2, q, expect):
    m = _two_geom_model(g1, g2)
    o = oracle.Oracle(m)
    assert len(o.pairs()) == 1
    assert o.pair_distance(np.array(q, float), 0) == pytest.approx(expect, abs=1e-9)
    # the batched check agrees (pairs removed by the conservative sphere cull report +1e30)
    v, d, _ = o.check(np.array([q], float), oracle.CHECK_COLLISION, want_dist=True)
    assert v[0] and (d[0] > 1e29 or d[0] == pytest.approx(expect, abs=1e-9))


def test_penetration_depths():
    m = _two_geom_model('<geom type="box" size="0.1 0.1 0.1"/>', '<geom type="box" size="0.2 0.2 0.2" pos="0.2999 0 0"/>')
    o = oracle.Oracle(m)
    assert o.pair_distance(np.zeros(4), 0) == pytest.approx(-1e-4, abs=1e-9)   # shallow: exact depth from EPA
    v, d, _ = o.check(np.zeros((1, 4)), 2, want_dist=True)
    assert not v[0] and d[0] == pytest.approx(-1e-4, abs=1e-9)
    assert o.pair_distance(np.array([0.1, 0, 0, 0.0]), 0) <= -oracle.DEPTH_CAP  # deep: capped
    # touching counts as contact (dist <= margin with margin 0)
    m = _two_geom_model('<geom type="sphere" size="0.25"/>', '<geom type="sphere" size="0.25" pos="0.5 0 0"/>')
    assert not oracle.Oracle(m).check(np.zeros((1, 4)), 2)[0]


def test_geom_margin_is_max_of_pair():
    m = _two_geom_model('<geom type="sphere" size="0.1" margin="0.05"/>', '<geom type="sphere" size="0.1" pos="0.23 0 0" margin="0.01"/>')
    o = oracle.Oracle(m)
    assert o.pair_distance(np.zeros(4), 0) == pytest.approx(0.03 - 0.05, abs=1e-12)
    assert not o.check(np.zeros((1, 4)), 2)[0]
    # 0.23 + 0.021 apart: gap 0.051 > margin 0.05 -> no contact
    assert o.check(np.array([[-0.021, 0, 0, 0]]), 2)[0]


def test_fk_against_hand_computation():
    m = mjcf.from_xml_string(toys.PRIMITIVE_ARM)
    o = oracle.Oracle(m)
    q = np.array([0.3, -0.4, 0.5, 0.2, 0.01])
    xpos, xquat = o.fk(q)
    # l1 rotates about world z at (0,0,0.25); l2 origin is 0.3 above, unaffected by j1
    np.testing.assert_allclose(xpos[0, m.body("l1").id], [0, 0, 0.25], atol=1e-15)
    np.testing.assert_allclose(xpos[0, m.body("l2").id], [0, 0, 0.55], atol=1e-15)
    # l2 frame: rotz(j1) * rotx(90deg) * rotz(j2); l3 sits 0.3 along l2's x axis
    def rz(a): return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    def rx(a): return np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    R2 = rz(0.3) @ rx(np.pi / 2) @ rz(-0.4)
    np.testing.assert_allclose(xpos[0, m.body("l3").id], [0, 0, 0.55] + R2 @ [0.3, 0, 0], atol=1e-12)
    R3 = R2 @ rz(0.5)
    p_wrist = xpos[0, m.body("l3").id] + R3 @ [0.25, 0, 0]
    np.testing.assert_allclose(xpos[0, m.body("wrist").id], p_wrist, atol=1e-12)
    R4 = R3 @ rx(0.2)
    np.testing.assert_allclose(xpos[0, m.body("finger").id], p_wrist + R4 @ [0.1, 0.01, 0], atol=1e-12)
    assert np.allclose(np.linalg.norm(xquat[0], axis=1), 1.0, atol=1e-14)


# ---------------------------------------------------------------- golden vectors + second implementation
@pytest.mark.parametrize("name,mname,allowed", [
    ("franka_scene", "franka_scene", []),
    ("franka_obstacles", "franka_scene_with_obstacles", [("left_finger", "right_finger")]),
    ("ur5e_scene", "ur5e_scene", []),
    ("two_dof_ball", "two_dof_ball", []),
])
def test_oracle_reproduces_golden(name, mname, allowed):
    g = np.load(GOLDEN / f"{name}.npz")
    o = oracle.Oracle(models.load(mname), allowed)
    valid, dist, _ = o.check(g["q"].astype(np.float64), 3, want_dist=True)
    np.testing.assert_array_equal(valid, g["valid"])
    np.testing.assert_allclose(dist, g["dist"], atol=1e-9)
    xpos, xquat = o.fk(g["q"][: len(g["xpos"])].astype(np.float64))
    np.testing.assert_allclose(xpos, g["xpos"], atol=1e-12)
    mj = GOLDEN / f"{name}_mujoco.npz"
    if mj.exists():  # real-MuJoCo vectors, when somebody could generate them (tools/make_golden.py --mujoco)
        r = np.load(mj)
        v2, d2, _ = o.check(r["q"].astype(np.float64), 3, want_dist=True)
        bad = v2 != r["valid"]
        assert not (bad & (np.abs(d2) >= 1e-5)).any()
        assert np.abs(o.fk(r["q"][: len(r["xpos"])].astype(np.float64))[0] - r["xpos"]).max() < 1e-9


@pytest.mark.parametrize("mname,allowed,n", [
    ("franka_scene", [], 6000),
    ("franka_scene_with_obstacles", [("left_finger", "right_finger")], 4000),
    ("ur5e_scene", [], 20000),
])
def test_oracle_vs_kernel_core_on_cpu(mname, allowed, n):
    """Two independent implementations (oracle: Voronoi-region GJK + EPA in C; kernel core:
    signed-volume GJK with certified verdicts + OBB mid-phase, compiled for the host) must give
    the same validity on seeded rows, and the same static pair list."""
    m = models.load(mname)
    o, h = oracle.Oracle(m, allowed), HostSim(m, allowed)
    assert set(map(tuple, o.pairs().tolist())) == set(map(tuple, h.pairs().tolist()))
    rng = np.random.default_rng(7)
    Q = rng.uniform(m.jnt_range[:, 0], m.jnt_range[:, 1], size=(n, m.nq)).astype(np.float32)
    v, d, _ = o.check(Q.astype(np.float64), 3, want_dist=True)
    for obb in (False, True):
        hv, st = h.check(Q, 3, obb=obb)
        bad = hv.astype(bool) != v
        assert not (bad & (np.abs(d) >= 1e-5)).any(), st
    xp, xq = h.fk(Q[:500])
    op, oq = o.fk(Q[:500].astype(np.float64))
    assert np.abs(xp - op).max() < 1e-5  # north-star FK tolerance, fp32 vs fp64
    assert np.minimum(np.abs(xq - oq).max(-1), np.abs(xq + oq).max(-1)).max() < 1e-5


def test_primitive_zoo_core_vs_oracle():
    m = mjcf.from_xml_string(toys.PRIMITIVE_ARM)
    o, h = oracle.Oracle(m), HostSim(m)
    rng = np.random.default_rng(3)
    Q = rng.uniform(m.jnt_range[:, 0], m.jnt_range[:, 1], size=(20000, m.nq)).astype(np.float32)
    v, d, _ = o.check(Q.astype(np.float64), 3, want_dist=True)
    hv, _ = h.check(Q, 3)
    assert not ((hv.astype(bool) != v) & (np.abs(d) >= 1e-5)).any()
    assert 0.2 < v.mean() < 0.95


def test_unsupported_models_are_rejected():
    m = mjcf.from_xml_string(toys.JOINT_ZOO)
    with pytest.raises(ValueError, match="hinge"):
        oracle.Oracle(m)
    with pytest.raises(ValueError, match="hinge and slide"):
        HostSim(m)


def test_hill_climbing_support_equals_full_scan(monkeypatch):
    """The hull graphs (MuJoCo mesh_graph layout) and the hill-climbing support query return a true
    support vertex for every mesh and 3000 directions (cold and warm starts)."""
    monkeypatch.setenv("MJB_HILL_MIN", "16")
    m = models.load("franka_scene_with_obstacles")
    assert (m.mesh_graphadr >= 0).sum() == 12
    h = HostSim(m, [("left_finger", "right_finger")])
    nshape, gap, evals_hill, evals_scan = h.support_check(3000, seed=5)
    assert nshape == 13 and gap == 0.0
    assert evals_hill * 3 < evals_scan
    monkeypatch.setenv("MJB_HILL_MIN", "100000")
    assert HostSim(m).support_check(10)[0] == 0  # switched off -> no shape uses a graph
