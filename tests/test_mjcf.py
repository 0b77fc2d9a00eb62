"""MJCF-subset compiler: structure, defaults, ordering, fromto, meshes."""

import numpy as np
import pytest

from mjpl_b200 import mjcf, models
from mjpl_b200.model import GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_PLANE, GEOM_SPHERE, Model

from . import toy_models as toys


def test_toy_ball_models():
    m = mjcf.from_xml_string(toys.TWO_DOF_BALL)
    assert (m.nq, m.nv, m.nbody, m.njnt, m.ngeom, m.nsite) == (2, 2, 2, 2, 3, 1)
    assert m.body("ball").id == 1
    # world geoms first (plane, wall), then the ball's sphere: MuJoCo groups geoms by body id
    assert m.geom_type.tolist() == [GEOM_PLANE, GEOM_BOX, GEOM_SPHERE]
    assert m.geom("wall_obstacle").id == 1
    np.testing.assert_allclose(m.geom("wall_obstacle").pos, [0.6, 0, 1])
    np.testing.assert_allclose(m.geom("wall_obstacle").size, [0.1, 0.5, 0.5])
    np.testing.assert_allclose(m.jnt_range, [[-2, 2], [-2, 2]])
    assert m.jnt_limited.tolist() == [1, 1]
    assert m.body_weldid.tolist() == [0, 1]
    with pytest.raises(KeyError):
        m.body("nope")


def test_joint_zoo_indexing():
    from mjpl_b200.utils import all_joints, qpos_idx, qvel_idx

    m = mjcf.from_xml_string(toys.JOINT_ZOO)
    assert (m.nq, m.nv) == (13, 11)
    assert all_joints(m) == ["slide_joint", "free_joint", "hinge_joint", "ball_joint"]
    # reference: test/test_utils.py index expectations for the same joint layout
    assert qpos_idx(m, ["slide_joint"]) == [0]
    assert qpos_idx(m, ["free_joint"]) == [1, 2, 3, 4, 5, 6, 7]
    assert qpos_idx(m, ["hinge_joint"]) == [8]
    assert qpos_idx(m, ["ball_joint"]) == [9, 10, 11, 12]
    assert qvel_idx(m, ["ball_joint"]) == [8, 9, 10]
    assert qvel_idx(m, ["ball_joint", "hinge_joint", "free_joint"]) == [8, 9, 10, 7, 1, 2, 3, 4, 5, 6]
    assert qvel_idx(m, []) == []


def test_defaults_fromto_degrees_exclude():
    m = mjcf.from_xml_string(toys.PRIMITIVE_ARM)
    assert m.nq == 5 and m.nbody == 7
    # degree -> radian on hinge ranges, slide range untouched
    np.testing.assert_allclose(m.jnt_range[0], np.deg2rad([-170, 170]))
    np.testing.assert_allclose(m.jnt_range[1], np.deg2rad([-120, 120]))
    np.testing.assert_allclose(m.jnt_range[4], [0, 0.04])
    g = m.geom("l1_geom")
    assert g.type == GEOM_CAPSULE
    np.testing.assert_allclose(g.size[:2], [0.04, 0.15])
    np.testing.assert_allclose(g.pos, [0, 0, 0.15])
    g = m.geom("l3_geom")
    np.testing.assert_allclose(g.size[:2], [0.03, 0.125])  # explicit size overrides the class default
    assert m.geom("tip_geom").type == GEOM_BOX
    assert m.geom("base_geom").type == GEOM_CYLINDER
    ghost = [i for i in range(m.ngeom) if m.geom_contype[i] == 0]
    assert len(ghost) == 1 and m.geom_conaffinity[ghost[0]] == 0
    # euler "0 0 30" degrees about z
    np.testing.assert_allclose(m.geom("crate").quat, [np.cos(np.pi / 12), 0, 0, np.sin(np.pi / 12)], atol=1e-12)
    b1, b3 = m.body("l1").id, m.body("l3").id
    assert m.exclude_signature.tolist() == [(b1 << 16) + b3]
    # base has no joint: welded to the world; wrist/finger move
    assert m.body_weldid[m.body("base").id] == 0
    assert m.body_weldid[m.body("finger").id] == m.body("finger").id
    np.testing.assert_allclose(m.keyframe("home").qpos, [0, 0.5, -1.0, 0, 0.02])


def test_save_load_roundtrip(tmp_path):
    m = mjcf.from_xml_string(toys.PRIMITIVE_ARM)
    m.save(tmp_path / "m.npz")
    m2 = Model.load(tmp_path / "m.npz")
    for f in ("body_pos", "body_quat", "jnt_range", "geom_size", "geom_quat", "mesh_vert", "key_qpos"):
        np.testing.assert_array_equal(getattr(m, f), getattr(m2, f))
    assert m2.body_names == m.body_names and m2.nq == m.nq
    assert m2.joint("j3").id == 2


def test_bundled_models_census():
    """SURVEY.md Appendix B census of the reference's models."""
    f = models.load("franka_scene")
    assert (f.nq, f.njnt, f.nbody, f.ngeom) == (9, 9, 12, 82)
    coll = (f.geom_contype != 0) | (f.geom_conaffinity != 0)
    assert coll.sum() == 24
    assert len(f.mesh_vert) == 1234
    assert sorted(f.mesh_vertnum[f.mesh_vertnum > 0].tolist()) == sorted([102, 152, 152, 152, 152, 64, 41, 64, 102, 102, 102, 49])
    np.testing.assert_allclose(f.keyframe("home").qpos, [0, 0, 0, -1.57079, 0, 1.57079, -0.7853, 0.04, 0.04])
    assert f.body_weldid[f.body("link0").id] == 0
    assert f.body_weldid[f.body("hand").id] == f.body("link7").id
    o = models.load("franka_scene_with_obstacles")
    assert ((o.geom_contype != 0) | (o.geom_conaffinity != 0)).sum() == 31
    u = models.load("ur5e_scene")
    assert (u.nq, u.nbody) == (6, 8)
    assert ((u.geom_contype != 0) | (u.geom_conaffinity != 0)).sum() == 10


def test_bundled_tables_match_reference_xml(reference_dir):
    """The committed .npz tables are exactly what the compiler produces from the reference's XML."""
    import sys

    sys.path.insert(0, str(reference_dir.parent / "repo" / "tools"))
    from tools.compile_models import MODELS

    for name, rel in MODELS.items():
        fresh = mjcf.from_xml_path(reference_dir / rel)
        stored = models.load(name)
        for fld in ("body_parentid", "body_pos", "body_quat", "jnt_axis", "jnt_range", "geom_type", "geom_size",
                    "geom_pos", "geom_quat", "geom_bodyid", "mesh_vert", "mesh_vertnum", "key_qpos"):
            np.testing.assert_array_equal(getattr(fresh, fld), getattr(stored, fld), err_msg=f"{name}.{fld}")
