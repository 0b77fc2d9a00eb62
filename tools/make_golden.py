"""Golden vectors for the validity path.

Two modes:

* default (what was run to produce tests/golden/*.npz, committed): the vectors come from the
  in-repo fp64 oracle.  They pin the oracle (and through it the CUDA path) against regressions;
  they are NOT MuJoCo outputs -- MuJoCo cannot be installed in this image ("parity unpinned",
  see oracle/oracle.h).
* ``--mujoco``: when a real MuJoCo >= 3 is importable, run the reference's own code path
  (data.qpos = q; mj_kinematics; mj_collision; CollisionRuleset rule -- reference:
  src/mjpl/constraint/collision_constraint.py:26-30, 83-95) on the same seeded q and write
  ``*_mujoco.npz`` next to the oracle vectors, with body poses and the signed distance of the
  closest tested pair from ``mj_geomDistance``.  tests/test_golden.py picks those files up
  automatically and holds both the oracle and the CUDA path to them.

    python tools/make_golden.py [--mujoco /path/to/reference]
"""

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CASES = {
    # name: (bundled model, allowed body pairs, rows, reference MJCF)
    "franka_scene": ("franka_scene", [], 2000, "examples/models/franka_emika_panda/scene.xml"),
    "franka_obstacles": ("franka_scene_with_obstacles", [("left_finger", "right_finger")], 2000,
                         "examples/models/franka_emika_panda/scene_with_obstacles.xml"),
    "ur5e_scene": ("ur5e_scene", [], 2000, "examples/models/universal_robots_ur5e/scene.xml"),
    "two_dof_ball": ("two_dof_ball", [], 500, "test/models/two_dof_ball.xml"),
}
NFK = 64


def seeded_rows(model, n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(model.jnt_range[:, 0], model.jnt_range[:, 1], size=(n, model.nq)).astype(np.float32)


def from_oracle(out: Path):
    import oracle
    from mjpl_b200 import models

    for name, (mname, allowed, n, _) in CASES.items():
        m = models.load(mname)
        o = oracle.Oracle(m, allowed)
        q = seeded_rows(m, n)
        q[0] = m.key_qpos[0] if m.nkey else q[0]
        valid, dist, pair = o.check(q.astype(np.float64), oracle.CHECK_LIMITS | oracle.CHECK_COLLISION, want_dist=True)
        xpos, xquat = o.fk(q[:NFK].astype(np.float64))
        np.savez_compressed(out / f"{name}.npz", q=q, valid=valid, dist=dist, pair=pair, xpos=xpos, xquat=xquat,
                            pairs=o.pairs(), source=np.array(["oracle"]))
        print(f"{name}: {n} rows, valid={valid.mean():.3f}, in-band={(np.abs(dist) < 1e-5).sum()}")


def from_mujoco(out: Path, ref: Path):
    import mujoco

    from mjpl_b200.model import Model

    for name, (_, allowed, n, rel) in CASES.items():
        mj = mujoco.MjModel.from_xml_path(str(ref / rel))
        data = mujoco.MjData(mj)
        q = seeded_rows(Model.from_mjmodel(mj), n)
        allowed_ids = {tuple(sorted((mj.body(a).id, mj.body(b).id))) for a, b in allowed}
        valid = np.zeros(n, bool)
        xpos = np.zeros((NFK, mj.nbody, 3))
        xquat = np.zeros((NFK, mj.nbody, 4))
        for i in range(n):
            data.qpos = q[i].astype(np.float64)
            mujoco.mj_kinematics(mj, data)
            mujoco.mj_collision(mj, data)
            bodies = {tuple(sorted(mj.geom_bodyid[g])) for g in data.contact.geom}
            lim = np.all((q[i] >= mj.jnt_range[:, 0]) & (q[i] <= mj.jnt_range[:, 1]))
            valid[i] = lim and bodies <= allowed_ids
            if i < NFK:
                xpos[i], xquat[i] = data.xpos, data.xquat
        np.savez_compressed(out / f"{name}_mujoco.npz", q=q, valid=valid, xpos=xpos, xquat=xquat,
                            source=np.array([f"mujoco {mujoco.__version__}"]))
        print(f"{name}: {n} rows from MuJoCo {mujoco.__version__}, valid={valid.mean():.3f}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mujoco", metavar="REFERENCE_DIR", default=None)
    a = ap.parse_args()
    out = ROOT / "tests" / "golden"
    out.mkdir(exist_ok=True)
    if a.mujoco:
        from_mujoco(out, Path(a.mujoco))
    else:
        from_oracle(out)
