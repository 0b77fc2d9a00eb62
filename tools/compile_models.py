"""Compile the reference's MJCF models into the constant tables the engine consumes.

The GPU box has no /root/reference, so the compiled tables (hull vertices, kinematic tree,
geom list; ~20 KB each) are committed under mjpl_b200/models/*.npz.  They are DERIVED DATA
produced by mjpl_b200.mjcf from the reference's model files; rerun this script to regenerate:

    python tools/compile_models.py [/root/reference]
"""

import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from mjpl_b200 import mjcf  # noqa: E402

MODELS = {
    "franka_scene": "examples/models/franka_emika_panda/scene.xml",
    "franka_scene_with_obstacles": "examples/models/franka_emika_panda/scene_with_obstacles.xml",
    "ur5e_scene": "examples/models/universal_robots_ur5e/scene.xml",
    "one_dof_ball": "test/models/one_dof_ball.xml",
    "two_dof_ball": "test/models/two_dof_ball.xml",
}

if __name__ == "__main__":
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    out = Path(__file__).resolve().parent.parent / "mjpl_b200" / "models"
    out.mkdir(exist_ok=True)
    for name, rel in MODELS.items():
        m = mjcf.from_xml_path(ref / rel)
        m.save(out / f"{name}.npz")
        print(f"{name}: nq={m.nq} nbody={m.nbody} ngeom={m.ngeom} hullverts={len(m.mesh_vert)} -> {out / (name + '.npz')}")
