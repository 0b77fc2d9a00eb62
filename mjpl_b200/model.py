"""MuJoCo-free model container with the ``MjModel`` field names the hot path reads.

The reference keeps all of this inside ``mujoco.MjModel`` (a third-party C struct);
the hot path only ever reads a small set of constant tables from it (SURVEY.md
section 8b, "Model accessors").  :class:`Model` carries exactly those tables as numpy
arrays under MuJoCo's own field names, so reference-style callers keep working
unchanged:

* ``model.body(name).id``                -> ``collision_constraint.py:61``
* ``model.jnt_range``                    -> ``joint_limit_constraint.py:16-17``, ``rrt.py:206``
* ``model.geom_bodyid``                  -> ``collision_constraint.py:93``
* ``model.joint(j).id/.name``, ``jnt_type``, ``jnt_qposadr``, ``jnt_dofadr`` -> ``utils.py:19-55``
* ``model.keyframe("home").qpos``        -> ``examples/benchmark.py:54-55``
* ``model.geom(name).pos/.size``         -> ``test/test_planning_utils.py:247-248``

A real ``mujoco.MjModel`` can be converted with :func:`Model.from_mjmodel` when MuJoCo is
installed (it is not in this image).
"""

from __future__ import annotations

import io
from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np

# mjtGeom
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE = 0, 1, 2, 3
GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 4, 5, 6, 7
GEOM_TYPE_NAMES = {
    "plane": GEOM_PLANE,
    "hfield": GEOM_HFIELD,
    "sphere": GEOM_SPHERE,
    "capsule": GEOM_CAPSULE,
    "ellipsoid": GEOM_ELLIPSOID,
    "cylinder": GEOM_CYLINDER,
    "box": GEOM_BOX,
    "mesh": GEOM_MESH,
}
# mjtJoint
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
JNT_TYPE_NAMES = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}
JNT_QPOS_WIDTH = {JNT_FREE: 7, JNT_BALL: 4, JNT_SLIDE: 1, JNT_HINGE: 1}
JNT_DOF_WIDTH = {JNT_FREE: 6, JNT_BALL: 3, JNT_SLIDE: 1, JNT_HINGE: 1}

_ARRAY_FIELDS = (
    "body_parentid body_weldid body_jntadr body_jntnum body_geomadr body_geomnum "
    "body_pos body_quat jnt_type jnt_qposadr jnt_dofadr jnt_bodyid jnt_pos jnt_axis "
    "jnt_range jnt_limited qpos0 geom_type geom_bodyid geom_contype geom_conaffinity "
    "geom_size geom_pos geom_quat geom_rbound geom_margin geom_gap geom_dataid "
    "mesh_vertadr mesh_vertnum mesh_vert mesh_graphadr mesh_graph site_bodyid site_pos site_quat key_qpos "
    "exclude_signature"
).split()
_NAME_FIELDS = "body_names jnt_names geom_names site_names mesh_names key_names".split()


class _Named(SimpleNamespace):
    """Tiny stand-in for MuJoCo's named-access views (``model.body("x")``)."""


@dataclass
class Model:
    nq: int = 0
    nv: int = 0
    nbody: int = 0
    njnt: int = 0
    ngeom: int = 0
    nsite: int = 0
    nmesh: int = 0
    nkey: int = 0
    body_parentid: np.ndarray = None
    body_weldid: np.ndarray = None
    body_jntadr: np.ndarray = None
    body_jntnum: np.ndarray = None
    body_geomadr: np.ndarray = None
    body_geomnum: np.ndarray = None
    body_pos: np.ndarray = None
    body_quat: np.ndarray = None
    jnt_type: np.ndarray = None
    jnt_qposadr: np.ndarray = None
    jnt_dofadr: np.ndarray = None
    jnt_bodyid: np.ndarray = None
    jnt_pos: np.ndarray = None
    jnt_axis: np.ndarray = None
    jnt_range: np.ndarray = None
    jnt_limited: np.ndarray = None
    qpos0: np.ndarray = None
    geom_type: np.ndarray = None
    geom_bodyid: np.ndarray = None
    geom_contype: np.ndarray = None
    geom_conaffinity: np.ndarray = None
    geom_size: np.ndarray = None
    geom_pos: np.ndarray = None
    geom_quat: np.ndarray = None
    geom_rbound: np.ndarray = None
    geom_margin: np.ndarray = None
    geom_gap: np.ndarray = None
    geom_dataid: np.ndarray = None
    mesh_vertadr: np.ndarray = None
    mesh_vertnum: np.ndarray = None
    mesh_vert: np.ndarray = None  # (nmeshvert, 3) float64: convex-hull vertices only
    mesh_graphadr: np.ndarray = None  # (nmesh,) start of each mesh's hull graph in mesh_graph, -1 = none
    mesh_graph: np.ndarray = None  # MuJoCo layout per mesh: nvert, nface, vert_edgeadr[nvert], vert_globalid[nvert],
    #                                edge_localid[nvert+3*nface] (neighbour lists, -1 terminated), face_globalid[3*nface]
    site_bodyid: np.ndarray = None
    site_pos: np.ndarray = None
    site_quat: np.ndarray = None
    key_qpos: np.ndarray = None
    exclude_signature: np.ndarray = None  # (nexclude,) int: (b1<<16)+b2, b1<b2
    body_names: list = field(default_factory=list)
    jnt_names: list = field(default_factory=list)
    geom_names: list = field(default_factory=list)
    site_names: list = field(default_factory=list)
    mesh_names: list = field(default_factory=list)
    key_names: list = field(default_factory=list)
    disable_contact: bool = False
    disable_filterparent: bool = False
    timestep: float = 0.002
    name: str = ""

    # ---- MuJoCo-style named access ------------------------------------------------
    @property
    def opt(self):
        return _Named(
            timestep=self.timestep,
            disableflags=(16 if self.disable_contact else 0)
            | (512 if self.disable_filterparent else 0),
        )

    @staticmethod
    def _resolve(key, names, kind) -> int:
        if isinstance(key, (int, np.integer)):
            if key < 0 or key >= len(names):
                raise IndexError(f"Invalid {kind} index {key}")
            return int(key)
        for i, n in enumerate(names):
            if n == key and n:
                return i
        # MuJoCo's bindings raise KeyError for unknown names (collision_constraint.py:61)
        raise KeyError(f"Invalid name '{key}'. Valid names: {[n for n in names if n]}")

    def body(self, key):
        i = self._resolve(key, self.body_names, "body")
        return _Named(
            id=i,
            name=self.body_names[i],
            parentid=int(self.body_parentid[i]),
            pos=self.body_pos[i],
            quat=self.body_quat[i],
            geomadr=int(self.body_geomadr[i]),
            geomnum=int(self.body_geomnum[i]),
        )

    def joint(self, key):
        i = self._resolve(key, self.jnt_names, "joint")
        return _Named(
            id=i,
            name=self.jnt_names[i],
            type=int(self.jnt_type[i]),
            range=self.jnt_range[i],
            qposadr=int(self.jnt_qposadr[i]),
            dofadr=int(self.jnt_dofadr[i]),
            bodyid=int(self.jnt_bodyid[i]),
        )

    def geom(self, key):
        i = self._resolve(key, self.geom_names, "geom")
        return _Named(
            id=i,
            name=self.geom_names[i],
            type=int(self.geom_type[i]),
            pos=self.geom_pos[i],
            quat=self.geom_quat[i],
            size=self.geom_size[i],
            bodyid=int(self.geom_bodyid[i]),
        )

    def site(self, key):
        i = self._resolve(key, self.site_names, "site")
        return _Named(
            id=i,
            name=self.site_names[i],
            pos=self.site_pos[i],
            quat=self.site_quat[i],
            bodyid=int(self.site_bodyid[i]),
        )

    def keyframe(self, key):
        i = self._resolve(key, self.key_names, "keyframe")
        return _Named(id=i, name=self.key_names[i], qpos=self.key_qpos[i])

    key = keyframe

    # ---- (de)serialisation: compiled tables travel as one .npz ---------------------
    def save(self, path) -> None:
        arrays = {f: getattr(self, f) for f in _ARRAY_FIELDS}
        for f in _NAME_FIELDS:
            arrays[f] = np.array(getattr(self, f), dtype="U")
        arrays["_scalars"] = np.array(
            [self.nq, self.nv, self.nbody, self.njnt, self.ngeom, self.nsite, self.nmesh,
             self.nkey, int(self.disable_contact), int(self.disable_filterparent)],
            dtype=np.int64,
        )
        arrays["_timestep"] = np.array([self.timestep])
        arrays["_name"] = np.array([self.name], dtype="U")
        np.savez_compressed(path, **arrays)

    @classmethod
    def load(cls, path) -> "Model":
        if isinstance(path, (bytes, bytearray)):
            path = io.BytesIO(path)
        with np.load(path, allow_pickle=False) as z:
            m = cls()
            for f in _ARRAY_FIELDS:
                setattr(m, f, z[f].copy())
            for f in _NAME_FIELDS:
                setattr(m, f, [str(s) for s in z[f]])
            s = z["_scalars"]
            (m.nq, m.nv, m.nbody, m.njnt, m.ngeom, m.nsite, m.nmesh, m.nkey) = (
                int(v) for v in s[:8]
            )
            m.disable_contact, m.disable_filterparent = bool(s[8]), bool(s[9])
            m.timestep = float(z["_timestep"][0])
            m.name = str(z["_name"][0])
        return m

    @classmethod
    def from_mjmodel(cls, mj) -> "Model":
        """Duck-typed conversion from a real ``mujoco.MjModel`` (same field names).

        Mesh geoms use MuJoCo's own (re-centred) ``mesh_vert`` restricted to the hull
        vertices listed in ``mesh_graph``; the world-space hull is identical.
        """
        import mujoco  # noqa: F401  (only reachable when MuJoCo exists)

        m = cls()
        for f in ("nq", "nv", "nbody", "njnt", "ngeom", "nsite", "nmesh", "nkey"):
            setattr(m, f, int(getattr(mj, f)))
        for f in _ARRAY_FIELDS:
            if f in ("mesh_vert", "mesh_vertadr", "mesh_vertnum", "exclude_signature", "mesh_graphadr", "mesh_graph"):
                continue
            setattr(m, f, np.array(getattr(mj, f)).copy())
        m.jnt_limited = m.jnt_limited.astype(np.int32)
        # hull vertex subsets
        verts, adr, num, graphs, gadr = [], [], [], [], []
        for i in range(m.nmesh):
            va, vn = int(mj.mesh_vertadr[i]), int(mj.mesh_vertnum[i])
            v = np.array(mj.mesh_vert[va : va + vn], dtype=np.float64)
            ga = int(mj.mesh_graphadr[i])
            if ga >= 0:
                g = np.array(mj.mesh_graph[ga:], dtype=np.int32)
                nhv, nhf = int(g[0]), int(g[1])
                idx = g[2 + nhv : 2 + 2 * nhv]  # vert_globalid
                v = v[idx]
                size = 2 + 3 * nhv + 6 * nhf
                gg = g[:size].copy()
                gg[2 + nhv : 2 + 2 * nhv] = np.arange(nhv)  # mesh_vert now holds hull vertices only, in local order
                gadr.append(sum(len(x) for x in graphs))
                graphs.append(gg)
            else:
                gadr.append(-1)
            adr.append(sum(num))
            num.append(len(v))
            verts.append(v)
        m.mesh_graphadr = np.array(gadr, dtype=np.int32)
        m.mesh_graph = np.concatenate(graphs).astype(np.int32) if graphs else np.zeros(0, np.int32)
        m.mesh_vert = np.concatenate(verts) if verts else np.zeros((0, 3))
        m.mesh_vertadr = np.array(adr, dtype=np.int32)
        m.mesh_vertnum = np.array(num, dtype=np.int32)
        m.exclude_signature = np.array(mj.exclude_signature, dtype=np.int64)
        m.body_names = [mj.body(i).name for i in range(m.nbody)]
        m.jnt_names = [mj.joint(i).name for i in range(m.njnt)]
        m.geom_names = [mj.geom(i).name for i in range(m.ngeom)]
        m.site_names = [mj.site(i).name for i in range(m.nsite)]
        m.mesh_names = [mj.mesh(i).name for i in range(m.nmesh)]
        m.key_names = [mj.key(i).name for i in range(m.nkey)]
        m.disable_contact = bool(mj.opt.disableflags & 16)
        m.disable_filterparent = bool(mj.opt.disableflags & 512)
        m.timestep = float(mj.opt.timestep)
        return m
