"""MJCF-subset compiler: XML -> :class:`mjpl_b200.model.Model`, with no MuJoCo installed.

The reference obtains its model with ``mujoco.MjModel.from_xml_path`` (e.g.
``examples/benchmark.py:28``, ``test/test_collision_constraint.py:18``).  MuJoCo is absent
from this image, so this module compiles the MJCF subset the in-scope models use (census in
SURVEY.md Appendix B) into the constant tables the validity path needs:

* ``<include file>`` (textual splice of the included root's children);
* ``<compiler angle meshdir autolimits>``; nested ``<default class>`` + ``childclass``;
* bodies (DFS pre-order ids, ``pos``/``quat``/``euler``), hinge/slide (and, for the index
  helpers only, free/ball) joints, geoms incl. ``fromto``, sites, ``<mesh>`` assets,
  ``<contact><exclude>``, ``<keyframe>``;
* binary STL / OBJ ``v`` lines -> convex hull (``scipy.spatial.ConvexHull``) for every mesh
  that a colliding geom references; visual-only meshes are never opened.

Ordering follows MuJoCo's compiler: bodies depth-first in document order; joints, geoms and
sites grouped by body id, document order within a body.
"""

from __future__ import annotations

import copy
import struct
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

from .model import (
    GEOM_BOX,
    GEOM_CAPSULE,
    GEOM_CYLINDER,
    GEOM_ELLIPSOID,
    GEOM_MESH,
    GEOM_PLANE,
    GEOM_SPHERE,
    GEOM_TYPE_NAMES,
    JNT_DOF_WIDTH,
    JNT_FREE,
    JNT_HINGE,
    JNT_QPOS_WIDTH,
    JNT_SLIDE,
    JNT_TYPE_NAMES,
    Model,
)


# --------------------------------------------------------------------- small math
def _floats(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and len(v) != n:
        raise ValueError(f"expected {n} numbers, got '{s}'")
    return v


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ]
    )


def _normalize_quat(q):
    n = np.linalg.norm(q)
    if n < 1e-15:
        return np.array([1.0, 0.0, 0.0, 0.0])
    return q / n


def _z2quat(vec):
    """Quaternion rotating (0,0,1) onto ``vec`` (MuJoCo ``mjuu_z2quat``)."""
    v = vec / np.linalg.norm(vec)
    axis = np.cross([0.0, 0.0, 1.0], v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0.0, 0.0])
    else:
        axis = axis / s
    ang = np.arctan2(s, v[2])
    q = np.array([np.cos(ang / 2), *(axis * np.sin(ang / 2))])
    return _normalize_quat(q)


def _euler2quat(e, seq="xyz"):
    q = np.array([1.0, 0.0, 0.0, 0.0])
    for ang, ax in zip(e, seq):
        r = np.zeros(4)
        r[0] = np.cos(ang / 2)
        r[1 + "xyz".index(ax.lower())] = np.sin(ang / 2)
        # lower case: intrinsic (post-multiply); upper case: extrinsic (pre-multiply)
        q = quat_mul(q, r) if ax.islower() else quat_mul(r, q)
    return q


# --------------------------------------------------------------------- mesh files
def load_mesh_vertices(path: Path) -> np.ndarray:
    """Vertex positions of a binary STL or an OBJ file, as float64 (n,3)."""
    path = Path(path)
    suf = path.suffix.lower()
    if suf == ".stl":
        raw = path.read_bytes()
        (ntri,) = struct.unpack_from("<I", raw, 80)
        if 84 + 50 * ntri != len(raw):
            raise ValueError(f"{path}: not a binary STL")
        rec = np.frombuffer(raw, dtype=np.uint8, count=50 * ntri, offset=84).reshape(ntri, 50)
        tri = rec[:, 12:48].copy().view("<f4").reshape(ntri * 3, 3)
        return tri.astype(np.float64)
    if suf == ".obj":
        vs = []
        with open(path, "r") as f:
            for line in f:
                if line.startswith("v "):
                    p = line.split()
                    vs.append((float(p[1]), float(p[2]), float(p[3])))
        return np.array(vs, dtype=np.float64).reshape(-1, 3)
    raise ValueError(f"unsupported mesh format: {path}")


def convex_hull_vertices(v: np.ndarray) -> np.ndarray:
    """Hull vertex subset (MuJoCo runs qhull on collision meshes; support = hull vertices)."""
    return convex_hull(v)[0]


def convex_hull(v: np.ndarray):
    """-> (hull vertices (n,3), hull graph in MuJoCo's ``mesh_graph`` layout or None).

    ``mesh_graph`` per mesh: ``nvert, nface, vert_edgeadr[nvert], vert_globalid[nvert],
    edge_localid[nvert + 3*nface]`` (for every hull vertex the list of its neighbours, as local
    ids, terminated by -1) ``, face_globalid[3*nface]``.  The engine uses the neighbour lists for
    hill-climbing support queries.
    """
    from scipy.spatial import ConvexHull

    v = np.unique(np.asarray(v, dtype=np.float64), axis=0)
    if len(v) < 4:
        return v, None
    hull = ConvexHull(v)  # qhull 'Qt': triangulated facets
    keep = np.sort(hull.vertices)
    local = -np.ones(len(v), dtype=np.int64)
    local[keep] = np.arange(len(keep))
    faces = local[hull.simplices]
    nv, nf = len(keep), len(faces)
    nbr = [set() for _ in range(nv)]
    for a, b, c in faces:
        nbr[a].update((b, c)); nbr[b].update((a, c)); nbr[c].update((a, b))
    edge_adr, edges = [], []
    for i in range(nv):
        edge_adr.append(len(edges))
        edges.extend(sorted(nbr[i]))
        edges.append(-1)
    assert len(edges) == nv + 3 * nf  # Euler: 2E = 3F for a triangulated closed surface
    graph = np.concatenate([[nv, nf], edge_adr, np.arange(nv), edges, faces.reshape(-1)]).astype(np.int32)
    return v[keep], graph


# --------------------------------------------------------------------- XML handling
def _expand_includes(root: ET.Element, base: Path) -> None:
    def rec(parent):
        out = []
        for child in list(parent):
            if child.tag == "include":
                inc = ET.parse(base / child.attrib["file"]).getroot()
                _expand_includes(inc, base)
                out.extend(list(inc))
            else:
                rec(child)
                out.append(child)
        parent[:] = out

    rec(root)


class _Defaults:
    """Default-class tree: class name -> {element tag -> attribute dict} (inherited)."""

    def __init__(self):
        self.classes: dict[str, dict[str, dict[str, str]]] = {"main": {}}

    def parse(self, elem: ET.Element, parent: str | None) -> None:
        name = elem.attrib.get("class", "main" if parent is None else None)
        if name is None:
            raise ValueError("nested <default> needs a class name")
        merged = copy.deepcopy(self.classes[parent]) if parent is not None else {}
        if name == "main" and "main" in self.classes:
            merged = copy.deepcopy(self.classes["main"])
        for child in elem:
            if child.tag == "default":
                continue
            merged.setdefault(child.tag, {}).update(child.attrib)
        self.classes[name] = merged
        for child in elem:
            if child.tag == "default":
                self.parse(child, name)

    def resolve(self, tag: str, elem: ET.Element, childclass: str | None) -> dict[str, str]:
        cls = elem.attrib.get("class", childclass or "main")
        if cls not in self.classes:
            raise ValueError(f"unknown default class '{cls}'")
        attrs = dict(self.classes[cls].get(tag, {}))
        attrs.update(elem.attrib)
        return attrs


def _orientation(a: dict[str, str], degree: bool, eulerseq: str) -> np.ndarray:
    if "quat" in a:
        return _normalize_quat(_floats(a["quat"], 4))
    if "euler" in a:
        e = _floats(a["euler"], 3)
        return _euler2quat(np.deg2rad(e) if degree else e, eulerseq)
    if "axisangle" in a:
        aa = _floats(a["axisangle"], 4)
        ang = np.deg2rad(aa[3]) if degree else aa[3]
        ax = aa[:3] / np.linalg.norm(aa[:3])
        return np.array([np.cos(ang / 2), *(ax * np.sin(ang / 2))])
    if "zaxis" in a:
        return _z2quat(_floats(a["zaxis"], 3))
    if "xyaxes" in a:
        xy = _floats(a["xyaxes"], 6)
        x = xy[:3] / np.linalg.norm(xy[:3])
        y = xy[3:] - x * np.dot(x, xy[3:])
        y /= np.linalg.norm(y)
        z = np.cross(x, y)
        return _mat2quat(np.stack([x, y, z], axis=1))
    return np.array([1.0, 0.0, 0.0, 0.0])


def _mat2quat(R):
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
    return _normalize_quat(np.array(q))


def from_xml_path(path) -> Model:
    """Compile an MJCF file (drop-in for ``mujoco.MjModel.from_xml_path``)."""
    path = Path(path)
    root = ET.parse(path).getroot()
    return _compile(root, path.parent, name=root.attrib.get("model", path.stem))


def from_xml_string(xml: str, base_dir=".") -> Model:
    root = ET.fromstring(xml)
    return _compile(root, Path(base_dir), name=root.attrib.get("model", ""))


def _compile(root: ET.Element, base: Path, name: str) -> Model:
    if root.tag != "mujoco":
        raise ValueError("root element must be <mujoco>")
    _expand_includes(root, base)

    # ---- compiler / option ------------------------------------------------------
    degree, autolimits, meshdir, eulerseq = True, True, "", "xyz"
    for c in root.findall("compiler"):
        if "angle" in c.attrib:
            degree = c.attrib["angle"] == "degree"
        if "autolimits" in c.attrib:
            autolimits = c.attrib["autolimits"] == "true"
        if "meshdir" in c.attrib:
            meshdir = c.attrib["meshdir"]
        if "assetdir" in c.attrib and "meshdir" not in c.attrib:
            meshdir = c.attrib["assetdir"]
        if "eulerseq" in c.attrib:
            eulerseq = c.attrib["eulerseq"]
    timestep, dis_contact, dis_filterparent = 0.002, False, False
    for o in root.findall("option"):
        if "timestep" in o.attrib:
            timestep = float(o.attrib["timestep"])
        for fl in o.findall("flag"):
            dis_contact |= fl.attrib.get("contact") == "disable"
            dis_filterparent |= fl.attrib.get("filterparent") == "disable"

    defaults = _Defaults()
    for d in root.findall("default"):
        defaults.parse(d, None)

    # ---- mesh assets (name -> file); vertices are loaded lazily ---------------------
    mesh_files: dict[str, tuple[Path, np.ndarray]] = {}
    mesh_order: list[str] = []
    for asset in root.findall("asset"):
        for me in asset.findall("mesh"):
            a = defaults.resolve("mesh", me, None)
            if "file" not in a:
                raise ValueError("<mesh> without file= is not supported")
            mname = a.get("name", Path(a["file"]).stem)
            scale = _floats(a["scale"], 3) if "scale" in a else np.ones(3)
            mesh_files[mname] = (base / meshdir / a["file"], scale)
            mesh_order.append(mname)

    # ---- kinematic tree ----------------------------------------------------------------
    bodies = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]))]
    joints, geoms, sites = [], [], []

    def add_children(elem: ET.Element, bid: int, childclass: str | None):
        for child in elem:
            if child.tag in ("joint", "freejoint"):
                if bid == 0:
                    raise ValueError("joint in worldbody")
                if child.tag == "freejoint":
                    a = dict(child.attrib)
                    a["type"] = "free"
                else:
                    a = defaults.resolve("joint", child, childclass)
                joints.append((bid, a))
            elif child.tag == "geom":
                geoms.append((bid, defaults.resolve("geom", child, childclass)))
            elif child.tag == "site":
                sites.append((bid, defaults.resolve("site", child, childclass)))
        for child in elem:
            if child.tag == "body":
                cc = child.attrib.get("childclass", childclass)
                new_id = len(bodies)
                bodies.append(
                    dict(
                        name=child.attrib.get("name", ""),
                        parent=bid,
                        pos=_floats(child.attrib.get("pos", "0 0 0"), 3),
                        quat=_orientation(child.attrib, degree, eulerseq),
                    )
                )
                add_children(child, new_id, cc)

    # MuJoCo assigns body ids depth-first: a body's subtree is numbered before its next
    # sibling.  add_children() above registers a body's own joints/geoms/sites first and then
    # recurses into child bodies in document order, which yields the same numbering.
    for wb in root.findall("worldbody"):
        add_children(wb, 0, wb.attrib.get("childclass"))

    m = Model(name=name, timestep=timestep)
    m.disable_contact, m.disable_filterparent = dis_contact, dis_filterparent
    nb = m.nbody = len(bodies)
    m.body_names = [b["name"] for b in bodies]
    m.body_parentid = np.array([b["parent"] for b in bodies], dtype=np.int32)
    m.body_pos = np.array([b["pos"] for b in bodies], dtype=np.float64)
    m.body_quat = np.array([b["quat"] for b in bodies], dtype=np.float64)

    # ---- joints (grouped by body, document order inside a body) ---------------------
    joints.sort(key=lambda t: t[0])  # stable
    m.njnt = len(joints)
    jt, jq, jd, jb, jp, ja, jr, jl, q0, jn = [], [], [], [], [], [], [], [], [], []
    nq = nv = 0
    for bid, a in joints:
        t = JNT_TYPE_NAMES[a.get("type", "hinge")]
        jt.append(t)
        jq.append(nq)
        jd.append(nv)
        jb.append(bid)
        jn.append(a.get("name", ""))
        jp.append(_floats(a.get("pos", "0 0 0"), 3))
        ax = _floats(a.get("axis", "0 0 1"), 3)
        ja.append(ax / np.linalg.norm(ax) if t in (JNT_HINGE, JNT_SLIDE) else np.array([0, 0, 1.0]))
        rng = _floats(a["range"], 2) if "range" in a else np.zeros(2)
        if t == JNT_HINGE and degree:
            rng = np.deg2rad(rng)
        lim = a.get("limited", "auto")
        limited = (lim == "true") or (lim == "auto" and autolimits and "range" in a)
        jr.append(rng)
        jl.append(int(limited))
        ref = float(a.get("ref", "0"))
        if t == JNT_HINGE and degree:
            ref = np.deg2rad(ref)
        if t == JNT_FREE:
            q0.extend([*bodies[bid]["pos"], *bodies[bid]["quat"]])
        elif t in (JNT_HINGE, JNT_SLIDE):
            q0.append(ref)
        else:
            q0.extend([1.0, 0.0, 0.0, 0.0])
        nq += JNT_QPOS_WIDTH[t]
        nv += JNT_DOF_WIDTH[t]
    m.nq, m.nv = nq, nv
    m.jnt_type = np.array(jt, dtype=np.int32)
    m.jnt_qposadr = np.array(jq, dtype=np.int32)
    m.jnt_dofadr = np.array(jd, dtype=np.int32)
    m.jnt_bodyid = np.array(jb, dtype=np.int32)
    m.jnt_pos = np.array(jp, dtype=np.float64).reshape(-1, 3)
    m.jnt_axis = np.array(ja, dtype=np.float64).reshape(-1, 3)
    m.jnt_range = np.array(jr, dtype=np.float64).reshape(-1, 2)
    m.jnt_limited = np.array(jl, dtype=np.int32)
    m.qpos0 = np.array(q0, dtype=np.float64)
    m.jnt_names = jn
    m.body_jntadr = np.full(nb, -1, dtype=np.int32)
    m.body_jntnum = np.zeros(nb, dtype=np.int32)
    for j, b in enumerate(jb):
        if m.body_jntadr[b] < 0:
            m.body_jntadr[b] = j
        m.body_jntnum[b] += 1
    # weld id: a body without joints is welded to its parent's weld body
    m.body_weldid = np.zeros(nb, dtype=np.int32)
    for b in range(1, nb):
        m.body_weldid[b] = b if m.body_jntnum[b] > 0 else m.body_weldid[m.body_parentid[b]]

    # ---- geoms ---------------------------------------------------------------------------
    geoms.sort(key=lambda t: t[0])
    m.ngeom = len(geoms)
    used_meshes: dict[str, int] = {}
    g_type, g_body, g_ct, g_ca, g_size, g_pos, g_quat, g_margin, g_gap, g_data, g_name = (
        [] for _ in range(11)
    )
    for bid, a in geoms:
        if "mesh" in a and "type" not in a:
            t = GEOM_MESH
        else:
            t = GEOM_TYPE_NAMES[a.get("type", "sphere")]
        size = np.zeros(3)
        if "size" in a:
            s = _floats(a["size"])
            size[: len(s)] = s
        pos = _floats(a.get("pos", "0 0 0"), 3)
        quat = _orientation(a, degree, eulerseq)
        if "fromto" in a:
            if t not in (GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX, GEOM_ELLIPSOID):
                raise ValueError("fromto only valid for capsule/cylinder/box/ellipsoid")
            ft = _floats(a["fromto"], 6)
            vec = ft[:3] - ft[3:]
            half = np.linalg.norm(vec) / 2
            if t in (GEOM_CAPSULE, GEOM_CYLINDER):
                size[1] = half
            else:
                size[2] = half
                size[1] = size[0]
            pos = (ft[:3] + ft[3:]) / 2
            quat = _z2quat(vec)
        ct, ca = int(a.get("contype", "1")), int(a.get("conaffinity", "1"))
        data = -1
        if t == GEOM_MESH:
            mname = a["mesh"]
            if mname not in mesh_files:
                raise ValueError(f"unknown mesh '{mname}'")
            data = mesh_order.index(mname)
            if ct or ca:
                used_meshes[mname] = data
        g_type.append(t)
        g_body.append(bid)
        g_ct.append(ct)
        g_ca.append(ca)
        g_size.append(size)
        g_pos.append(pos)
        g_quat.append(quat)
        g_margin.append(float(a.get("margin", "0")))
        g_gap.append(float(a.get("gap", "0")))
        g_data.append(data)
        g_name.append(a.get("name", ""))
    m.geom_type = np.array(g_type, dtype=np.int32)
    m.geom_bodyid = np.array(g_body, dtype=np.int32)
    m.geom_contype = np.array(g_ct, dtype=np.int32)
    m.geom_conaffinity = np.array(g_ca, dtype=np.int32)
    m.geom_size = np.array(g_size, dtype=np.float64).reshape(-1, 3)
    m.geom_pos = np.array(g_pos, dtype=np.float64).reshape(-1, 3)
    m.geom_quat = np.array(g_quat, dtype=np.float64).reshape(-1, 4)
    m.geom_margin = np.array(g_margin, dtype=np.float64)
    m.geom_gap = np.array(g_gap, dtype=np.float64)
    m.geom_dataid = np.array(g_data, dtype=np.int32)
    m.geom_names = g_name
    m.body_geomadr = np.full(nb, -1, dtype=np.int32)
    m.body_geomnum = np.zeros(nb, dtype=np.int32)
    for g, b in enumerate(g_body):
        if m.body_geomadr[b] < 0:
            m.body_geomadr[b] = g
        m.body_geomnum[b] += 1

    # ---- meshes: hull vertices for colliding meshes only ---------------------------------
    m.nmesh = len(mesh_order)
    m.mesh_names = list(mesh_order)
    m.mesh_vertadr = np.zeros(m.nmesh, dtype=np.int32)
    m.mesh_vertnum = np.zeros(m.nmesh, dtype=np.int32)
    m.mesh_graphadr = np.full(m.nmesh, -1, dtype=np.int32)
    verts, graphs = [], []
    n = ng = 0
    for i, mname in enumerate(mesh_order):
        m.mesh_vertadr[i] = n
        if mname in used_meshes:
            f, scale = mesh_files[mname]
            hv, graph = convex_hull(load_mesh_vertices(f) * scale)
            verts.append(hv)
            m.mesh_vertnum[i] = len(hv)
            n += len(hv)
            if graph is not None:
                m.mesh_graphadr[i] = ng
                graphs.append(graph)
                ng += len(graph)
    m.mesh_vert = np.concatenate(verts) if verts else np.zeros((0, 3))
    m.mesh_graph = np.concatenate(graphs).astype(np.int32) if graphs else np.zeros(0, np.int32)

    # ---- rbound (conservative, geom-frame origin) ---------------------------------------------
    rb = np.zeros(m.ngeom)
    for g in range(m.ngeom):
        t, s = m.geom_type[g], m.geom_size[g]
        if t == GEOM_SPHERE:
            rb[g] = s[0]
        elif t == GEOM_CAPSULE:
            rb[g] = s[0] + s[1]
        elif t == GEOM_CYLINDER:
            rb[g] = np.hypot(s[0], s[1])
        elif t == GEOM_BOX:
            rb[g] = np.linalg.norm(s)
        elif t == GEOM_ELLIPSOID:
            rb[g] = np.max(s)
        elif t == GEOM_MESH and m.mesh_vertnum[m.geom_dataid[g]] > 0:
            d = m.geom_dataid[g]
            v = m.mesh_vert[m.mesh_vertadr[d] : m.mesh_vertadr[d] + m.mesh_vertnum[d]]
            rb[g] = np.max(np.linalg.norm(v, axis=1))
    m.geom_rbound = rb

    # ---- sites -------------------------------------------------------------------------------------
    sites.sort(key=lambda t: t[0])
    m.nsite = len(sites)
    m.site_bodyid = np.array([b for b, _ in sites], dtype=np.int32)
    m.site_pos = np.array([_floats(a.get("pos", "0 0 0"), 3) for _, a in sites]).reshape(-1, 3)
    m.site_quat = np.array([_orientation(a, degree, eulerseq) for _, a in sites]).reshape(-1, 4)
    m.site_names = [a.get("name", "") for _, a in sites]

    # ---- contact excludes ------------------------------------------------------------------------
    sig = []
    for c in root.findall("contact"):
        for e in c.findall("exclude"):
            b1 = m.body(e.attrib["body1"]).id
            b2 = m.body(e.attrib["body2"]).id
            lo, hi = min(b1, b2), max(b1, b2)
            sig.append((lo << 16) + hi)
        if c.findall("pair"):
            raise ValueError("<contact><pair> is not supported")
    m.exclude_signature = np.array(sig, dtype=np.int64)

    # ---- keyframes -------------------------------------------------------------------------------
    keys, knames = [], []
    for kf in root.findall("keyframe"):
        for k in kf.findall("key"):
            q = _floats(k.attrib["qpos"]) if "qpos" in k.attrib else m.qpos0.copy()
            if len(q) != m.nq:
                raise ValueError("keyframe qpos has wrong size")
            keys.append(q)
            knames.append(k.attrib.get("name", ""))
    m.nkey = len(keys)
    m.key_qpos = np.array(keys, dtype=np.float64).reshape(-1, m.nq)
    m.key_names = knames
    return m
