"""MJCF-subset compiler for the Open Duck Mini V2 scenes.

The reference loads its model with ``mujoco.MjModel.from_xml_string`` and
``mjx.put_model`` (reference open_duck_mini_v2/base.py:53-61).  Neither MuJoCo nor
MJX exists in this image, so this module compiles exactly the MJCF features the
three scenes use into a :class:`CompiledModel` -- the numpy
mirror of ``OduckModel`` in include/oduck.h:

* ``<include>``, nested ``<default class=...>`` with ``childclass`` inheritance
* body tree, ``<inertial fullinertia=...>`` -> principal inertia + iquat
* free / hinge joints, ``range`` with MuJoCo's default ``autolimits``
* ``<position>`` actuators with ``inheritrange``
* sites, the ``home`` keyframe, the two convex foot meshes (binary STL -> hull)
* compile-time constants MuJoCo derives at qpos0: ``dof_invweight0``,
  ``body_invweight0`` and ``stat.meaninertia``

Ids follow MuJoCo's depth-first numbering so that the reference's hard-coded
``FLOOR_GEOM_ID = 0`` / ``TORSO_BODY_ID = 1`` (common/randomize.py:22-23) keep
their (quirky) meaning.
"""
from __future__ import annotations

import copy
import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

JNT_FREE = 0
JNT_HINGE = 3

MAX_BODY, MAX_JNT, MAX_NQ, MAX_NV, MAX_NU, MAX_SITE, MAX_VERT, MAX_FACE = 20, 28, 36, 32, 16, 8, 32, 64


# --------------------------------------------------------------------------- math helpers
def _vec(s: Optional[str], default):
    if s is None:
        return np.array(default, dtype=np.float64)
    return np.array([float(x) for x in s.split()], dtype=np.float64)


def quat_to_mat(q):
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


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


def mat_to_quat(m):
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    return q / np.linalg.norm(q)


# --------------------------------------------------------------------------- model container
@dataclass
class CompiledModel:
    """Numpy mirror of ``OduckModel`` (include/oduck.h) plus name tables."""

    nbody: int = 0
    njnt: int = 0
    nq: int = 0
    nv: int = 0
    nu: int = 0
    nsite: int = 0
    body_names: List[str] = field(default_factory=list)
    joint_names: List[str] = field(default_factory=list)
    actuator_names: List[str] = field(default_factory=list)
    site_names: List[str] = field(default_factory=list)
    geom_names: List[str] = field(default_factory=list)  # every geom in MuJoCo id order (visual ones included)
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)
    xml_path: str = ""

    def __getattr__(self, k):
        arrays = self.__dict__.get("arrays", {})
        if k in arrays:
            return arrays[k]
        raise AttributeError(k)

    # -- persistence (the shipped blobs under open_duck_playground_b200/data/)
    def save(self, path: str) -> None:
        meta = dict(
            nbody=self.nbody, njnt=self.njnt, nq=self.nq, nv=self.nv, nu=self.nu, nsite=self.nsite,
            body_names=self.body_names, joint_names=self.joint_names, actuator_names=self.actuator_names,
            site_names=self.site_names, geom_names=self.geom_names, xml_path=os.path.basename(self.xml_path),
        )
        np.savez_compressed(path, __meta__=np.array(repr(meta)), **self.arrays)

    @staticmethod
    def load(path: str) -> "CompiledModel":
        import ast

        z = np.load(path, allow_pickle=False)
        meta = ast.literal_eval(str(z["__meta__"]))
        m = CompiledModel(**meta)
        m.arrays = {k: z[k] for k in z.files if k != "__meta__"}
        return m

    # -- MuJoCo-like name lookups used by the env shim
    def body_id(self, name):
        return self.body_names.index(name)

    def joint_id(self, name):
        return self.joint_names.index(name) if name in self.joint_names else -1

    def site_id(self, name):
        return self.site_names.index(name)

    def geom_id(self, name):
        return self.geom_names.index(name)


# --------------------------------------------------------------------------- XML front end
def _load_xml(path: str) -> ET.Element:
    root = ET.parse(path).getroot()
    base = os.path.dirname(path)

    def expand(elem):
        out = []
        for ch in list(elem):
            if ch.tag == "include":
                sub = _load_xml(os.path.join(base, ch.attrib["file"]))
                out.extend(list(sub))
            else:
                expand(ch)
                out.append(ch)
        elem[:] = out

    expand(root)
    return root


class _Defaults:
    """Default-class tree: class name -> {tag -> attrib dict}, inherited from the parent class."""

    def __init__(self):
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {"main": {}}

    def add(self, elem: ET.Element, parent: str):
        name = elem.attrib.get("class", "main") if parent else "main"
        if name not in self.classes:
            self.classes[name] = copy.deepcopy(self.classes[parent]) if parent else {}
        cur = self.classes[name]
        for ch in elem:
            if ch.tag == "default":
                continue
            cur.setdefault(ch.tag, {}).update(ch.attrib)
        for ch in elem:
            if ch.tag == "default":
                self.add(ch, name)

    def resolve(self, tag: str, elem: ET.Element, childclass: Optional[str]) -> Dict[str, str]:
        cls = elem.attrib.get("class", childclass or "main")
        if cls not in self.classes:
            raise KeyError(f"unknown default class {cls!r}")
        out = dict(self.classes[cls].get(tag, {}))
        out.update(elem.attrib)
        return out


def _read_png_gray(path: str) -> np.ndarray:
    """Minimal PNG decoder (8-bit, non-interlaced, grey / grey+alpha / RGB / RGBA) -> uint8 [height, width].
    Colour images are reduced to their RED channel, which is what MuJoCo's height-field loader gets from lodepng's LCT_GREY
    conversion (user_objects.cc mjCHField::LoadPNG) **[upstream-memory]**.  PIL is not available in this image."""
    import struct
    import zlib
    raw = open(path, "rb").read()
    if raw[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos < len(raw):
        ln, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + ln]
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
        pos += 12 + ln
    w, h, depth, ctype, _, _, interlace = hdr
    nch = {0: 1, 2: 3, 4: 2, 6: 4}.get(ctype)
    if depth != 8 or nch is None or interlace:
        raise ValueError(f"{path}: only 8-bit non-interlaced grey/RGB(A) PNGs are supported")
    data = zlib.decompress(b"".join(idat))
    stride = w * nch
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    for r in range(h):
        f = data[r * (stride + 1)]
        line = np.frombuffer(data, np.uint8, stride, r * (stride + 1) + 1).astype(np.int32)
        cur = np.zeros(stride, np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 255
        else:                                   # Sub / Average / Paeth depend on the pixel to the left: sequential per channel byte
            for i in range(stride):
                a = cur[i - nch] if i >= nch else 0
                b = prev[i]
                c = prev[i - nch] if i >= nch else 0
                if f == 1:
                    pr = a
                elif f == 3:
                    pr = (a + b) >> 1
                elif f == 4:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    pr = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                else:
                    raise ValueError(f"{path}: bad PNG filter {f}")
                cur[i] = (line[i] + pr) & 255
        out[r] = cur
        prev = cur
    return out.reshape(h, w, nch)[:, :, 0].copy()


def load_hfield_png(path: str) -> np.ndarray:
    """Height-field elevation data as MuJoCo stores it: image rows flipped (row 0 = minimum y), normalised to [0, 1] by the
    data range (mjCHField: ``data = (data - min) / (max - min)``) **[upstream-memory]**.  float32 [nrow, ncol]."""
    img = _read_png_gray(path).astype(np.float64)[::-1]
    lo, hi = img.min(), img.max()
    return ((img - lo) / (hi - lo) if hi > lo else np.zeros_like(img)).astype(np.float32)


def _read_stl(path: str) -> np.ndarray:
    b = open(path, "rb").read()
    n = struct.unpack("<I", b[80:84])[0]
    if len(b) != 84 + 50 * n:
        raise ValueError(f"{path}: not a binary STL")
    tris = np.zeros((n, 3, 3))
    for i in range(n):
        vals = struct.unpack("<12f", b[84 + 50 * i: 84 + 50 * i + 48])
        tris[i] = np.array(vals[3:]).reshape(3, 3)
    return tris


def _convex_hull(tris: np.ndarray):
    """Unique vertices in first-appearance order -> hull vertex subset (same order) + outward triangles."""
    from scipy.spatial import ConvexHull

    flat = tris.reshape(-1, 3).astype(np.float32).astype(np.float64)  # STL stores float32
    uniq: List[np.ndarray] = []
    for v in flat:
        if not any(np.array_equal(v, u) for u in uniq):
            uniq.append(v)
    uniq = np.array(uniq)
    hull = ConvexHull(uniq)
    keep = sorted(set(hull.vertices.tolist()))
    remap = {old: new for new, old in enumerate(keep)}
    verts = uniq[keep]
    centre = verts.mean(axis=0)
    faces = []
    for simplex, eq in zip(hull.simplices, hull.equations):
        a, b, c = (remap[int(s)] for s in simplex)
        n = np.cross(verts[b] - verts[a], verts[c] - verts[a])
        if np.dot(n, verts[a] - centre) < 0:
            b, c = c, b
        faces.append((a, b, c))
    return verts, np.array(faces, dtype=np.int32)


def _hull_features(verts: np.ndarray, faces: np.ndarray):
    """Polygon faces (coplanar hull triangles merged, vertices counter-clockwise seen from outside) and the edges between
    two different polygon faces, with the ids of both -- the face / edge features convex-convex SAT works on."""
    n = np.cross(verts[faces[:, 1]] - verts[faces[:, 0]], verts[faces[:, 2]] - verts[faces[:, 0]])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    planes = []
    for i, nn in enumerate(n):
        for u in planes:
            if np.dot(u["n"], nn) > 0.99999 and abs(np.dot(u["n"], verts[faces[i, 0]]) - u["d"]) < 1e-6:
                u["tri"].append(i)
                break
        else:
            planes.append(dict(n=nn, d=float(np.dot(nn, verts[faces[i, 0]])), tri=[i]))
    tri_plane = {t: k for k, u in enumerate(planes) for t in u["tri"]}
    for u in planes:
        vs = sorted(set(faces[u["tri"]].ravel().tolist()))
        c = verts[vs].mean(axis=0)
        e1 = verts[vs[0]] - c
        e1 /= np.linalg.norm(e1)
        e2 = np.cross(u["n"], e1)
        ang = [np.arctan2(np.dot(verts[v] - c, e2), np.dot(verts[v] - c, e1)) for v in vs]
        u["loop"] = [v for _, v in sorted(zip(ang, vs))]
    edge_planes: Dict[tuple, set] = {}
    for t, f in enumerate(faces):
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            edge_planes.setdefault((int(min(a, b)), int(max(a, b))), set()).add(tri_plane[t])
    edges = [(e, sorted(ps)) for e, ps in sorted(edge_planes.items()) if len(ps) == 2]
    return planes, edges


# --------------------------------------------------------------------------- compile
def compile_mjcf(xml_path: str, timestep: float = 0.002) -> CompiledModel:
    """Compile a scene XML.  ``timestep`` mirrors ``mj_model.opt.timestep = sim_dt`` (base.py:56)."""
    root = _load_xml(xml_path)
    xml_dir = os.path.dirname(xml_path)
    compiler = {}
    for c in root.findall("compiler"):
        compiler.update(c.attrib)
    if compiler.get("angle", "degree") != "radian":
        raise ValueError("only angle=radian is supported")
    autolimits = compiler.get("autolimits", "true") == "true"
    meshdir = os.path.join(xml_dir, compiler.get("meshdir", ""))

    defaults = _Defaults()
    for d in root.findall("default"):
        defaults.add(d, "")

    option = {}
    flags = {}
    for o in root.findall("option"):
        option.update(o.attrib)
        for f in o.findall("flag"):
            flags.update(f.attrib)
    if flags.get("eulerdamp", "enable") != "disable":
        raise ValueError("only eulerdamp=disable is supported (the reference scenes set it)")

    mesh_files = {}
    hfield = None
    for a in root.findall("asset"):
        for m in a.findall("mesh"):
            name = m.attrib.get("name", os.path.splitext(os.path.basename(m.attrib["file"]))[0])
            mesh_files[name] = os.path.join(meshdir, m.attrib["file"])
        for h in a.findall("hfield"):
            hfield = dict(h.attrib)

    bodies = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), ipos=np.zeros(3),
                   iquat=np.array([1.0, 0, 0, 0]), mass=0.0, inertia=np.zeros(3), joints=[])]
    joints, sites, geoms = [], [], []

    def walk(elem, parent_id, childclass):
        for b in elem.findall("body"):
            cc = b.attrib.get("childclass", childclass)
            bid = len(bodies)
            body = dict(name=b.attrib.get("name", f"body{bid}"), parent=parent_id, pos=_vec(b.attrib.get("pos"), [0, 0, 0]),
                        quat=_vec(b.attrib.get("quat"), [1, 0, 0, 0]), joints=[], mass=0.0, inertia=np.zeros(3),
                        ipos=np.zeros(3), iquat=np.array([1.0, 0, 0, 0]))
            body["quat"] = body["quat"] / np.linalg.norm(body["quat"])
            bodies.append(body)
            inert = b.find("inertial")
            if inert is not None:
                body["mass"] = float(inert.attrib["mass"])
                body["ipos"] = _vec(inert.attrib.get("pos"), [0, 0, 0])
                if "fullinertia" in inert.attrib:
                    f = _vec(inert.attrib["fullinertia"], None)
                    full = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
                    w, v = np.linalg.eigh(full)
                    order = np.argsort(-w)  # MuJoCo sorts principal moments in decreasing order
                    w, v = w[order], v[:, order]
                    if np.linalg.det(v) < 0:
                        v[:, 2] = -v[:, 2]
                    body["inertia"] = w
                    body["iquat"] = mat_to_quat(v)
                else:
                    body["inertia"] = _vec(inert.attrib.get("diaginertia"), [0, 0, 0])
                    body["iquat"] = _vec(inert.attrib.get("quat"), [1, 0, 0, 0])
            for j in list(b):
                if j.tag == "freejoint":
                    joints.append(dict(name=j.attrib.get("name", ""), type=JNT_FREE, body=bid, pos=np.zeros(3),
                                       axis=np.array([0, 0, 1.0]), range=np.zeros(2), limited=0, damping=0.0,
                                       armature=0.0, frictionloss=0.0))
                    body["joints"].append(len(joints) - 1)
                elif j.tag == "joint":
                    a = defaults.resolve("joint", j, cc)
                    jtype = a.get("type", "hinge")
                    if jtype not in ("hinge", "free"):
                        raise ValueError(f"unsupported joint type {jtype}")
                    rng = _vec(a.get("range"), [0, 0])
                    lim = a.get("limited", "auto")
                    limited = (lim == "true") or (lim == "auto" and autolimits and "range" in a)
                    axis = _vec(a.get("axis"), [0, 0, 1])
                    joints.append(dict(name=a.get("name", ""), type=JNT_HINGE if jtype == "hinge" else JNT_FREE, body=bid,
                                       pos=_vec(a.get("pos"), [0, 0, 0]), axis=axis / np.linalg.norm(axis), range=rng,
                                       limited=int(limited), damping=float(a.get("damping", 0)),
                                       armature=float(a.get("armature", 0)), frictionloss=float(a.get("frictionloss", 0))))
                    body["joints"].append(len(joints) - 1)
                elif j.tag == "site":
                    a = defaults.resolve("site", j, cc)
                    q = _vec(a.get("quat"), [1, 0, 0, 0])
                    sites.append(dict(name=a.get("name", ""), body=bid, pos=_vec(a.get("pos"), [0, 0, 0]), quat=q / np.linalg.norm(q)))
                elif j.tag == "geom":
                    a = defaults.resolve("geom", j, cc)
                    q = _vec(a.get("quat"), [1, 0, 0, 0])
                    geoms.append(dict(name=a.get("name", ""), body=bid, type=a.get("type", "sphere"), mesh=a.get("mesh"),
                                      pos=_vec(a.get("pos"), [0, 0, 0]), quat=q / np.linalg.norm(q),
                                      contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
                                      priority=int(a.get("priority", 0)), condim=int(a.get("condim", 3)),
                                      friction=_vec(a.get("friction"), [1, 0.005, 0.0001]), hfield=a.get("hfield")))
            walk(b, bid, cc)

    for wb in root.findall("worldbody"):
        # geoms / sites directly under <worldbody> belong to the world body (none in these scenes)
        walk(wb, 0, None)

    m = CompiledModel(xml_path=xml_path)
    nbody, njnt = len(bodies), len(joints)
    if nbody > MAX_BODY or njnt > MAX_JNT:
        raise ValueError("model exceeds ODUCK_MAX_* limits")
    A = m.arrays
    A["body_parentid"] = np.zeros(MAX_BODY, np.int32)
    A["body_jntadr"] = np.full(MAX_BODY, -1, np.int32)
    A["body_jntnum"] = np.zeros(MAX_BODY, np.int32)
    A["body_dofadr"] = np.full(MAX_BODY, -1, np.int32)
    A["body_dofnum"] = np.zeros(MAX_BODY, np.int32)
    for k, shp in dict(body_pos=3, body_quat=4, body_ipos=3, body_iquat=4, body_inertia=3, body_invweight0=2).items():
        A[k] = np.zeros((MAX_BODY, shp))
    A["body_quat"][:, 0] = 1
    A["body_iquat"][:, 0] = 1
    A["body_mass"] = np.zeros(MAX_BODY)
    A["jnt_type"] = np.zeros(MAX_JNT, np.int32)
    A["jnt_qposadr"] = np.zeros(MAX_JNT, np.int32)
    A["jnt_dofadr"] = np.zeros(MAX_JNT, np.int32)
    A["jnt_bodyid"] = np.zeros(MAX_JNT, np.int32)
    A["jnt_limited"] = np.zeros(MAX_JNT, np.int32)
    A["jnt_pos"] = np.zeros((MAX_JNT, 3))
    A["jnt_axis"] = np.zeros((MAX_JNT, 3))
    A["jnt_range"] = np.zeros((MAX_JNT, 2))
    A["qpos0"] = np.zeros(MAX_NQ)
    A["dof_bodyid"] = np.zeros(MAX_NV, np.int32)
    A["dof_jntid"] = np.zeros(MAX_NV, np.int32)
    A["dof_parentid"] = np.full(MAX_NV, -1, np.int32)
    for k in ("dof_armature", "dof_damping", "dof_frictionloss", "dof_invweight0"):
        A[k] = np.zeros(MAX_NV)

    nq = nv = 0
    last_dof_of_body = {}
    for bid, b in enumerate(bodies):
        A["body_parentid"][bid] = b["parent"]
        A["body_pos"][bid], A["body_quat"][bid] = b["pos"], b["quat"]
        A["body_ipos"][bid], A["body_iquat"][bid] = b["ipos"], b["iquat"]
        A["body_mass"][bid], A["body_inertia"][bid] = b["mass"], b["inertia"]
        A["body_jntnum"][bid] = len(b["joints"])
        if b["joints"]:
            A["body_jntadr"][bid] = b["joints"][0]
            A["body_dofadr"][bid] = nv
        # dof parent = last dof of the nearest ancestor that has dofs
        anc = b["parent"]
        while anc > 0 and anc not in last_dof_of_body:
            anc = bodies[anc]["parent"]
        parent_dof = last_dof_of_body.get(anc, -1)
        for jid in b["joints"]:
            j = joints[jid]
            A["jnt_type"][jid], A["jnt_bodyid"][jid] = j["type"], bid
            A["jnt_qposadr"][jid], A["jnt_dofadr"][jid] = nq, nv
            A["jnt_limited"][jid] = j["limited"]
            A["jnt_pos"][jid], A["jnt_axis"][jid], A["jnt_range"][jid] = j["pos"], j["axis"], j["range"]
            nd = 6 if j["type"] == JNT_FREE else 1
            if j["type"] == JNT_FREE:
                A["qpos0"][nq:nq + 3] = b["pos"]
                A["qpos0"][nq + 3:nq + 7] = b["quat"]
                nq += 7
            else:
                nq += 1
            for k in range(nd):
                A["dof_bodyid"][nv], A["dof_jntid"][nv] = bid, jid
                A["dof_parentid"][nv] = parent_dof
                A["dof_armature"][nv], A["dof_damping"][nv] = j["armature"], j["damping"]
                A["dof_frictionloss"][nv] = j["frictionloss"]
                parent_dof = nv
                nv += 1
        if b["joints"]:
            A["body_dofnum"][bid] = nv - A["body_dofadr"][bid]
            last_dof_of_body[bid] = nv - 1
    if nq > MAX_NQ or nv > MAX_NV:
        raise ValueError("model exceeds ODUCK_MAX_NQ/NV")
    m.nbody, m.njnt, m.nq, m.nv = nbody, njnt, nq, nv
    m.body_names = [b["name"] for b in bodies]
    m.joint_names = [j["name"] for j in joints]
    m.geom_names = [g["name"] for g in geoms]

    # actuators ----------------------------------------------------------
    acts = []
    for sec in root.findall("actuator"):
        for a in sec:
            if a.tag != "position":
                raise ValueError(f"unsupported actuator <{a.tag}>")
            acts.append(defaults.resolve("position", a, None))
    nu = len(acts)
    m.nu = nu
    m.actuator_names = [a["name"] for a in acts]
    A["act_jntid"] = np.zeros(MAX_NU, np.int32)
    A["act_kp"], A["act_kv"] = np.zeros(MAX_NU), np.zeros(MAX_NU)
    A["act_ctrlrange"], A["act_forcerange"] = np.zeros((MAX_NU, 2)), np.zeros((MAX_NU, 2))
    for i, a in enumerate(acts):
        jid = m.joint_names.index(a["joint"])
        A["act_jntid"][i] = jid
        A["act_kp"][i], A["act_kv"][i] = float(a.get("kp", 1)), float(a.get("kv", 0))
        if "inheritrange" in a:
            r = A["jnt_range"][jid]
            mean, rad = 0.5 * (r[0] + r[1]), 0.5 * (r[1] - r[0]) * float(a["inheritrange"])
            A["act_ctrlrange"][i] = [mean - rad, mean + rad]
        else:
            A["act_ctrlrange"][i] = _vec(a.get("ctrlrange"), [-np.inf, np.inf])
        A["act_forcerange"][i] = _vec(a.get("forcerange"), [-np.inf, np.inf])

    # sites ----------------------------------------------------------------
    m.nsite = len(sites)
    m.site_names = [s["name"] for s in sites]
    A["site_bodyid"] = np.zeros(MAX_SITE, np.int32)
    A["site_pos"], A["site_quat"] = np.zeros((MAX_SITE, 3)), np.zeros((MAX_SITE, 4))
    A["site_quat"][:, 0] = 1
    for i, s in enumerate(sites):
        A["site_bodyid"][i], A["site_pos"][i], A["site_quat"][i] = s["body"], s["pos"], s["quat"]
    A["imu_site"] = np.array(m.site_id("imu"), np.int32)
    A["foot_site"] = np.array([m.site_id("left_foot"), m.site_id("right_foot")], np.int32)

    # collision geoms --------------------------------------------------------
    col = [g for g in geoms if g["contype"] or g["conaffinity"]]
    floor = [g for g in col if g["type"] in ("plane", "hfield")]
    feet = [g for g in col if g["type"] == "mesh"]
    if len(floor) != 1 or len(feet) != 2:
        raise ValueError("expected exactly one floor geom and two convex foot geoms")
    floor = floor[0]
    if np.any(bodies[floor["body"]]["pos"] != 0) or np.any(floor["pos"] != 0):
        raise ValueError("floor must sit at the world origin")
    A["floor_is_hfield"] = np.array(int(floor["type"] == "hfield"), np.int32)
    A["foot_body"] = np.array([g["body"] for g in feet], np.int32)
    A["foot_vert"] = np.zeros((2, MAX_VERT, 3))
    A["foot_face"] = np.zeros((MAX_FACE, 3), np.int32)
    for k, g in enumerate(feet):
        verts, faces = _convex_hull(_read_stl(mesh_files[g["mesh"]]))
        if len(verts) > MAX_VERT or len(faces) > MAX_FACE:
            raise ValueError("foot hull too large")
        R = quat_to_mat(g["quat"])
        A["foot_vert"][k, : len(verts)] = g["pos"] + verts @ R.T  # geom frame -> body frame
        A["foot_nvert"] = np.array(len(verts), np.int32)
        A["foot_nface"] = np.array(len(faces), np.int32)
        A["foot_face"][: len(faces)] = faces
        planes, edges = _hull_features(verts, faces)
        if len(planes) > 32 or len(edges) > 48 or max(len(u["loop"]) for u in planes) > 8:
            raise ValueError("foot hull has too many polygon faces / edges")
        if k == 0:
            A["foot_nplane"] = np.array(len(planes), np.int32)
            A["foot_plane_nvert"] = np.zeros(32, np.int32)
            A["foot_plane_vert"] = np.zeros((32, 8), np.int32)
            A["foot_plane_normal"] = np.zeros((2, 32, 3))
            A["foot_nedge"] = np.array(len(edges), np.int32)
            A["foot_edge_vert"] = np.zeros((48, 2), np.int32)
            A["foot_edge_plane"] = np.zeros((48, 2), np.int32)
            A["foot_center"] = np.zeros((2, 3))
            for q, u in enumerate(planes):
                A["foot_plane_nvert"][q] = len(u["loop"])
                A["foot_plane_vert"][q, : len(u["loop"])] = u["loop"]
            for q, (e, ps) in enumerate(edges):
                A["foot_edge_vert"][q], A["foot_edge_plane"][q] = e, ps
        for q, u in enumerate(planes):
            A["foot_plane_normal"][k, q] = R @ u["n"]
        body_v = A["foot_vert"][k, : len(verts)]
        A["foot_center"][k] = body_v.mean(axis=0)
        A["foot_radius"] = np.array(max(float(A["foot_radius"]) if "foot_radius" in A else 0.0, float(np.linalg.norm(body_v - body_v.mean(axis=0), axis=1).max())))
    # contact parameter mixing: higher priority wins, else max friction (MuJoCo mj_contactParam)
    f0 = feet[0]
    A["floor_friction"] = np.array(floor["friction"][0] if floor["priority"] > f0["priority"] else
                                   (f0["friction"][0] if f0["priority"] > floor["priority"] else max(floor["friction"][0], f0["friction"][0])))
    A["foot_friction"] = np.array(max(feet[0]["friction"][0], feet[1]["friction"][0]))
    ff = (feet[0]["contype"] & feet[1]["conaffinity"]) or (feet[1]["contype"] & feet[0]["conaffinity"])
    A["enable_foot_foot"] = np.array(int(bool(ff)), np.int32)
    if any(g["condim"] != 3 for g in (floor, *feet)):
        raise ValueError("only condim=3 contacts are supported")
    if hfield is not None and A["floor_is_hfield"]:
        A["hfield_size"] = _vec(hfield["size"], None)
        A["hfield_file"] = np.array(os.path.join(xml_dir, hfield["file"]))
        A["hfield_data"] = load_hfield_png(os.path.join(xml_dir, hfield["file"]))

    # options ----------------------------------------------------------------
    A["timestep"] = np.array(float(timestep))
    A["gravity"] = _vec(option.get("gravity"), [0, 0, -9.81])
    A["tolerance"] = np.array(float(option.get("tolerance", 1e-8)))
    A["ls_tolerance"] = np.array(float(option.get("ls_tolerance", 0.01)))
    A["impratio"] = np.array(float(option.get("impratio", 1.0)))
    A["iterations"] = np.array(int(option.get("iterations", 100)), np.int32)
    A["ls_iterations"] = np.array(int(option.get("ls_iterations", 50)), np.int32)
    A["solref"] = np.array([0.02, 1.0])
    A["solimp"] = np.array([0.9, 0.95, 0.001, 0.5, 2.0])

    # keyframe ---------------------------------------------------------------
    A["key_qpos"], A["key_ctrl"] = A["qpos0"].copy(), np.zeros(MAX_NU)
    for sec in root.findall("keyframe"):
        for k in sec.findall("key"):
            if k.attrib.get("name") == "home":
                q = _vec(k.attrib["qpos"], None)
                if len(q) != nq:
                    raise ValueError(f"keyframe qpos has {len(q)} entries, model nq={nq}")
                A["key_qpos"][:nq] = q
                A["key_ctrl"][:nu] = _vec(k.attrib.get("ctrl"), np.zeros(nu))

    _set_const(m)
    return m


# --------------------------------------------------------------------------- compile-time constants
def world_kinematics(m: CompiledModel, qpos: np.ndarray):
    """Plain forward kinematics: body frames, joint anchors/axes in the world (MuJoCo mj_kinematics)."""
    A = m.arrays
    xpos, xmat = np.zeros((m.nbody, 3)), np.zeros((m.nbody, 3, 3))
    xmat[0] = np.eye(3)
    anchor, axis = np.zeros((m.njnt, 3)), np.zeros((m.njnt, 3))
    for b in range(1, m.nbody):
        p = A["body_parentid"][b]
        pos = xpos[p] + xmat[p] @ A["body_pos"][b]
        quat = quat_mul(mat_to_quat(xmat[p]), A["body_quat"][b])
        for k in range(A["body_jntnum"][b]):
            j = A["body_jntadr"][b] + k
            qa = A["jnt_qposadr"][j]
            if A["jnt_type"][j] == JNT_FREE:
                pos = qpos[qa:qa + 3].copy()
                quat = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
                anchor[j], axis[j] = pos, [0, 0, 1]
            else:
                R = quat_to_mat(quat)
                anchor[j] = pos + R @ A["jnt_pos"][j]
                axis[j] = R @ A["jnt_axis"][j]
                ang = qpos[qa] - A["qpos0"][qa]
                ql = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * A["jnt_axis"][j]])
                quat = quat_mul(quat, ql)
                pos = anchor[j] - quat_to_mat(quat) @ A["jnt_pos"][j]
        xpos[b], xmat[b] = pos, quat_to_mat(quat)
    return xpos, xmat, anchor, axis


def body_jacobians(m: CompiledModel, qpos: np.ndarray):
    """Per-body (translational at the body COM, rotational) Jacobians, nv columns."""
    A = m.arrays
    xpos, xmat, anchor, axis = world_kinematics(m, qpos)
    xipos = np.array([xpos[b] + xmat[b] @ A["body_ipos"][b] for b in range(m.nbody)])
    jacp, jacr = np.zeros((m.nbody, 3, m.nv)), np.zeros((m.nbody, 3, m.nv))
    for b in range(1, m.nbody):
        anc = b
        while anc > 0:
            for k in range(A["body_jntnum"][anc]):
                j = A["body_jntadr"][anc] + k
                d = A["jnt_dofadr"][j]
                if A["jnt_type"][j] == JNT_FREE:
                    jacp[b, :, d:d + 3] = np.eye(3)
                    for c in range(3):
                        a = xmat[anc][:, c]
                        jacr[b, :, d + 3 + c] = a
                        jacp[b, :, d + 3 + c] = np.cross(a, xipos[b] - xpos[anc])
                else:
                    jacr[b, :, d] = axis[j]
                    jacp[b, :, d] = np.cross(axis[j], xipos[b] - anchor[j])
            anc = A["body_parentid"][anc]
    return xpos, xmat, xipos, jacp, jacr


def mass_matrix(m: CompiledModel, qpos: np.ndarray, body_mass=None, armature=None) -> np.ndarray:
    """Dense joint-space inertia from body Jacobians (an independent check on CRBA in oracle/ and csrc/)."""
    A = m.arrays
    mass = A["body_mass"] if body_mass is None else body_mass
    arm = A["dof_armature"] if armature is None else armature
    _, xmat, _, jacp, jacr = body_jacobians(m, qpos)
    M = np.diag(arm[: m.nv]).astype(np.float64)
    for b in range(1, m.nbody):
        Ri = xmat[b] @ quat_to_mat(A["body_iquat"][b])
        Iw = Ri @ np.diag(A["body_inertia"][b]) @ Ri.T
        M += mass[b] * jacp[b].T @ jacp[b] + jacr[b].T @ Iw @ jacr[b]
    return M


def _set_const(m: CompiledModel) -> None:
    """dof_invweight0 / body_invweight0 / meaninertia at qpos0 (MuJoCo engine_setconst.c set0)."""
    A = m.arrays
    q0 = A["qpos0"][: m.nq]
    M = mass_matrix(m, q0)
    Minv = np.linalg.inv(M)
    A["meaninertia"] = np.array(np.trace(M) / m.nv)
    _, _, _, jacp, jacr = body_jacobians(m, q0)
    # body is static (welded to the world) iff no ancestor has a dof
    for b in range(1, m.nbody):
        anc, moving = b, False
        while anc > 0:
            moving |= A["body_dofnum"][anc] > 0
            anc = A["body_parentid"][anc]
        if not moving:
            continue
        A["body_invweight0"][b, 0] = np.trace(jacp[b] @ Minv @ jacp[b].T) / 3
        A["body_invweight0"][b, 1] = np.trace(jacr[b] @ Minv @ jacr[b].T) / 3
    d = np.diag(Minv)
    for j in range(m.njnt):
        a = A["jnt_dofadr"][j]
        if A["jnt_type"][j] == JNT_FREE:
            A["dof_invweight0"][a:a + 3] = d[a:a + 3].mean()
            A["dof_invweight0"][a + 3:a + 6] = d[a + 3:a + 6].mean()
        else:
            A["dof_invweight0"][a] = d[a]
