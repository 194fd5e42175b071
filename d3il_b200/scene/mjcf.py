"""Minimal MJCF reader + numpy rigid-body helpers used by the offline scene compiler.

Dev-time only (reads the reference's XML assets, which do not travel to the GPU
box); the compiled tables under ``d3il_b200/scenes/`` are what the runtime loads.

Restates the parts of ``MjModel.from_xml_string`` the hot path depends on
(SURVEY.md App. B.3): body tree, explicit ``<inertial>`` or geom-derived
inertia, childclass defaults for the gripper, and the welding of joint-less
bodies into their moving ancestor (dynamically exact).
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np


# ----------------------------------------------------------------------------- math
def quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.linalg.norm(q)


def quat2mat(q):
    w, x, y, z = quat_normalize(q)
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


def mat2quat(R):
    """Rotation matrix -> unit quaternion (w,x,y,z), w >= 0 branch-stable."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    if q[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def rpy2mat(rpy):
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix (Rz(y) Ry(p) Rx(r))."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _floats(s, n=None, default=None):
    if s is None:
        return None if default is None else np.asarray(default, dtype=np.float64)
    v = np.array([float(t) for t in s.split()], dtype=np.float64)
    if n is not None and len(v) != n:
        raise ValueError(f"expected {n} floats, got {s!r}")
    return v


# ----------------------------------------------------------------------------- geom inertia
def geom_volume(gtype, size):
    if gtype == "box":
        return 8 * size[0] * size[1] * size[2]
    if gtype == "sphere":
        return 4.0 / 3.0 * np.pi * size[0] ** 3
    if gtype == "cylinder":
        return np.pi * size[0] ** 2 * 2 * size[1]
    raise ValueError(gtype)


def geom_inertia_diag(gtype, size, mass):
    """Principal inertia of a primitive in its own frame (SURVEY App. B.3 [EXT])."""
    if gtype == "box":
        a, b, c = size
        return mass / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    if gtype == "sphere":
        return np.full(3, 0.4 * mass * size[0] ** 2)
    if gtype == "cylinder":
        r, h = size[0], 2 * size[1]
        ip = mass * (3 * r * r + h * h) / 12.0
        return np.array([ip, ip, 0.5 * mass * r * r])
    raise ValueError(gtype)


def geom_rbound(gtype, size):
    if gtype == "box":
        return float(np.linalg.norm(size))
    if gtype == "sphere":
        return float(size[0])
    if gtype == "cylinder":
        return float(np.hypot(size[0], size[1]))
    if gtype == "plane":
        return 0.0
    raise ValueError(gtype)


# ----------------------------------------------------------------------------- tree
GEOM_DEFAULTS = dict(
    contype=1, conaffinity=1, condim=3, priority=0,
    friction=(1.0, 0.005, 0.0001), solref=(0.02, 1.0),
    solimp=(0.9, 0.95, 0.001, 0.5, 2.0), margin=0.0, gap=0.0, solmix=1.0,
    density=1000.0,
)


@dataclass
class Geom:
    name: str
    type: str
    size: np.ndarray
    pos: np.ndarray            # in body frame
    quat: np.ndarray
    mass: float | None
    mesh: str | None
    params: dict               # contype, conaffinity, condim, priority, friction, solref, solimp, margin, gap, solmix
    body: "Body" = None


@dataclass
class Joint:
    name: str
    type: str                  # hinge | slide | free
    axis: np.ndarray
    limited: bool
    range: np.ndarray
    damping: float


@dataclass
class Body:
    name: str
    pos: np.ndarray
    quat: np.ndarray
    parent: "Body | None"
    inertial: dict | None = None       # explicit: mass, pos, quat, diag
    joints: list = field(default_factory=list)
    geoms: list = field(default_factory=list)
    children: list = field(default_factory=list)
    # filled by finalize():
    mass: float = 0.0
    ipos: np.ndarray = None            # CoM in body frame
    inertia: np.ndarray = None         # 3x3 about CoM, body-frame axes


def _full_solimp(v):
    out = list(GEOM_DEFAULTS["solimp"])
    for i, x in enumerate(v):
        out[i] = x
    return tuple(out)


def _geom_params(elem, cls):
    p = dict(GEOM_DEFAULTS)
    src = {}
    src.update(cls)
    src.update({k: v for k, v in elem.attrib.items()})
    for k in ("contype", "conaffinity", "condim", "priority"):
        if k in src:
            p[k] = int(src[k])
    for k in ("margin", "gap", "solmix", "density"):
        if k in src:
            p[k] = float(src[k])
    if "friction" in src:
        f = list(GEOM_DEFAULTS["friction"])
        for i, x in enumerate(_floats(src["friction"])):
            f[i] = x
        p["friction"] = tuple(f)
    if "solref" in src:
        p["solref"] = tuple(_floats(src["solref"], 2))
    if "solimp" in src:
        p["solimp"] = _full_solimp(_floats(src["solimp"]))
    return p, src


def parse_bodies(worldbody: ET.Element, defaults: dict[str, dict], rename=lambda s: s) -> list[Body]:
    """Depth-first list of bodies under ``worldbody`` (world itself excluded)."""
    out: list[Body] = []

    def rec(elem, parent, childclass):
        cc = elem.get("childclass", childclass)
        b = Body(
            name=rename(elem.get("name", "")),
            pos=_floats(elem.get("pos"), 3, (0, 0, 0)),
            quat=quat_normalize(_floats(elem.get("quat"), 4, (1, 0, 0, 0))),
            parent=parent,
        )
        out.append(b)
        if parent is not None:
            parent.children.append(b)
        ine = elem.find("inertial")
        if ine is not None:
            di = ine.get("diaginertia")
            b.inertial = dict(
                mass=float(ine.get("mass")),
                pos=_floats(ine.get("pos"), 3, (0, 0, 0)),
                quat=quat_normalize(_floats(ine.get("quat"), 4, (1, 0, 0, 0))),
                diag=_floats(di, 3),
            )
        for j in list(elem.findall("joint")) + list(elem.findall("freejoint")):
            cls = defaults.get(cc, {}).get("joint", {}) if cc else {}
            src = dict(cls)
            src.update(j.attrib)
            jtype = "free" if j.tag == "freejoint" else src.get("type", "hinge")
            b.joints.append(
                Joint(
                    name=rename(src.get("name", "")),
                    type=jtype,
                    axis=quat_normalize(_floats(src.get("axis"), 3, (0, 0, 1))) if jtype != "free" else np.zeros(3),
                    limited=src.get("limited", "false") == "true",
                    range=_floats(" ".join(src.get("range", "0 0").split()), 2),
                    damping=float(src.get("damping", 0.0)),
                )
            )
        for g in elem.findall("geom"):
            cls = defaults.get(cc, {}).get("geom", {}) if cc else {}
            p, src = _geom_params(g, cls)
            gtype = src.get("type", "sphere")
            geom = Geom(
                name=rename(src.get("name", "")),
                type=gtype,
                size=_floats(src.get("size"), None, (0, 0, 0)),
                pos=_floats(src.get("pos"), 3, (0, 0, 0)),
                quat=quat_normalize(_floats(src.get("quat"), 4, (1, 0, 0, 0))),
                mass=float(src["mass"]) if "mass" in src else None,
                mesh=src.get("mesh"),
                params=p,
                body=b,
            )
            b.geoms.append(geom)
        for c in elem.findall("body"):
            rec(c, b, cc)

    for top in worldbody.findall("body"):
        rec(top, None, None)
    return out


def parse_defaults(root: ET.Element) -> dict[str, dict]:
    out = {}
    for d in root.iter("default"):
        cls = d.get("class")
        if cls is None:
            continue
        out[cls] = {child.tag: dict(child.attrib) for child in d if child.tag != "default"}
    return out


def finalize_inertia(b: Body):
    """Body mass / CoM / inertia: explicit <inertial> wins, else sum of primitive geoms."""
    if b.inertial is not None:
        R = quat2mat(b.inertial["quat"])
        b.mass = b.inertial["mass"]
        b.ipos = b.inertial["pos"].copy()
        b.inertia = R @ np.diag(b.inertial["diag"]) @ R.T
        return
    m_tot, com = 0.0, np.zeros(3)
    parts = []
    for g in b.geoms:
        if g.type in ("mesh", "plane"):
            continue
        m = g.mass if g.mass is not None else g.params["density"] * geom_volume(g.type, g.size)
        parts.append((m, g))
        m_tot += m
        com += m * g.pos
    if m_tot <= 0:
        b.mass, b.ipos, b.inertia = 0.0, np.zeros(3), np.zeros((3, 3))
        return
    com /= m_tot
    I = np.zeros((3, 3))
    for m, g in parts:
        R = quat2mat(g.quat)
        Ig = R @ np.diag(geom_inertia_diag(g.type, g.size, m)) @ R.T
        d = g.pos - com
        I += Ig + m * (d @ d * np.eye(3) - np.outer(d, d))
    b.mass, b.ipos, b.inertia = m_tot, com, I


@dataclass
class Link:
    """A moving rigid body after welding joint-less descendants into it."""
    name: str
    parent: int                # link index, -1 = world
    pos: np.ndarray            # frame rel. parent link frame
    quat: np.ndarray
    jtype: int                 # 0 hinge, 1 slide, 2 free
    axis: np.ndarray
    range: np.ndarray
    limited: bool
    damping: float
    mass: float
    ipos: np.ndarray
    inertia: np.ndarray        # 3x3 about CoM in link frame
    members: dict              # original body name -> (pos, R) of that body frame in link frame


def weld(bodies: list[Body]) -> tuple[list[Link], dict]:
    """Fold joint-less bodies into their nearest jointed ancestor.

    Returns (links, static) where ``static`` maps names of world-welded bodies
    to their world (pos, R).
    """
    for b in bodies:
        finalize_inertia(b)
    links: list[Link] = []
    link_of: dict[str, int] = {}
    static: dict[str, tuple] = {}
    rel: dict[str, tuple] = {}     # body name -> (link index or -1, pos, R) in that link/world frame

    acc = {}                        # link index -> list of (mass, com_in_link, I_in_link_axes)
    for b in bodies:
        Rb = quat2mat(b.quat)
        if b.parent is None:
            pl, pp, pR = -1, np.zeros(3), np.eye(3)
        else:
            pl, pp, pR = rel[b.parent.name]
        pos_in = pp + pR @ b.pos
        R_in = pR @ Rb
        if b.joints:
            assert len(b.joints) == 1, "one joint per body on this path"
            j = b.joints[0]
            jt = {"hinge": 0, "slide": 1, "free": 2}[j.type]
            idx = len(links)
            links.append(
                Link(
                    name=b.name, parent=pl, pos=pos_in, quat=mat2quat(R_in), jtype=jt, axis=j.axis.copy(),
                    range=j.range.copy(), limited=j.limited, damping=j.damping,
                    mass=0.0, ipos=np.zeros(3), inertia=np.zeros((3, 3)), members={},
                )
            )
            link_of[b.name] = idx
            rel[b.name] = (idx, np.zeros(3), np.eye(3))
            acc[idx] = []
        else:
            rel[b.name] = (pl, pos_in, R_in)
            if pl == -1:
                static[b.name] = (pos_in, R_in)
        li, p, R = rel[b.name]
        if li >= 0:
            links[li].members[b.name] = (p.copy(), R.copy())
            if b.mass > 0:
                acc[li].append((b.mass, p + R @ b.ipos, R @ b.inertia @ R.T))
    for li, parts in acc.items():
        m = sum(x[0] for x in parts)
        com = sum(x[0] * x[1] for x in parts) / m
        I = np.zeros((3, 3))
        for mi, ci, Ii in parts:
            d = ci - com
            I += Ii + mi * (d @ d * np.eye(3) - np.outer(d, d))
        links[li].mass, links[li].ipos, links[li].inertia = m, com, I
    return links, {"static": static, "rel": rel, "link_of": link_of}


# ----------------------------------------------------------------------------- numpy dynamics (compile-time checks + invweight0)
def link_fk(links: list[Link], qpos_by_link):
    """World (p, R) of every link frame. qpos_by_link[i] is scalar (hinge/slide) or 7-vector (free)."""
    P, Rm = [], []
    for i, L in enumerate(links):
        if L.parent < 0:
            pp, pR = np.zeros(3), np.eye(3)
        else:
            pp, pR = P[L.parent], Rm[L.parent]
        q = qpos_by_link[i]
        if L.jtype == 2:
            p = np.asarray(q[:3], dtype=np.float64)
            R = quat2mat(q[3:7])
        else:
            p = pp + pR @ L.pos
            R = pR @ quat2mat(L.quat)
            if L.jtype == 0:
                a = L.axis
                K = skew(a)
                R = R @ (np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * K @ K)
            else:
                p = p + R @ (L.axis * q)
        P.append(p)
        Rm.append(R)
    return P, Rm


def dof_layout(links):
    adr, n = [], 0
    for L in links:
        adr.append(n)
        n += 6 if L.jtype == 2 else 1
    return adr, n


def motion_subspace(links, P, Rm):
    """Columns S[:, d] = [omega; v_at_world_origin] per dof, world axes."""
    adr, nv = dof_layout(links)
    S = np.zeros((6, nv))
    for i, L in enumerate(links):
        p, R = P[i], Rm[i]
        if L.jtype == 0:
            a = R @ L.axis
            S[:3, adr[i]] = a
            S[3:, adr[i]] = np.cross(p, a)
        elif L.jtype == 1:
            S[3:, adr[i]] = R @ L.axis
        else:
            for k in range(3):
                S[3 + k, adr[i] + k] = 1.0                  # world-frame linear dofs
                a = R[:, k]                                  # body-local angular dofs
                S[:3, adr[i] + 3 + k] = a
                S[3:, adr[i] + 3 + k] = np.cross(p, a)
    return S, adr, nv


def spatial_inertia_world(L: Link, p, R):
    c = p + R @ L.ipos
    I = R @ L.inertia @ R.T
    cx = skew(c)
    out = np.zeros((6, 6))
    out[:3, :3] = I - L.mass * cx @ cx
    out[:3, 3:] = L.mass * cx
    out[3:, :3] = -L.mass * cx
    out[3:, 3:] = L.mass * np.eye(3)
    return out


def mass_matrix(links, qpos_by_link):
    P, Rm = link_fk(links, qpos_by_link)
    S, adr, nv = motion_subspace(links, P, Rm)
    n = len(links)
    Ic = [spatial_inertia_world(links[i], P[i], Rm[i]) for i in range(n)]
    for i in range(n - 1, -1, -1):
        if links[i].parent >= 0:
            Ic[links[i].parent] = Ic[links[i].parent] + Ic[i]
    # ancestor sets
    anc = []
    for i in range(n):
        s = {i}
        j = links[i].parent
        while j >= 0:
            s.add(j)
            j = links[j].parent
        anc.append(s)
    dof_link = []
    for i, L in enumerate(links):
        dof_link += [i] * (6 if L.jtype == 2 else 1)
    M = np.zeros((nv, nv))
    for a in range(nv):
        for b in range(nv):
            la, lb = dof_link[a], dof_link[b]
            if la in anc[lb]:
                deep = lb
            elif lb in anc[la]:
                deep = la
            else:
                continue
            M[a, b] = S[:, a] @ Ic[deep] @ S[:, b]
    return M, (P, Rm, S, adr, nv, anc, dof_link)


def point_jacobian(links, ctx, link_idx, point_world):
    """6 x nv Jacobian [lin; ang] of a point rigidly attached to ``link_idx``."""
    P, Rm, S, adr, nv, anc, dof_link = ctx
    J = np.zeros((6, nv))
    for d in range(nv):
        if dof_link[d] in anc[link_idx]:
            w, v0 = S[:3, d], S[3:, d]
            J[:3, d] = v0 + np.cross(w, point_world)
            J[3:, d] = w
    return J
