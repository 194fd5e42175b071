"""Binary scene-table layout ("D3SC" v1) shared by the scene compiler, the CUDA
library loader and the oracle.  One blob per task; everything the kernels need
that the reference gets from ``MjModel.from_xml_string`` + the ``.gin`` gains.

Layout: int32 header[32] then float64 sections  LINK | GEOM | PAIR | CTRL | TASK.
The same offsets are restated in ``include/d3il.h`` (D3SC_* macros).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

MAGIC = 0x43533344  # 'D3SC'
VERSION = 1
HDR_INTS = 32
LINK_W, GEOM_W, PAIR_W, CTRL_W = 32, 24, 24, 192

TASK_IDS = {"avoiding": 0, "pushing": 1, "aligning": 2, "sorting": 3, "stacking": 4, "inserting": 5}

# mjtGeom ids so the tables read like a MuJoCo model dump
GEOM_PLANE, GEOM_SPHERE, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 0, 2, 5, 6, 7
GEOM_TYPE_ID = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "mesh": GEOM_MESH}

HDR_FIELDS = [
    "magic", "version", "task_id", "nlink", "nobj", "nq", "nv", "ngeom", "npair", "n_substeps",
    "max_steps", "obs_dim", "act_dim", "ctx_dim", "info_dim", "ctrl_kind", "ntaskp", "nextra", "maxcon",
]

# CTRL section offsets
C_IK_ORIGIN = 0        # 7 x (pos3, R9)
C_IK_EE = 84           # pos3, R9
C_PGAIN_POS = 96
C_PGAIN_QUAT = 99
C_PGAIN_NULL = 102
C_REST = 109
C_JMIN = 116
C_JMAX = 123
C_PD_P = 130
C_PD_D = 137
C_JREG = 144
C_SVD_MIN = 145
C_SVD_MAX = 146
C_NUM_ITER = 147
C_LRATE = 148
C_DT = 149
C_INIT_QPOS = 150      # 7
C_TCP_POS = 157        # tcp body in link7 frame (physics chain)
C_TCP_QUAT = 160
C_GRAVITY = 164
C_IMPRATIO = 167
C_TOL = 168
C_JNT_SOLREF = 169     # 2
C_JNT_SOLIMP = 171     # 5
C_INIT_TCP = 176       # commanded start pose (3) — informational


@dataclass
class Scene:
    header: dict
    link: np.ndarray   # [nlink, LINK_W]
    geom: np.ndarray   # [ngeom, GEOM_W]
    pair: np.ndarray   # [npair, PAIR_W]
    ctrl: np.ndarray   # [CTRL_W]
    task: np.ndarray   # [ntaskp]

    def pack(self) -> bytes:
        hdr = [0] * HDR_INTS
        for i, k in enumerate(HDR_FIELDS):
            hdr[i] = int(self.header.get(k, 0))
        out = struct.pack(f"<{HDR_INTS}i", *hdr)
        for a in (self.link, self.geom, self.pair, self.ctrl, self.task):
            out += np.ascontiguousarray(a, dtype="<f8").tobytes()
        return out

    @staticmethod
    def unpack(buf: bytes) -> "Scene":
        hdr = struct.unpack_from(f"<{HDR_INTS}i", buf, 0)
        h = {k: hdr[i] for i, k in enumerate(HDR_FIELDS)}
        if h["magic"] != MAGIC or h["version"] != VERSION:
            raise ValueError("not a D3SC v1 scene blob")
        off = 4 * HDR_INTS

        def take(n, shape):
            nonlocal off
            a = np.frombuffer(buf, dtype="<f8", count=n, offset=off).reshape(shape).copy()
            off += 8 * n
            return a

        link = take(h["nlink"] * LINK_W, (h["nlink"], LINK_W))
        geom = take(h["ngeom"] * GEOM_W, (h["ngeom"], GEOM_W))
        pair = take(h["npair"] * PAIR_W, (h["npair"], PAIR_W))
        ctrl = take(CTRL_W, (CTRL_W,))
        task = take(h["ntaskp"], (h["ntaskp"],))
        if off != len(buf):
            raise ValueError("trailing bytes in scene blob")
        return Scene(h, link, geom, pair, ctrl, task)


def links_from_scene(scene: "Scene"):
    """Rebuild the compile-time ``mjcf.Link`` list from a scene blob (numpy cross-checks in tests)."""
    from .mjcf import Link

    out = []
    for r in scene.link:
        I = np.array([[r[16], r[19], r[20]], [r[19], r[17], r[21]], [r[20], r[21], r[18]]])
        out.append(Link(name="", parent=int(r[0]), pos=r[2:5].copy(), quat=r[5:9].copy(), jtype=int(r[1]), axis=r[9:12].copy(),
                        range=r[23:25].copy(), limited=bool(r[22]), damping=r[25], mass=r[12], ipos=r[13:16].copy(), inertia=I, members={}))
    return out


def load_scene(task: str) -> tuple[bytes, "Scene"]:
    import os

    path = os.path.join(os.path.dirname(__file__), "..", "scenes", f"{task}.d3sc")
    buf = open(path, "rb").read()
    return buf, Scene.unpack(buf)
