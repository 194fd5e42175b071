"""Offline scene compiler: reference XML/URDF/gin/pkl assets -> flat scene tables.

Restates what ``MjSceneParser.create_scene`` + ``MjModel.from_xml_string`` do
for the hot path (reference
``environments/d3il/d3il_sim/sims/mj_beta/mj_utils/mj_scene_parser.py:36-53``):
assemble surroundings + task objects + Panda, resolve defaults, derive inertias,
weld joint-less bodies, build the filtered contact-pair table with mixed contact
parameters, and compute ``body_invweight0`` / ``dof_invweight0`` at qpos0.  Also
restates the start-up offline IK that fixes each task's ``init_qpos`` (a13;
``controllers/TrajectoryTracking.py:331-447``).

Run (dev container only, the GPU box has no /root/reference):
    python -m d3il_b200.scene.compile --ref /root/reference --out d3il_b200/scenes
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import re
import xml.etree.ElementTree as ET

import numpy as np

from . import blob as B
from . import mjcf as M

D3IL = "environments/d3il"
DT = 0.001


# ----------------------------------------------------------------------------- gin
def read_gin(path):
    """20-line reader for the ``.gin`` gains file (gin-config is not installed)."""
    out = {}
    for line in open(path):
        line = line.split("#")[0].strip()
        if "=" not in line:
            continue
        k, v = line.split("=", 1)
        k = ".".join(k.strip().split(".")[-2:])
        out[k] = json.loads(v.strip().replace(" .", " 0.").replace("[.", "[0.").replace(". ", ".0 ").replace(".,", ".0,").replace(".]", ".0]"))
    return out


# ----------------------------------------------------------------------------- URDF kinematic chain (controller side, pinocchio model)
def read_urdf_chain(path):
    root = ET.parse(path).getroot()
    joints = {j.get("name"): j for j in root.findall("joint")}
    chain = []
    for i in range(1, 8):
        j = joints[f"panda_joint{i}"]
        o = j.find("origin")
        xyz = np.array([float(x) for x in o.get("xyz").split()])
        rpy = np.array([float(x) for x in o.get("rpy").split()])
        axis = np.array([float(x) for x in j.find("axis").get("xyz").split()])
        assert np.allclose(axis, [0, 0, 1])
        chain.append((xyz, M.rpy2mat(rpy)))
    p, R = np.zeros(3), np.eye(3)
    for name in ("panda_joint8", "panda_hand_joint", "panda_grasptarget_hand"):
        o = joints[name].find("origin")
        xyz = np.array([float(x) for x in o.get("xyz").split()])
        rpy = np.array([float(x) for x in o.get("rpy").split()])
        p = p + R @ xyz
        R = R @ M.rpy2mat(rpy)
    return chain, (p, R)


def ik_fk(chain, ee, q):
    """FK + 6x7 LOCAL_WORLD_ALIGNED Jacobian of ``panda_grasptarget`` (``core/Model.py:37-66``)."""
    p, R = np.zeros(3), np.eye(3)
    origins, axes = [], []
    for i in range(7):
        xyz, Ro = chain[i]
        p = p + R @ xyz
        R = R @ Ro
        c, s = np.cos(q[i]), np.sin(q[i])
        R = R @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        origins.append(p.copy())
        axes.append(R[:, 2].copy())
    pe = p + R @ ee[0]
    Re = R @ ee[1]
    J = np.zeros((6, 7))
    for i in range(7):
        J[:3, i] = np.cross(axes[i], pe - origins[i])
        J[3:, i] = axes[i]
    return pe, M.mat2quat(Re), J


def quat_error(c, d):
    """``utils/geometric_transformation.py:14-46`` (Siciliano 3.91)."""
    return np.array(
        [
            c[0] * d[1] - d[0] * c[1] - c[3] * d[2] + c[2] * d[3],
            c[0] * d[2] - d[0] * c[2] + c[3] * d[1] - c[1] * d[3],
            c[0] * d[3] - d[0] * c[3] - c[2] * d[1] + c[1] * d[2],
        ]
    )


Q0_DEFAULT = np.array([3.57795216e-09, 1.74532920e-01, 3.30500960e-08, -8.72664630e-01, -1.14096181e-07, 1.22173047e00, 7.85398126e-01])
JPOS_MIN = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])   # core/Robots.py:60-65
JPOS_MAX = np.array([2.8973, 1.7628, 2.0, -0.0698, 2.8973, 3.7525, 2.8973])


def offline_ik(chain, ee, target_pos, target_quat=(0, 1, 0, 0)):
    """``OfflineIKTrajectoryGenerator.generate_trajectory`` (``TrajectoryTracking.py:331-447``)."""
    des = np.concatenate([target_pos, target_quat]).astype(np.float64)
    pgain = np.array([33.9403713446798, 30.9403713446798, 33.9403713446798, 27.69370238555632, 33.98706171459314, 30.9185531893281])
    pgain_null = 5 * np.array([7.675519770796831, 2.676935478437176, 8.539040163444975, 1.270446361314313, 8.87752182480855, 2.186782233762969, 4.414432577659688])
    pgain_limit, J_reg, eps, IT_MAX, dt = 20, 1e-6, 1e-5, 1000, 1e-3
    target_null = Q0_DEFAULT.copy()
    q = Q0_DEFAULT.copy()
    qd = np.zeros(7)
    old_err = np.inf
    i = 0
    while True:
        old_q = q
        q = np.clip(q + dt * qd, JPOS_MIN, JPOS_MAX)
        pos, orient = ik_fk(chain, ee, q)[:2]
        cpos_err = des[:3] - pos
        if np.linalg.norm(orient - des[3:]) > np.linalg.norm(orient + des[3:]):
            orient = -orient
        cpos_err = np.clip(cpos_err, -0.1, 0.1)
        cquat_err = np.clip(quat_error(orient, des[3:]), -0.5, 0.5)
        err = np.hstack((cpos_err, cquat_err))
        err_norm = np.sum(cpos_err**2) + np.sum((orient - des[3:]) ** 2)
        if err_norm > old_err:
            q = old_q
            dt *= 0.7
            continue
        dt *= 1.025
        if err_norm < eps or i >= IT_MAX:
            break
        old_err = err_norm
        J = ik_fk(chain, ee, q)[2]
        A = J @ J.T + J_reg * np.eye(6)
        qd_null = pgain_null * (target_null - q)
        lim = np.zeros(7)
        hi = q > JPOS_MAX - 0.1
        lo = q < JPOS_MIN + 0.1
        lim[hi] += (pgain_limit * (JPOS_MAX - 0.1 - q))[hi]
        lim[lo] += (pgain_limit * (JPOS_MIN + 0.1 - q))[lo]
        qd_null = qd_null + lim
        qd = J.T @ np.linalg.solve(A, pgain * err - J @ qd_null) + qd_null
        i += 1
    return q, i, err_norm


# ----------------------------------------------------------------------------- robot + surroundings
def load_robot(ref, with_rod=True):
    xml = os.path.join(ref, D3IL, "models/mj/robot", "panda_rod_invisible.xml" if with_rod else "panda_invisible.xml")
    root = ET.parse(xml).getroot()
    defaults = M.parse_defaults(root)
    bodies = M.parse_bodies(root.find("worldbody"), defaults)
    force = {}
    for m in root.find("actuator").findall("motor"):
        lo, hi = [float(x) for x in m.get("forcerange").split()]
        assert lo == -hi
        force[m.get("joint")] = hi
    return bodies, force


def load_surroundings(ref):
    """The static geoms free objects can touch: ``table_plane`` and ``support_body`` of the frame, and the
    ground plane of ``base.xml`` (body ``ground`` at z = -0.94), which catches what is pushed off the table.  The plane is
    compiled as the top face of a big static axis-aligned slab: for the convex objects that land on it that is the same
    contact set, and it takes the slab fast path of the box narrow phase (one comparison per tick while nothing is near)."""
    xml = os.path.join(ref, D3IL, "models/mujoco/surroundings/lab_surrounding.xml")
    root = ET.parse(xml).getroot()
    bodies = M.parse_bodies(root.find("worldbody"), {})
    _, info = M.weld(bodies)
    out = []
    for b in bodies:
        if b.name in ("table_plane", "support_body"):
            p, R = info["static"][b.name]
            g = b.geoms[0]
            out.append(dict(name=b.name, type=g.type, size=g.size, pos=p + R @ g.pos, R=R @ M.quat2mat(g.quat), params=g.params))
    base = ET.parse(os.path.join(ref, D3IL, "models/mj/surroundings/base.xml")).getroot()
    for b in M.parse_bodies(base.find("worldbody"), {}):
        if b.name == "ground":
            g = b.geoms[0]
            assert g.type == "plane" and np.allclose(M.quat2mat(M.quat_normalize(b.quat)), np.eye(3)) and np.allclose(M.quat2mat(g.quat), np.eye(3))
            half = np.array([4.0, 4.0, 0.5])
            out.append(dict(name="ground", type="box", size=half, pos=np.asarray(b.pos, float) + g.pos - np.array([0.0, 0.0, half[2]]), R=np.eye(3), params=g.params))
    assert out[-1]["name"] == "ground"
    return out


# ----------------------------------------------------------------------------- task descriptions (objects as the envs build them)
def prim(name, gtype, size, pos, quat, mass=0.1, static=False, **params):
    p = dict(M.GEOM_DEFAULTS)
    p.update(params)
    return dict(name=name, parts=[dict(type=gtype, size=np.array(size, float), pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), mass=mass, params=p)],
                pos=np.array(pos, float), quat=np.array(quat, float), static=static)


def task_spec(task, ref):
    if task == "pushing":   # envs/gym_pushing_env/.../objects/pushing_objects.py:18-65, pushing.py:171-253
        return dict(
            rod=True, n_substeps=35, max_steps=400, init_tcp=[0.525, -0.28, 0.12], ctrl_kind=0,
            objects=[
                prim("push_box", "box", [0.03, 0.03, 0.03], [0.4, -0.3, -0.0072], [0, 1, 0, 0], mass=0.05),
                prim("push_box2", "box", [0.03, 0.03, 0.03], [0.5, -0.3, -0.0072], [0, 1, 0, 0], mass=0.05),
            ],
            obs_dim=8, act_dim=7, info_dim=4,
            taskp=[0.42, 0.3, 0.0, 0.63, 0.3, 0.0, 0.05],   # target_pos1, target_pos2, target_min_dist (pushing.py:253)
        )
    if task == "avoiding":  # envs/gym_avoiding_env/.../objects/avoiding_objects.py:8-63, avoiding.py:52-116
        mid, off, y1, dy = 0.5, 0.075, -0.1, 0.18
        cyl = lambda n, x, y, r, h: prim(n, "cylinder", [r, h], [x, y, 0], [1, 0, 0, 0], static=True)
        return dict(
            rod=True, n_substeps=35, max_steps=250, init_tcp=[0.525, -0.28, 0.12], ctrl_kind=0,
            objects=[
                cyl("l1_obs", mid, y1, 0.03, 0.07),
                cyl("l2_top_obs", mid - off, y1 + dy, 0.025, 0.1),
                cyl("l2_bottom_obs", mid + off, y1 + dy, 0.025, 0.1),
                cyl("l3_top_obs", mid - 2 * off, y1 + 2 * dy, 0.025, 0.1),
                cyl("l3_mid_obs", mid, y1 + 2 * dy, 0.025, 0.1),
                cyl("l3_bottom_obs", mid + 2 * off, y1 + 2 * dy, 0.025, 0.1),
            ],
            obs_dim=2, act_dim=7, info_dim=11,
            # l1_y, l2_y, l3_y, goal_y, l1_x, l2_top_x, l2_bottom_x, l3_top_x, l3_mid_x, l3_bottom_x  (avoiding.py:94-105)
            taskp=[y1, y1 + dy, y1 + 2 * dy, y1 + 2.5 * dy, mid, mid - off, mid + off, mid - 2 * off, mid, mid + 2 * off],
        )
    if task.startswith("sorting"):   # envs/gym_sorting_env/.../objects/sorting_objects.py:13-220, sorting.py:193-306
        k = int(task.split("_")[1])
        assert k in (2, 4, 6)
        box = lambda n: prim(n, "box", [0.03, 0.03, 0.03], [0.5, -0.1, 0.0], [0, 1, 0, 0], mass=0.05)
        wall = lambda n, pos, size: prim(n, "box", size, pos, [0, 1, 0, 0], mass=0.05, static=True)
        objs = [box(f"red_{i + 1}") for i in range(k // 2)] + [box(f"blue_{i + 1}") for i in range(k // 2)]
        objs += [
            wall("target_box_1", [0.4, 0.41, 0.0], [0.1, 0.01, 0.1]), wall("target_box_2", [0.3, 0.32, 0.0], [0.005, 0.1, 0.1]),
            wall("target_box_3", [0.5, 0.32, 0.0], [0.005, 0.1, 0.1]), wall("target_box_4", [0.4, 0.22, 0.0], [0.1, 0.005, 0.1]),
            wall("target_box_5", [0.625, 0.41, 0.0], [0.1, 0.01, 0.1]), wall("target_box_6", [0.525, 0.32, 0.0], [0.005, 0.1, 0.1]),
            wall("target_box_7", [0.725, 0.32, 0.0], [0.005, 0.1, 0.1]), wall("target_box_8", [0.625, 0.22, 0.0], [0.1, 0.005, 0.1]),
            # models/mj/common-objects/sorting/platform.xml, re-posed by SortingObject.mj_load (sorting_objects.py:15,73-79)
            prim("platform", "box", [0.3, 0.3, 0.1], [0.5, -0.1, 0.0], [1, 0, 0, 0], mass=10, static=True, friction=[0.3, 0.001, 0.0001], priority=1),
        ]
        # absent boxes are read through mj_name2id = -1 -> the LAST body of the model, which is the joint-less
        # `finger_joint2_tip_rb0` (robots are loaded last, mj_scene_parser.py:42): model.body_pos[-1] = its local offset
        bogus = [0.0, -0.0085]
        return dict(
            rod=True, n_substeps=35, max_steps={2: 500, 4: 700, 6: 1200}[k], init_tcp=[0.525, -0.3, 0.25], ctrl_kind=0,
            objects=objs, obs_dim=2 + 3 * k, act_dim=7, info_dim=4, spawn_z=0.05,
            maxcon={2: 28, 4: 40, 6: 64}[k],     # 4 per resting box + piles pushed together / boxes against the bin walls (a scripted push of one box into its bin reaches 21 with two boxes; six boxes pushed into a pile reached 52 on the GPU; overflow raises status 2)
            # red target xy, blue target xy, red bin x range, blue bin x range, bin y range, num_boxes, bogus xy  (sorting.py:300-306,477-536)
            taskp=[0.4, 0.32, 0.625, 0.32, 0.3, 0.5, 0.525, 0.725, 0.22, 0.41, k, bogus[0], bogus[1]],
        )
    if task == "aligning":  # envs/gym_aligning_env/.../objects/aligning_objects.py:19-67, models/mj/common-objects/robot_push_box/*.xml
        xml = os.path.join(ref, D3IL, "models/mj/common-objects/robot_push_box/robot_push_box.xml")
        body = ET.parse(xml).getroot().find("worldbody").find("body")
        parts = []
        for g in body.findall("geom"):
            prm = dict(M.GEOM_DEFAULTS)
            if g.get("friction"):
                prm["friction"] = [float(x) for x in g.get("friction").split()]
            if g.get("priority"):
                prm["priority"] = int(g.get("priority"))
            parts.append(dict(type=g.get("type"), size=np.array([float(x) for x in g.get("size").split()]), pos=np.array([float(x) for x in g.get("pos").split()]),
                              quat=np.array([1.0, 0, 0, 0]), mass=float(g.get("mass")), params=prm))
        assert len(parts) == 5 and body.find("joint").get("type") == "free"
        obj = dict(name="aligning_box", parts=parts, pos=np.array([0.6, 0.15, 0.0]), quat=np.array([1.0, 0, 0, 0]), static=False)
        return dict(
            rod=True, rod_hits_table=True, n_substeps=35, max_steps=400, init_tcp=[0.525, -0.35, 0.25], ctrl_kind=0,
            objects=[obj], obs_dim=17, act_dim=7, info_dim=4, nextra=7,
            # pos_min_dist, rot_min_dist, robot_box_dist (aligning.py:200-202); default target pose = XML pose of `target_box`
            taskp=[0.018, 0.048, 0.051, 0.6, 0.15, 0.0, 1.0, 0.0, 0.0, 0.0],
        )
    if task == "stacking":  # envs/gym_stacking_env/.../objects/stacking_objects.py:13-62, stacking.py:135-226
        box = lambda n, pos, size: prim(n, "box", size, pos, [0, 1, 0, 0], mass=0.05)
        return dict(
            rod=False, gripper=True, n_substeps=30, max_steps=1000, init_tcp=[0.525, 0.0, 0.3], ctrl_kind=1,
            objects=[box("red_box", [0.5, -0.1, 0.0], [0.03, 0.03, 0.03]), box("green_box", [0.5, 0.0, 0.0], [0.03, 0.03, 0.03]),
                     box("blue_box", [0.5, 0.0, 0.0], [0.03, 0.05, 0.03])],
            obs_dim=12, act_dim=8, info_dim=5, maxcon=40,
            # target_box xy (static, visual only), pos_min_dist, min z gap, gripper-open threshold (stacking.py:202,337-341,425-447)
            taskp=[0.5, 0.2, 0.06, 0.03, 0.075],
        )
    if task == "inserting":  # envs/gym_inserting_env/.../objects/gate_insertion_objects.py, gate_insertion.py:154-286
        src = open(os.path.join(ref, D3IL, "envs/gym_inserting_env/gym_inserting/envs/objects/gate_insertion_objects.py")).read()
        num = lambda txt: [float(x) for x in txt.split(",")]
        maze = {}
        for mm in re.finditer(r'maze_(\d+) = Box\(\s*name="maze_\d+",\s*init_pos=\[([^\]]*)\],\s*init_quat=\[([^\]]*)\],.*?size=\[([^\]]*)\],\s*static=True', src, re.S):
            maze[int(mm.group(1))] = (num(mm.group(2)), num(mm.group(3)), num(mm.group(4)))
        assert sorted(maze) == list(range(1, 20)), sorted(maze)
        box = lambda n, pos: prim(n, "box", [0.025, 0.025, 0.025], pos, [0, 1, 0, 0], mass=0.05)
        objs = [box("push_box1", [0.4, -0.3, -0.0072]), box("push_box2", [0.55, -0.3, -0.0072]), box("push_box3", [0.5, -0.35, -0.0072])]
        # the env adds maze_3 .. maze_19 to the scene; maze_1 and maze_2 are created but never added (gate_insertion.py:229-255)
        objs += [prim(f"maze_{k}", "box", maze[k][2], maze[k][0], maze[k][1], mass=0.05, static=True) for k in range(3, 20)]
        return dict(
            rod=True, n_substeps=35, max_steps=2000, init_tcp=[0.525, -0.28, 0.12], ctrl_kind=0,
            objects=objs, obs_dim=11, act_dim=7, info_dim=5, maxcon=36,
            # target_box1..3 positions (visual only), target_min_dist (gate_insertion.py:281)
            taskp=[0.3575, 0.276, 0.0, 0.525, 0.4535, 0.0, 0.6925, 0.276, 0.0, 0.01],
        )
    raise ValueError(f"task {task!r} not compiled yet")


# ----------------------------------------------------------------------------- contact parameter mixing (SURVEY App. B.5 [EXT])
def mix_params(p1, p2):
    if p1["priority"] != p2["priority"]:
        w = p1 if p1["priority"] > p2["priority"] else p2
        condim, fr, solref, solimp = w["condim"], np.array(w["friction"]), np.array(w["solref"]), np.array(w["solimp"])
    else:
        condim = max(p1["condim"], p2["condim"])
        fr = np.maximum(p1["friction"], p2["friction"])
        s1, s2 = p1["solmix"], p2["solmix"]
        mix = s1 / (s1 + s2)
        solref = mix * np.array(p1["solref"]) + (1 - mix) * np.array(p2["solref"])
        solimp = mix * np.array(p1["solimp"]) + (1 - mix) * np.array(p2["solimp"])
    solref = solref.copy()
    solref[0] = max(solref[0], 2 * DT)                      # refsafe
    solimp = solimp.copy()
    solimp[0] = np.clip(solimp[0], 1e-4, 0.9999)
    solimp[1] = np.clip(solimp[1], 1e-4, 0.9999)
    solimp[2] = max(solimp[2], 0.0)
    friction5 = np.array([fr[0], fr[0], fr[1], fr[2], fr[2]])
    return condim, friction5, solref, solimp, max(p1["margin"], p2["margin"]), max(p1["gap"], p2["gap"])


# ----------------------------------------------------------------------------- compile
ROBOT_LINKS = ["panda_link1", "panda_link2", "panda_link3", "panda_link4", "panda_link5", "panda_link6", "panda_link7", "panda_leftfinger", "panda_rightfinger"]


def compile_task(task, ref):
    spec = task_spec(task, ref)
    gin = read_gin(os.path.join(ref, D3IL, "d3il_sim/controllers/Config/mujoco_controller_config.gin"))
    chain, ee = read_urdf_chain(os.path.join(ref, D3IL, "models/common/robots/panda_arm_hand_pinocchio.urdf"))

    # ---- robot
    bodies, force = load_robot(ref, with_rod=spec["rod"])
    links, info = M.weld(bodies)
    assert [L.name for L in links] == ROBOT_LINKS, [L.name for L in links]
    jnames = {b.name: b.joints[0].name for b in bodies if b.joints}
    forcerange = [force[jnames[L.name]] for L in links]
    body_by_name = {b.name: b for b in bodies}

    # ---- free objects as extra links
    objs = [o for o in spec["objects"] if not o["static"]]
    statics = [o for o in spec["objects"] if o["static"]]
    for o in objs:
        fake = M.Body(name=o["name"], pos=o["pos"], quat=o["quat"], parent=None)
        for prt in o["parts"]:
            fake.geoms.append(M.Geom(name=o["name"] + ":geom", type=prt["type"], size=np.asarray(prt["size"], float), pos=prt["pos"], quat=prt["quat"], mass=prt["mass"], mesh=None, params=prt["params"]))
        M.finalize_inertia(fake)
        links.append(M.Link(name=o["name"], parent=-1, pos=o["pos"].copy(), quat=M.quat_normalize(o["quat"]), jtype=2, axis=np.zeros(3),
                            range=np.zeros(2), limited=False, damping=0.0, mass=fake.mass, ipos=fake.ipos, inertia=fake.inertia, members={o["name"]: (np.zeros(3), np.eye(3))}))
        forcerange.append(0.0)
    nlink, nobj = len(links), len(objs)

    # ---- qpos0 state, M(qpos0), invweight0   (SURVEY App. B.6 [EXT] engine_setconst set0)
    q0 = [0.0] * 9 + [np.concatenate([L.pos, L.quat]) for L in links[9:]]
    Mq0, ctx = M.mass_matrix(links, q0)
    P0, R0 = ctx[0], ctx[1]
    adr, nv = ctx[3], ctx[4]
    Minv = np.linalg.inv(Mq0)
    dof_invw = np.zeros((nlink, 2))
    for i, L in enumerate(links):
        if L.jtype == 2:
            d = np.diag(Minv)[adr[i]:adr[i] + 6]
            dof_invw[i] = [d[:3].mean(), d[3:].mean()]
        else:
            dof_invw[i] = [Minv[adr[i], adr[i]], 0.0]

    def body_invweight(link_idx, com_local):
        """invweight0 (trn, rot) of an original body whose CoM sits at ``com_local`` in link frame."""
        c = P0[link_idx] + R0[link_idx] @ com_local
        J = M.point_jacobian(links, ctx, link_idx, c)
        A = J @ Minv @ J.T
        return np.trace(A[:3, :3]) / 3.0, np.trace(A[3:, 3:]) / 3.0

    # ---- geoms (curated list: see DESIGN.md "collision scope")
    geoms = []   # dicts: name,type,link,pos,R,size,params,invw,tag

    def add_geom(name, gtype, link, pos, R, size, params, invw, tag=0):
        sz = np.zeros(3)
        sz[: len(size)] = size
        geoms.append(dict(name=name, type=gtype, link=link, pos=np.array(pos, float), R=np.array(R, float), size=sz, params=params, invw=invw, tag=tag))

    for s in load_surroundings(ref):
        add_geom(s["name"], s["type"], -1, s["pos"], s["R"], s["size"], s["params"], (0.0, 0.0))
    for k, o in enumerate(statics):
        prt = o["parts"][0]
        add_geom(o["name"], prt["type"], -1, o["pos"], M.quat2mat(o["quat"]), prt["size"], prt["params"], (0.0, 0.0), tag=100 + k)
    if spec["rod"]:
        rod = body_by_name["rod"]
        g = rod.geoms[0]
        li, p, R = info["rel"]["rod"]
        invw = body_invweight(li, p + R @ rod.ipos)
        add_geom("rod:geom_rb0", g.type, li, p + R @ g.pos, R @ M.quat2mat(g.quat), g.size, g.params, invw, tag=1)
    if spec.get("gripper"):
        # Collision geometry of the gripper.  The two finger-tip pads are real boxes; the finger and hand
        # MESH geoms are collided as the oriented bounding boxes of their convex hulls (DESIGN.md "deviations": no
        # general convex-hull narrow phase on this path yet).  Tag 2 = tip pad, 3 = finger hull box, 4 = hand hull box.
        def mesh_obb(name):
            import struct
            raw = open(os.path.join(ref, D3IL, "models/mj/robot/assets", name + ".stl"), "rb").read()
            ntri = struct.unpack("<I", raw[80:84])[0]
            tri = np.frombuffer(raw[84:84 + 50 * ntri], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
            v = tri["v"].reshape(-1, 3).astype(np.float64)
            return 0.5 * (v.min(0) + v.max(0)), 0.5 * (v.max(0) - v.min(0))
        for bname, tag in (("finger_joint1_tip", 2), ("finger_joint2_tip", 2), ("panda_leftfinger", 3), ("panda_rightfinger", 3), ("panda_hand", 4)):
            b = body_by_name[bname]
            li, p, R = info["rel"][bname]
            invw = body_invweight(li, p + R @ b.ipos)
            for g in b.geoms:
                if g.params["contype"] == 0 and g.params["conaffinity"] == 0:
                    continue                                       # visual copies (`*:geom1`)
                Rg = R @ M.quat2mat(g.quat)
                if g.type == "mesh":
                    c, half = mesh_obb(g.mesh)
                    add_geom(g.name + "_rb0", "box", li, p + R @ g.pos + Rg @ c, Rg, half, g.params, invw, tag=tag)
                else:
                    add_geom(g.name + "_rb0", g.type, li, p + R @ g.pos, Rg, g.size, g.params, invw, tag=tag)
    for k, o in enumerate(objs):
        li = 9 + k
        invw = body_invweight(li, links[li].ipos)
        for prt in o["parts"]:
            add_geom(o["name"] + ":geom", prt["type"], li, prt["pos"], M.quat2mat(prt["quat"]), prt["size"], prt["params"], invw, tag=10 + k)

    # ---- candidate pairs: geom filters of App. B.5 on the curated geom list
    pairs = []
    for i in range(len(geoms)):
        for j in range(i + 1, len(geoms)):
            a, b = geoms[i], geoms[j]
            if a["link"] == b["link"]:
                continue                                        # same body / both static
            if a["link"] >= 0 and b["link"] >= 0 and a["link"] < 9 and b["link"] < 9:
                # gripper self-pairs: finger bodies are children of the hand (parent filter); of the finger-finger pairs only
                # tip pad vs tip pad is kept (the hull boxes of two fingers touch face to face when the gripper is shut)
                if not (a["tag"] == 2 and b["tag"] == 2):
                    continue
            if min(a["link"], b["link"]) == -1 and max(a["link"], b["link"]) < 9 and 4 in (a["tag"], b["tag"]):
                continue                                        # hand vs table: unreachable before the fingers are 4 cm deep
            if "ground" in (a["name"], b["name"]) and max(a["link"], b["link"]) < 9:
                continue                                        # the ground, 0.92 m below the table top, is out of the arm's reach: free objects only
            pa, pb = a["params"], b["params"]
            if not ((pa["contype"] & pb["conaffinity"]) or (pb["contype"] & pa["conaffinity"])):
                continue
            if not spec.get("rod_hits_table"):
                # rod vs static geoms: unreachable at the harness' frozen tool height (Avoiding's obstacle cylinders excepted)
                if a["tag"] == 1 and a["link"] >= 0 and b["link"] == -1 and (b["tag"] < 100 or task != "avoiding"):
                    continue
                if b["tag"] == 1 and a["link"] == -1 and (a["tag"] < 100 or task != "avoiding"):
                    continue
            # order by geom type like mj_collideGeoms (lower type id first)
            g1, g2 = (i, j) if B.GEOM_TYPE_ID[a["type"]] <= B.GEOM_TYPE_ID[b["type"]] else (j, i)
            condim, fr5, solref, solimp, margin, gap = mix_params(geoms[g1]["params"], geoms[g2]["params"])
            flags = 1 if (task == "avoiding" and {a["tag"], b["tag"]} & {1} and max(a["tag"], b["tag"]) >= 100) else 0
            pairs.append(dict(g1=g1, g2=g2, condim=condim, friction=fr5, solref=solref, solimp=solimp, margin=margin, gap=gap, flags=flags))

    # ---- start-up IK (a13)
    init_qpos, ik_iters, ik_err = offline_ik(chain, ee, np.array(spec["init_tcp"], float))

    # ---- tables
    nq = 9 + 7 * nobj
    link_tab = np.zeros((nlink, B.LINK_W))
    qadr = 0
    for i, L in enumerate(links):
        r = link_tab[i]
        r[0], r[1] = L.parent, L.jtype
        r[2:5], r[5:9], r[9:12] = L.pos, L.quat, L.axis
        r[12], r[13:16] = L.mass, L.ipos
        I = L.inertia
        r[16:22] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
        r[22], r[23:25], r[25], r[26] = float(L.limited), L.range, L.damping, forcerange[i]
        r[27], r[28] = dof_invw[i]
        r[29], r[30] = qadr, adr[i]
        qadr += 7 if L.jtype == 2 else 1
    geom_tab = np.zeros((len(geoms), B.GEOM_W))
    for i, g in enumerate(geoms):
        r = geom_tab[i]
        r[0], r[1] = B.GEOM_TYPE_ID[g["type"]], g["link"]
        r[2:5], r[5:9], r[9:12] = g["pos"], M.mat2quat(g["R"]), g["size"]
        r[12] = M.geom_rbound(g["type"], g["size"])
        r[13], r[14], r[15] = g["invw"][0], g["invw"][1], g["tag"]
    pair_tab = np.zeros((len(pairs), B.PAIR_W))
    for i, p in enumerate(pairs):
        r = pair_tab[i]
        r[0], r[1], r[2] = p["g1"], p["g2"], p["condim"]
        r[3:8], r[8:10], r[10:15], r[15], r[16], r[17] = p["friction"], p["solref"], p["solimp"], p["margin"], p["gap"], p["flags"]

    cq = "CartPosQuatControllerConfig."
    ctrl = np.zeros(B.CTRL_W)
    for i in range(7):
        ctrl[B.C_IK_ORIGIN + 12 * i: B.C_IK_ORIGIN + 12 * i + 3] = chain[i][0]
        ctrl[B.C_IK_ORIGIN + 12 * i + 3: B.C_IK_ORIGIN + 12 * i + 12] = chain[i][1].reshape(-1)
    ctrl[B.C_IK_EE: B.C_IK_EE + 3] = ee[0]
    ctrl[B.C_IK_EE + 3: B.C_IK_EE + 12] = ee[1].reshape(-1)
    ctrl[B.C_PGAIN_POS: B.C_PGAIN_POS + 3] = gin[cq + "pgain_pos"]
    ctrl[B.C_PGAIN_QUAT: B.C_PGAIN_QUAT + 3] = gin[cq + "pgain_quat"]
    ctrl[B.C_PGAIN_NULL: B.C_PGAIN_NULL + 7] = gin[cq + "pgain_null"]
    ctrl[B.C_REST: B.C_REST + 7] = gin[cq + "rest_posture"]
    ctrl[B.C_JMIN: B.C_JMIN + 7] = JPOS_MIN
    ctrl[B.C_JMAX: B.C_JMAX + 7] = JPOS_MAX
    ctrl[B.C_PD_P: B.C_PD_P + 7] = gin["JointPDGains.pgain"]
    ctrl[B.C_PD_D: B.C_PD_D + 7] = gin["JointPDGains.dgain"]
    ctrl[B.C_JREG], ctrl[B.C_SVD_MIN], ctrl[B.C_SVD_MAX] = gin[cq + "J_reg"], gin[cq + "min_svd_values"], gin[cq + "max_svd_values"]
    ctrl[B.C_NUM_ITER], ctrl[B.C_LRATE], ctrl[B.C_DT] = gin[cq + "num_iter"], gin[cq + "learningRate"], DT
    assert gin[cq + "joint_filter_coefficient"] == 1.0 and all(w == 1 for w in gin[cq + "W"])
    ctrl[B.C_INIT_QPOS: B.C_INIT_QPOS + 7] = init_qpos
    li, p, R = info["rel"]["tcp"]
    assert li == 6
    ctrl[B.C_TCP_POS: B.C_TCP_POS + 3] = p
    ctrl[B.C_TCP_QUAT: B.C_TCP_QUAT + 4] = M.mat2quat(R)
    ctrl[B.C_GRAVITY: B.C_GRAVITY + 3] = [0, 0, -9.81]
    ctrl[B.C_IMPRATIO], ctrl[B.C_TOL] = 3.0, 1e-10              # models/mj/surroundings/base.xml:3-5
    ctrl[B.C_JNT_SOLREF: B.C_JNT_SOLREF + 2] = M.GEOM_DEFAULTS["solref"]
    ctrl[B.C_JNT_SOLIMP: B.C_JNT_SOLIMP + 5] = M.GEOM_DEFAULTS["solimp"]
    ctrl[B.C_INIT_TCP: B.C_INIT_TCP + 3] = spec["init_tcp"]
    ctrl[179] = np.trace(Mq0) / nv                               # stat.meaninertia

    header = dict(magic=B.MAGIC, version=B.VERSION, task_id=B.TASK_IDS[task.split("_")[0]], nlink=nlink, nobj=nobj, nq=nq, nv=nv, ngeom=len(geoms),
                  npair=len(pairs), n_substeps=spec["n_substeps"], max_steps=spec["max_steps"], obs_dim=spec["obs_dim"], act_dim=spec["act_dim"],
                  ctx_dim=7 * nobj + spec.get("nextra", 0), info_dim=spec["info_dim"], ctrl_kind=spec["ctrl_kind"], ntaskp=len(spec["taskp"]),
                  nextra=spec.get("nextra", 0), maxcon=spec.get("maxcon", 0))
    scene = B.Scene(header, link_tab, geom_tab, pair_tab, ctrl, np.array(spec["taskp"], float))
    report = dict(
        task=task, header=header, links=[L.name for L in links], link_mass=[L.mass for L in links],
        geoms=[g["name"] for g in geoms], pairs=[(geoms[p["g1"]]["name"], geoms[p["g2"]]["name"], int(p["condim"])) for p in pairs],
        init_qpos=init_qpos.tolist(), offline_ik_iters=ik_iters, offline_ik_err=ik_err,
        tcp_at_init=ik_fk(chain, ee, init_qpos)[0].tolist(), meaninertia=float(ctrl[179]),
        geom_invweight0={g["name"]: g["invw"] for g in geoms},
    )
    return scene, report


def export_contexts(ref, out_dir):
    """Evaluation contexts shipped by the reference -> small committed .npy fixtures ([n, nobj, 7] = xyz + quat wxyz per object)."""
    os.makedirs(out_dir, exist_ok=True)
    c = pickle.load(open(os.path.join(ref, "environments/dataset/data/pushing/test_contexts.pkl"), "rb"))
    # BlockContextManager.set_context (pushing.py:99-113): boxes placed at z = 0.0 with the context quaternion
    arr = np.array([[[r[0], r[1], 0.0, *q], [g[0], g[1], 0.0, *q2]] for r, q, g, q2 in c], dtype=np.float64)
    np.save(os.path.join(out_dir, "pushing_test_contexts.npy"), arr)
    raw = np.array([[*r, *q, *g, *q2] for r, q, g, q2 in c], dtype=np.float64)
    np.save(os.path.join(out_dir, "pushing_test_contexts_raw.npy"), raw)
    # Aligning (aligning.py:109-123): box at [x, y, 0] + quat, then the target pose [tx, ty, 0] + quat (written to model.body_pos/quat)
    c = pickle.load(open(os.path.join(ref, "environments/dataset/data/aligning/test_contexts.pkl"), "rb"))
    arr = np.array([[[p[0], p[1], 0.0, *q], [tp[0], tp[1], 0.0, *tq]] for p, q, tp, tq in c], dtype=np.float64)
    np.save(os.path.join(out_dir, "aligning_test_contexts.npy"), arr)
    # Stacking (stacking.py:99-129): red / green / blue boxes at [x, y, 0] + quat; the 4th entry (target) is never applied
    c = pickle.load(open(os.path.join(ref, "environments/dataset/data/stacking/test_contexts.pkl"), "rb"))
    arr = np.array([[[e[0][0], e[0][1], 0.0, *e[1]] for e in ctx[:3]] for ctx in c], dtype=np.float64)
    np.save(os.path.join(out_dir, "stacking_test_contexts.npy"), arr)
    # Inserting: no dataset / contexts in the reference; drawn from BlockContextManager's three boxes (gate_insertion.py:49-62), z = 0
    rng = np.random.default_rng(4242)
    lows, highs = np.array([[0.35, -0.2], [0.55, -0.1], [0.35, 0.0]]), np.array([[0.5, -0.15], [0.7, -0.05], [0.5, 0.05]])
    out = np.zeros((60, 3, 7))
    for n in range(60):
        xy = rng.uniform(lows, highs).astype(np.float32)
        ang = rng.uniform(-90, 90, 3).astype(np.float32) * np.pi / 180
        for i in range(3):
            out[n, i] = [xy[i, 0], xy[i, 1], 0.0, np.cos(ang[i] / 2), 0.0, 0.0, np.sin(ang[i] / 2)]
    np.save(os.path.join(out_dir, "inserting_contexts.npy"), out)
    # Sorting: `<k>_test_contexts.pkl` are NOT shipped -> drawn here from the six BlockContextManager boxes
    # (sorting.py:52-74,88-119) with our own seeded generator; boxes are placed at z = 0.05 (sorting.py:130-181)
    lows = np.array([[0.4, -0.15], [0.4, -0.05], [0.4, 0.05], [0.55, -0.15], [0.55, -0.05], [0.55, 0.05]])
    highs = np.array([[0.5, -0.1], [0.5, 0.0], [0.5, 0.1], [0.65, -0.1], [0.65, 0.0], [0.65, 0.1]])
    for k in (2, 4, 6):
        rng = np.random.default_rng(42 + k)
        out = np.zeros((60, k, 7))
        for n in range(60):
            xy = rng.uniform(lows, highs).astype(np.float32)
            ang = rng.uniform(-90, 90, 6).astype(np.float32)
            order = rng.permutation(6)
            for i in range(k):
                j = order[i]
                a = float(ang[j]) * np.pi / 180
                out[n, i] = [xy[j, 0], xy[j, 1], 0.05, np.cos(a / 2), 0.0, 0.0, np.sin(a / 2)]     # euler2quat([0, 0, a]) (geometric_transformation.py:73-89)
        np.save(os.path.join(out_dir, f"sorting_{k}_contexts.npy"), out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "scenes"))
    ap.add_argument("--tasks", nargs="*", default=["avoiding", "pushing", "sorting_2", "sorting_4", "sorting_6", "aligning", "stacking", "inserting"])
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for t in a.tasks:
        scene, report = compile_task(t, a.ref)
        with open(os.path.join(a.out, f"{t}.d3sc"), "wb") as f:
            f.write(scene.pack())
        with open(os.path.join(a.out, f"{t}.json"), "w") as f:
            json.dump(report, f, indent=1, default=lambda o: o.tolist() if hasattr(o, "tolist") else o)
        print(t, "nq/nv", report["header"]["nq"], report["header"]["nv"], "pairs", len(report["pairs"]), "ik iters", report["offline_ik_iters"], "tcp", np.round(report["tcp_at_init"], 6))
    export_contexts(a.ref, os.path.join(a.out, "..", "data"))


if __name__ == "__main__":
    main()
