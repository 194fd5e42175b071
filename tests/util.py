"""Shared helpers for the parity tests: scripted action streams and oracle roll-outs."""
import numpy as np

from d3il_b200.scene.blob import load_scene
from oracle.oracle import OracleEnv


def scripted_push_actions(ctx, tcp0, n_steps=110, speed=0.006, approach_steps=60):
    """Drive the rod behind box 1 (-y side), then push it towards +y: exercises rod-box and box-table contacts."""
    des = np.array(tcp0, dtype=np.float64).copy()
    box = ctx[0, :2]
    behind = np.array([box[0], box[1] - 0.08])
    out = []
    for k in range(n_steps):
        goal = behind if k < approach_steps else np.array([box[0], 0.3])
        d = goal - des[:2]
        n = np.linalg.norm(d)
        if n > speed:
            d = d / n * speed
        des[:2] += d
        out.append(np.concatenate([des, [0, 1, 0, 0]]))
    return np.array(out)


def random_walk_actions(tcp0, n_steps, seed, lo=(0.3, -0.45), hi=(0.8, 0.45), step=0.01):
    """BASELINE.md §3 synthetic stream: dxy ~ U(-0.01, 0.01)^2 accumulated on the desired xy, clipped to the workspace."""
    rng = np.random.default_rng(seed)
    des = np.array(tcp0, dtype=np.float64).copy()
    out = []
    for _ in range(n_steps):
        des[:2] = np.clip(des[:2] + rng.uniform(-step, step, 2), lo, hi)
        out.append(np.concatenate([des, [0, 1, 0, 0]]))
    return np.array(out)


def oracle_rollout_states(task, ctx, actions, every=1):
    """Roll the oracle and return the flat state BEFORE each env step plus the step outputs."""
    blob, sc = load_scene(task)
    env = OracleEnv(blob, sc.header)
    env.reset(ctx)
    states, outs = [], []
    for k, a in enumerate(actions):
        if k % every == 0:
            states.append(env.get_state())
        outs.append(env.step(a))
    return env, np.array(states), outs


def with_setpoint(state, sc, action, grip=0.04):
    """Flat state with the Cartesian set-point installed (what GymEnvWrapper.step does before its substeps)."""
    s = state.copy()
    o = sc.header["nq"] + 2 * sc.header["nv"]
    s[o + 23:o + 26] = action[:3]
    s[o + 26:o + 30] = action[3:] / np.linalg.norm(action[3:])
    s[o + 44 + 1] = 1      # ctrl_mode
    s[o + 44 + 2] = grip   # gripper set-point
    s[o + 44 + 3] = 0      # grasp flag
    return s


def with_joint_setpoint(state, sc, action, open_thresh=0.075):
    """Flat state with the joint-space set-point + gripper command of CubeStacking_Env.step installed (stacking.py:337-346):
    joint targets as float32 (the device action buffer), zero desired velocity, joint-PD mode 2."""
    s = state.copy()
    o = sc.header["nq"] + 2 * sc.header["nv"]
    s[o + 30:o + 37] = np.asarray(action[:7], dtype=np.float32).astype(np.float64)
    s[o + 37:o + 44] = 0
    is_open = float(np.float32(action[7])) > open_thresh
    s[o + 44 + 1] = 2
    s[o + 44 + 2] = 0.04 if is_open else 0.0
    s[o + 44 + 3] = 0.0 if is_open else 1.0
    return s


TASK_CONTEXT_FILES = {"pushing": "pushing_test_contexts", "sorting_2": "sorting_2_contexts", "sorting_4": "sorting_4_contexts",
                      "sorting_6": "sorting_6_contexts", "aligning": "aligning_test_contexts"}


def task_contexts(task):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return np.load(os.path.join(root, "d3il_b200", "data", TASK_CONTEXT_FILES[task] + ".npy"))


def scripted_task_actions(task, ctx, tcp0, n_steps=60, speed=0.008):
    """Drive the rod from the start pose towards a point 10 cm beyond the first object (+y), lowering the tool for
    Aligning (3-D action) so the rod meets the box walls: free motion, impact, pushing, box-box contacts."""
    des = np.array(tcp0, dtype=np.float64).copy()
    goal = np.asarray(ctx, dtype=np.float64).reshape(-1, 7)[0, :2] + np.array([0.0, 0.1])
    out = []
    for _ in range(n_steps):
        d = goal - des[:2]
        n = np.linalg.norm(d)
        if n > 1e-12:
            des[:2] += d / n * min(speed, n)
        if task == "aligning":
            des[2] += np.clip(0.13 - des[2], -speed, speed)
        out.append(np.concatenate([des, [0, 1, 0, 0]]))
    return np.array(out)


def step_errors(ref, got, nq, nv):
    """(qpos error in units of rtol 1e-4 + atol 5e-6, qvel error in units of rtol 1e-3 + atol 2e-4)."""
    dq = np.abs(got[:nq] - ref[:nq]) / (1e-4 * np.abs(ref[:nq]) + 5e-6)
    dv = np.abs(got[nq:nq + nv] - ref[nq:nq + nv]) / (1e-3 * np.abs(ref[nq:nq + nv]) + 2e-4)
    return dq.max(), dv.max()


TASK_CONTEXT_FILES["stacking"] = "stacking_test_contexts"
TASK_CONTEXT_FILES["inserting"] = "inserting_contexts"


def _panda_ik(sc):
    """Damped-least-squares IK on the scene blob's URDF chain (tool pointing down, yaw about z): joint targets for the
    scripted Stacking grasp."""
    from d3il_b200.scene import compile as CP
    ctrl = sc.ctrl
    chain = [(ctrl[12 * i:12 * i + 3], ctrl[12 * i + 3:12 * i + 12].reshape(3, 3)) for i in range(7)]
    ee = (ctrl[84:87], ctrl[87:96].reshape(3, 3))

    def ik(target, q0, yaw):
        q = q0.copy()
        dq = np.array([0.0, np.cos(yaw / 2), np.sin(yaw / 2), 0.0])
        for _ in range(100):
            p, quat, J = CP.ik_fk(chain, ee, q)
            if np.linalg.norm(quat - dq) > np.linalg.norm(quat + dq):
                quat = -quat
            err = np.concatenate([target - p, CP.quat_error(quat, dq)])
            if np.linalg.norm(err) < 1e-10:
                break
            q = q + J.T @ np.linalg.solve(J @ J.T + 1e-8 * np.eye(6), err)
        return q
    return ik


def scripted_grasp_actions(sc, ctx, tcp0, q0, obs0, lift=0.2):
    """Stacking: move above the red box, descend with the fingers aligned to its yaw, close the gripper, lift.
    8-D actions (7 joint set-points + gripper command).  Exercises pad / finger-hull contacts (condim 4) and a held box."""
    ik = _panda_ik(sc)
    red = np.asarray(ctx, dtype=np.float64).reshape(-1, 7)[0, :3]
    yaw = float(np.arctan(obs0[3]))
    wps = [(np.array([red[0], red[1], 0.15]), 0.08, 14), (np.array([red[0], red[1], 0.018]), 0.08, 22),
           (np.array([red[0], red[1], 0.018]), 0.0, 10), (np.array([red[0], red[1], lift]), 0.0, 20)]
    p, q, out = np.array(tcp0, float).copy(), np.array(q0, float).copy(), []
    for tgt, grip, n in wps:
        for k in range(n):
            p = p + (tgt - p) / (n - k)
            q = ik(p, q, yaw)
            out.append(np.concatenate([q, [grip]]))
    return np.array(out)


# ---- tail analysis: fp32 kernel vs fp64 oracle on ill-conditioned steps -----------------------------------------------
def units(ref, got, nq, nv, start=None):
    """Error of one state in units of the tolerance box: qpos (rtol 1e-4 + atol 5e-6), qvel (rtol 1e-3 + atol 2e-4; with
    `start` given — a single tick — the velocity box also scales with the tick's largest velocity change, see
    tests/test_gpu_parity.py::test_single_tick_teacher_forced)."""
    dq = np.abs(got[:nq] - ref[:nq]) / (1e-4 * np.abs(ref[:nq]) + 5e-6)
    if start is None:
        dv = np.abs(got[nq:nq + nv] - ref[nq:nq + nv]) / (1e-3 * np.abs(ref[nq:nq + nv]) + 2e-4)
    else:
        acc = np.abs(ref[nq:nq + nv] - start[nq:nq + nv]).max()
        dv = np.abs(got[nq:nq + nv] - ref[nq:nq + nv]) / (1e-4 * np.abs(ref[nq:nq + nv]) + 2e-4 * acc + 5e-6)
    return float(dq.max()), float(dv.max())


def explain_tail(task, starts, actions, refs, gots, nq, nv, single_tick=False):
    """For every teacher-forced sample whose fp32 result leaves the tolerance box, run the SAME kernel source compiled for
    fp64 on the host (tests/emu, -DD3IL_REAL=double) from the same start state.  Returns a list of dicts
    (index, fp32 units, fp64 units, absolute fp32 qpos / qvel error): if the fp64 build sits inside the box where the fp32
    build does not, the excursion is precision (conditioning of that step), not logic."""
    from d3il_b200.scene.blob import load_scene
    from tests.emu.emu import EmuEnv
    blob, sc = load_scene(task)
    emu = None
    out = []
    for i, (s0, ref, got) in enumerate(zip(starts, refs, gots)):
        u32 = units(ref, got, nq, nv, s0 if single_tick else None)
        if max(u32) <= 1.0:
            continue
        if emu is None:
            emu = EmuEnv(blob, sc.header, "f64")
            emu.reset(None if not sc.header["ctx_dim"] else np.tile([0.5, 0.0, 0.05, 1.0, 0.0, 0.0, 0.0], sc.header["ctx_dim"] // 7))
        emu.set_state(s0)
        if single_tick:
            emu.substep(1)
        else:
            emu.step(actions[i])
        g64 = emu.get_state()
        out.append(dict(i=i, u32=u32, u64=units(ref, g64, nq, nv, s0 if single_tick else None),
                        dq=float(np.abs(got[:nq] - ref[:nq]).max()), dv=float(np.abs(got[nq:nq + nv] - ref[nq:nq + nv]).max())))
    return out


# ---- protocol (iv): scripted FEEDBACK policies (functions of the observation only, vectorised over envs) --------------
def push_policy(tcp_xy, des_xy, box_xy, goal_xy, speed=0.008, standoff=0.075):
    """One step of a two-phase pusher for N envs: get behind the box on the line goal -> box (going around it when the rod
    is on the wrong side), then push it towards the goal.  All arguments [N, 2]; returns the new desired xy.  A pure
    function of (observation, last desired xy): the GPU batch and the oracle envs run the same law closed loop."""
    u = goal_xy - box_xy
    dist = np.linalg.norm(u, axis=1, keepdims=True)
    u = u / np.maximum(dist, 1e-9)
    behind = box_xy - standoff * u
    rel = tcp_xy - box_xy
    along = np.sum(rel * u, axis=1, keepdims=True)               # > 0: rod between box and goal (wrong side)
    side = rel - along * u
    side_n = np.linalg.norm(side, axis=1, keepdims=True)
    perp = np.concatenate([-u[:, 1:2], u[:, 0:1]], 1)
    sgn = np.where(np.sum(side * perp, axis=1, keepdims=True) >= 0, 1.0, -1.0)
    behind_box = along < -0.02                                     # contact happens at along ~ -0.045 (box half width + rod radius)
    lined_up = behind_box & (side_n < 0.025)
    beside = box_xy - standoff * u + sgn * perp * 0.1              # way-point beside-and-behind: clears the box while going around
    goal_pt = np.where(lined_up, box_xy + 0.03 * u, np.where(behind_box | (side_n > 0.085), behind, beside))
    d = goal_pt - des_xy
    n = np.linalg.norm(d, axis=1, keepdims=True)
    return des_xy + d / np.maximum(n, 1e-9) * np.minimum(n, speed)


def _iv_plan(task, ci):
    """Which (box, goal) sequence the scripted policy follows in context `ci` (spreads the rollouts over the task's modes)."""
    if task == "pushing":          # pushing.py:341-377: modes = visiting order of (box, target) pairs
        G1, G2 = np.array([0.42, 0.3]), np.array([0.63, 0.3])
        return [[(0, G1), (1, G2)], [(1, G2), (0, G1)], [(0, G2), (1, G1)], [(1, G1), (0, G2)]][ci % 4]
    if task.startswith("sorting"):   # sorting.py: red box -> red bin (x 0.3..0.5), blue box -> blue bin (x 0.525..0.725), bins at y 0.22..0.41
        k = int(task.split("_")[1])   # observation order: k/2 red boxes, then k/2 blue ones
        R, B = np.array([0.4, 0.33]), np.array([0.625, 0.33])
        return [[(0, R)], [(k // 2, B)]][ci % 2]      # one box into its bin, then hold (getting behind the next box between the bin walls needs a planner)
    if task == "aligning":
        return None
    raise ValueError(task)


def iv_policy_step(task, ci, obs, des, phase):
    """One closed-loop policy step for ONE env (numpy); returns (new desired pose, new phase).  obs layouts: SURVEY a10."""
    tcp = obs[None, 0:2].astype(np.float64)
    if task == "aligning":
        box, goal = obs[None, 3:5].astype(np.float64), obs[None, 10:12].astype(np.float64)
        d2 = push_policy(tcp, des[None, :2], box, goal)[0]
        z = des[2] + np.clip(0.13 - des[2], -0.008, 0.008)
        return np.array([d2[0], d2[1], z]), phase
    plan = _iv_plan(task, ci)
    stride = 3
    nbox = int(task.split("_")[1]) if task.startswith("sorting") else 2
    boxes = [obs[2 + stride * b:4 + stride * b].astype(np.float64) for b in range(nbox)]
    while phase < len(plan) and np.linalg.norm(boxes[plan[phase][0]] - plan[phase][1]) < 0.03:
        phase += 1
    if phase >= len(plan):
        return des.copy(), phase
    b, g = plan[phase]
    d2 = push_policy(tcp, des[None, :2], boxes[b][None], g[None])[0]
    return np.concatenate([d2, des[2:]]), phase


def iv_oracle_episode(args):
    """Closed-loop scripted episode of one context on the fp64 oracle (process-pool worker).  Returns (info row, steps)."""
    task, ci, max_steps = args
    from d3il_b200.scene.blob import load_scene
    from oracle.oracle import OracleEnv
    blob, sc = load_scene(task)
    o = OracleEnv(blob, sc.header)
    obs = o.reset(task_contexts(task)[ci])
    des, phase = o.robot_state().copy(), 0
    info = None
    for k in range(max_steps):
        des, phase = iv_policy_step(task, ci, obs, des, phase)
        obs, r, done, info = o.step(np.concatenate([des, [0, 1, 0, 0]]))
        if done:
            break
    return np.array(info), k + 1


def iv_stacking_episode(ci):
    """Scripted grasp-and-lift of the red box in Stacking context `ci` on the fp64 oracle (process-pool worker): the
    open-loop joint-space script is a function of the context and the reset observation only.  Returns (actions, info row
    after the last step, final red-box height)."""
    blob, sc = load_scene("stacking")
    ctx = task_contexts("stacking")[ci]
    o = OracleEnv(blob, sc.header)
    obs0 = o.reset(ctx)
    acts = scripted_grasp_actions(sc, ctx, o.robot_state(), o.joint_state()[:7], obs0)
    info = None
    for a in acts:
        _, _, _, info = o.step(a)
    return acts, np.array(info), float(o.get_state()[9 + 2])
