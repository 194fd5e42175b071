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
