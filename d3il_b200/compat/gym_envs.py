"""Gym-shaped single-env classes with the reference's names and call signatures, thin over ``BatchedEnv(task, 1)`` through the
host-buffer C ABI: ``Env(render=False).start()``, ``reset(random=True, context=None)``, ``step(action)``,
``robot_state()``, ``robot.current_c_pos``, ``manager.sample()``.  With ``d3il_b200/compat`` on ``sys.path`` they are importable
at the reference's own paths (``envs.gym_pushing_env.gym_pushing.envs.pushing.Block_Push_Env`` ...), so the reference's
*unmodified* ``simulation/*_sim.py`` rollout loops run on this backend (one env per process, as the reference does).

Contexts arrive in the reference's pickle format — per object ``[x, y, yaw_deg], quat(wxyz)`` (``pushing.py:99-113``,
``aligning.py:109-123``, ``stacking.py:99-129``, ``sorting.py:121-187``) — and are converted to the env's ``[n_obj, 7]`` pose
rows (xyz + quat; z is the task's spawn height).  ``reset(random=True)`` draws from the task's ``BlockContextManager`` boxes
(``pushing.py:53-58``, ``aligning.py:62-68``, ``sorting.py:52-74``, ``stacking.py:52-66``) with ``numpy.random`` (the
reference's ``gym.spaces.Box.sample`` is unseeded; here ``seed`` is honoured).
"""
from __future__ import annotations

import math

import numpy as np


def euler2quat_z(yaw_deg: float) -> np.ndarray:
    """``euler2quat([0, 0, yaw])`` (utils/geometric_transformation.py:73-89), wxyz."""
    a = float(yaw_deg) * math.pi / 180.0
    return np.array([math.cos(a / 2), 0.0, 0.0, math.sin(a / 2)])


class _Box:
    def __init__(self, low, high, rng):
        self.low, self.high, self.rng = np.asarray(low, float), np.asarray(high, float), rng

    def sample(self):
        return self.rng.uniform(self.low, self.high).astype(np.float32)


class ContextManager:
    """``BlockContextManager``: ``sample()`` returns a context in the reference's format for the task."""

    SPACES = {
        "pushing": [([0.4, -0.15, -90], [0.5, 0, 90]), ([0.55, -0.15, -90], [0.65, 0, 90])],
        "aligning": [([0.4, -0.25, -90], [0.6, -0.1, 90]), ([0.4, 0.2, -90], [0.6, 0.35, 90])],
        "stacking": [([0.35, -0.25, -90], [0.45, -0.15, 90]), ([0.35, -0.1, -90], [0.45, 0, 90]), ([0.55, -0.2, -90], [0.6, 0, 90]),
                     ([0.4, 0.15, -90], [0.6, 0.25, 90])],
        "sorting": [([0.4, -0.15, -90], [0.5, -0.1, 90]), ([0.4, -0.05, -90], [0.5, 0.0, 90]), ([0.4, 0.05, -90], [0.5, 0.1, 90]),
                    ([0.55, -0.15, -90], [0.65, -0.1, 90]), ([0.55, -0.05, -90], [0.65, 0.0, 90]), ([0.55, 0.05, -90], [0.65, 0.1, 90])],
        "inserting": [([0.4, -0.2, -90], [0.45, -0.15, 90]), ([0.5, -0.2, -90], [0.55, -0.15, 90]), ([0.6, -0.2, -90], [0.65, -0.15, 90])],
    }

    def __init__(self, kind: str, n_obj: int, seed: int = 42, index: int = 0):
        self.kind, self.n_obj, self.index = kind, n_obj, index
        self.rng = np.random.default_rng(seed)
        self.spaces = [_Box(lo, hi, self.rng) for lo, hi in self.SPACES[kind]]
        self.context = None

    def set_index(self, index):
        self.index = index

    def sample(self):
        pq = []
        for sp in self.spaces:
            p = sp.sample()
            pq.append([p, euler2quat_z(p[-1])])
        if self.kind == "sorting":                       # sorting.py:88-119: six draws, shuffled; the env uses the first num_boxes
            self.rng.shuffle(pq)
            return pq
        if self.kind in ("stacking", "inserting"):       # stacking.py:68-97, gate_insertion.py: [[pos, quat], ...] (Stacking incl. the target)
            return pq
        return [x for p, q in pq for x in (p, q)]        # pushing / aligning: flat [pos, quat, pos, quat]

    def to_poses(self, context, z: float) -> np.ndarray:
        """Reference-format context -> [n_obj (+ target), 7] rows xyz + quat."""
        if isinstance(context, np.ndarray) and context.ndim == 2 and context.shape[1] == 7:
            return context.astype(np.float64)
        items = list(context)
        if len(items) and np.ndim(items[0]) == 1 and len(items[0]) == 3 and self.kind in ("pushing", "aligning"):
            items = [[items[2 * i], items[2 * i + 1]] for i in range(len(items) // 2)]
        rows = [[float(p[0]), float(p[1]), z, *np.asarray(q, float)] for p, q in items]
        return np.array(rows[: self.n_obj], dtype=np.float64)


class _Robot:
    def __init__(self, env):
        self._env = env

    @property
    def current_c_pos(self):
        return self._env._benv.robot_state_host()[0].astype(np.float64)

    @property
    def current_j_pos(self):
        return self._env._benv.joint_state_host()[0, :7].astype(np.float64)

    @property
    def gripper_width(self):
        return float(self._env._benv.joint_state_host()[0, 7])


class _Scene:
    def __init__(self, env):
        self._env = env

    def get_obj_pos(self, obj=None, index: int = 0):
        return self._env._poses()[index, :3]

    def get_obj_quat(self, obj=None, index: int = 0):
        return self._env._poses()[index, 3:]


class SingleEnv:
    """Common part of the six env classes.  ``task`` = compiled scene, ``kind`` = context family."""
    task, kind, spawn_z = "pushing", "pushing", 0.0

    def __init__(self, render: bool = False, device: int | None = None, seed: int = 42, **_ignored):
        if render:
            raise NotImplementedError("the batched CUDA backend has no viewer (render=True)")
        self._device, self._seed = device, seed
        self._benv = None
        self.robot, self.scene = _Robot(self), _Scene(self)
        self.manager = None
        self.success = False
        self.env_step_counter = 0
        self.episode = 0

    # -- lifecycle
    def start(self):
        import torch
        from ..batched_env import BatchedEnv
        dev = torch.cuda.current_device() if self._device is None else self._device
        self._benv = BatchedEnv(self.task, 1, dev)
        n_obj = self._benv.scene.header["nobj"]
        self.manager = ContextManager(self.kind, n_obj, self._seed)
        self.max_steps_per_episode = self._benv.max_steps_per_episode
        return self

    def close(self):
        if self._benv is not None:
            self._benv.close()
            self._benv = None

    def _poses(self):
        import torch
        p = self._benv.object_poses()
        torch.cuda.synchronize()
        return p[0].cpu().numpy().astype(np.float64)

    def _ctx_rows(self, context):
        rows = self.manager.to_poses(context, self.spawn_z)
        return rows

    def reset(self, random: bool = True, context=None):
        if self._benv is None:
            self.start()
        self.episode += 1
        self.env_step_counter = 0
        self.success = False
        if self._benv.ctx_dim == 0:
            return self._benv.reset_host()[0]
        ctx = self.manager.sample() if random else context
        self.manager.context = ctx
        rows = self._ctx_rows(ctx)
        return self._benv.reset_host(rows.reshape(1, -1).astype(np.float32))[0]

    def robot_state(self):
        return self._benv.robot_state_host()[0].astype(np.float64)

    def _info(self, info_row):
        return {"mode": int(info_row[1]), "success": bool(info_row[0]), "mean_distance": float(info_row[2])}

    def step(self, action, gripper_width=None, desired_vel=None, desired_acc=None):
        a = np.asarray(action, dtype=np.float32).reshape(1, -1)
        obs, rew, done, info = self._benv.step_host(a)
        self.env_step_counter += 1
        self.success = bool(info[0, 0])
        self._status = int(info[0, -1])
        return obs[0], float(rew[0]), bool(done[0]), self._info(info[0])


class Block_Push_Env(SingleEnv):
    """``envs/gym_pushing_env/gym_pushing/envs/pushing.py:171`` — info = {mode, success, mean_distance}."""
    task, kind = "pushing", "pushing"


class ObstacleAvoidanceEnv(SingleEnv):
    """``envs/gym_avoiding_env/.../avoiding.py:52`` — step returns info = (mode_encoding[9], success) (:168-171)."""
    task, kind = "avoiding", "pushing"

    def _info(self, info_row):
        return (np.asarray(info_row[1:10], dtype=np.float32).copy(), bool(info_row[0]))


class Robot_Push_Env(SingleEnv):
    """``envs/gym_aligning_env/.../aligning.py:129`` — 3-D action; the context carries the target pose as well."""
    task, kind = "aligning", "aligning"

    def _ctx_rows(self, context):
        items = list(context)
        pos, quat, tpos, tquat = items
        return np.array([[pos[0], pos[1], 0.0, *np.asarray(quat, float)], [tpos[0], tpos[1], 0.0, *np.asarray(tquat, float)]], dtype=np.float64)


class Sorting_Env(SingleEnv):
    """``envs/gym_sorting_env/.../sorting.py:193`` — info = {mode (packed bits), success, min_inds}."""
    kind, spawn_z = "sorting", 0.05

    def __init__(self, render: bool = False, num_boxes: int = 2, max_steps_per_episode: int = 500, if_vision: bool = False, **kw):
        if if_vision:
            raise NotImplementedError("vision observations are outside the batched state-based path")
        super().__init__(render, **kw)
        self.num_boxes = num_boxes
        self.task = f"sorting_{num_boxes}"

    def _ctx_rows(self, context):
        return np.array([[p[0], p[1], self.spawn_z, *np.asarray(q, float)] for p, q in list(context)[: self.num_boxes]], dtype=np.float64)

    def _info(self, info_row):
        return {"mode": int(info_row[1]), "success": bool(info_row[0]), "min_inds": int(info_row[2])}


class CubeStacking_Env(SingleEnv):
    """``envs/gym_stacking_env/.../stacking.py:135`` — 8-D action (7 joint set-points + gripper command); ``robot_state()``
    returns (joint positions + gripper width, joint positions, tcp quaternion) (:218-226)."""
    task, kind = "stacking", "stacking"

    def _ctx_rows(self, context):
        return np.array([[p[0], p[1], 0.0, *np.asarray(q, float)] for p, q in list(context)[:3]], dtype=np.float64)

    def robot_state(self):
        j = self._benv.joint_state_host()[0].astype(np.float64)
        return j, j[:7].copy(), np.array([0.0, 1.0, 0.0, 0.0])

    def _info(self, info_row):
        code, length = int(info_row[1]), int(info_row[3])
        mode = "".join("rgb"[(code // 4 ** k) % 4 - 1] for k in range(length))
        return {"mode": mode, "success": bool(info_row[0]), "success_1": len(mode) > 0, "success_2": len(mode) > 1, "mean_distance": float(info_row[2])}


class Gate_Insertion_Env(SingleEnv):
    """``envs/gym_inserting_env/.../gate_insertion.py:154`` — info = {success, mode, mean_distance}."""
    task, kind = "inserting", "inserting"

    def _ctx_rows(self, context):
        if isinstance(context, np.ndarray):
            return context.reshape(-1, 7).astype(np.float64)
        return np.array([[p[0], p[1], 0.0, *np.asarray(q, float)] for p, q in list(context)[:3]], dtype=np.float64)
