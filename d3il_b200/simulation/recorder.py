"""Trajectory recording in the reference's dataset wire format.

The datasets the agents train on are pickled ``env_state`` dicts written by the reference's loggers while a human
teleoperates (``core/logger.py``), read back by ``environments/dataset/*_dataset.py``:

* every task: ``env_state['robot']['des_c_pos' | 'c_pos']`` [T, 3] and ``env_state['<object>']['pos' | 'quat']`` [T, 3 | 4]
  (``pushing_dataset.py:52-77``, ``sorting_dataset.py``, ``avoiding_dataset.py``);
* Aligning additionally ``env_state['target-box']['pos' | 'quat']`` (``aligning_dataset.py:62-70``) — the per-context target pose;
* Stacking additionally ``env_state['robot']['des_j_pos', 'des_j_vel', 'des_c_quat', 'c_quat', 'j_pos', 'j_vel', 'gripper_width']``
  (``stacking_dataset.py:92-104``).

``TrajectoryRecorder`` collects the same arrays from a ``BatchedEnv`` rollout so GPU rollouts can be fed to the existing
``*_Dataset`` classes (one dict per env instance).  In the joint-space scene the commanded Cartesian pose is the forward
kinematics of the commanded joints (URDF chain of the scene blob).
"""
from __future__ import annotations

import pickle

import numpy as np
import torch

# logger keys per task: free objects in qpos order, then static targets (name -> xyz)
OBJECT_KEYS = {
    "pushing": (["red-box", "green-box"], {"red-target": [0.42, 0.3, 0.0], "green-target": [0.63, 0.3, 0.0]}),     # pushing.py:220-253
    "aligning": (["push-box"], {}),
    "sorting_2": (["red-box1", "blue-box1"], {}),
    "sorting_4": (["red-box1", "red-box2", "blue-box1", "blue-box2"], {}),
    "sorting_6": (["red-box1", "red-box2", "red-box3", "blue-box1", "blue-box2", "blue-box3"], {}),
    "stacking": (["red-box", "green-box", "blue-box"], {"target-box": [0.5, 0.2, 0.0]}),
    "inserting": (["push-box1", "push-box2", "push-box3"], {}),
    "avoiding": ([], {}),
}


def chain_fk(ctrl: np.ndarray, q: np.ndarray):
    """Forward kinematics of the URDF chain in the scene's controller table (``core/Model.py:37-66``) for joint angles
    q [..., 7]: returns (pos [..., 3], quat wxyz [..., 4]), vectorised over the leading axes."""
    q = np.asarray(q, dtype=np.float64)
    if q.ndim == 1:
        p1, q1 = chain_fk(ctrl, q[None])
        return p1[0], q1[0]
    lead = q.shape[:-1]
    R = np.broadcast_to(np.eye(3), lead + (3, 3)).copy()
    p = np.zeros(lead + (3,))
    for i in range(7):
        o, oR = ctrl[12 * i:12 * i + 3], ctrl[12 * i + 3:12 * i + 12].reshape(3, 3)
        p = p + R @ o
        c, s = np.cos(q[..., i]), np.sin(q[..., i])
        Rz = np.zeros(lead + (3, 3))
        Rz[..., 0, 0], Rz[..., 0, 1], Rz[..., 1, 0], Rz[..., 1, 1], Rz[..., 2, 2] = c, -s, s, c, 1.0
        R = R @ oR @ Rz
    p = p + R @ ctrl[84:87]
    R = R @ ctrl[87:96].reshape(3, 3)
    w = np.sqrt(np.maximum(1 + R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2], 1e-12)) / 2      # tool pointing down: trace > -1 away from the branch cut
    small = w < 1e-3
    quat = np.stack([w, (R[..., 2, 1] - R[..., 1, 2]) / (4 * np.where(small, 1, w)), (R[..., 0, 2] - R[..., 2, 0]) / (4 * np.where(small, 1, w)),
                     (R[..., 1, 0] - R[..., 0, 1]) / (4 * np.where(small, 1, w))], -1)
    if small.any():       # rotation by ~pi (the Panda's tool-down pose: quat ~ [0, 1, 0, 0]): take the dominant axis from the diagonal
        d = np.stack([R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]], -1)
        k = d.argmax(-1)
        for idx in zip(*np.nonzero(small)):
            Rm, j = R[idx], int(k[idx])
            v = np.zeros(4)
            v[1 + j] = np.sqrt(max(1 + 2 * Rm[j, j] - np.trace(Rm), 0)) / 2
            a, b = (j + 1) % 3, (j + 2) % 3
            v[0] = (Rm[b, a] - Rm[a, b]) / (4 * v[1 + j])
            v[1 + a] = (Rm[a, j] + Rm[j, a]) / (4 * v[1 + j])
            v[1 + b] = (Rm[b, j] + Rm[j, b]) / (4 * v[1 + j])
            quat[idx] = v
    quat = quat / np.linalg.norm(quat, axis=-1, keepdims=True)
    return p, quat


class TrajectoryRecorder:
    def __init__(self, env):
        self.env = env
        self.keys, self.statics = OBJECT_KEYS[env.task]
        self.joint_space = env.act_dim == 8
        self._act, self._kin, self._obj, self._alive, self._target = [], [], [], [], []

    def record(self, action: torch.Tensor, alive: torch.Tensor | None = None):
        """Call once per env step, before ``env.step(action)``: logs the command, the measured robot state and the object poses."""
        self._act.append(action.detach().clone())
        self._kin.append(self.env.robot_kinematics().detach().clone())
        self._obj.append(self.env.object_poses().detach().clone())
        if self.env.task == "aligning":
            self._target.append(self.env.obs[:, 10:17].detach().clone())          # target pose of the context (aligning.py:205-235)
        self._alive.append(torch.ones(self.env.n_envs, dtype=torch.bool, device=action.device) if alive is None else alive.detach().clone().bool())

    def env_states(self) -> list[dict]:
        """One ``env_state`` dict per env, truncated to the steps in which that env was alive."""
        act, kin = torch.stack(self._act, 1).cpu().numpy().astype(np.float64), torch.stack(self._kin, 1).cpu().numpy().astype(np.float64)
        obj, alive = torch.stack(self._obj, 1).cpu().numpy().astype(np.float64), torch.stack(self._alive, 1).cpu().numpy()
        tgt = torch.stack(self._target, 1).cpu().numpy().astype(np.float64) if self._target else None
        if self.joint_space:
            des_c_pos, des_c_quat = chain_fk(np.asarray(self.env.scene.ctrl, np.float64), act[..., :7])
        else:
            des_c_pos, des_c_quat = act[..., :3], act[..., 3:7]
        out = []
        for e in range(self.env.n_envs):
            T = int(alive[e].sum())
            robot = {"des_c_pos": des_c_pos[e, :T], "des_c_quat": des_c_quat[e, :T], "c_pos": kin[e, :T, 0:3], "c_quat": kin[e, :T, 3:7],
                     "j_pos": kin[e, :T, 7:14], "j_vel": kin[e, :T, 14:21], "gripper_width": kin[e, :T, 21]}
            if self.joint_space:
                robot["des_j_pos"], robot["des_j_vel"] = act[e, :T, :7], np.zeros((T, 7))
            st = {"robot": robot}
            for k, name in enumerate(self.keys):
                st[name] = {"pos": obj[e, :T, k, :3], "quat": obj[e, :T, k, 3:]}
            for name, p in self.statics.items():
                st[name] = {"pos": np.tile(np.asarray(p, np.float64), (T, 1)), "quat": np.tile([0.0, 1.0, 0.0, 0.0], (T, 1))}
            if tgt is not None:
                st["target-box"] = {"pos": tgt[e, :T, :3], "quat": tgt[e, :T, 3:]}
            out.append(st)
        return out

    def save(self, path_pattern: str):
        """``path_pattern`` e.g. ``'out/env_{:04d}_00.pkl'`` (the reference's ``env_XXX_YY.pkl`` naming)."""
        for e, st in enumerate(self.env_states()):
            with open(path_pattern.format(e), "wb") as f:
                pickle.dump(st, f)
