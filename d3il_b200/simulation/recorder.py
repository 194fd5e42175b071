"""Trajectory recording in the reference's dataset wire format.

The datasets the agents train on are pickled ``env_state`` dicts written by the reference's loggers while a human
teleoperates (``core/logger.py``; read back by ``environments/dataset/pushing_dataset.py:52-77``):
``env_state['robot']['des_c_pos' | 'c_pos']`` [T, 3] and ``env_state['<object>']['pos' | 'quat']`` [T, 3 | 4].
``TrajectoryRecorder`` collects the same arrays from a ``BatchedEnv`` rollout so GPU rollouts can be fed to the existing
``*_Dataset`` classes (one dict per env instance).
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


class TrajectoryRecorder:
    def __init__(self, env):
        self.env = env
        self.keys, self.statics = OBJECT_KEYS[env.task]
        self._des, self._cpos, self._obj, self._alive = [], [], [], []

    def record(self, action: torch.Tensor, alive: torch.Tensor | None = None):
        """Call once per env step, before ``env.step(action)``: logs the commanded pose, the measured tcp and the object poses."""
        self._des.append(action[:, :3].detach().clone())
        self._cpos.append(self.env.robot_state().detach().clone())
        self._obj.append(self.env.object_poses().detach().clone())
        self._alive.append(torch.ones(self.env.n_envs, dtype=torch.bool, device=action.device) if alive is None else alive.detach().clone().bool())

    def env_states(self) -> list[dict]:
        """One ``env_state`` dict per env, truncated to the steps in which that env was alive."""
        des, cpos = torch.stack(self._des, 1).cpu().numpy(), torch.stack(self._cpos, 1).cpu().numpy()
        obj, alive = torch.stack(self._obj, 1).cpu().numpy(), torch.stack(self._alive, 1).cpu().numpy()
        out = []
        for e in range(self.env.n_envs):
            T = int(alive[e].sum())
            st = {"robot": {"des_c_pos": des[e, :T].astype(np.float64), "c_pos": cpos[e, :T].astype(np.float64)}}
            for k, name in enumerate(self.keys):
                st[name] = {"pos": obj[e, :T, k, :3].astype(np.float64), "quat": obj[e, :T, k, 3:].astype(np.float64)}
            for name, p in self.statics.items():
                st[name] = {"pos": np.tile(np.asarray(p, np.float64), (T, 1)), "quat": np.tile([0.0, 1.0, 0.0, 0.0], (T, 1))}
            out.append(st)
        return out

    def save(self, path_pattern: str):
        """``path_pattern`` e.g. ``'out/env_{:04d}_00.pkl'`` (the reference's ``env_XXX_YY.pkl`` naming)."""
        for e, st in enumerate(self.env_states()):
            with open(path_pattern.format(e), "wb") as f:
                pickle.dump(st, f)
