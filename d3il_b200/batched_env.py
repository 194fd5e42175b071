"""BatchedEnv: N lock-stepped instances of one D3IL task on one GPU, Gym-shaped ``reset`` / ``step``.

Host-side mirror of ``GymEnvWrapper`` (reference ``environments/d3il/d3il_sim/gyms/gym_env_wrapper.py:11-189``) and of
the task envs' ``reset(random, context)`` / ``step(action)`` (``envs/gym_pushing_env/.../pushing.py:335-339,461-488``),
batched: every argument gains a leading ``n_envs`` axis.  All arithmetic happens in ``libd3il.so``; torch tensors are
only the device buffers handed to the C ABI (``tensor.data_ptr()``), and kernels are enqueued on torch's current stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .scene.blob import load_scene


class BatchedEnv:
    def __init__(self, task: str, n_envs: int, device: int | str | torch.device = 0):
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise RuntimeError("BatchedEnv runs on CUDA devices only (there is no CPU path)")
        self.device = dev
        self.task = task
        self.n_envs = int(n_envs)
        blob, scene = load_scene(task)
        self.scene = scene
        self._L = _lib.lib()
        h = C.c_void_p()
        if dev.index is None:                   # plain "cuda": the CURRENT device, not device 0 (the output tensors below live there too)
            dev = torch.device("cuda", torch.cuda.current_device())
            self.device = dev
        _lib.check(self._L.d3il_create(C.byref(h), blob, len(blob), self.n_envs, dev.index), "d3il_create")
        self._h = h
        dims = (C.c_int32 * 8)()
        _lib.check(self._L.d3il_dims(self._h, dims), "d3il_dims")
        self.dims = dict(zip(_lib.DIM_NAMES, list(dims)))
        self.obs_dim, self.act_dim, self.ctx_dim, self.info_dim = (self.dims[k] for k in ("obs", "act", "ctx", "info"))
        self.n_substeps, self.max_steps_per_episode = self.dims["n_substeps"], self.dims["max_steps"]
        n = self.n_envs
        with torch.cuda.device(dev):
            self.obs = torch.zeros(n, self.obs_dim, device=dev)
            self.reward = torch.zeros(n, device=dev)
            self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
            self.info = torch.zeros(n, self.info_dim, device=dev)
            self.tcp = torch.zeros(n, 3, device=dev)
            self.joints = torch.zeros(n, 8, device=dev)

    def close(self):
        if getattr(self, "_h", None):
            self._L.d3il_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device-resident API (torch CUDA tensors in / out, asynchronous on the current stream)
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, contexts: torch.Tensor | None = None, mask: torch.Tensor | None = None) -> torch.Tensor:
        """contexts: [n_envs, n_obj, 7] (xyz + quat wxyz per free object) float32 CUDA; mask: uint8/bool [n_envs] or None."""
        ctx_p = None
        if self.ctx_dim:
            contexts = contexts.to(self.device, torch.float32).reshape(self.n_envs, self.ctx_dim).contiguous()
            ctx_p = C.c_void_p(contexts.data_ptr())
        mask_p = None
        if mask is not None:
            mask = mask.to(self.device, torch.uint8).contiguous()
            mask_p = C.c_void_p(mask.data_ptr())
        _lib.check(self._L.d3il_reset(self._h, ctx_p, mask_p, C.c_void_p(self.obs.data_ptr()), self._stream()), "d3il_reset")
        return self.obs

    def step(self, action: torch.Tensor):
        """action: [n_envs, act_dim] (float32 CUDA): desired tcp xyz + quat wxyz, or 7 joint set-points + gripper command (Stacking). Returns (obs, reward, done, info) tensors (views
        of the env's output buffers, overwritten by the next call)."""
        action = action.to(self.device, torch.float32).contiguous()
        assert action.shape == (self.n_envs, self.act_dim)
        _lib.check(self._L.d3il_step(self._h, C.c_void_p(action.data_ptr()), C.c_void_p(self.obs.data_ptr()), C.c_void_p(self.reward.data_ptr()),
                                      C.c_void_p(self.done.data_ptr()), C.c_void_p(self.info.data_ptr()), self._stream()), "d3il_step")
        return self.obs, self.reward, self.done, self.info

    def robot_state(self) -> torch.Tensor:
        _lib.check(self._L.d3il_robot_state(self._h, C.c_void_p(self.tcp.data_ptr()), self._stream()), "d3il_robot_state")
        return self.tcp

    def joint_state(self) -> torch.Tensor:
        """[n_envs, 8]: joint positions + gripper width (``CubeStacking_Env.robot_state``, stacking.py:218-226)."""
        _lib.check(self._L.d3il_joint_state(self._h, C.c_void_p(self.joints.data_ptr()), self._stream()), "d3il_joint_state")
        return self.joints

    def robot_kinematics(self) -> torch.Tensor:
        """[n_envs, 22]: tcp pos (3) + quat (4), joint positions (7), joint velocities (7), gripper width — the robot fields of the
        reference's dataset pickles (``MjRobot.receiveState``, MjRobot.py:133-184)."""
        if not hasattr(self, "_kin"):
            self._kin = torch.zeros(self.n_envs, 22, device=self.device)
        _lib.check(self._L.d3il_robot_kinematics(self._h, C.c_void_p(self._kin.data_ptr()), self._stream()), "d3il_robot_kinematics")
        return self._kin

    def object_poses(self) -> torch.Tensor:
        """[n_envs, n_obj, 7] xyz + quat wxyz of the free objects (``Scene.get_obj_pos`` / ``get_obj_quat``)."""
        n_obj = self.scene.header["nobj"]
        if not hasattr(self, "_obj_poses"):
            self._obj_poses = torch.zeros(self.n_envs, n_obj, 7, device=self.device)
        if n_obj:
            _lib.check(self._L.d3il_object_poses(self._h, C.c_void_p(self._obj_poses.data_ptr()), self._stream()), "d3il_object_poses")
        return self._obj_poses

    # ---- host-buffer API (numpy in / numpy out; H2D + D2H inside the call) — the reference-facing end-to-end path
    def reset_host(self, contexts: np.ndarray | None = None, mask: np.ndarray | None = None) -> np.ndarray:
        obs = np.zeros((self.n_envs, self.obs_dim), dtype=np.float32)
        ctx_p = None
        if self.ctx_dim:
            c = np.ascontiguousarray(contexts, dtype=np.float32).reshape(self.n_envs, self.ctx_dim)
            ctx_p = c.ctypes.data_as(C.c_void_p)
        mask_p = None
        if mask is not None:
            mk = np.ascontiguousarray(mask, dtype=np.uint8)
            mask_p = mk.ctypes.data_as(C.c_void_p)
        _lib.check(self._L.d3il_reset_host(self._h, ctx_p, mask_p, obs.ctypes.data_as(C.c_void_p)), "d3il_reset_host")
        return obs

    def step_host(self, action: np.ndarray):
        a = np.ascontiguousarray(action, dtype=np.float32).reshape(self.n_envs, self.act_dim)
        obs = np.empty((self.n_envs, self.obs_dim), dtype=np.float32)
        rew = np.empty(self.n_envs, dtype=np.float32)
        done = np.empty(self.n_envs, dtype=np.uint8)
        info = np.empty((self.n_envs, self.info_dim), dtype=np.float32)
        _lib.check(self._L.d3il_step_host(self._h, a.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p), rew.ctypes.data_as(C.c_void_p),
                                           done.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p)), "d3il_step_host")
        return obs, rew, done, info

    def robot_state_host(self) -> np.ndarray:
        t = np.empty((self.n_envs, 3), dtype=np.float32)
        _lib.check(self._L.d3il_robot_state_host(self._h, t.ctypes.data_as(C.c_void_p)), "d3il_robot_state_host")
        return t

    def joint_state_host(self) -> np.ndarray:
        j = np.empty((self.n_envs, 8), dtype=np.float32)
        _lib.check(self._L.d3il_joint_state_host(self._h, j.ctypes.data_as(C.c_void_p)), "d3il_joint_state_host")
        return j

    # ---- parity-test hooks
    def substep(self, n: int = 1):
        _lib.check(self._L.d3il_substep(self._h, int(n), self._stream()), "d3il_substep")

    def get_state(self, env: int) -> np.ndarray:
        s = np.zeros(self.dims["state"], dtype=np.float64)
        _lib.check(self._L.d3il_get_state(self._h, s.ctypes.data_as(C.POINTER(C.c_double)), int(env)), "d3il_get_state")
        return s

    def set_state(self, env: int, state: np.ndarray):
        s = np.ascontiguousarray(state, dtype=np.float64)
        assert s.size == self.dims["state"]
        _lib.check(self._L.d3il_set_state(self._h, s.ctypes.data_as(C.POINTER(C.c_double)), int(env)), "d3il_set_state")

    def set_solver(self, tolerance: float, max_iterations: int):
        _lib.check(self._L.d3il_set_solver(self._h, float(tolerance), int(max_iterations)), "d3il_set_solver")

    def set_profiling(self, on: bool):
        _lib.check(self._L.d3il_set_profiling(self._h, int(on)), "d3il_set_profiling")

    def get_profile(self):
        """(ms spent in the IK kernel, ms spent in the env-step kernel, number of profiled env steps)."""
        ms = (C.c_double * 2)()
        n = C.c_longlong()
        _lib.check(self._L.d3il_get_profile(self._h, ms, C.byref(n)), "d3il_get_profile")
        return ms[0], ms[1], n.value

    @property
    def kernel_launches(self) -> int:
        return int(self._L.d3il_kernel_launches(self._h))
