"""ctypes binding of the fp64 CPU oracle (``oracle/d3il_oracle.c``).

TEST INFRASTRUCTURE: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libd3il_oracle.so")
    src = os.path.join(_HERE, "d3il_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, fp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.d3o_create.restype = vp
        L.d3o_create.argtypes = [C.c_char_p, C.c_size_t]
        L.d3o_destroy.argtypes = [vp]
        L.d3o_last_error.restype = C.c_char_p
        L.d3o_reset.argtypes = [vp, dp]
        L.d3o_step.argtypes = [vp, dp, fp, dp, ip, dp]
        L.d3o_substep.argtypes = [vp, C.c_int]
        L.d3o_robot_state.argtypes = [vp, dp]
        L.d3o_joint_state.argtypes = [vp, dp]
        L.d3o_get_obs.argtypes = [vp, fp]
        L.d3o_state_dim.argtypes = [vp]
        L.d3o_state_dim.restype = C.c_int
        L.d3o_get_state.argtypes = [vp, dp]
        L.d3o_set_state.argtypes = [vp, dp]
        L.d3o_forward.argtypes = [vp]
        L.d3o_probe.argtypes = [vp, C.c_char_p, dp, C.c_int]
        L.d3o_probe.restype = C.c_int
        L.d3o_ik_fk.argtypes = [vp, dp, dp, dp, dp]
        L.d3o_collide.argtypes = [C.c_int, dp, dp, dp, C.c_int, dp, dp, dp, dp]
        L.d3o_collide.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleEnv:
    """One fp64 env instance; mirrors the Gym-shaped reset/step of ``GymEnvWrapper`` (gym_env_wrapper.py:45-100)."""

    def __init__(self, scene_blob: bytes, header: dict):
        self.L = lib()
        self.h = self.L.d3o_create(scene_blob, len(scene_blob))
        if not self.h:
            raise RuntimeError(self.L.d3o_last_error().decode())
        self.hdr = header
        self.state_dim = self.L.d3o_state_dim(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.d3o_destroy(self.h)
            self.h = None

    def reset(self, ctx=None):
        if ctx is None:
            self.L.d3o_reset(self.h, None)
        else:
            c = np.ascontiguousarray(ctx, dtype=np.float64).reshape(-1)
            assert c.size == self.hdr["ctx_dim"]
            self.L.d3o_reset(self.h, _dp(c))
        return self.obs()

    def obs(self):
        o = np.zeros(self.hdr["obs_dim"], dtype=np.float32)
        self.L.d3o_get_obs(self.h, o.ctypes.data_as(C.POINTER(C.c_float)))
        return o

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        assert a.size == self.hdr["act_dim"]
        o = np.zeros(self.hdr["obs_dim"], dtype=np.float32)
        r = C.c_double()
        d = C.c_int()
        info = np.zeros(self.hdr["info_dim"], dtype=np.float64)
        self.L.d3o_step(self.h, _dp(a), o.ctypes.data_as(C.POINTER(C.c_float)), C.byref(r), C.byref(d), _dp(info))
        return o, r.value, bool(d.value), info

    def substep(self, n=1):
        self.L.d3o_substep(self.h, n)

    def joint_state(self):
        j = np.zeros(8)
        self.L.d3o_joint_state(self.h, _dp(j))
        return j

    def robot_state(self):
        t = np.zeros(3)
        self.L.d3o_robot_state(self.h, _dp(t))
        return t

    def get_state(self):
        s = np.zeros(self.state_dim)
        self.L.d3o_get_state(self.h, _dp(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.size == self.state_dim
        self.L.d3o_set_state(self.h, _dp(s))

    def forward(self):
        self.L.d3o_forward(self.h)

    def probe(self, what, cap=1 << 16):
        buf = np.zeros(cap)
        n = self.L.d3o_probe(self.h, what.encode(), _dp(buf), cap)
        if n < 0:
            raise RuntimeError("probe buffer too small")
        return buf[:n].copy()

    def ik_fk(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        p, qt, J = np.zeros(3), np.zeros(4), np.zeros((6, 7))
        self.L.d3o_ik_fk(self.h, _dp(q), _dp(p), _dp(qt), _dp(J))
        return p, qt, J


def collide(t1, p1, q1, s1, t2, p2, q2, s2):
    """Stand-alone narrow phase: returns [n, 7] rows (pos3, normal3, dist)."""
    out = np.zeros(8 * 7)
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (p1, q1, np.resize(np.asarray(s1, float), 3), p2, q2, np.resize(np.asarray(s2, float), 3))]
    n = lib().d3o_collide(t1, _dp(arrs[0]), _dp(arrs[1]), _dp(arrs[2]), t2, _dp(arrs[3]), _dp(arrs[4]), _dp(arrs[5]), _dp(out))
    return out[: 7 * n].reshape(n, 7)
