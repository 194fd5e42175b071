"""ctypes binding of the TEST-ONLY host build (G = 1) of the CUDA kernel core — see tests/emu/emu.cpp."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib(prec="f32"):
    if prec not in _LIBS:
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(_HERE, f"libd3il_emu_{prec}.so"))
        dp, fp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.emu_create.restype = vp
        L.emu_create.argtypes = [C.c_char_p, C.c_size_t]
        L.emu_destroy.argtypes = [vp]
        L.emu_set_solver.argtypes = [vp, C.c_double, C.c_int]
        for f in ("emu_ws_floats", "emu_n_state", "emu_state_dim"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = C.c_int
        L.emu_reset.argtypes = [vp, dp]
        L.emu_step.argtypes = [vp, dp, fp, dp, ip, dp]
        L.emu_substep.argtypes = [vp, C.c_int]
        L.emu_robot_state.argtypes = [vp, dp]
        L.emu_get_obs.argtypes = [vp, fp]
        L.emu_get_state.argtypes = [vp, dp]
        L.emu_probe.argtypes = [vp, C.c_char_p, dp, C.c_int]
        L.emu_probe.restype = C.c_int
        L.emu_set_state.argtypes = [vp, dp]
        _LIBS[prec] = L
    return _LIBS[prec]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class EmuEnv:
    def __init__(self, scene_blob: bytes, header: dict, prec="f32"):
        self.L = lib(prec)
        self.h = self.L.emu_create(scene_blob, len(scene_blob))
        if not self.h:
            raise RuntimeError("emu_create failed")
        self.hdr = header
        self.state_dim = self.L.emu_state_dim(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emu_destroy(self.h)
            self.h = None

    def set_solver(self, tol, max_iter):
        self.L.emu_set_solver(self.h, tol, max_iter)

    def reset(self, ctx=None):
        if ctx is None:
            self.L.emu_reset(self.h, None)
        else:
            c = np.ascontiguousarray(ctx, dtype=np.float64).reshape(-1)
            self.L.emu_reset(self.h, _dp(c))
        return self.obs()

    def obs(self):
        o = np.zeros(self.hdr["obs_dim"], dtype=np.float32)
        self.L.emu_get_obs(self.h, o.ctypes.data_as(C.POINTER(C.c_float)))
        return o

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        o = np.zeros(self.hdr["obs_dim"], dtype=np.float32)
        r, d = C.c_double(), C.c_int()
        info = np.zeros(self.hdr["info_dim"], dtype=np.float64)
        self.L.emu_step(self.h, _dp(a), o.ctypes.data_as(C.POINTER(C.c_float)), C.byref(r), C.byref(d), _dp(info))
        return o, r.value, bool(d.value), info

    def substep(self, n=1):
        self.L.emu_substep(self.h, n)

    def robot_state(self):
        t = np.zeros(3)
        self.L.emu_robot_state(self.h, _dp(t))
        return t

    def get_state(self):
        s = np.zeros(self.state_dim)
        self.L.emu_get_state(self.h, _dp(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        self.L.emu_set_state(self.h, _dp(s))

    def probe(self, what, cap=1 << 16):
        buf = np.zeros(cap)
        n = self.L.emu_probe(self.h, what.encode(), _dp(buf), cap)
        if n < 0:
            raise RuntimeError("bad probe")
        return buf[:n].copy()
