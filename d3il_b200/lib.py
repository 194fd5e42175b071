"""ctypes loader for ``libd3il.so`` (CUDA kernels + the C ABI declared in ``include/d3il.h``).

The library is built in-tree (``d3il_b200/csrc/Makefile``; ``__graft_entry__.build()``).  There is no CPU fallback:
if the shared object is missing or cannot be loaded this module raises, and ``d3il_create`` itself fails without a
usable CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "libd3il.so")
_LIB = None

DIM_NAMES = ("obs", "act", "ctx", "info", "state", "n_envs", "n_substeps", "max_steps")

EXPORTS = (
    "d3il_create", "d3il_destroy", "d3il_last_error", "d3il_dims", "d3il_reset", "d3il_step", "d3il_robot_state",
    "d3il_reset_host", "d3il_step_host", "d3il_robot_state_host", "d3il_substep", "d3il_get_state", "d3il_set_state",
    "d3il_set_solver", "d3il_kernel_launches", "d3il_set_profiling", "d3il_get_profile", "d3il_joint_state", "d3il_joint_state_host", "d3il_object_poses", "d3il_robot_kinematics",
)


def build(force: bool = False) -> str:
    """Compile ``libd3il.so`` for sm_100a with nvcc (cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc")]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return SO_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        L = C.CDLL(SO_PATH)
        vp, fp, dp, u8 = C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p
        L.d3il_create.argtypes = [C.POINTER(vp), C.c_char_p, C.c_size_t, C.c_int, C.c_int]
        L.d3il_destroy.argtypes = [vp]
        L.d3il_destroy.restype = None
        L.d3il_last_error.restype = C.c_char_p
        L.d3il_dims.argtypes = [vp, C.POINTER(C.c_int32)]
        L.d3il_reset.argtypes = [vp, fp, u8, fp, vp]
        L.d3il_step.argtypes = [vp, fp, fp, fp, u8, fp, vp]
        L.d3il_robot_state.argtypes = [vp, fp, vp]
        L.d3il_reset_host.argtypes = [vp, fp, u8, fp]
        L.d3il_step_host.argtypes = [vp, fp, fp, fp, u8, fp]
        L.d3il_robot_state_host.argtypes = [vp, fp]
        L.d3il_joint_state.argtypes = [vp, fp, vp]
        L.d3il_joint_state_host.argtypes = [vp, fp]
        L.d3il_object_poses.argtypes = [vp, fp, vp]
        L.d3il_robot_kinematics.argtypes = [vp, fp, vp]
        L.d3il_substep.argtypes = [vp, C.c_int, vp]
        L.d3il_get_state.argtypes = [vp, dp, C.c_int]
        L.d3il_set_state.argtypes = [vp, dp, C.c_int]
        L.d3il_set_solver.argtypes = [vp, C.c_double, C.c_int]
        L.d3il_kernel_launches.argtypes = [vp]
        L.d3il_kernel_launches.restype = C.c_longlong
        L.d3il_set_profiling.argtypes = [vp, C.c_int]
        L.d3il_get_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        _LIB = L
    return _LIB


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().d3il_last_error().decode()}")
