"""CPU checks of the CUDA kernel core through its TEST-ONLY one-lane host build (tests/emu): the same source that nvcc
compiles for sm_100a, executed with G = 1, against the fp64 oracle.  fp64 build = logic equivalence (1e-9), fp32 build =
the precision the GPU path can deliver (tolerances as in tests/test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest

from d3il_b200.scene.blob import load_scene
from oracle.oracle import OracleEnv
from tests.emu.emu import EmuEnv, lib
from tests.util import random_walk_actions, scripted_push_actions


@pytest.mark.parametrize("ctx_id", [0, 7])
def test_core_logic_equals_oracle_fp64(pushing_contexts, ctx_id):
    """Teacher-forced per env step along a contact-rich push (rod-box, box-table, coupled and block-diagonal solves)."""
    blob, sc = load_scene("pushing")
    nq, nv = sc.header["nq"], sc.header["nv"]
    ctx = pushing_contexts[ctx_id]
    o, e = OracleEnv(blob, sc.header), EmuEnv(blob, sc.header, "f64")
    o.reset(ctx); e.reset(ctx)
    assert np.abs(o.get_state()[:-4] - e.get_state()[:-4]).max() < 1e-6     # contexts pass through float32 on the kernel side; last 4 words = kernel-side cost counters
    saw_coupled = False
    import json, os
    rod = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "d3il_b200", "scenes", "pushing.json")))["geoms"].index("rod:geom_rb0")
    for a in scripted_push_actions(ctx, o.robot_state(), n_steps=100):
        e.set_state(o.get_state())
        ro, re = o.step(a), e.step(a)
        so, se = o.get_state(), e.get_state()
        assert np.abs(so[:nq] - se[:nq]).max() < 1e-9
        assert np.abs(so[nq:nq + nv] - se[nq:nq + nv]).max() < 1e-7 * (1 + np.abs(so[nq:nq + nv]).max())
        assert np.allclose(ro[0], re[0], atol=1e-6) and ro[2] == re[2] and np.allclose(ro[3], re[3], atol=1e-6)
        con = o.probe("contacts").reshape(-1, 12)
        saw_coupled |= bool(((con[:, 8] == rod) & (con[:, 11] >= 0)).any())     # rod geom in an active contact
    assert saw_coupled


def test_core_fp32_single_env_step(pushing_contexts):
    """fp32 build, one env step closed loop from oracle states on the random-walk workload."""
    blob, sc = load_scene("pushing")
    nq, nv = sc.header["nq"], sc.header["nv"]
    ctx = pushing_contexts[3]
    o, e = OracleEnv(blob, sc.header), EmuEnv(blob, sc.header, "f32")
    o.reset(ctx)
    for a in random_walk_actions(o.robot_state(), 40, seed=1):
        e.set_state(o.get_state())
        ro, re = o.step(a), e.step(a)
        so, se = o.get_state(), e.get_state()
        assert np.allclose(se[:nq], so[:nq], rtol=1e-4, atol=5e-6)
        assert np.allclose(se[nq:nq + nv], so[nq:nq + nv], rtol=1e-3, atol=2e-4)
        assert np.allclose(ro[0], re[0], rtol=1e-4, atol=1e-5)


def test_avoiding_closed_loop_fp32():
    """Contact-free closed-loop episode (robot only): fp32 core tracks the oracle to ~1e-6 over thousands of ticks;
    an obstacle hit terminates both in the same env step with the same mode bits."""
    blob, sc = load_scene("avoiding")
    for tgt_x, want_success in ((0.304, 1.0), (0.345, 0.0)):
        o, e = OracleEnv(blob, sc.header), EmuEnv(blob, sc.header, "f32")
        o.reset(); e.reset()
        des = o.robot_state().copy()
        rng = np.random.default_rng(0)
        touched = False
        for k in range(250):
            des[0] += np.clip(tgt_x - des[0], -0.004, 0.004) + rng.uniform(-0.001, 0.001)
            des[1] += 0.004 + rng.uniform(-0.002, 0.002)
            a = np.concatenate([des, [0, 1, 0, 0]])
            ro, re = o.step(a), e.step(a)
            if not touched:
                assert np.allclose(ro[0], re[0], rtol=1e-4, atol=2e-5)
            touched |= o.get_state()[27 + 44 + 7] != 0
            assert ro[2] == re[2] and np.array_equal(ro[3][:10], re[3][:10])
            if ro[2]:
                break
        assert ro[2] and ro[3][0] == want_success


def test_slab_fast_path_equals_general_box_box():
    """The table/support fast path returns exactly what the general SAT + clipping routine returns whenever its
    preconditions hold (random poses: flat, tilted, flipped and arbitrary orientations, up to 2.5 cm deep)."""
    L = lib("f64")
    dp = C.POINTER(C.c_double)
    L.emu_collide_boxes.argtypes = [dp, dp, dp, dp, dp, C.c_int, dp]
    L.emu_collide_boxes.restype = C.c_int
    rng = np.random.default_rng(0)
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)   # noqa: E731
    used = 0
    for trial in range(20000):
        pA, hA = [(np.array([0.4, 0, -0.02]), np.array([0.49, 0.98, 0.001])), (np.array([0.4, 0, -0.42]), np.array([0.49, 0.98, 0.4]))][trial % 2]
        hB = np.array([0.03, 0.03, 0.03]) if trial % 3 else np.array([0.03, 0.05, 0.03])
        pB = np.array([rng.uniform(-0.1, 0.9), rng.uniform(-1, 1), pA[2] + hA[2] + rng.uniform(-0.03, 0.06)])
        if trial % 5 == 0:
            q = rng.normal(size=4)
        else:
            yaw, tilt = rng.uniform(-np.pi, np.pi), rng.normal(scale=10 ** rng.uniform(-4, -0.5), size=2)
            q = np.array([np.cos(yaw / 2), tilt[0], tilt[1], np.sin(yaw / 2)])
            if trial % 7 == 0:
                q = np.array([q[1], q[0], q[3], q[2]])
        q /= np.linalg.norm(q)
        o1, o2 = np.zeros(56), np.zeros(56)
        n1 = L.emu_collide_boxes(d(pA), d(hA), d(pB), d(q), d(hB), 1, d(o1))
        if n1 < 0:
            continue
        n2 = L.emu_collide_boxes(d(pA), d(hA), d(pB), d(q), d(hB), 0, d(o2))
        used += 1
        assert n1 == n2 and np.allclose(o1[:7 * n2], o2[:7 * n2], atol=1e-12)
    assert used > 5000


def test_ik_fast_path_and_clipped_spectrum_fallback():
    """ik_solve_unclipped (Cholesky of A and of A - svd_min I) replaces the eigen-decomposition whenever the spectrum of
    J J^T lies inside the clip range; towards the edge of the arm's reach lambda_min drops below svd_min = 1e-2 and the
    Jacobi path with the clipped spectrum takes over.  Both regimes (and the hand-over) must track the oracle, which
    always eigen-decomposes (IKControllers.py:239-262)."""
    blob, sc = load_scene("avoiding")
    o, e = OracleEnv(blob, sc.header), EmuEnv(blob, sc.header, "f64")
    o.reset(); e.reset()
    des = o.robot_state().copy()
    lam_min = []
    for k in range(140):
        des[0] += 0.004                      # straight out along +x: reach limit at ~0.85 m
        des[1] += 0.001
        a = np.concatenate([des, [0, 1, 0, 0]])
        ro, re = o.step(a), e.step(a)
        so, se = o.get_state(), e.get_state()
        assert np.abs(so[:9] - se[:9]).max() < 1e-9, k
        off = 9 + 2 * 9
        ikq = so[off + 16: off + 23]
        _, _, J = o.ik_fk(ikq)
        lam_min.append(np.linalg.eigvalsh(J @ J.T + 1e-12 * np.eye(6))[0])
        if ro[2]:
            break
    lam_min = np.array(lam_min)
    assert lam_min[0] > 1e-2 and lam_min.min() < 5e-3, (lam_min[0], lam_min.min())     # both regimes were visited


def test_ik_clipped_solve_known_answers():
    """The three routes of the IK's clipped-spectrum solve (Cholesky when nothing is clipped, one-eigenvalue deflation by
    inverse iteration, Jacobi eigen-decomposition) against numpy's restatement of IKControllers.py:239-262
    (np.linalg.svd of J J^T + reg I, singular values clipped to [1e-2, 1e2])."""
    L = lib("f64")
    dp = C.POINTER(C.c_double)
    L.emu_ik_clipped_solve.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, dp]
    L.emu_ik_clipped_solve.restype = C.c_int
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)   # noqa: E731
    rng = np.random.default_rng(5)
    lo, hi, reg = 1e-2, 1e2, 1e-12
    seen = {0: 0, 1: 0, 2: 0}
    for trial in range(600):
        U, _ = np.linalg.qr(rng.normal(size=(6, 6)))
        W, _ = np.linalg.qr(rng.normal(size=(7, 7)))
        sv = np.sqrt(rng.uniform(0.02, 4.0, 6))
        kind = trial % 3                       # 0: all inside the clip range, 1: one eigenvalue below, 2: two below
        if kind >= 1:
            sv[0] = np.sqrt(10 ** rng.uniform(-4, -2.05))
        if kind == 2:
            sv[1] = np.sqrt(10 ** rng.uniform(-4, -2.05))
        J = (U * sv) @ W[:6]
        rhs = rng.normal(size=6)
        A = J @ J.T + reg * np.eye(6)
        u, s, vt = np.linalg.svd(A)
        want = (u * (1.0 / np.clip(s, lo, hi))) @ (vt @ rhs)
        x0, x1 = np.zeros(6), np.zeros(6)
        ok0 = L.emu_ik_clipped_solve(d(J), d(rhs), reg, lo, hi, 0, d(x0))
        assert L.emu_ik_clipped_solve(d(J), d(rhs), reg, lo, hi, 1, d(x1)) == 1
        assert np.allclose(x1, want, rtol=1e-9, atol=1e-9 * np.abs(want).max()), (trial, kind)
        assert bool(ok0) == (kind < 2), (trial, kind)          # two small eigenvalues are left to the eigen-decomposition
        if ok0:
            assert np.allclose(x0, want, rtol=1e-10, atol=1e-10 * np.abs(want).max()), (trial, kind)
        seen[kind] += 1
    assert min(seen.values()) >= 190
