"""Oracle pinned by invariants — the reference ships no golden vectors for this path,
MuJoCo/pinocchio are not installable, so these are the checks that stand in for them ("parity unpinned")."""
import numpy as np
import pytest

from d3il_b200.scene import mjcf
from d3il_b200.scene.blob import links_from_scene
from oracle.oracle import OracleEnv, collide

G_CYL, G_BOX = 5, 6


def make(scene):
    blob, sc = scene
    return OracleEnv(blob, sc.header), sc


def set_q(env, sc, qarm, qvel_arm=None, objs=None):
    s = env.get_state()
    nq, nv = sc.header["nq"], sc.header["nv"]
    s[:7] = qarm
    s[nq:nq + nv] = 0
    if qvel_arm is not None:
        s[nq:nq + 7] = qvel_arm
    if objs is not None:
        s[9:nq] = np.asarray(objs).reshape(-1)
    env.set_state(s)


def test_fk_known_answers(pushing_scene):
    """B.9(1): URDF-chain FK KATs from SURVEY §8c."""
    env, sc = make(pushing_scene)
    q0 = np.array([0, 0.1745, 0, -0.8727, 0, 1.2217, 0.7854])
    p, quat, J = env.ik_fk(q0)
    assert np.allclose(p, [0.550900, 0, 0.699822], atol=2e-5)
    assert np.allclose(np.abs(quat), [0, 0.996195, 0, 0.087156], atol=2e-5)
    p, _, _ = env.ik_fk(np.zeros(7))
    assert np.allclose(p, [0.088, 0, 0.821], atol=1e-9)
    # Jacobian = finite difference of FK
    rng = np.random.default_rng(0)
    q = rng.uniform(-1, 1, 7)
    p0, _, J = env.ik_fk(q)
    for i in range(7):
        dq = np.zeros(7); dq[i] = 1e-6
        p1, _, _ = env.ik_fk(q + dq)
        assert np.allclose((p1 - p0) / 1e-6, J[:3, i], atol=1e-5)


def test_mjcf_chain_matches_urdf_chain(pushing_scene, pushing_contexts):
    """B.9(1): physics-side tcp (MJCF chain) == controller-side grasptarget (URDF chain) up to XML rounding."""
    env, sc = make(pushing_scene)
    env.reset(pushing_contexts[0])
    rng = np.random.default_rng(1)
    for _ in range(5):
        q = rng.uniform(-1.5, 1.5, 7); q[3] = -abs(q[3]) - 0.2
        set_q(env, sc, q); env.forward()
        p, _, _ = env.ik_fk(q)
        assert np.allclose(env.robot_state(), p, atol=2e-6)


def test_mass_matrix_matches_numpy_crba(pushing_scene, pushing_contexts):
    """B.9(2): oracle CRBA vs the independent numpy composite-inertia build; symmetric positive definite."""
    env, sc = make(pushing_scene)
    links = links_from_scene(sc)
    rng = np.random.default_rng(2)
    env.reset(pushing_contexts[3])
    for _ in range(3):
        q = rng.uniform(-1.5, 1.5, 7); q[3] = -abs(q[3]) - 0.2
        objs = pushing_contexts[rng.integers(60)].copy()
        objs[:, 2] = 0.011
        set_q(env, sc, q, objs=objs); env.forward()
        nv = sc.header["nv"]
        M = env.probe("M").reshape(nv, nv)
        s = env.get_state()
        qpl = list(s[:9]) + [s[9 + 7 * k: 16 + 7 * k] for k in range(sc.header["nobj"])]
        Mref, _ = mjcf.mass_matrix(links, qpl)
        assert np.allclose(M, Mref, atol=1e-10)
        assert np.allclose(M, M.T) and np.linalg.eigvalsh(M).min() > 0


def test_bias_is_lagrangian(pushing_scene, pushing_contexts):
    """B.9(2): RNE bias == Christoffel terms of M(q) + potential gradient (finite differences on the numpy model)."""
    env, sc = make(pushing_scene)
    links = links_from_scene(sc)[:9]
    rng = np.random.default_rng(3)
    env.reset(pushing_contexts[0])
    q = np.concatenate([rng.uniform(-1, 1, 7), [0.01, 0.02]]); q[3] = -1.5
    qd = np.concatenate([rng.uniform(-1, 1, 7), [0.05, -0.03]])
    s = env.get_state(); nq, nv = sc.header["nq"], sc.header["nv"]
    s[:9] = q; s[nq:nq + nv] = 0; s[nq:nq + 9] = qd
    env.set_state(s); env.forward()
    bias = env.probe("bias")[:9]

    def Mof(qq):
        return mjcf.mass_matrix(links, list(qq))[0]

    def V(qq):
        P, R = mjcf.link_fk(links, list(qq))
        return sum(L.mass * 9.81 * (P[i] + R[i] @ L.ipos)[2] for i, L in enumerate(links))

    h = 1e-6
    dM = np.zeros((9, 9, 9)); g = np.zeros(9)
    for k in range(9):
        e = np.zeros(9); e[k] = h
        dM[:, :, k] = (Mof(q + e) - Mof(q - e)) / (2 * h)
        g[k] = (V(q + e) - V(q - e)) / (2 * h)
    c = np.einsum("ijk,j,k->i", dM, qd, qd) - 0.5 * np.einsum("jki,j,k->i", dM, qd, qd)
    assert np.allclose(bias, c + g, atol=2e-6)


def test_free_fall(pushing_scene, pushing_contexts):
    """B.9(3): a box above the table falls with -g (semi-implicit Euler: z_n = z0 - g h^2 n(n+1)/2)."""
    env, sc = make(pushing_scene)
    ctx = pushing_contexts[0].copy(); ctx[:, 2] = 0.5
    env.reset(ctx)
    z0 = env.get_state()[9 + 2]
    env.substep(100)
    z = env.get_state()[9 + 2]
    n = 100
    # reset already took one tick (v = -g h), so ticks 2..n+1 contribute -g h^2 k each
    assert abs((z - z0) + 9.81 * 1e-6 * ((n + 1) * (n + 2) / 2 - 1)) < 1e-12


def test_box_rest_force_balance(pushing_scene, pushing_contexts):
    """B.9(4,5): at rest the table carries m*g, friction stays inside the (regularised) cone, no drift."""
    env, sc = make(pushing_scene)
    env.reset(pushing_contexts[5])
    env.substep(600)
    s0 = env.get_state()
    env.substep(200)
    s1 = env.get_state()
    nq, nv = sc.header["nq"], sc.header["nv"]
    assert np.abs(s1[9:nq] - s0[9:nq]).max() < 1e-7          # boxes do not creep
    assert np.abs(s1[nq + 9:nq + nv]).max() < 1e-5
    f = env.probe("efc_force"); con = env.probe("contacts").reshape(-1, 12)
    J = env.probe("efc_J").reshape(-1, nv)
    qc = J.T @ f
    # vertical constraint force on each box == weight (0.05 kg)
    assert abs(qc[9 + 2] - 0.05 * 9.81) < 1e-6 and abs(qc[15 + 2] - 0.05 * 9.81) < 1e-6
    for c in con:
        e0 = int(c[11])
        if e0 < 0:
            continue
        fn, ft = f[e0], np.hypot(f[e0 + 1], f[e0 + 2])
        assert fn >= -1e-12 and ft <= 1.0 * fn + 1e-9        # friction coefficient 1 (max of the pair)
    assert 0.0109 < s1[9 + 2] < 0.011                          # rests ~16 um inside the table top (soft contact)


def test_gravity_compensated_arm_holds(avoiding_scene):
    """B.9(3): with the joint-PD + stale-bias gravity compensation the arm stays at init_qpos."""
    env, sc = make(avoiding_scene)
    env.reset()
    q0 = env.get_state()[:7].copy()
    env.substep(300)
    assert np.abs(env.get_state()[:7] - q0).max() < 2e-5


def test_solver_optimality_and_cone(pushing_scene, pushing_contexts):
    """B.9(6): Newton result satisfies the KKT stationarity of the primal problem; forces lie in the dual cone."""
    env, sc = make(pushing_scene)
    env.reset(pushing_contexts[7])                  # deep initial penetration: 16 contacts, 48 rows
    nv = sc.header["nv"]
    M = env.probe("M").reshape(nv, nv); J = env.probe("efc_J").reshape(-1, nv)
    f = env.probe("efc_force"); qacc = env.probe("qacc"); qs = env.probe("qacc_smooth")
    assert J.shape[0] == 48
    grad = M @ (qacc - qs) - J.T @ f
    assert np.abs(grad).max() < 1e-7
    con = env.probe("contacts").reshape(-1, 12)
    for c in con:
        e0 = int(c[11]); mu = c[10]
        assert f[e0] >= 0
        # regularised cone: |f_t| <= mu_reg * f_n * (f0/mu) ... in scaled variables  f_t/f0 <= f_n/mu
        assert np.hypot(f[e0 + 1], f[e0 + 2]) / 1.0 <= f[e0] / mu + 1e-9


def test_determinism_and_state_roundtrip(pushing_scene, pushing_contexts):
    """B.9(8): set_state(get_state()) is idempotent and reset reproduces bit-identical trajectories."""
    env, sc = make(pushing_scene)
    traj = []
    for rep in range(2):
        env.reset(pushing_contexts[11])
        des = env.robot_state().copy()
        out = []
        for k in range(10):
            des[1] += 0.005
            o, r, d, info = env.step(np.concatenate([des, [0, 1, 0, 0]]))
            if k == 4:
                env.set_state(env.get_state())
            out.append(env.get_state())
        traj.append(np.array(out))
    assert np.array_equal(traj[0], traj[1])


def _numpy_ik_tick(sc, q, des_pos, des_quat):
    """numpy restatement of IKControllers.py:197-276 (np.linalg.svd / solve exactly as the reference calls them)."""
    from d3il_b200.scene import blob as B
    from d3il_b200.scene.compile import ik_fk, quat_error
    C = sc.ctrl
    chain = [(C[12 * i: 12 * i + 3], C[12 * i + 3: 12 * i + 12].reshape(3, 3)) for i in range(7)]
    ee = (C[B.C_IK_EE: B.C_IK_EE + 3], C[B.C_IK_EE + 3: B.C_IK_EE + 12].reshape(3, 3))
    q = q.copy(); dq = np.array(des_quat, float)
    for _ in range(3):
        p, cq, J = ik_fk(chain, ee, q)
        if np.linalg.norm(cq - dq) > np.linalg.norm(cq + dq):
            dq = -dq
        acc = np.hstack((C[B.C_PGAIN_POS:B.C_PGAIN_POS + 3] * np.clip(des_pos - p, -0.01, 0.01),
                         C[B.C_PGAIN_QUAT:B.C_PGAIN_QUAT + 3] * np.clip(quat_error(cq, dq), -0.1, 0.1)))
        A = J @ J.T + C[B.C_JREG] * np.eye(6)
        u, sv, v = np.linalg.svd(A, full_matrices=False)
        A = u @ np.diag(np.clip(sv, C[B.C_SVD_MIN], C[B.C_SVD_MAX])) @ v
        qd_null = C[B.C_PGAIN_NULL:B.C_PGAIN_NULL + 7] * np.clip(C[B.C_REST:B.C_REST + 7] - q, -0.2, 0.2)
        qd = J.T @ np.linalg.solve(A, acc - J @ qd_null) + qd_null
        if np.linalg.norm(qd) > 3:
            qd = qd * 3 / np.linalg.norm(qd)
        q = np.clip(q + C[B.C_LRATE] * qd, C[B.C_JMIN:B.C_JMIN + 7], C[B.C_JMAX:B.C_JMAX + 7])
    return q


def test_ik_tick_matches_numpy_restatement(avoiding_scene):
    """a5: the oracle's Jacobi-eigen IK iteration == numpy svd/solve restatement of IKControllers.py:197-276."""
    env, sc = make(avoiding_scene)
    env.reset()
    rng = np.random.default_rng(5)
    nq, nv = 9, 9
    o = nq + 2 * nv
    for trial in range(4):
        tcp = env.robot_state().copy()
        des = tcp + rng.uniform(-0.02, 0.02, 3)
        s = env.get_state()
        s[o + 23: o + 26] = des; s[o + 26: o + 30] = [0, 1, 0, 0]
        s[o + 44 + 1] = 1                 # ctrl_mode = Cartesian
        env.set_state(s)
        valid = s[o + 44] != 0
        q_start = s[o + 16: o + 23].copy() if valid else s[:7].copy()
        env.substep(1)
        ikq = env.get_state()[o + 16: o + 23]
        ref = _numpy_ik_tick(sc, q_start, des, [0, 1, 0, 0])
        assert np.allclose(ikq, ref, atol=1e-12), np.abs(ikq - ref).max()
        env.substep(20)


# ------------------------------------------------------------------ narrow phase known answers
def test_box_box_flat_rest_gives_four_corners():
    q = [1, 0, 0, 0]
    c = collide(G_BOX, [0.4, 0, -0.02], q, [0.49, 0.98, 0.001], G_BOX, [0.5, 0.1, 0.0109], [np.cos(0.3), 0, 0, np.sin(0.3)], [0.03, 0.03, 0.03])
    assert c.shape == (4, 7)
    assert np.allclose(c[:, 3:6], [0, 0, 1])
    assert np.allclose(c[:, 6], -0.0001, atol=1e-12)
    assert np.allclose(np.sort(np.hypot(c[:, 0] - 0.5, c[:, 1] - 0.1)), 0.03 * np.sqrt(2))


def test_box_box_separated_and_edge():
    q = [1, 0, 0, 0]
    assert collide(G_BOX, [0, 0, 0], q, [0.03] * 3, G_BOX, [0.07, 0, 0], q, [0.03] * 3).shape[0] == 0
    # edge-edge: B rotated 45deg about x and 45deg about y-ish, touching A's top edge
    qa = mjcf.mat2quat(mjcf.rpy2mat([0.0, 0.0, np.pi / 4]))
    qb = mjcf.mat2quat(mjcf.rpy2mat([np.pi / 4, 0.0, 0.0]))
    c = collide(G_BOX, [0, 0, 0], qa, [0.03] * 3, G_BOX, [0.0, 0.0, 0.03 + 0.03 * np.sqrt(2) - 0.001], qb, [0.03] * 3)
    assert c.shape[0] >= 1 and np.all(c[:, 6] < 0) and np.all(c[:, 5] > 0.9)


def test_cyl_box_side_face_and_corner():
    q = [1, 0, 0, 0]
    # vertical rod flush against the +x face of a box: normal -x (cyl -> box), depth 2 mm, point at mid overlap height
    c = collide(G_CYL, [0.038, 0.0, 0.15], q, [0.01, 0.15], G_BOX, [0, 0, 0.011], q, [0.03] * 3)
    assert c.shape == (1, 7)
    assert np.allclose(c[0, 3:6], [-1, 0, 0]) and abs(c[0, 6] + 0.002) < 1e-12
    assert abs(c[0, 2] - 0.5 * (0.0 + 0.041)) < 1e-9 and abs(c[0, 0] - 0.029) < 1e-12
    # rod near the vertical edge (corner in top view): normal along the diagonal
    d = 0.03 + (0.01 - 0.001) / np.sqrt(2)
    c = collide(G_CYL, [d, d, 0.15], q, [0.01, 0.15], G_BOX, [0, 0, 0.011], q, [0.03] * 3)
    assert c.shape == (1, 7)
    assert np.allclose(c[0, 3:6], [-np.sqrt(0.5), -np.sqrt(0.5), 0], atol=1e-9) and abs(c[0, 6] + 0.001) < 1e-9
    # separated
    assert collide(G_CYL, [0.041, 0.0, 0.15], q, [0.01, 0.15], G_BOX, [0, 0, 0.011], q, [0.03] * 3).shape[0] == 0


def test_cyl_cyl_parallel():
    q = [1, 0, 0, 0]
    c = collide(G_CYL, [0.5, -0.136, 0.152], q, [0.01, 0.15], G_CYL, [0.5, -0.1, 0.0], q, [0.03, 0.07])
    assert c.shape == (1, 7) and abs(c[0, 6] + 0.004) < 1e-12 and np.allclose(c[0, 3:6], [0, 1, 0])
    assert collide(G_CYL, [0.5, -0.1401, 0.152], q, [0.01, 0.15], G_CYL, [0.5, -0.1, 0.0], q, [0.03, 0.07]).shape[0] == 0


def _yaw_tilt_quat(yaw, tx, ty):
    q = np.array([np.cos(yaw / 2), tx / 2, ty / 2, np.sin(yaw / 2)])
    return q / np.linalg.norm(q)


def test_cyl_box_contact_point_is_continuous_in_tilt():
    """The contact point of the cylinder-box rule must not jump when the relative tilt moves through zero (the first rule
    classified features with 1e-4 thresholds and jumped by centimetres; a rod-end fallback flipped ends with the sign of a
    1e-17 dot product).  Vertical rod pushing a box face / a box edge, rod end pressing on a box top."""
    from oracle.oracle import collide
    rng = np.random.default_rng(5)
    G_CYL, G_BOX = 5, 6
    worst = 0.0
    for trial in range(300):
        kind = trial % 3
        yaw = rng.uniform(-np.pi, np.pi) if kind != 1 else np.pi / 4 + rng.uniform(-0.02, 0.02)      # kind 1: rod against a vertical box edge
        box_p = np.array([0.5, 0.0, 0.13])
        if kind < 2:        # side contact: rod beside the box, 0.3 .. 1.5 mm deep, overlapping the box's z range partially
            d = np.array([np.cos(yaw), np.sin(yaw)]) if kind == 0 else np.array([1.0, 0.0])
            reach = 0.03 if kind == 0 else 0.03 * np.sqrt(2) * np.cos(yaw - np.pi / 4)
            rod_p = np.array([*(box_p[:2] + d * (reach + 0.01 - rng.uniform(3e-4, 1.5e-3))), 0.13 + 0.15 - rng.uniform(0.0, 0.03)])
        else:               # cap contact: rod end pressing 1 .. 3 mm into the box top
            rod_p = np.array([box_p[0] + rng.uniform(-0.015, 0.015), box_p[1] + rng.uniform(-0.015, 0.015), 0.16 + 0.15 - rng.uniform(1e-3, 3e-3)])
        tilts = [(-2e-5, 1e-5), (0.0, 0.0), (2e-5, -1e-5), (1e-4, 5e-5), (2e-4, 1e-4)]
        pts = []
        for tx, ty in tilts:
            c = collide(G_CYL, rod_p, _yaw_tilt_quat(0.0, tx, ty), [0.01, 0.15], G_BOX, box_p, _yaw_tilt_quat(yaw, 0, 0), [0.03, 0.03, 0.03])
            assert len(c) == 1, (trial, kind, tx, c)
            pts.append(c[0, :3])
        pts = np.array(pts)
        # tilt steps of <= 1e-4 rad move the contact point by at most a few centimetres per radian x the lever arms involved:
        # bound 2 mm (the old rules jumped by 10 - 300 mm)
        step = np.abs(np.diff(pts, axis=0)).max()
        worst = max(worst, step)
        assert step < 2e-3, (trial, kind, pts)
    assert worst > 0      # the point does move, smoothly


def test_box_box_nearly_parallel_faces_give_face_contacts_in_both_precisions():
    """Pad-on-face configurations (a small box pressed flat on a big one with a tilt of 1e-5 .. 5e-3 rad, penetrating or
    inside the margin): the SAT must pick the FACE (4 contacts) in the fp64 oracle and in the fp32 kernel core alike -
    an edge axis within 2 degrees of a face normal is that face contact in disguise and used to be chosen by rounding noise."""
    import ctypes as C
    from oracle.oracle import collide
    from tests.emu.emu import lib
    L = lib("f32")
    dp = C.POINTER(C.c_double)
    L.emu_collide_boxes.argtypes = [dp, dp, dp, dp, dp, C.c_int, dp]
    L.emu_collide_boxes.restype = C.c_int
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)   # noqa: E731
    rng = np.random.default_rng(11)
    for trial in range(400):
        pA, hA = np.array([0.4, -0.2, 0.03]), np.array([0.03, 0.03, 0.03])
        hB = np.array([0.008, 0.004, 0.008])
        tilt = 10 ** rng.uniform(-5, -2.3) * rng.choice([-1, 1], 2)
        pB = pA + np.array([rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01), hA[2] + hB[1] - rng.uniform(5e-5, 6e-4)])
        w, x, y, z = np.cos(np.pi / 4), np.sin(np.pi / 4), 0.0, 0.0                  # pad lying on its broad face: its thin axis points up

        tq = _yaw_tilt_quat(rng.uniform(-0.3, 0.3), *tilt)
        q = np.array([tq[0] * w - tq[1] * x - tq[2] * y - tq[3] * z, tq[0] * x + tq[1] * w + tq[2] * z - tq[3] * y,
                      tq[0] * y - tq[1] * z + tq[2] * w + tq[3] * x, tq[0] * z + tq[1] * y - tq[2] * x + tq[3] * w])
        ref = collide(6, pA, [1, 0, 0, 0], hA, 6, pB, q, hB)
        out = np.zeros(56)
        n32 = L.emu_collide_boxes(d(pA), d(hA), d(pB), d(q), d(hB), 0, d(out))
        assert len(ref) == n32 and len(ref) >= 3, (trial, len(ref), n32, tilt)
        assert np.allclose(out[:7 * n32].reshape(-1, 7)[:, :3], ref[:, :3], atol=2e-6)


def test_grasp_holds_the_box_kkt_and_torsional_cone():
    """Stacking, condim-4 rows: after the scripted grasp-and-lift the box hangs between the pads — KKT stationarity of the
    Newton result, every contact force inside its (regularised) elliptic cone incl. the torsional component, and the
    contact forces on the box balance its weight and inertia (J^T f restricted to the box's linear dofs = m (a + g))."""
    from d3il_b200.scene.blob import load_scene
    from tests.util import scripted_grasp_actions, task_contexts
    blob, sc = load_scene("stacking")
    nv = sc.header["nv"]
    o = OracleEnv(blob, sc.header)
    ctx = task_contexts("stacking")[1]
    obs0 = o.reset(ctx)
    for a in scripted_grasp_actions(sc, ctx, o.robot_state(), o.joint_state()[:7], obs0):
        obs, r, d, info = o.step(a)
    assert obs[2] > 0.14                                   # red box in the air
    o.substep(1)
    M = o.probe("M").reshape(nv, nv); J = o.probe("efc_J").reshape(-1, nv)
    f = o.probe("efc_force"); qacc = o.probe("qacc"); qs = o.probe("qacc_smooth")
    assert np.abs(M @ (qacc - qs) - J.T @ f).max() < 1e-6
    con = o.probe("contacts").reshape(-1, 12)
    fr4 = np.array([2.0, 2.0, 0.05])                       # pad-box pair: friction (tangent, tangent, torsion) = max of the geoms'
    n4 = 0
    for c in con:
        e0, dim, mu = int(c[11]), int(c[7]), c[10]
        if e0 < 0:
            continue
        assert f[e0] >= -1e-12
        if dim == 4 and f[e0] > 1e-6 and (int(c[8]) in (2, 3)):      # finger-tip pads vs the red box
            n4 += 1
            t = np.sqrt(sum((f[e0 + 1 + j] / fr4[j]) ** 2 for j in range(3)))
            assert t <= f[e0] / mu * (1 + 1e-9) + 1e-9
    assert n4 >= 4
    # Newton's second law for the held box (dofs 9..11 = world-frame linear acceleration of the red box)
    lin = slice(9, 12)
    assert np.allclose((J.T @ f)[lin], 0.05 * (qacc[lin] + np.array([0, 0, 9.81])), atol=1e-6)
    assert abs(qacc[11]) < 2.0                             # and it is (nearly) held: far from free fall


# ---- analytic known answers of MuJoCo's soft-contact model [EXT], restated independently of oracle/d3il_oracle.c ---------
def _mj_impedance(solimp, r):
    """mj_makeImpedance / getimpedance (MuJoCo 2.3 engine_core_constraint.c): d(r) for |dist - margin| = r, power 2."""
    d0, dmax, width, mid, power = solimp
    x = min(abs(r) / width, 1.0)
    y = x ** power / mid ** (power - 1) if x <= mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
    return d0 + y * (dmax - d0)


def test_rest_penetration_is_the_soft_contact_known_answer(pushing_scene, pushing_contexts):
    """A box at rest on the table: every corner contact carries f = D * k * imp(r) * r with k = 1 / (dmax^2 tau^2 zeta^2)
    (solref = (tau, zeta)), D = imp / ((1 - imp) * invweight sum) — and the four of them carry m g: the penetration depth is
    r = (m g / 4) (1 - imp) w / (k imp^2).  Checked against the pair's solref / solimp from the compiled MJCF, not against
    anything the oracle computed for itself."""
    env, sc = make(pushing_scene)
    env.reset(pushing_contexts[5])
    env.substep(800)
    f = env.probe("efc_force"); con = env.probe("contacts").reshape(-1, 12); D = env.probe("efc_D")
    pair = sc.pair[0]                                   # table_plane - push box: same solref / solimp for both boxes
    tau, zeta = pair[8], pair[9]
    solimp = pair[10:15]
    k = 1.0 / (solimp[1] ** 2 * tau ** 2 * zeta ** 2)
    rows = [c for c in con if c[11] >= 0 and c[6] < 0]
    assert len(rows) == 8                                # 2 boxes x 4 corners
    for c in rows:
        r, e0 = -c[6], int(c[11])
        imp = _mj_impedance(solimp, r)
        assert abs(f[e0] - D[e0] * k * imp * r) < 1e-6 * f[e0]          # normal force = D * (-aref) at rest (a = v = 0)
        # D = 1 / R, R = (1 - imp) / imp * w with w = translational invweight0 of the two bodies = 1 / m + 0 (free box, world)
        w = sc.geom[int(pair[0])][13] + sc.geom[int(pair[1])][13]
        assert abs(w - 1 / 0.05) < 1e-9
        assert abs(D[e0] - imp / ((1 - imp) * w)) < 1e-9 * D[e0]
        r_closed = (0.05 * 9.81 / 4) * (1 - imp) * w / (k * imp * imp)    # closed form of the rest depth for a quarter of the weight
        assert abs(r - r_closed) < 0.02 * r, (r, r_closed)              # corners share the weight evenly to 2 % (box CoM is central)
    assert abs(sum(f[int(c[11])] for c in rows) - 2 * 0.05 * 9.81) < 1e-6


def test_sliding_friction_stays_on_the_elliptic_cone_and_dissipates():
    """A body given a horizontal velocity on the table (Aligning's 1.004 kg tray, priority-1 friction 0.3): while its
    contacts are active the friction force sits ON the elliptic cone, |f_t| = mu f_n, the sliding speed never increases, and
    the tray comes to rest.  (A textbook mu g deceleration is NOT a known answer of this model: in MuJoCo's convex contact
    formulation a sliding contact also produces extra normal force — the tray hops — which this restatement reproduces.)"""
    from d3il_b200.scene.blob import load_scene
    from tests.util import task_contexts
    blob, sc = load_scene("aligning")
    env = OracleEnv(blob, sc.header)
    env.reset(task_contexts("aligning")[0])
    env.substep(600)
    s = env.get_state()
    nq = sc.header["nq"]
    mu = 0.3
    s[nq + 9] = 0.6                                     # tray x velocity (free joint: world-frame linear dofs first)
    env.set_state(s)
    v_prev, n_active = 0.6, 0
    for k in range(400):
        env.substep(1)
        st = env.get_state()
        f = env.probe("efc_force"); con = env.probe("contacts").reshape(-1, 12)
        for c in con:
            e0 = int(c[11])
            if e0 >= 0 and f[e0] > 1e-9 and st[nq + 9] > 0.05:
                assert abs(np.hypot(f[e0 + 1], f[e0 + 2]) - mu * f[e0]) <= 1e-6 * f[e0]      # sliding: on the cone
                n_active += 1
        assert st[nq + 9] <= v_prev + 1e-7 or abs(st[nq + 9]) < 1e-4                                             # friction only dissipates (free flight between hops: unchanged to rounding)
        v_prev = st[nq + 9]
    assert n_active > 100 and abs(v_prev) < 1e-3                                       # it slid in contact, and it stopped
