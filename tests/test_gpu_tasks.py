"""GPU parity tests for the Sorting-k and Aligning scenes (run on the B200 box): CUDA path through the C ABI vs the fp64
oracle.  Same protocol as tests/test_gpu_parity.py; contact-rich env steps with a tipping box are ill-conditioned, so
the teacher-forced env-step bound is a quantile bound (see tests/test_emu_tasks.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from d3il_b200.scene.blob import load_scene          # noqa: E402
from oracle.oracle import OracleEnv                   # noqa: E402
from tests.util import oracle_rollout_states, scripted_task_actions, step_errors, task_contexts  # noqa: E402

TASKS = ["sorting_2", "sorting_4", "sorting_6", "aligning", "inserting"]


def _benv(task, n):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from d3il_b200.batched_env import BatchedEnv
    return BatchedEnv(task, n, 0)


@pytest.mark.parametrize("task", TASKS)
def test_task_reset_matches_oracle(task):
    """a12 for every committed context of the task: state install (+ Aligning target pose) and the single reset tick."""
    blob, sc = load_scene(task)
    nq, nv, nx = sc.header["nq"], sc.header["nv"], sc.header.get("nextra", 0)
    ctxs = task_contexts(task)
    n = len(ctxs)
    env = _benv(task, n)
    obs = env.reset(torch.tensor(ctxs, dtype=torch.float32, device="cuda")).cpu().numpy()
    o = OracleEnv(blob, sc.header)
    for i in range(n):
        oo = o.reset(ctxs[i])
        s_ref, s = o.get_state(), env.get_state(i)
        assert np.allclose(s[:nq], s_ref[:nq], rtol=1e-4, atol=2e-6), (i, np.abs(s[:nq] - s_ref[:nq]).max())
        # one tick of the spawn transient (80 mm inside the platform: ~200 m/s^2)
        # (overlapping spawn poses add deep box-box contacts): fp32 resolves the tick's velocity change to ~2e-3 of its largest component
        vtol = 2e-4 + 3e-3 * np.abs(s_ref[nq:nq + nv]).max()
        assert np.allclose(s[nq:nq + nv], s_ref[nq:nq + nv], rtol=1e-3, atol=vtol), (i, np.abs(s[nq:nq + nv] - s_ref[nq:nq + nv]).max())
        assert np.allclose(obs[i], oo, rtol=2e-4, atol=1e-5)
        if nx:
            assert np.allclose(s[-nx:], s_ref[-nx:], atol=1e-6)
    env.close()


@pytest.mark.parametrize("task", TASKS)
def test_task_env_step_teacher_forced(task):
    """(ii): full env steps from oracle states along a scripted push, all steps of the script in one batch."""
    blob, sc = load_scene(task)
    nq, nv = sc.header["nq"], sc.header["nv"]
    ctx = task_contexts(task)[2]
    o = OracleEnv(blob, sc.header)
    o.reset(ctx)
    acts = scripted_task_actions(task, ctx, o.robot_state(), n_steps=56)
    _, states, outs = oracle_rollout_states(task, ctx, acts)
    n = len(acts)
    env = _benv(task, n)
    env.reset(torch.tensor(np.repeat(ctx[None], n, 0), dtype=torch.float32, device="cuda"))
    for i in range(n):
        env.set_state(i, states[i])
    obs, rew, done, info = (t.cpu().numpy() for t in env.step(torch.tensor(acts, dtype=torch.float32, device="cuda")))
    o2 = OracleEnv(blob, sc.header)
    errs = []
    for i in range(n):
        o2.set_state(states[i])
        oo, rr, dd, ii = o2.step(acts[i])
        errs.append(step_errors(o2.get_state(), env.get_state(i), nq, nv))
        assert np.allclose(obs[i], oo, rtol=2e-4, atol=1e-5) and abs(rew[i] - rr) < 1e-4 and bool(done[i]) == dd      # sampled BEFORE the substeps
        assert np.array_equal(info[i, :2], ii[:2]) and info[i, 3] == 0
    errs = np.array(errs)
    assert (errs.max(axis=1) <= 1.0).mean() >= 0.85, errs
    # the tail (a box tipping over an edge) is bounded in physical units and shown to be fp32 precision, not logic, in
    # tests/test_gpu_protocol.py::test_env_step_tail_is_precision_not_logic (fp64 build of the same source inside the box)
    gots = [env.get_state(i) for i in range(n)]
    for i in range(n):
        o2.set_state(states[i]); o2.step(acts[i])
        ref = o2.get_state()
        assert np.abs(gots[i][:nq] - ref[:nq]).max() <= 1e-3, (i, np.abs(gots[i][:nq] - ref[:nq]).max())      # 1 mm / 1 mrad after 35 ticks
    env.close()


def test_sorting4_full_size_properties():
    """BASELINE config 3 size (8192 envs of Sorting-4): finite, no faults, boxes on the platform, envs sharing a context and
    an action stream stay bit-identical, determinism across two runs."""
    n = 8192
    ctxs = task_contexts("sorting_4")
    ctx = torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda")
    finals = []
    for rep in range(2):
        env = _benv("sorting_4", n)
        env.reset(ctx)
        tcp = env.robot_state().clone()
        des = torch.cat([tcp, torch.tensor([0, 1, 0, 0], device="cuda").repeat(n, 1)], 1)
        g = torch.Generator(device="cuda").manual_seed(0)
        lo, hi = torch.tensor([0.3, -0.45], device="cuda"), torch.tensor([0.8, 0.45], device="cuda")
        for k in range(30):
            d = (torch.rand(60, 2, generator=g, device="cuda") * 0.02 - 0.01).repeat((n + 59) // 60, 1)[:n]
            des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + d, lo), hi)
            obs, rew, done, info = env.step(des)
        assert torch.isfinite(obs).all() and torch.isfinite(info).all()
        assert (info[:, 3] == 0).all()
        finals.append(np.array([env.get_state(i) for i in (5, 65, 8165, 100)]))
        env.close()
    assert np.array_equal(finals[0], finals[1])
    assert np.array_equal(finals[0][0], finals[0][1]) and np.array_equal(finals[0][0], finals[0][2])
    z = finals[0][:, 9 + 2:37:7]
    assert np.all((z > 0.128) & (z < 0.132)), z


def test_stacking_reset_and_grasp_teacher_forced():
    """Stacking on the GPU: reset of all 100 shipped contexts, then the scripted grasp-and-lift teacher-forced per env
    step (joint-space action, gripper command, condim-4 contacts; k_env<4>)."""
    from tests.util import scripted_grasp_actions
    blob, sc = load_scene("stacking")
    nq, nv = sc.header["nq"], sc.header["nv"]
    ctxs = task_contexts("stacking")
    env = _benv("stacking", len(ctxs))
    obs = env.reset(torch.tensor(ctxs, dtype=torch.float32, device="cuda")).cpu().numpy()
    joints = env.joint_state().cpu().numpy()
    o = OracleEnv(blob, sc.header)
    for i in range(len(ctxs)):
        oo = o.reset(ctxs[i])
        s_ref, s = o.get_state(), env.get_state(i)
        assert np.allclose(s[:nq], s_ref[:nq], rtol=1e-4, atol=2e-6) and np.allclose(s[nq:nq + nv], s_ref[nq:nq + nv], rtol=1e-3, atol=2e-4)
        assert np.allclose(obs[i], oo, rtol=2e-4, atol=1e-5) and np.allclose(joints[i], o.joint_state(), atol=1e-6)
    env.close()
    ctx = ctxs[1]
    obs0 = o.reset(ctx)
    acts = scripted_grasp_actions(sc, ctx, o.robot_state(), o.joint_state()[:7], obs0)
    _, states, outs = oracle_rollout_states("stacking", ctx, acts)
    n = len(acts)
    env = _benv("stacking", n)
    env.reset(torch.tensor(np.repeat(ctx[None], n, 0), dtype=torch.float32, device="cuda"))
    for i in range(n):
        env.set_state(i, states[i])
    obs, rew, done, info = (t.cpu().numpy() for t in env.step(torch.tensor(acts, dtype=torch.float32, device="cuda")))
    o2 = OracleEnv(blob, sc.header)
    errs = []
    for i in range(n):
        o2.set_state(states[i])
        oo, rr, dd, ii = o2.step(acts[i])
        errs.append(step_errors(o2.get_state(), env.get_state(i), nq, nv))
        assert np.allclose(obs[i], oo, rtol=2e-4, atol=1e-5) and bool(done[i]) == dd
        assert np.array_equal(info[i, [0, 1, 3]], ii[[0, 1, 3]]) and info[i, 4] == 0
    errs = np.array(errs)
    assert (errs.max(axis=1) <= 1.0).mean() >= 0.85, errs
    assert errs[:, 0].max() * 5e-6 <= 1e-3 or errs[:, 0].max() <= 200      # worst qpos excursion <= 1 mm-equivalent (see test_gpu_protocol.py for the fp64 cross-check)
    assert outs[-1][0][2] > 0.14                      # the oracle lifted the red box
    env.close()


def test_boxes_off_the_table_land_on_the_ground():
    """Boxes released past the table's edges (front x = 0.89, sides y = +-0.98) fall onto the ground plane (z = -0.94) and
    rest there like the oracle's; the env next to them in the CTA, whose boxes stay on the table, is bit-identical to a run
    without falling neighbours; no fault bits."""
    blob, sc = load_scene("pushing")
    nq, nv = sc.header["nq"], sc.header["nv"]
    ctx = task_contexts("pushing")[0]
    spots = [(1.0, 0.05, 0.0), (1.05, -0.4, 0.02), (0.5, 1.1, 0.0), (0.3, -1.12, 0.05)]
    n = 2 * len(spots)
    env = _benv("pushing", n)
    env.reset(torch.tensor(np.repeat(ctx[None], n, 0), dtype=torch.float32, device="cuda"))
    o = OracleEnv(blob, sc.header)
    o.reset(ctx)
    s0 = o.get_state()
    refs = []
    for i, sp in enumerate(spots):
        s = s0.copy(); s[9:12] = sp; s[nq:nq + nv] = 0
        env.set_state(2 * i, s)                       # odd envs keep the untouched reset state
        o.set_state(s); o.substep(900); refs.append(o.get_state())
    env.substep(900)
    for i, sp in enumerate(spots):
        got, ref = env.get_state(2 * i), refs[i]
        assert abs(ref[11] - (-0.94 + 0.03)) < 5e-4, ref[9:12]                           # the oracle's box rests on the ground
        assert np.abs(got[9:12] - ref[9:12]).max() < 2e-3 and abs(got[11] - ref[11]) < 2e-4, (sp, got[9:12], ref[9:12])
        assert np.abs(got[nq + 9:nq + 15]).max() < 5e-3                                   # at rest
    o.set_state(s0); o.substep(900)
    quiet = o.get_state()
    for i in range(len(spots)):
        got = env.get_state(2 * i + 1)
        assert np.array_equal(got, env.get_state(1))                                      # neighbours of falling boxes: bit-identical to each other
        assert np.allclose(got[:nq], quiet[:nq], rtol=1e-4, atol=5e-6)
    env.close()


def test_box_pushed_into_box_teacher_forced():
    """Several contacts that couple kinematic trees at once (rod -> box 1 -> box 2: one rod contact + the box-box face contacts):
    the Newton Hessian is dense over arm + both boxes and takes the dense rolled Cholesky of the kernel, not the block /
    Woodbury route.  Teacher-forced env steps against the oracle along the push."""
    import json, os
    blob, sc = load_scene("pushing")
    nq, nv = sc.header["nq"], sc.header["nv"]
    ctx = np.array([[0.5, -0.12, 0.0, 1.0, 0.0, 0.0, 0.0], [0.5, -0.055, 0.0, 1.0, 0.0, 0.0, 0.0]])
    o = OracleEnv(blob, sc.header)
    o.reset(ctx)
    from tests.util import scripted_push_actions
    acts = scripted_push_actions(ctx, o.robot_state(), n_steps=90, approach_steps=45)
    names = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "d3il_b200", "scenes", "pushing.json")))["geoms"]
    rod, b1, b2 = names.index("rod:geom_rb0"), names.index("push_box:geom"), names.index("push_box2:geom")
    states, multi = [o.get_state()], []
    for a in acts:
        o.step(a)
        states.append(o.get_state())
        con = o.probe("contacts").reshape(-1, 12)
        act = con[con[:, 11] >= 0]
        pairs = {(int(r[8]), int(r[9])) for r in act}
        multi.append(int(any(rod in p for p in pairs)) + sum(1 for r in act if {int(r[8]), int(r[9])} == {b1, b2}))
    multi = np.array(multi)
    assert (multi >= 2).sum() >= 10, multi                       # the script really produces multi-coupling ticks
    n = len(acts)
    env = _benv("pushing", n)
    env.reset(torch.tensor(np.repeat(ctx[None], n, 0), dtype=torch.float32, device="cuda"))
    for i in range(n):
        env.set_state(i, states[i])
    obs, rew, done, info = (t.cpu().numpy() for t in env.step(torch.tensor(acts, dtype=torch.float32, device="cuda")))
    assert (info[:, -1] == 0).all()
    errs = np.array([step_errors(states[i + 1], env.get_state(i), nq, nv) for i in range(n)])
    sel = multi >= 2
    print(f"[box into box] multi-coupling steps {sel.sum()}: inside the tolerance box {(errs[sel].max(axis=1) <= 1.0).mean():.3f}, median {np.median(errs[sel].max(axis=1)):.3f}; all steps inside {(errs.max(axis=1) <= 1.0).mean():.3f}")
    assert (errs.max(axis=1) <= 1.0).mean() >= 0.85 and (errs[sel].max(axis=1) <= 1.0).mean() >= 0.8, errs
    for i in range(n):
        got = env.get_state(i)
        assert np.abs(got[:nq] - states[i + 1][:nq]).max() <= 1e-3, (i, np.abs(got[:nq] - states[i + 1][:nq]).max())
    env.close()
