"""GPU parity protocol of SURVEY.md §8c for EVERY scene (run on the B200 box), CUDA path through the C ABI vs the fp64
oracle:
  (i)  teacher-forced single physics ticks along contact-rich scripted episodes;
  (ii) teacher-forced env steps — with the tail of ill-conditioned steps (a box tipping over an edge, a closing gripper)
       bounded in physical units AND shown to be precision, not logic: on exactly those samples the fp64 host build of the
       same kernel source (tests/emu) sits inside the tolerance box;
  (iv) contact-rich closed-loop episodes with a scripted feedback policy over all shipped contexts: trajectories are
       chaotic beyond a few env steps, so the task metrics (success, mode, mean_distance) are compared as distributions.
Tolerance box ("units"): qpos rtol 1e-4 + atol 5e-6 (north_star's 1e-4), qvel rtol 1e-3 + atol 2e-4 — velocities carry
h x acceleration of stiff contacts (up to 1e3 rad/s^2), which fp32 resolves to ~1e-4 of the LARGEST acceleration in the
system, hence the wider velocity box."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from d3il_b200.scene.blob import load_scene          # noqa: E402
from oracle.oracle import OracleEnv                   # noqa: E402
from tests.util import (explain_tail, oracle_rollout_states, scripted_grasp_actions, scripted_task_actions, task_contexts, units,  # noqa: E402
                        with_joint_setpoint, with_setpoint)

CART_TASKS = ["sorting_2", "sorting_4", "sorting_6", "aligning", "inserting"]


def _benv(task, n):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from d3il_b200.batched_env import BatchedEnv
    return BatchedEnv(task, n, 0)


def _script(task):
    blob, sc = load_scene(task)
    ctx = task_contexts(task)[2 if task != "stacking" else 1]
    o = OracleEnv(blob, sc.header)
    obs0 = o.reset(ctx)
    if task == "stacking":
        acts = scripted_grasp_actions(sc, ctx, o.robot_state(), o.joint_state()[:7], obs0)
    else:
        acts = scripted_task_actions(task, ctx, o.robot_state(), n_steps=56)
    return blob, sc, ctx, acts


@pytest.mark.parametrize("task", CART_TASKS + ["stacking"])
def test_single_tick_teacher_forced_scene(task):
    """(i) for every scene: two ticks (first and middle) of every env step of the scripted episode, one GPU tick vs one
    oracle tick from the same state."""
    blob, sc, ctx, acts = _script(task)
    nq, nv = sc.header["nq"], sc.header["nv"]
    _, states, _ = oracle_rollout_states(task, ctx, acts)
    o2 = OracleEnv(blob, sc.header)
    starts = []
    for k in range(len(acts)):
        s = with_setpoint(states[k], sc, acts[k]) if sc.header["act_dim"] == 7 else with_joint_setpoint(states[k], sc, acts[k])
        o2.set_state(s)
        for t in range(18):
            if t in (0, 17):
                starts.append(o2.get_state())
            o2.substep(1)
    n = len(starts)
    env = _benv(task, n)
    env.reset(torch.tensor(np.repeat(ctx[None], n, 0), dtype=torch.float32, device="cuda"))
    for i, s in enumerate(starts):
        env.set_state(i, s)
    env.substep(1)
    refs, gots = [], []
    for i, s in enumerate(starts):
        o2.set_state(s); o2.substep(1)
        refs.append(o2.get_state()); gots.append(env.get_state(i))
    env.close()
    u = np.array([units(r, g, nq, nv, s) for s, r, g in zip(starts, refs, gots)])
    tail = explain_tail(task, starts, None, refs, gots, nq, nv, single_tick=True)
    print(f"[{task}] single tick: n={n} inside={np.mean(u.max(1) <= 1):.3f} worst q {u[:, 0].max():.2f} v {u[:, 1].max():.2f}; tail {[(t['i'], round(max(t['u32']), 1), round(max(t['u64']), 3), t['dq'], t['dv']) for t in tail]}")
    assert u[:, 0].max() <= 1.0, u[:, 0].max()                  # positions after one tick: always inside (rtol 1e-4, atol 5e-6)
    assert np.mean(u.max(1) <= 1.0) >= 0.95
    for t in tail:
        assert max(t["u64"]) <= 0.05, t                         # same source in fp64: 20x inside the box on the very samples fp32 leaves it
        assert t["dv"] <= 2e-2, t                               # physical bound on the excursion: 2 cm/s (rad/s) after one tick


@pytest.mark.parametrize("task", CART_TASKS + ["stacking", "pushing"])
def test_env_step_tail_is_precision_not_logic(task):
    """(ii) with an explained tail: every env step of the scripted episode teacher-forced; samples outside the box are
    bounded physically and reproduced inside the box by the fp64 build of the same source."""
    if task == "pushing":
        from tests.util import scripted_push_actions
        blob, sc = load_scene(task)
        ctx = task_contexts(task)[0]
        o = OracleEnv(blob, sc.header); o.reset(ctx)
        acts = scripted_push_actions(ctx, o.robot_state())
    else:
        blob, sc, ctx, acts = _script(task)
    nq, nv = sc.header["nq"], sc.header["nv"]
    _, states, _ = oracle_rollout_states(task, ctx, acts)
    n = len(acts)
    env = _benv(task, n)
    env.reset(torch.tensor(np.repeat(ctx[None], n, 0), dtype=torch.float32, device="cuda"))
    for i in range(n):
        env.set_state(i, states[i])
    obs, rew, done, info = (t.cpu().numpy() for t in env.step(torch.tensor(acts, dtype=torch.float32, device="cuda")))
    o2 = OracleEnv(blob, sc.header)
    refs, gots = [], []
    for i in range(n):
        o2.set_state(states[i]); o2.step(acts[i])
        refs.append(o2.get_state()); gots.append(env.get_state(i))
    env.close()
    u = np.array([units(r, g, nq, nv) for r, g in zip(refs, gots)])
    tail = explain_tail(task, states, acts, refs, gots, nq, nv)
    print(f"[{task}] env step: n={n} inside={np.mean(u.max(1) <= 1):.3f} median {np.median(u.max(1)):.3f} worst q {u[:, 0].max():.1f} v {u[:, 1].max():.1f}; tail {[(t['i'], round(max(t['u32']), 1), round(max(t['u64']), 3), t['dq'], t['dv']) for t in tail]}")
    assert (info[:, -1] == 0).all()                              # no fault bits (incl. the Newton iteration cap)
    assert np.mean(u.max(1) <= 1.0) >= 0.85
    for t in tail:
        assert max(t["u64"]) <= 0.1, t                           # fp64 build of the same source: inside the box
        assert t["dq"] <= 1e-3 and t["dv"] <= 0.5, t             # physical bound on the fp32 excursion after 35 ticks of a tipping / tumbling box: 1 mm (mrad), 0.5 rad/s


@pytest.mark.parametrize("task,max_steps", [("pushing", 400), ("sorting_2", 160), ("sorting_4", 200), ("sorting_6", 200), ("aligning", 300)])
def test_closed_loop_task_metrics_match_oracle(task, max_steps):
    """(iv): the same scripted feedback policy closed loop on the GPU batch (one env per shipped context) and on the fp64
    oracle.  Contact-rich trajectories decorrelate after a few env steps, so what is compared is what the benchmark
    reports: success, mode and mean_distance, context by context and as distributions."""
    import multiprocessing as mp
    from tests.util import iv_oracle_episode, iv_policy_step
    blob, sc = load_scene(task)
    ctxs = task_contexts(task)
    n = len(ctxs)
    with mp.get_context("fork").Pool(min(16, mp.cpu_count())) as pool:
        job = pool.map_async(iv_oracle_episode, [(task, ci, max_steps) for ci in range(n)])
        env = _benv(task, n)
        obs = env.reset(torch.tensor(ctxs, dtype=torch.float32, device="cuda")).cpu().numpy()
        des = env.robot_state().cpu().numpy().astype(np.float64)
        phase = np.zeros(n, int)
        rows, steps, active = np.zeros((n, env.info_dim)), np.zeros(n, int), np.ones(n, bool)
        quat = np.tile([0.0, 1.0, 0.0, 0.0], (n, 1))
        for k in range(max_steps):
            for i in range(n):
                if active[i]:
                    des[i], phase[i] = iv_policy_step(task, i, obs[i], des[i], phase[i])
            o_t, _, done, info = env.step(torch.tensor(np.concatenate([des, quat], 1), dtype=torch.float32, device="cuda"))
            obs, done, info = o_t.cpu().numpy(), done.cpu().numpy().astype(bool), info.cpu().numpy()
            fin = active & (done | (k == max_steps - 1))
            rows[fin], steps[fin] = info[fin], k + 1
            active &= ~fin
            if not active.any():
                break
        env.close()
        ref = job.get(timeout=600)
    ref_rows, ref_steps = np.array([r[0] for r in ref]), np.array([r[1] for r in ref])
    assert (rows[:, -1] == 0).all(), "fault bits raised on the GPU"
    succ, succ_ref = rows[:, 0], ref_rows[:, 0]
    print(f"[{task}] success GPU {succ.mean():.3f} oracle {succ_ref.mean():.3f}; modes GPU {np.bincount(rows[:, 1].astype(int) + 1)} oracle {np.bincount(ref_rows[:, 1].astype(int) + 1)}; "
          f"mean_distance GPU {rows[:, 2].mean():.4f} oracle {ref_rows[:, 2].mean():.4f}; steps GPU {steps.mean():.1f} oracle {ref_steps.mean():.1f}; per-context success agreement {(succ == succ_ref).mean():.3f}")
    assert abs(succ.mean() - succ_ref.mean()) <= 0.1                                   # success rate (binomial sigma at n = 60: 0.06)
    # ... and context by context (Aligning's 1.8 cm / 8.6 degree success box at the end of a 250-step push is a coin flip for
    # rollouts that end near its boundary: distribution-level agreement is the claim there)
    assert (succ == succ_ref).mean() >= (0.7 if task == "aligning" else 0.85)
    both = (succ == 1) & (succ_ref == 1)
    if task == "pushing":
        assert (rows[both, 1] == ref_rows[both, 1]).all()                              # successful rollouts: the mode is the visiting order the script dictates
        assert succ_ref.mean() >= 0.8                                                  # the script does solve the task (the comparison is not vacuous)
    if task.startswith("sorting"):      # one box per context is pushed into its bin: packed mode word and mode_step (sorting.py:464-543)
        assert ((rows[:, 1] == ref_rows[:, 1]) & (rows[:, 2] == ref_rows[:, 2])).mean() >= 0.85
        assert (ref_rows[:, 2] >= 1).mean() >= 0.5
    if task in ("pushing", "aligning"):
        assert abs(rows[:, 2].mean() - ref_rows[:, 2].mean()) <= 0.1 * ref_rows[:, 2].mean() + 2e-3      # mean_distance, distribution mean
        assert np.abs(rows[both, 2] - ref_rows[both, 2]).max(initial=0) <= 0.02        # successful rollouts end within 2 cm of the same configuration
    assert abs(steps.mean() - ref_steps.mean()) <= 0.1 * ref_steps.mean()              # episode lengths


def test_stacking_scripted_grasps_all_contexts_match_oracle():
    """(iv) for Stacking: the scripted grasp-and-lift of the red box (joint-space actions, gripper command, condim-4 pad contacts,
    a held box) on ALL 100 shipped contexts, one env per context on the GPU and one oracle episode each.  The script lifts the
    box in ~60 % of the contexts (the others present it at a yaw the open-loop approach cannot close on): which ones, how high,
    and the info rows must agree."""
    import multiprocessing as mp
    from tests.util import iv_stacking_episode
    ctxs = task_contexts("stacking")
    n = len(ctxs)
    with mp.get_context("fork").Pool(min(16, mp.cpu_count())) as pool:
        ref = pool.map(iv_stacking_episode, range(n))
    acts = np.stack([r[0] for r in ref])                      # [n, steps, 8]
    ref_info, ref_z = np.array([r[1] for r in ref]), np.array([r[2] for r in ref])
    env = _benv("stacking", n)
    env.reset(torch.tensor(ctxs, dtype=torch.float32, device="cuda"))
    info = None
    for k in range(acts.shape[1]):
        _, _, _, info = env.step(torch.tensor(acts[:, k], dtype=torch.float32, device="cuda"))
    info = info.cpu().numpy()
    z = np.array([env.get_state(i)[9 + 2] for i in range(n)])
    env.close()
    lifted, lifted_ref = z > 0.1, ref_z > 0.1
    both = lifted & lifted_ref
    print(f"[stacking] lifted GPU {lifted.mean():.2f} oracle {lifted_ref.mean():.2f}, per-context agreement {(lifted == lifted_ref).mean():.2f}; "
          f"height difference among both-lifted max {np.abs(z - ref_z)[both].max(initial=0):.2e} m; mean_distance GPU {info[:, 2].mean():.4f} oracle {ref_info[:, 2].mean():.4f}")
    assert (info[:, -1] == 0).all(), "fault bits raised on the GPU"
    assert lifted_ref.mean() >= 0.4                                                    # the comparison is not vacuous
    assert (lifted == lifted_ref).mean() >= 0.9 and abs(lifted.mean() - lifted_ref.mean()) <= 0.08
    assert np.abs(z - ref_z)[both].max(initial=0) <= 2e-3                              # held boxes hang at the same height
    assert np.array_equal(info[:, :2], ref_info[:, :2]) and np.array_equal(info[:, 3], ref_info[:, 3])      # success, mode digits, len(mode)
    assert abs(info[:, 2].mean() - ref_info[:, 2].mean()) <= 0.02 * ref_info[:, 2].mean() + 1e-3


def test_faults_are_reported_per_env_and_never_hang():
    """Fault handling inside multi-env CTAs: a NaN action, a zero quaternion and a NaN state in three envs of
    one batch.  The step must RETURN (the Newton loop is CTA-uniform: an env that stops voting would dead-lock its CTA),
    the faulty envs carry their status bits (16 bad action: set-point held; 4 / 8 solver), and every other env of the batch —
    including the CTA neighbours — is bit-identical to a clean run."""
    from tests.util import random_walk_actions
    ctxs = task_contexts("pushing")
    n = 21
    ctx = torch.tensor(ctxs[:n], dtype=torch.float32, device="cuda")
    finals = []
    for faulty in (False, True):
        env = _benv("pushing", n)
        env.reset(ctx)
        tcp = env.robot_state().cpu().numpy()
        acts = np.concatenate([tcp, np.tile([0, 1, 0, 0], (n, 1))], 1).astype(np.float32)
        acts[:, 0] += 0.004
        if faulty:
            s = env.get_state(9)
            s[23 + 12] = np.nan                       # a box velocity
            env.set_state(9, s)
        infos = []
        for k in range(3):
            a = acts.copy()
            a[:, 1] += 0.003 * k
            if faulty and k == 1:
                a[3, 0] = np.nan
                a[5, 3:] = 0
            obs, rew, done, info = env.step(torch.tensor(a, device="cuda"))
            torch.cuda.synchronize()
            infos.append(info.cpu().numpy().copy())
        finals.append((np.array([env.get_state(i) for i in range(n)]), infos))
        env.close()
    (clean, ci), (dirty, di) = finals
    status = di[-1][:, -1].astype(int)
    assert status[3] & 16 and status[5] & 16, status
    assert status[9] & (4 | 8 | 1), status
    ok = [i for i in range(n) if i not in (3, 5, 9)]
    assert (status[ok] == 0).all() and (ci[-1][:, -1] == 0).all()
    assert np.array_equal(clean[ok], dirty[ok])
    # a held set-point: envs 3 and 5 skipped one set-point update but stayed finite
    assert np.isfinite(dirty[3]).all() and np.isfinite(dirty[5]).all()
