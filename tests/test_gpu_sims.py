"""GPU smoke tests of the drop-in rollout harnesses (reference class names / signatures) with synthetic policies, and of
the mixed-task batch."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_sims_run_with_synthetic_policies():
    _need_gpu()
    from d3il_b200.simulation import Aligning_Sim, Avoiding_Sim, Pushing_Sim, Sorting_Sim, Stacking_Sim
    from d3il_b200.simulation.policies import SyntheticBCPolicy, SyntheticDDPMPolicy

    s, m, d = Pushing_Sim(seed=0, device="cuda:0", render=False, n_contexts=4, n_trajectories_per_context=2).test_agent(SyntheticBCPolicy(10, 2))
    assert s.shape == (4, 2) and m.shape == (4, 2) and torch.isfinite(d).all() and (d > 0).all()
    s, e = Avoiding_Sim(seed=0, device="cuda:0", render=False, n_trajectories=6).test_agent(SyntheticBCPolicy(4, 2))
    assert s.shape == (6,)
    s, m, d = Aligning_Sim(seed=0, device="cuda:0", render=False, n_contexts=3, n_trajectories_per_context=2).test_agent(SyntheticBCPolicy(20, 3))
    assert s.shape == (3, 2) and set(m.flatten().tolist()) <= {0.0, 1.0}
    sr, m = Sorting_Sim(seed=0, device="cuda:0", render=False, n_contexts=3, n_trajectories_per_context=2, num_box=4, max_steps_per_episode=60).test_agent(
        SyntheticDDPMPolicy(16, 2))
    assert m.shape == (3, 2) and 0.0 <= sr <= 1.0 and (m == 240).all()          # nothing sorted in 60 random steps: mode bits all "unset"
    from d3il_b200.simulation import Inserting_Sim
    s, m, d = Inserting_Sim(seed=0, device="cuda:0", render=False, n_contexts=3, n_trajectories_per_context=2, max_steps_per_episode=40).test_agent(SyntheticBCPolicy(13, 2))
    assert s.shape == (3, 2) and (s == 0).all() and (d > 0.2).all()
    s, m = Stacking_Sim(seed=0, device="cuda:0", render=False, n_contexts=3, n_trajectories_per_context=2, max_steps_per_episode=40).test_agent(
        SyntheticBCPolicy(20, 8, width=256, n_hidden_layers=8))
    assert s.shape == (3, 2) and (s == 0).all()


def test_mixed_batch_steps_all_seven_configs():
    _need_gpu()
    from d3il_b200.mixed import SEVEN_CONFIGS, MixedBatch
    from tests.util import TASK_CONTEXT_FILES, task_contexts

    mb = MixedBatch(70, 0)
    assert mb.tasks == SEVEN_CONFIGS and sum(mb.counts) == 70
    ctxs = [torch.tensor(task_contexts(t)[np.arange(e.n_envs) % 60], dtype=torch.float32, device="cuda") if t in TASK_CONTEXT_FILES else None
            for t, e in zip(mb.tasks, mb.envs)]
    mb.reset(ctxs)
    acts = []
    for e in mb.envs:
        if e.act_dim == 8:
            a = e.joint_state().clone(); a[:, 7] = 0.08
        else:
            a = torch.cat([e.robot_state().clone(), torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(e.n_envs, 1)], 1)
        acts.append(a.contiguous())
    for _ in range(3):
        outs = mb.step(acts)
    torch.cuda.synchronize()
    for (obs, rew, done, info), e in zip(outs, mb.envs):
        assert obs.shape == (e.n_envs, e.obs_dim) and torch.isfinite(obs).all() and (info[:, -1] == 0).all()
    mb.close()


def test_trajectory_recorder_matches_dataset_format():
    """The recorded env_state dicts feed the reference's Pushing dataset arithmetic (pushing_dataset.py:55-77) unchanged."""
    _need_gpu()
    from d3il_b200.batched_env import BatchedEnv
    from d3il_b200.simulation.recorder import TrajectoryRecorder
    from tests.util import task_contexts

    n = 4
    env = BatchedEnv("pushing", n, 0)
    ctx = task_contexts("pushing")[:n]
    env.reset(torch.tensor(ctx, dtype=torch.float32, device="cuda"))
    rec = TrajectoryRecorder(env)
    des = torch.cat([env.robot_state().clone(), torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
    for k in range(6):
        des[:, 1] += 0.005
        rec.record(des)
        env.step(des)
    states = rec.env_states()
    assert len(states) == n
    st = states[1]
    assert st["robot"]["des_c_pos"].shape == (6, 3) and st["red-box"]["quat"].shape == (6, 4) and st["green-target"]["pos"].shape == (6, 3)
    assert np.allclose(st["red-box"]["pos"][0, :2], ctx[1, 0, :2], atol=1e-3) and abs(st["red-box"]["pos"][-1, 2] - 0.011) < 1e-3
    vel = st["robot"]["des_c_pos"][1:, :2] - st["robot"]["des_c_pos"][:-1, :2]           # the dataset's action definition
    assert np.allclose(vel, [[0.0, 0.005]] * 5, atol=1e-6)
    env.close()
