"""Host-side pieces that need no GPU: C-ABI exports, scene tables, metrics, sharding + gather (gloo, world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    from d3il_b200 import lib
    lib.build()
    L = C.CDLL(lib.SO_PATH)
    header = open(os.path.join(ROOT, "include", "d3il.h")).read()
    declared = set(re.findall(r"\b(d3il_[a-z_]+)\s*\(", header))
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: on a machine without CUDA d3il_create must return an error, never a working handle."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from d3il_b200 import lib
    from d3il_b200.scene.blob import load_scene
    L = lib.lib()
    blob, _ = load_scene("pushing")
    h = C.c_void_p()
    rc = L.d3il_create(C.byref(h), blob, len(blob), 4, 0)
    assert rc != 0 and not h.value and b"CUDA" in L.d3il_last_error()
    with pytest.raises(RuntimeError):
        from d3il_b200.batched_env import BatchedEnv
        BatchedEnv("pushing", 4, "cpu")


def test_scene_tables_known_answers():
    """Compiler output vs SURVEY §8c known answers (offline IK start pose) and App. A sizes."""
    from d3il_b200.scene import blob as B
    _, sc = B.load_scene("pushing")
    h = sc.header
    assert (h["nq"], h["nv"], h["nlink"], h["nobj"], h["n_substeps"], h["max_steps"], h["obs_dim"]) == (23, 21, 11, 2, 35, 400, 8)
    init_qpos = sc.ctrl[B.C_INIT_QPOS:B.C_INIT_QPOS + 7]
    assert np.allclose(init_qpos, [-0.356409, 0.429445, -0.135011, -2.054652, 0.093578, 2.480718, 0.232507], atol=2e-6)
    # free boxes: invweight0 = (1/m, 1/I) for a 0.05 kg, 6 cm cube
    box = sc.geom[4]      # table_plane, support_body, ground, rod, then the boxes
    assert abs(box[13] - 20.0) < 1e-9 and abs(box[14] - 1.0 / (0.05 / 3 * 2 * 0.03 ** 2)) < 1e-3
    # mixed box-table contact parameters (App. B.5): solref[0] = (0.002 + 0.02)/2, friction 1
    pair = sc.pair[0]
    assert abs(pair[8] - 0.011) < 1e-12 and pair[3] == 1.0 and int(pair[2]) == 3
    _, av = B.load_scene("avoiding")
    assert (av.header["nq"], av.header["nv"], av.header["npair"], av.header["max_steps"]) == (9, 9, 6, 250)


def test_mode_entropy_matches_reference_loops():
    """metrics.mode_entropy == the per-context Python loops of simulation/pushing_sim.py:140-167."""
    from d3il_b200.simulation.metrics import avoiding_entropy, mode_entropy
    g = torch.Generator().manual_seed(0)
    n_ctx, n_traj, n_modes = 30, 8, 4
    modes = torch.randint(-1, 4, (n_ctx, n_traj), generator=g).float()
    succ = (torch.rand(n_ctx, n_traj, generator=g) > 0.4).float()
    probs, ent = mode_entropy(modes, succ, n_modes)
    ref = torch.zeros(n_ctx, n_modes)
    for c in range(n_ctx):
        ref[c, :] = torch.tensor([sum(modes[c, succ[c, :] == 1] == k) / n_traj for k in range(n_modes)])
    ref /= (ref.sum(1).reshape(-1, 1) + 1e-12)
    ent_ref = -(ref * torch.log(ref + 1e-12) / torch.log(torch.tensor(float(n_modes)))).sum(1).mean()
    assert torch.allclose(probs, ref) and torch.allclose(ent, ent_ref)
    # avoiding: numpy restatement of avoiding_sim.py:128-134
    enc = (torch.rand(40, 9, generator=g) > 0.6).float()
    s = (torch.rand(40, generator=g) > 0.3).float()
    data = enc[s == 1].numpy()
    dec = data.dot(1 << np.arange(9))
    _, counts = np.unique(dec, return_counts=True)
    dist = counts / counts.sum()
    assert abs(float(avoiding_entropy(enc, s)[1]) - float(-np.sum(dist * (np.log(dist) / np.log(24))))) < 1e-6


def test_agent_adapter_matches_single_sample_predict():
    """Batched BC-style path == looping the agent's own predict() (bc_agent.py:241-271 restated on a stub agent)."""
    from d3il_b200.simulation.agent_adapter import predict_batch

    class Scaler:
        def __init__(self):
            self.x_mean, self.x_std = torch.linspace(-0.2, 0.3, 10), torch.linspace(0.5, 1.5, 10)
            self.y_mean, self.y_std = torch.tensor([0.001, -0.002]), torch.tensor([0.004, 0.005])

        def scale_input(self, x):
            return ((x - self.x_mean) / (self.x_std + 1e-12)).float()

        def inverse_scale_output(self, y):
            return y * (self.y_std + 1e-12) + self.y_mean

    class BC_Agent:                    # dispatch is by class name (agents/bc_agent.py)
        def __init__(self):
            torch.manual_seed(0)
            self.model = torch.nn.Sequential(torch.nn.Linear(10, 32), torch.nn.Mish(), torch.nn.Linear(32, 2))
            self.scaler, self.device = Scaler(), "cpu"
            self.min_action, self.max_action = torch.tensor([-1.5, -1.5]), torch.tensor([1.5, 1.5])

        @torch.no_grad()
        def predict(self, state):
            st = torch.from_numpy(state).float().unsqueeze(0).unsqueeze(0)
            out = self.model(self.scaler.scale_input(st)).clamp_(self.min_action, self.max_action)
            return self.scaler.inverse_scale_output(out).numpy()[0]

        def reset(self):
            pass

    agent = BC_Agent()
    obs = torch.randn(17, 10)
    batched = predict_batch(agent, obs)
    looped = np.stack([agent.predict(o.numpy())[0] for o in obs])
    assert np.allclose(batched.numpy(), looped, atol=1e-6)


def _gloo_worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from d3il_b200.simulation.base_sim import BaseSim
    lo, hi = BaseSim.shard_range(n_items, rank, world)
    rows = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1) * torch.tensor([[1.0, 10.0, 100.0]])
    out = BaseSim.gather_rows(rows, n_items)
    q.put((rank, lo, hi, out.numpy()))
    dist.destroy_process_group()


def test_sharding_and_gather_world_size_2():
    """N>1 path of the rollout harness on CPU: contiguous (context, rollout) shards per rank + one gather of the
    per-env result rows (gloo, world_size 2) — the replacement of the share_memory_() tensors in pushing_sim.py:97-99."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_items, world, port = 37, 2, 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    expect = np.arange(n_items, dtype=np.float32)[:, None] * np.array([[1.0, 10.0, 100.0]], dtype=np.float32)
    ranges = sorted((lo, hi) for _, lo, hi, _ in res)
    assert ranges[0][0] == 0 and ranges[-1][1] == n_items and ranges[0][1] == ranges[1][0]
    for _, _, _, out in res:
        assert np.array_equal(out, expect)


def test_shard_ranges_are_balanced_and_an_empty_shard_still_gathers():
    """ceil-sized shards left trailing ranks empty (9 items on 8 ranks) and an empty rank raised in d3il_create
    while the others blocked in all_gather.  Ranges are balanced now; with fewer items than ranks the empty rank contributes a
    [0, k] tensor to the same collective (gloo, world_size 2, ONE item)."""
    import torch.multiprocessing as mp
    from d3il_b200.simulation.base_sim import BaseSim
    for n_items, world in ((9, 8), (37, 2), (64, 8), (5, 8), (0, 4)):
        rs = [BaseSim.shard_range(n_items, r, world) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == n_items and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        sizes = [hi - lo for lo, hi in rs]
        assert max(sizes) - min(sizes) <= 1 and (min(sizes) >= 1 or n_items < world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 30100 + os.getpid() % 500
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, 1, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted((lo, hi) for _, lo, hi, _ in res) == [(0, 1), (1, 1)]
    for _, _, _, out in res:
        assert out.shape == (1, 3) and np.array_equal(out, np.zeros((1, 3), np.float32))


def test_mode_kl_matches_reference_loops():
    """``Sorting_Sim`` metrics (sorting_sim.py:191-213) restated with the reference's Python loops vs the vectorised form."""
    from d3il_b200.simulation.metrics import mode_kl

    torch.manual_seed(1)
    prior = {192: 0.5, 64: 0.3, 128: 0.2}
    keys = torch.tensor(list(prior.keys()), dtype=torch.float32)
    n_ctx, n_traj = 7, 12
    mode_encoding = keys[torch.randint(0, 3, (n_ctx, n_traj))]
    successes = (torch.rand(n_ctx, n_traj) > 0.4).float()
    successes[3] = 0                                    # a context without any success is dropped
    probs, entropy, KL = mode_kl(mode_encoding, successes, prior)
    n_mode = 3
    ref = torch.zeros(n_ctx, n_mode)
    for c in range(n_ctx):
        for num in range(n_mode):
            ref[c, num] = sum(mode_encoding[c, successes[c, :] == 1] == keys[num]) / n_traj
    ref /= (ref.sum(1).reshape(-1, 1) + 1e-12)
    ref = ref[torch.nonzero(ref.sum(1), as_tuple=True)[0]]
    prior_t = torch.tensor(list(prior.values()))
    ent_ref = -(ref * torch.log(ref + 1e-12) / torch.log(torch.tensor(float(n_mode)))).sum(1).mean()
    log_ = (ref * torch.log(prior_t + 1e-12) / torch.log(torch.tensor(float(n_mode)))).sum(1).mean()
    assert torch.allclose(probs, ref, atol=1e-6) and abs(entropy - ent_ref.item()) < 1e-6 and abs(KL - (-ent_ref - log_).item()) < 1e-6


def test_stacking_mode_string_round_trip_and_synthetic_policies():
    from d3il_b200.simulation.policies import SyntheticBCPolicy, SyntheticDDPMPolicy
    from d3il_b200.simulation.stacking_sim import MODE_3, decode_mode

    for s in MODE_3:
        code = sum(("rgb".index(ch) + 1) * 4 ** k for k, ch in enumerate(s))
        assert decode_mode(code, 3) == s and decode_mode(code, 2) == s[:2]
    assert decode_mode(0, 0) == ""
    p = SyntheticDDPMPolicy(16, 2, device="cpu", seed=3)
    assert sum(x.numel() for x in p.eps_net.parameters()) == 533762      # DDPM-MLP of scripts/sorting_4/ddpm_benchmark.sh (~0.53 M)
    a = p.predict_batch(torch.randn(9, 16))
    assert a.shape == (9, 2) and torch.isfinite(a).all() and a.abs().max() <= 0.01 + 1e-9
    b = SyntheticBCPolicy(4, 2, device="cpu")
    assert b.predict(np.zeros(4, np.float32)).shape == (1, 2)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle port on the host cores) prints ONE JSON line with the contract's keys; a
    non-zero rank under torchrun prints nothing and exits 0."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(line) == 1
    d = json.loads(line[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_agent_adapter_ddpm_window_and_ema_and_refusals():
    """DiffusionAgent path (ddpm_agent.py:214-274 restated on a stub): a window of past observations per env (deque, cleared by
    reset), EMA parameters swapped in for the forward pass and restored afterwards — batched == N independent single-env agents.
    Agents with per-episode state and no batched path are refused instead of silently sharing one state across envs."""
    from collections import deque
    from d3il_b200.simulation.agent_adapter import predict_batch

    class Ema:
        def __init__(self, params):
            self.shadow = [p.detach().clone() * 0.5 for p in params]
            self.saved = None
            self.swaps = 0

        def store(self, params):
            self.saved = [p.detach().clone() for p in params]

        def copy_to(self, params):
            self.swaps += 1
            for p, s in zip(params, self.shadow):
                p.data.copy_(s)

        def restore(self, params):
            for p, s in zip(params, self.saved):
                p.data.copy_(s)

    class Net(torch.nn.Module):          # deterministic stand-in for the reverse-diffusion sampler: consumes [B, T, obs], returns [B, T, act]
        def __init__(self):
            super().__init__()
            torch.manual_seed(1)
            self.lin = torch.nn.Linear(6, 2)

        def forward(self, state, goal):
            return torch.cumsum(self.lin(state), dim=1)          # depends on the WHOLE window

    class Scaler:
        def scale_input(self, x):
            return (x * 2.0 - 0.1).float()

        def inverse_scale_output(self, y):
            return y * 0.01

    class DiffusionAgent:
        def __init__(self):
            self.model, self.scaler, self.device = Net(), Scaler(), "cpu"
            self.window_size, self.use_ema, self.diffusion_kde = 3, True, False
            self.obs_context = deque(maxlen=self.window_size)
            self.ema_helper = Ema(list(self.model.parameters()))

        def reset(self):
            self.obs_context.clear()

        @torch.no_grad()
        def predict(self, state):                                   # the reference's single-env law
            st = self.scaler.scale_input(torch.from_numpy(state).float().unsqueeze(0))
            self.obs_context.append(st)
            inp = torch.stack(tuple(self.obs_context), dim=1)
            self.ema_helper.store(self.model.parameters()); self.ema_helper.copy_to(self.model.parameters())
            pred = self.model(inp, None)[:, -1, :]
            self.ema_helper.restore(self.model.parameters())
            return self.scaler.inverse_scale_output(pred).numpy()

    n, steps = 5, 6
    rng = np.random.default_rng(0)
    obs_seq = rng.normal(size=(steps, n, 6)).astype(np.float32)
    batched_agent = DiffusionAgent(); batched_agent.reset()
    w_before = [p.detach().clone() for p in batched_agent.model.parameters()]
    singles = [DiffusionAgent() for _ in range(n)]
    for a in singles:
        a.reset()
    for k in range(steps):
        out = predict_batch(batched_agent, torch.from_numpy(obs_seq[k]))
        ref = np.stack([singles[i].predict(obs_seq[k, i])[0] for i in range(n)])
        assert np.allclose(out.numpy(), ref, atol=1e-6), k
    assert batched_agent.ema_helper.swaps == steps
    assert all(torch.equal(a, b) for a, b in zip(w_before, batched_agent.model.parameters()))      # training weights restored
    assert len(batched_agent.obs_context) == 3

    class ACT_Agent:                     # action chunking: per-episode state, no batched law here
        def __init__(self):
            self.model, self.scaler, self.action_counter = Net(), Scaler(), 0

        def predict(self, state):
            return np.zeros((1, 2))

    with pytest.raises(NotImplementedError):
        predict_batch(ACT_Agent(), torch.zeros(3, 6))
