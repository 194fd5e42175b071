#!/usr/bin/env python
"""bench.py — env steps/sec of the batched rollout hot path (BASELINE.json metric).

Our arm (default): Pushing, 4096 envs per GPU, synthetic random-walk action stream, contexts
``pushing/test_contexts.pkl[i % 60]``, 400-step episodes with auto-reset.  One "step" = one env step of every env in
the batch (= 35 physics ticks each, incl. the IK controller, observation, termination/mode bookkeeping and the masked
reset of finished envs).  ``value`` is measured with everything resident in HBM, ``e2e`` through the host-buffer C ABI
(numpy in / numpy out, H2D + D2H inside the timed region).

``--impl reference`` times the CPU implementation of the same path — the fp64 oracle port (``oracle/``; the real
reference needs mujoco/pinocchio, which cannot be installed here) — one env per worker process on all host cores, the
reference's own sharding scheme (``simulation/pushing_sim.py:114-135``).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKSPACE_LO, WORKSPACE_HI = (0.3, -0.45), (0.8, 0.45)

# BASELINE.json configs -> workloads.  The default (`pushing`) is the configuration the metric is quoted on (configs[1]:
# Pushing, 4096 envs, random actions, 1 GPU); the others are the wider configs, run on request (--workload).
WORKLOADS = {
    "pushing": dict(task="pushing", envs=4096, ctx="pushing_test_contexts", policy=None, n_act=2),
    "avoiding": dict(task="avoiding", envs=4096, ctx=None, policy=None, n_act=2),
    "aligning": dict(task="aligning", envs=4096, ctx="aligning_test_contexts", policy=None, n_act=3),
    "sorting2": dict(task="sorting_2", envs=4096, ctx="sorting_2_contexts", policy=None, n_act=2),
    "sorting4-ddpm": dict(task="sorting_4", envs=8192, ctx="sorting_4_contexts", policy="ddpm", n_act=2),      # configs[2]
    "sorting4": dict(task="sorting_4", envs=8192, ctx="sorting_4_contexts", policy=None, n_act=2),
    "sorting6": dict(task="sorting_6", envs=4096, ctx="sorting_6_contexts", policy=None, n_act=2),
    "inserting": dict(task="inserting", envs=4096, ctx="inserting_contexts", policy=None, n_act=2),
    "stacking": dict(task="stacking", envs=4096, ctx="stacking_test_contexts", policy=None, n_act=7),          # configs[3]: 4096 envs / GPU
}
MIXED7 = ["avoiding", "aligning", "pushing", "sorting2", "sorting4", "sorting6", "stacking"]      # configs[4]: 8192 envs / GPU over the 7 state-based configs
TASK = "pushing"
N_ENVS = WORKLOADS["pushing"]["envs"]


def load_contexts(name="pushing_test_contexts"):
    return np.load(os.path.join(ROOT, "d3il_b200", "data", name + ".npy"))


# ------------------------------------------------------------------------------------------------ CPU (oracle) arm
def _cpu_worker(args):
    """One env on one core: `n_steps` env steps of the workload; returns (env steps done, seconds)."""
    workload, wid, n_steps, seed = args
    try:
        os.sched_setaffinity(0, {wid % os.cpu_count()})
    except Exception:
        pass
    from d3il_b200.scene.blob import load_scene
    from oracle.oracle import OracleEnv          # bench.py's cpu_baseline / reference leg: allowed user of oracle/

    if workload == "mixed7":                    # BASELINE.json configs[4]: the workers cycle through the seven state-based configs
        workload = MIXED7[wid % len(MIXED7)]
    wl = WORKLOADS[workload]
    blob, sc = load_scene(wl["task"])
    ctxs = load_contexts(wl["ctx"]) if wl["ctx"] else None
    env = OracleEnv(blob, sc.header)
    rng = np.random.default_rng(seed)
    n_act, joint_space = wl["n_act"], sc.header["act_dim"] == 8
    lo, hi = np.array([*WORKSPACE_LO, 0.02])[:n_act], np.array([*WORKSPACE_HI, 0.35])[:n_act]

    def fresh(k):
        env.reset(ctxs[(wid + k) % len(ctxs)] if ctxs is not None else None)
        return np.concatenate([env.joint_state()[:7], [0.08]]) if joint_space else np.concatenate([env.robot_state(), [0.0, 1.0, 0.0, 0.0]])
    des = fresh(0)
    t0 = time.perf_counter()
    for k in range(n_steps):
        if joint_space:
            des[:7] += rng.uniform(-0.01, 0.01, 7)
            des[7] = 0.08 if ((k + 37 * wid) // 50) % 2 == 0 else 0.0       # same per-env toggling phase as the GPU arm
        else:
            des[:n_act] = np.clip(des[:n_act] + rng.uniform(-0.01, 0.01, n_act), lo, hi)
        _, _, done, _ = env.step(des)
        if done:
            des = fresh(k)
    return n_steps, time.perf_counter() - t0


def cpu_sample(workload: str, n_workers: int, steps_per_worker: int, pool=None):
    """Bounded sample of the workload on `n_workers` cores; returns (env-steps/s aggregate, wall seconds)."""
    jobs = [(workload, w, steps_per_worker, 1000 + w) for w in range(n_workers)]
    t0 = time.perf_counter()
    if n_workers == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res) / wall, wall


def run_reference(args):
    """--impl reference: the oracle port on all host cores (kind "port"); each step is a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle.oracle as oo
    oo.build()
    cores = os.cpu_count() or 1
    task = "mixed7" if args.workload == "mixed7" else WORKLOADS[args.workload]["task"]
    steps_per_worker = 256 if task in ("pushing", "avoiding", "aligning", "sorting_2") else 64
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_sample(args.workload, cores, 32, pool)
        t0 = time.perf_counter()
        tot = 0
        for _ in range(args.steps):
            v, wall = cpu_sample(args.workload, cores, steps_per_worker, pool)
            tot += steps_per_worker * cores
        dt = time.perf_counter() - t0
    value = tot / dt
    sample = f"{cores} worker processes x {steps_per_worker} env steps of the {args.workload} workload per step (one fp64 oracle env per core)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}-randomwalk (bounded CPU sample: one env per host core)", "task": task},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "native_so": os.path.join(ROOT, "oracle", "libd3il_oracle.so") + " (fp64 CPU oracle, loaded in the forked worker processes; no repo CUDA library on this arm)",
    }))


# Facts taken from committed ncu captures (profiles/): dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum of
# ONE k_env launch at the workload's default size.  None / absent = not captured for that workload.
NCU_FACTS = {
    "pushing": {"dram_bytes": 28.183040e6 + 21.610240e6, "warp_inst_per_env_step": 1992753372 / 4096,
                "source": "profiles/r2_summary.md (k_env<3>, 4096 envs, one launch, ncu --set full: cold caches, so the 12 MB set-point hand-off that stays in L2 in steady state is counted as DRAM traffic)"},
}


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ------------------------------------------------------------------------------------------------ GPU arm
def _alg_bytes_per_env_step(env) -> int:
    """Algorithmic HBM bytes of one env step of one env: persistent fp32 state words in + out, action in,
    obs / reward / done / info out."""
    h = env.scene.header
    n_state = h["nq"] + h["nv"] * 2 + 9 + 9 + 7 + 16 + h.get("nextra", 0) + 7 * 2 + 7 + 21      # + IK reference (7 doubles), desired pose, joint set-point
    return 2 * 4 * n_state + 4 * env.act_dim + 4 * env.obs_dim + 4 + 1 + 4 * env.info_dim


class EnvStream:
    """One task's env batch with its synthetic action stream, episode bookkeeping and auto-reset: everything
    on the device, no host synchronisation, capturable in a CUDA graph (the env step's launch number lives on the device)."""

    def __init__(self, workload: str, n: int, rank: int, local: int, max_iter: int = 0):
        import torch
        from d3il_b200.batched_env import BatchedEnv
        self.torch = torch
        wl = WORKLOADS[workload]
        self.workload, self.task, self.n, self.n_act = workload, wl["task"], n, wl["n_act"]
        self.dev = dev = torch.device(f"cuda:{local}")
        self.env = env = BatchedEnv(self.task, n, local)
        if max_iter:
            env.set_solver(1e-6, max_iter)
        self.ctxs = load_contexts(wl["ctx"]) if wl["ctx"] else None
        self.ctx_ids = (np.arange(n) + rank * n) % (len(self.ctxs) if self.ctxs is not None else 1)
        self.ctx_t = torch.tensor(self.ctxs[self.ctx_ids], dtype=torch.float32, device=dev) if self.ctxs is not None else None
        env.reset(self.ctx_t)
        self.joint_space = env.act_dim == 8
        if self.joint_space:      # Stacking: q_des += U(-0.01, 0.01)^7, gripper command toggled every 50 steps
            start = env.joint_state().clone(); start[:, 7] = 0.08
        else:
            start = torch.cat([env.robot_state().clone(), torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev).repeat(n, 1)], 1)
        self.start, self.des = start.contiguous(), start.clone().contiguous()
        self.lo = torch.tensor([WORKSPACE_LO[0], WORKSPACE_LO[1], 0.02], device=dev)[:self.n_act]
        self.hi = torch.tensor([WORKSPACE_HI[0], WORKSPACE_HI[1], 0.35], device=dev)[:self.n_act]
        self.policy = None
        if wl["policy"] == "ddpm":
            from d3il_b200.simulation.policies import SyntheticDDPMPolicy
            self.policy = SyntheticDDPMPolicy(self.n_act + env.obs_dim, self.n_act, width=256, n_hidden_layers=8, n_timesteps=4, t_dim=8, device=dev, seed=rank)
        self.last_info = torch.zeros(n, env.info_dim, device=dev)     # info row of every env at its latest episode end
        self.episodes = torch.zeros((), dtype=torch.long, device=dev)
        self.fault_steps = torch.zeros((), dtype=torch.long, device=dev)      # env steps with a non-zero status word
        self.bit_counts = torch.zeros(5, dtype=torch.long, device=dev)        # ... per status bit 1, 2, 4, 8, 16
        self.bits = torch.tensor([1, 2, 4, 8, 16], dtype=torch.int32, device=dev)
        self.step_no = torch.zeros((), dtype=torch.long, device=dev)
        self.grip_phase = (torch.arange(n, device=dev) * 37) % 100      # every env toggles its gripper every 50 steps, on its own phase: the cost of a step does not depend on when it is timed
        self.last_obs = env.obs.clone()
        self.ep_len = env.max_steps_per_episode

    def advance(self, force=None, events=None):
        torch, env, n_act = self.torch, self.env, self.n_act
        # synthetic action stream: policy output (config 3) or U(-0.01, 0.01) deltas integrated on the last DESIRED set-point
        if self.policy is not None:
            delta = self.policy.predict_batch(torch.cat([self.des[:, :n_act], self.last_obs], 1))
        else:
            delta = torch.rand(self.n, n_act, device=self.dev) * 0.02 - 0.01
        if self.joint_space:
            self.des[:, :7] += delta
            self.des[:, 7] = torch.where(((self.step_no + self.grip_phase) // 50) % 2 == 0, 0.08, 0.0)
        else:
            self.des[:, :n_act] = torch.minimum(torch.maximum(self.des[:, :n_act] + delta, self.lo), self.hi)
        if events:
            events[0].record()
        obs, rew, done, info = env.step(self.des)
        if events:
            events[1].record()
        self.last_obs.copy_(obs)
        self.step_no.add_(1)
        st = info[:, -1].to(torch.int32)
        self.fault_steps.add_((st != 0).sum())
        self.bit_counts.add_(((st.unsqueeze(1) & self.bits) != 0).sum(0))      # per-bit counts: an OR over envs, not a max
        # episode bookkeeping + auto-reset of finished envs (masked reset kernel; the set-point snaps back to the start pose)
        m = done if force is None else force
        mb = m.bool().unsqueeze(1)
        self.last_info.copy_(torch.where(mb, info, self.last_info))
        self.episodes.add_(m.sum())
        env.reset(self.ctx_t, m)
        self.des.copy_(torch.where(mb, self.start, self.des))

    def result_rows(self):
        """Per-env result rows in the layout the task's ``*_Sim.test_agent`` gathers."""
        torch, info = self.torch, self.last_info
        if self.task == "stacking":
            from d3il_b200.simulation.metrics import stacking_rows
            return stacking_rows(info)                                           # mode_3, mode_1, mode_2, success, success_1, success_2
        if self.task == "avoiding":
            return torch.cat([info[:, 1:10], info[:, 0:1]], 1)                   # 9 mode bits + success
        return torch.stack([info[:, 1], info[:, 0], info[:, 2]], 1)             # mode, success, mean_distance / mode_step


def device_metrics(task: str, rows, n_ctx: int):
    """Behaviour metrics of the gathered result rows, computed ON THE DEVICE (simulation/metrics.py): per-context mode
    histograms over successful rollouts -> entropy (KL for Sorting / Stacking).  Rows are grouped as [context, rollout] by the
    env index modulo the number of contexts."""
    import torch
    from d3il_b200.simulation import metrics as M
    n = rows.shape[0]
    if task == "avoiding":
        probs, ent = M.avoiding_entropy(rows[:, :9], rows[:, 9])
        return {"success": float(rows[:, 9].mean()), "entropy": float(ent)}
    per = n // n_ctx
    if per == 0:
        return {}
    idx = torch.arange(n, device=rows.device)
    grid = rows[idx[: per * n_ctx].reshape(per, n_ctx).t().reshape(-1)].reshape(n_ctx, per, -1)      # env i -> context i % n_ctx
    if task == "stacking":
        from d3il_b200.simulation.stacking_sim import MODE_3, MODE_PROB
        prior = {MODE_3[k]: v for k, v in MODE_PROB.items()}
        _, ent, kl = M.mode_kl(grid[:, :, 0], grid[:, :, 3], prior)
        return {"success": float(grid[:, :, 3].mean()), "success_1_box": float(grid[:, :, 4].mean()), "success_2_boxes": float(grid[:, :, 5].mean()), "entropy_3": ent, "KL_3": kl}
    mode, succ = grid[:, :, 0], grid[:, :, 1]
    if task.startswith("sorting"):
        _, ent, kl = M.mode_kl(mode, succ, None)
        return {"success": float(succ.mean()), "entropy": ent, "KL": kl}
    n_modes = {"pushing": 4, "aligning": 2, "inserting": 6}[task]
    _, ent = M.mode_entropy(mode, succ, n_modes)
    return {"success": float(succ.mean()), "entropy": float(ent), "mean_distance": float(grid[:, :, 2].mean())}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # stdout carries the one JSON line only: NCCL prints its version banner to fd 1 when the communicator is created
        # (NCCL_DEBUG=VERSION in the box environment), so fd 1 points at stderr until the line is printed
        sys.stdout.flush()
        run_gpu.saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    torch.manual_seed(1234 + rank)
    K, W = args.steps, max(args.warmup, 3)
    mixed = args.workload == "mixed7"
    if mixed:      # BASELINE.json configs[4]: 8192 envs per GPU split over the seven state-based configs, one env batch + stream per task
        n_total = args.envs or 8192
        counts = [n_total // len(MIXED7)] * len(MIXED7)
        counts[0] += n_total - sum(counts)
        streams = [EnvStream(w, c, rank, local, args.max_iter) for w, c in zip(MIXED7, counts)]
    else:
        streams = [EnvStream(args.workload, args.envs or WORKLOADS[args.workload]["envs"], rank, local, args.max_iter)]
    n = sum(s.n for s in streams)
    side = [torch.cuda.Stream(device=dev) for _ in streams] if mixed else None

    def step_all(force_k=None, events=None):
        if not mixed:
            s = streams[0]
            s.advance(None if force_k is None else (torch.arange(s.n, device=dev) % s.ep_len == force_k).to(torch.uint8), events)
            return
        cur = torch.cuda.current_stream(dev)
        if events:
            events[0].record()
        for s, st in zip(streams, side):                    # the per-task step kernels of one env step overlap on the device
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                s.advance(None if force_k is None else (torch.arange(s.n, device=dev) % s.ep_len == force_k).to(torch.uint8))
        for st in side:
            cur.wait_stream(st)
        if events:
            events[1].record()

    # pre-roll (untimed, not part of W): spread the envs uniformly over the episode so the timed region sees the steady-state
    # mix of episode phases (fresh resets, free motion, contact) instead of n synchronised envs; env i is force-reset once at
    # pre-roll step i % ep_len.
    preroll = 0 if args.no_preroll else min(max(s.ep_len for s in streams), 400 if mixed else 10 ** 9)
    for k in range(preroll):
        step_all(force_k=k)
    for k in range(W):
        step_all()
    torch.cuda.synchronize()
    launches_per_step = 0
    l0 = sum(s.env.kernel_launches for s in streams)
    step_all()
    torch.cuda.synchronize()
    launches_per_step = sum(s.env.kernel_launches for s in streams) - l0      # k_sched + k_ik + k_env (+ k_reset) per task
    graph = None
    if not args.no_graph:
        # one env step of the whole job ([action stream / policy -> step -> bookkeeping -> masked reset] of every task) as ONE graph
        # launch: no per-kernel launch gaps, no Python between the ~25 small kernels of the step
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            for _ in range(2):
                step_all()
        torch.cuda.current_stream(dev).wait_stream(cap)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            step_all()
        torch.cuda.synchronize()
    run_step = graph.replay if graph is not None else step_all
    for s in streams:
        s.fault_steps.zero_(); s.bit_counts.zero_(); s.episodes.zero_()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("D3IL_NO_CLOCKS") else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    seg = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)] if os.environ.get("D3IL_SEG") and graph is None else None
    ev0.record()
    for k in range(K):
        if seg:
            step_all(events=seg[k])
        else:
            run_step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if seg:      # debug: per-step device time of env.step and of the gaps between consecutive env.step calls
        st = np.array([a.elapsed_time(b) for a, b in seg]); gap = np.array([seg[k][1].elapsed_time(seg[k + 1][0]) for k in range(K - 1)])
        print("[seg] env.step ms by step: " + " ".join(f"{x:.1f}" for x in st), file=sys.stderr)
        print(f"[seg] env.step mean {st.mean():.3f} p50 {np.median(st):.3f} p90 {np.percentile(st, 90):.3f} max {st.max():.3f}; between steps mean {gap.mean():.3f} p50 {np.median(gap):.3f} p90 {np.percentile(gap, 90):.3f} max {gap.max():.3f}", file=sys.stderr)
    n_faults = int(sum(int(s.fault_steps.item()) for s in streams))
    bit_counts = sum(s.bit_counts for s in streams).tolist()
    episodes = int(sum(int(s.episodes.item()) for s in streams))
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    # the path's only exchange: ONE gather of the per-env episode result rows at rollout end (replaces the shared-memory result
    # tensors of simulation/pushing_sim.py:97-99; Stacking: the six columns of stacking_sim.py:122-141), then the behaviour
    # metrics on the gathered rows, on the device
    metrics = {}
    for s in streams:
        rows = s.result_rows().contiguous()
        if world > 1:
            parts = [torch.zeros_like(rows) for _ in range(world)]
            dist.all_gather(parts, rows)
            rows = torch.cat(parts, 0)
        metrics[s.task] = dict(device_metrics(s.task, rows, len(s.ctxs) if s.ctxs is not None else 1), rows=list(rows.shape))
    clocks = sampler.stop() if sampler else None
    value = world * n * K / (ms * 1e-3)
    if os.environ.get("D3IL_TL"):      # timing build only (profiles/build_variant.sh ... -DD3IL_PHASE_TIMING): where the LAST timed step spent its time
        import ctypes as C
        from d3il_b200 import lib as _lib
        buf = (C.c_ulonglong * (4 * 4096))(); _lib.lib().d3il_debug_timeline(buf)
        a = np.array(buf, dtype=np.uint64).reshape(4096, 4).astype(np.int64)
        t0 = a[a[:, 3] == 3][0, 0]
        for kind, name in ((3, "k_sched"), (1, "k_ik"), (2, "k_env"), (5, "k_reset (CTAs without work)"), (4, "k_reset (CTAs that reset an env)")):
            r = a[a[:, 3] == kind]
            if len(r):
                print(f"[timeline] {name:34s} blocks {len(r):4d} start {(r[:, 0].min() - t0) / 1e6:7.3f} .. {(r[:, 0].max() - t0) / 1e6:7.3f} ms, end {(r[:, 1].min() - t0) / 1e6:7.3f} .. {(r[:, 1].max() - t0) / 1e6:7.3f} ms", file=sys.stderr)
        rows = np.nonzero(a[:, 3] == 2)[0]
        late = rows[np.argsort(-a[rows, 1])[:8]]
        print("[timeline] last k_env CTAs to end (block, start ms, end ms): " + ", ".join(f"({b}, {(a[b, 0] - t0) / 1e6:.2f}, {(a[b, 1] - t0) / 1e6:.2f})" for b in late), file=sys.stderr)
        prev = a[a[:, 3] == 6]
        print(f"[timeline] mean period {ms / K:.3f} ms per step; the previous step's k_sched started {(t0 - prev[0, 0]) / 1e6 if len(prev) else float('nan'):.3f} ms before this one's", file=sys.stderr)

    # ---- device time of the step's own kernels (k_sched + k_ik + k_env of every task): CUDA events recorded on the launching
    # stream around env.step, NOT synchronised step by step (the overlap of k_ik and k_env stays intact), read back at the end
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(32)]
    for e in evs:
        step_all(events=e)
    torch.cuda.synchronize()
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    if os.environ.get("D3IL_SEG"):
        print("[seg] kernel_ms pass by step: " + " ".join(f"{a.elapsed_time(b):.1f}" for a, b in evs), file=sys.stderr)
    alg_bytes_launch = sum(_alg_bytes_per_env_step(s.env) * s.n for s in streams)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes_launch / (kernel_ms * 1e-3) / 1e9
    key = "mixed7" if mixed else streams[0].task
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_FACTS.get(key, {}).get("dram_bytes"), "traffic_source": NCU_FACTS.get(key, {}).get("source"),
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                "kernel": "k_env (+ overlapped k_ik, k_sched)" + (" x 7 tasks on 7 streams" if mixed else ""), "kernel_ms": kernel_ms,
                "alg_bytes_per_launch": alg_bytes_launch,
                "note": "algorithmic bytes = state row in + out, action, obs, reward, done, info; the fused n_substeps-tick kernel keeps the state in shared memory and is issue / latency bound (second object)"}
    inst = NCU_FACTS.get(key, {}).get("warp_inst_per_env_step")
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    roofline_issue = None
    if inst:
        ach = inst * n / (kernel_ms * 1e-3)
        pk = 148 * 4 * sm_mhz * 1e6
        roofline_issue = {"bound": "issue", "achieved": ach, "peak": pk, "unit": "warp-inst/s", "frac": ach / pk, "warp_inst_per_env_step": inst,
                          "source": NCU_FACTS[key]["source"], "note": "148 SMs x 4 schedulers x 1 warp instruction per cycle at the SM clock sampled during the timed region"}

    # ---- e2e: same workload through the host-buffer C ABI (numpy in/out, H2D + D2H every step)
    e2e_steps = max(8, min(K, 64))
    host = []
    for s in streams:
        host.append(dict(des=s.des.cpu().numpy().astype(np.float32), start=s.start.cpu().numpy().astype(np.float32),
                         ctx=s.ctxs[s.ctx_ids].astype(np.float32) if s.ctxs is not None else None, obs=s.last_obs.cpu().numpy(),
                         lo=s.lo.cpu().numpy(), hi=s.hi.cpu().numpy(), rng=np.random.default_rng(7 + rank)))
    h2d = d2h = 0

    def host_step(k, count):
        nonlocal h2d, d2h
        for s, hb in zip(streams, host):
            na = s.n_act
            if s.policy is not None:      # policy on the GPU: observations go up, deltas come down, every step
                pin = torch.from_numpy(np.concatenate([hb["des"][:, :na], hb["obs"]], 1)).to(dev)
                delta = s.policy.predict_batch(pin).cpu().numpy()
                if count:
                    h2d += pin.numel() * 4; d2h += delta.nbytes
            else:
                delta = hb["rng"].uniform(-0.01, 0.01, (s.n, na)).astype(np.float32)
            if s.joint_space:
                hb["des"][:, :7] += delta
                hb["des"][:, 7] = np.where(((k + (np.arange(s.n) * 37) % 100) // 50) % 2 == 0, 0.08, 0.0)
            else:
                hb["des"][:, :na] = np.clip(hb["des"][:, :na] + delta, hb["lo"], hb["hi"])
            hb["obs"], rew_h, done_h, info_h = s.env.step_host(hb["des"])
            if count:
                h2d += hb["des"].nbytes
                d2h += hb["obs"].nbytes + rew_h.nbytes + done_h.nbytes + info_h.nbytes
            if done_h.any():
                s.env.reset_host(hb["ctx"], done_h)
                if count:
                    h2d += (hb["ctx"].nbytes if hb["ctx"] is not None else 0) + done_h.nbytes
                hb["des"][done_h.astype(bool)] = hb["start"][done_h.astype(bool)]

    for k in range(3):
        host_step(k, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        host_step(k, True)
    dt = time.perf_counter() - t0
    if world > 1:          # every rank drives its own GPU through the host API at the same time; slowest rank sets the rate
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": world * n * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": world * (h2d // e2e_steps), "d2h_bytes_per_step": world * (d2h // e2e_steps),
           "steps": e2e_steps, "n_gpus": world}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it: the oracle port on this box's host cores, bounded sample
    import oracle.oracle as oo
    oo.build()
    cores = os.cpu_count() or 1
    cpu_wl = args.workload
    with mp.get_context("fork").Pool(cores) as pool:
        probe_val, _ = cpu_sample(cpu_wl, cores, 32, pool)
        # bounded sample: ~20 s of CPU work in total (all cores for ~1.3 s each), sized from the probe's rate
        spw = int(min(4096, max(64, 1.3 * probe_val / cores)))
        cpu_val, cpu_wall = cpu_sample(cpu_wl, cores, spw, pool)
    one_val, one_wall = cpu_sample(cpu_wl, 1, 2 * spw)
    cpu_baseline = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{cores} processes x {spw} env steps of the {args.workload} workload (one fp64 oracle env per core, {cpu_wall:.1f} s); single core: {one_val:.0f} env-steps/s over {2 * spw} steps",
                    "single_core_value": one_val}

    s0 = streams[0]
    config = {"workload": (f"mixed7-{n}env-per-gpu-randomwalk" if mixed else
                           f"{args.workload}-{n}env-per-gpu-" + ("ddpm-mlp-in-loop" if s0.policy is not None else "randomwalk")),
              "task": "mixed7" if mixed else s0.task, "envs_per_gpu": n, "auto_reset": True, "preroll_steps": preroll,
              "cuda_graph": graph is not None,
              "l2_note": "every env step rewrites the whole state working set (no cross-step reuse of inputs); timing is launch-to-launch on one stream"}
    if mixed:
        config.update(tasks=[s.task for s in streams], envs_per_task=[s.n for s in streams])
    else:
        config.update(n_substeps=s0.env.n_substeps, episode_len=s0.ep_len)
    if getattr(run_gpu, "saved_stdout", None) is not None:
        sys.stdout.flush()
        os.dup2(run_gpu.saved_stdout, 1)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * K), "gpu_launches_per_step": int(launches_per_step),
        "env_step_faults": n_faults, "env_step_fault_bit_counts": dict(zip(["1_M_not_PD", "2_overflow", "4_H_not_PD", "8_iter_cap", "16_bad_action"], bit_counts)),
        "episodes_finished": episodes, "behaviour_metrics": metrics,
        "roofline": roofline, "roofline_issue": roofline_issue, "cpu_baseline": cpu_baseline,
        "target": {"env_steps_per_sec": 1.0e6, "met": bool(value >= 1.0e6)},
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's BASELINE.json size)")
    ap.add_argument("--workload", default="pushing", choices=sorted(WORKLOADS) + ["mixed7"], help="default = the configuration the metric is quoted on")
    ap.add_argument("--no-preroll", action="store_true", help="skip the 400-step episode-phase pre-roll (debug)")
    ap.add_argument("--max-iter", type=int, default=0, help="override the Newton iteration cap (diagnostics; default: the library's)")
    ap.add_argument("--no-graph", action="store_true", help="launch the env step kernel by kernel instead of replaying one CUDA graph per env step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
