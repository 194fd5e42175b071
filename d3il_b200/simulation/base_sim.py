"""``BaseSim`` — same constructor and abstract ``test_agent(agent)`` as the reference (``simulation/base_sim.py:8-30``).

What changes underneath: instead of one MuJoCo env per spawned CPU process (``pushing_sim.py:114-135``) every
(context, rollout) pair is one env instance of a single ``BatchedEnv`` on this rank's GPU; with ``torch.distributed``
initialised, env instances are sharded by contiguous (context, rollout) ranges across ranks and the per-env result
rows are gathered once at rollout end (the replacement of the ``share_memory_()`` result tensors).
"""
from __future__ import annotations

import abc
import logging
import os

import numpy as np
import torch

log = logging.getLogger(__name__)


def _wandb_log(d):
    try:
        import wandb
        if wandb.run is not None:
            wandb.log(d)
    except Exception:
        pass


class BaseSim(abc.ABC):
    def __init__(self, seed: int, device: str, render: bool = True, n_cores: int = 1, if_vision: bool = False):
        self.seed = seed
        self.device = device
        self.render = render
        self.n_cores = n_cores          # kept for config compatibility; parallelism is the env batch, not CPU cores
        self.if_vision = if_vision
        self.working_dir = os.getcwd()
        self.env_name = "BaseEnvironment"
        if if_vision:
            raise NotImplementedError("vision observations need a rasteriser and are outside the batched state-based path")

    @abc.abstractmethod
    def test_agent(self, agent):
        pass

    # ---- sharding helpers shared by the task sims
    @staticmethod
    def dist_info():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    @staticmethod
    def shard_range(n_items: int, rank: int, world: int):
        """Contiguous item range of this rank (the reference's ``contexts[i*workload:(i+1)*workload]`` split), balanced:
        the first ``n_items % world`` ranks get one item more, so no rank is empty while ``n_items >= world``.  A rank whose
        range IS empty (fewer items than ranks) skips the rollout but still takes part in ``gather_rows``."""
        base, extra = divmod(n_items, world)
        lo = rank * base + min(rank, extra)
        return lo, lo + base + (1 if rank < extra else 0)

    @staticmethod
    def gather_rows(local_rows: torch.Tensor, n_items: int) -> torch.Tensor:
        """All-gather per-env result rows ([n_local, k]) into [n_items, k] on every rank (one collective per rollout)."""
        import torch.distributed as dist
        rank, world = BaseSim.dist_info()
        if world == 1:
            return local_rows
        per = (n_items + world - 1) // world
        pad = torch.zeros(per, local_rows.shape[1], dtype=local_rows.dtype, device=local_rows.device)
        pad[: local_rows.shape[0]] = local_rows
        out = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(out, pad)
        counts = [BaseSim.shard_range(n_items, r, world) for r in range(world)]
        return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, counts)], 0)

    def _cuda_index(self) -> int:
        d = torch.device(self.device) if not isinstance(self.device, torch.device) else self.device
        if d.type != "cuda":
            raise RuntimeError("the batched simulation runs on CUDA devices only")
        return d.index if d.index is not None else torch.cuda.current_device()


STATUS_BITS = {1: "mass matrix not positive definite / NaN state", 2: "contact or constraint-row budget exceeded (contacts dropped)",
               4: "Newton Hessian not positive definite", 8: "Newton iteration cap reached", 16: "unusable action (set-point held)"}


def report_faults(status: torch.Tensor, what: str) -> int:
    """``status``: per-env OR of the env's fault word (last ``info`` column) over a rollout.  Logs which bits were raised by how
    many envs; metrics computed from such rollouts are physically suspect and the caller should know.  Returns the count."""
    n_bad = int((status != 0).sum())
    if n_bad:
        bits = {name: int(((status & b) != 0).sum()) for b, name in STATUS_BITS.items() if int(((status & b) != 0).sum())}
        log.warning("%s: %d of %d env instances raised fault bits during the rollout: %s", what, n_bad, status.numel(), bits)
    return n_bad


@torch.no_grad()
def episode_loop(env, policy_step, cap: int, sync_every: int = 8, graph: bool = False):
    """The lock-step episode loop shared by every ``*_Sim``: ``policy_step(active) -> action`` computes the batch action from the
    caller's own state (last desired pose, observation; updated IN PLACE), the env steps, and each env's ``info`` row is latched
    at ITS final step.  The host looks at the device only every ``sync_every`` env steps (``active.any()``), not every step; the
    fault words are OR-ed over the steps of each env.

    ``graph=True`` (policies whose batched forward is pure tensor code, ``agent_adapter.graph_safe``): after two eager steps
    the body [policy -> env step -> bookkeeping] is captured once in a CUDA graph and replayed — one launch per env step
    instead of ~25 kernel launches plus the Python between them (the env step's launch number lives on the device, so nothing in
    the graph changes from step to step).  Returns (info rows [n, info_dim], status [n] int32)."""
    n, dev = env.n_envs, env.device
    rows = torch.zeros(n, env.info_dim, device=dev)
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    active = torch.ones(n, dtype=torch.bool, device=dev)
    step_no = torch.zeros((), dtype=torch.long, device=dev)

    def body():
        action = policy_step(active)
        obs, _, done, info = env.step(action)
        step_no.add_(1)
        fin = active & (done.bool() | (step_no >= cap))         # a shorter episode cap than the compiled one (config max_steps_per_episode)
        rows.copy_(torch.where(fin.unsqueeze(1), info, rows))
        status.bitwise_or_(torch.where(active, info[:, -1].to(torch.int32), torch.zeros_like(status)))
        active.logical_and_(~fin)

    run, k0 = body, 0
    if graph and cap > 4:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body(); body()                                       # real env steps 1 and 2, on the capture stream
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            body()                                               # captured, not executed
        run, k0 = g.replay, 2
    for k in range(k0, cap + 1):
        run()
        if k % sync_every == sync_every - 1 and not bool(active.any()):
            break
    return rows, status


@torch.no_grad()
def cartesian_rollout(agent, task: str, contexts: torch.Tensor | None, n: int, dev_index: int, seed: int, n_act: int, max_steps: int | None = None,
                      use_graph: bool | None = None):
    """The env loop shared by the Cartesian-action sims (``pushing_sim.py:55-85``, ``sorting_sim.py:99-136``,
    ``aligning_sim.py:105-120``, ``avoiding_sim.py:45-76``), all (context, rollout) pairs in lock-step:
    agent input = [last DESIRED tcp xy(z) || env obs], policy output = delta integrated on the last desired pose,
    z frozen at the reset tcp height when the action is 2-D.  Returns the ``info`` row of every env at ITS final step."""
    from ..batched_env import BatchedEnv
    from .agent_adapter import graph_safe, predict_batch

    dev = torch.device(f"cuda:{dev_index}")
    if n == 0:                                          # an empty shard (fewer items than ranks) still takes part in the gather
        from ..scene.blob import load_scene
        return torch.zeros(0, load_scene(task)[1].header["info_dim"], device=dev)
    env = BatchedEnv(task, n, dev_index)
    torch.manual_seed(seed)
    agent.reset()
    env.reset(contexts)
    tcp = env.robot_state().clone()
    action = torch.cat([tcp, torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev).repeat(n, 1)], 1).contiguous()      # [des xy(z) | frozen z | quat]
    des = action[:, :n_act]

    def policy_step(active):
        agent_in = torch.cat([des, env.obs], 1)
        delta = predict_batch(agent, agent_in)
        des.copy_(torch.where(active.unsqueeze(1), delta + agent_in[:, :n_act], des))
        return action

    cap = env.max_steps_per_episode if max_steps is None else min(int(max_steps), env.max_steps_per_episode)
    rows, status = episode_loop(env, policy_step, cap, graph=graph_safe(agent) if use_graph is None else use_graph)
    report_faults(status, task)
    env.close()
    return rows
