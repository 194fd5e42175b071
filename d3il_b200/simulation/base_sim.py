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


@torch.no_grad()
def cartesian_rollout(agent, task: str, contexts: torch.Tensor | None, n: int, dev_index: int, seed: int, n_act: int, max_steps: int | None = None):
    """The env loop shared by the Cartesian-action sims (``pushing_sim.py:55-85``, ``sorting_sim.py:99-136``,
    ``aligning_sim.py:105-120``, ``avoiding_sim.py:45-76``), all (context, rollout) pairs in lock-step:
    agent input = [last DESIRED tcp xy(z) || env obs], policy output = delta integrated on the last desired pose,
    z frozen at the reset tcp height when the action is 2-D.  Returns the ``info`` row of every env at ITS final step."""
    from ..batched_env import BatchedEnv
    from .agent_adapter import predict_batch

    dev = torch.device(f"cuda:{dev_index}")
    env = BatchedEnv(task, n, dev_index)
    torch.manual_seed(seed)
    agent.reset()
    obs = env.reset(contexts).clone()
    tcp = env.robot_state().clone()
    des = tcp[:, :n_act].clone()
    tail = torch.cat([tcp[:, n_act:], torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev).repeat(n, 1)], 1)
    rows = torch.zeros(n, env.info_dim, device=dev)
    active = torch.ones(n, dtype=torch.bool, device=dev)
    cap = env.max_steps_per_episode if max_steps is None else min(int(max_steps), env.max_steps_per_episode)
    for k in range(cap + 1):
        agent_in = torch.cat([des, obs], 1)
        delta = predict_batch(agent, agent_in)
        des = torch.where(active.unsqueeze(1), delta + agent_in[:, :n_act], des)
        obs_t, _, done, info = env.step(torch.cat([des, tail], 1))
        obs = obs_t.clone()
        done = done.bool() | (k >= cap - 1)             # a shorter episode cap than the compiled one (config max_steps_per_episode)
        rows = torch.where((active & done).unsqueeze(1), info, rows)
        active = active & ~done
        if not bool(active.any()):
            break
    env.close()
    return rows
