"""``Avoiding_Sim`` — drop-in for ``simulation/avoiding_sim.py:20-144`` on the batched CUDA env."""
from __future__ import annotations

import logging

import numpy as np
import torch

from ..batched_env import BatchedEnv
from .agent_adapter import predict_batch
from .base_sim import BaseSim, _wandb_log
from .metrics import avoiding_entropy

log = logging.getLogger(__name__)


class Avoiding_Sim(BaseSim):
    def __init__(self, seed: int, device: str, render: bool, n_cores: int = 1, n_trajectories: int = 30):
        super().__init__(seed, device, render, n_cores)
        self.n_trajectories = n_trajectories

    @torch.no_grad()
    def eval_agent(self, agent, n_trajectories: int):
        """[n, 10] rows: 9 mode bits + success (``avoiding_sim.py:33-76``), all rollouts in lock-step."""
        dev_index = self._cuda_index()
        dev = torch.device(f"cuda:{dev_index}")
        n = n_trajectories
        env = BatchedEnv("avoiding", n, dev_index)
        torch.manual_seed(self.seed)
        agent.reset()
        obs = env.reset().clone()
        pred_action = env.robot_state().clone()
        fixed_z = pred_action[:, 2:3].clone()
        quat = torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev).repeat(n, 1)
        rows = torch.zeros(n, 10, device=dev)
        active = torch.ones(n, dtype=torch.bool, device=dev)
        des_xy = pred_action[:, :2].clone()
        for _ in range(env.max_steps_per_episode + 1):
            agent_in = torch.cat([des_xy, obs], 1)
            delta = predict_batch(agent, agent_in)
            des_xy = torch.where(active.unsqueeze(1), delta + agent_in[:, :2], des_xy)
            obs_t, _, done, info = env.step(torch.cat([des_xy, fixed_z, quat], 1))
            obs = obs_t.clone()
            just_done = active & done.bool()
            rows = torch.where(just_done.unsqueeze(1), torch.cat([info[:, 1:10], info[:, 0:1]], 1), rows)
            active = active & ~done.bool()
            if not bool(active.any()):
                break
        env.close()
        return rows

    def test_agent(self, agent):
        log.info("Starting trained model evaluation")
        rank, world = self.dist_info()
        lo, hi = self.shard_range(self.n_trajectories, rank, world)
        rows = self.eval_agent(agent, hi - lo)
        rows = self.gather_rows(rows, self.n_trajectories).cpu()
        mode_encoding, successes = rows[:, :9].clone(), rows[:, 9].clone()
        success_rate = torch.mean(successes).item()
        _, entropy = avoiding_entropy(mode_encoding, successes)
        entropy = float(entropy)
        _wandb_log({"score": (success_rate * 0.8 + entropy * 0.2)})
        _wandb_log({"Metrics/successes": success_rate})
        _wandb_log({"Metrics/entropy": entropy})
        print(f"Successrate {success_rate}")
        print(f"entropy {entropy}")
        return successes, entropy
