"""``Avoiding_Sim`` — drop-in for ``simulation/avoiding_sim.py:20-144`` on the batched CUDA env."""
from __future__ import annotations

import logging

import numpy as np
import torch

from .base_sim import BaseSim, _wandb_log, cartesian_rollout
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
        info = cartesian_rollout(agent, "avoiding", None, n_trajectories, dev_index, self.seed, 2)
        return torch.cat([info[:, 1:10], info[:, 0:1]], 1)

    def test_agent(self, agent):
        log.info("Starting trained model evaluation")
        rank, world = self.dist_info()
        lo, hi = self.shard_range(self.n_trajectories, rank, world)
        rows = self.eval_agent(agent, hi - lo)
        rows = self.gather_rows(rows, self.n_trajectories)                  # metrics stay on the device; only scalars come back
        mode_encoding, successes = rows[:, :9].clone(), rows[:, 9].clone()
        success_rate = torch.mean(successes).item()
        _, entropy = avoiding_entropy(mode_encoding, successes)
        entropy = float(entropy)
        _wandb_log({"score": (success_rate * 0.8 + entropy * 0.2)})
        _wandb_log({"Metrics/successes": success_rate})
        _wandb_log({"Metrics/entropy": entropy})
        print(f"Successrate {success_rate}")
        print(f"entropy {entropy}")
        return successes.cpu(), entropy
