"""``Pushing_Sim`` — drop-in for ``simulation/pushing_sim.py:28-178`` on the batched CUDA env.

Same constructor kwargs as ``configs/pushing_config.yaml`` ``simulation:`` block, same ``test_agent(agent)`` return
``(successes, mode_encoding, mean_distance)`` as ``[n_contexts, n_trajectories_per_context]`` float tensors, same wandb
keys.  Every (context, rollout) pair is one env instance; the policy is evaluated on the whole batch per env step.
"""
from __future__ import annotations

import logging
import os

import numpy as np
import torch

from .base_sim import BaseSim, _wandb_log, cartesian_rollout
from .metrics import mode_entropy

log = logging.getLogger(__name__)
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data")


def load_test_contexts() -> np.ndarray:
    """[60, 2, 7] xyz + quat per box from ``environments/dataset/data/pushing/test_contexts.pkl`` (committed .npy)."""
    return np.load(os.path.join(_DATA, "pushing_test_contexts.npy"))


class Pushing_Sim(BaseSim):
    def __init__(self, seed: int, device: str, render: bool, n_cores: int = 1, n_contexts: int = 30, n_trajectories_per_context: int = 1):
        super().__init__(seed, device, render, n_cores)
        self.n_contexts = n_contexts
        self.n_trajectories_per_context = n_trajectories_per_context

    @torch.no_grad()
    def eval_agent(self, agent, items: np.ndarray):
        """Roll out the (context, rollout) pairs in ``items`` ([n, 2] ints) in lock-step; returns [n, 3] result rows
        (mode, success, mean_distance) — the batched ``eval_agent`` of ``pushing_sim.py:43-85`` (agent input = [last desired
        xy || env obs] :74, delta integrated on the last DESIRED xy :77, z frozen at the reset tcp height :69)."""
        dev_index = self._cuda_index()
        ctx = torch.tensor(load_test_contexts()[items[:, 0]], dtype=torch.float32, device=f"cuda:{dev_index}").reshape(len(items), -1)
        info = cartesian_rollout(agent, "pushing", ctx, len(items), dev_index, self.seed, 2)
        return torch.stack([info[:, 1], info[:, 0], info[:, 2]], 1)                     # :83-85

    def test_agent(self, agent):
        log.info("Starting trained model evaluation")
        n_items = self.n_contexts * self.n_trajectories_per_context
        items = np.stack(np.meshgrid(np.arange(self.n_contexts), np.arange(self.n_trajectories_per_context), indexing="ij"), -1).reshape(-1, 2)
        rank, world = self.dist_info()
        lo, hi = self.shard_range(n_items, rank, world)
        rows = self.eval_agent(agent, items[lo:hi])
        rows = self.gather_rows(rows, n_items)          # result rows and the metrics below stay on the device; only scalars and the returned tensors come back
        shape = (self.n_contexts, self.n_trajectories_per_context)
        mode_encoding, successes, mean_distance = (rows[:, k].reshape(shape).clone() for k in range(3))

        n_modes = 4
        success_rate = torch.mean(successes).item()
        mode_probs, entropy = mode_entropy(mode_encoding, successes, n_modes)
        print(f"p(m|c) {mode_probs}")
        _wandb_log({"score": 0.5 * (success_rate + entropy)})
        _wandb_log({"Metrics/successes": success_rate})
        _wandb_log({"Metrics/entropy": entropy})
        _wandb_log({"Metrics/distance": mean_distance.mean().item()})
        print(f"Mean Distance {mean_distance.mean().item()}")
        print(f"Successrate {success_rate}")
        print(f"entropy {entropy}")
        return successes.cpu(), mode_encoding.cpu(), mean_distance.cpu()
