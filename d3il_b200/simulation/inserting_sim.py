"""``Inserting_Sim`` — rollout harness for ``Gate_Insertion_Env`` (``envs/gym_inserting_env/.../gate_insertion.py``).

The reference ships the env but no ``simulation/`` wrapper or config for it; this class follows the shape of
``Pushing_Sim`` (2-D Cartesian action, agent input = [desired xy || env obs(11)]) so the same agents can be rolled out.
Contexts are committed draws from ``BlockContextManager``'s three boxes (``d3il_b200/data/inserting_contexts.npy``).
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


class Inserting_Sim(BaseSim):
    def __init__(self, seed: int, device: str, render: bool, n_cores: int = 1, n_contexts: int = 30, n_trajectories_per_context: int = 1,
                 max_steps_per_episode: int = 2000):
        super().__init__(seed, device, render, n_cores)
        self.n_contexts = n_contexts
        self.n_trajectories_per_context = n_trajectories_per_context
        self.max_steps_per_episode = max_steps_per_episode
        self.test_contexts = np.load(os.path.join(_DATA, "inserting_contexts.npy"))

    def eval_agent(self, agent, items: np.ndarray):
        """[n, 4] rows (mode id 1..6 or 0, success, mean_distance, boxes inserted)."""
        dev_index = self._cuda_index()
        ctx = torch.tensor(self.test_contexts[items[:, 0]], dtype=torch.float32, device=f"cuda:{dev_index}")
        info = cartesian_rollout(agent, "inserting", ctx, len(items), dev_index, self.seed, 2, self.max_steps_per_episode)
        return torch.stack([info[:, 1], info[:, 0], info[:, 2], info[:, 3]], 1)

    def test_agent(self, agent):
        log.info("Starting trained model evaluation")
        n_items = self.n_contexts * self.n_trajectories_per_context
        items = np.stack(np.meshgrid(np.arange(self.n_contexts), np.arange(self.n_trajectories_per_context), indexing="ij"), -1).reshape(-1, 2)
        rank, world = self.dist_info()
        lo, hi = self.shard_range(n_items, rank, world)
        rows = self.gather_rows(self.eval_agent(agent, items[lo:hi]), n_items)          # result rows and the metrics below stay on the device; only scalars and the returned tensors come back
        shape = (self.n_contexts, self.n_trajectories_per_context)
        mode_encoding, successes, mean_distance, n_boxes = (rows[:, k].reshape(shape).clone() for k in range(4))
        success_rate = torch.mean(successes).item()
        _, entropy = mode_entropy(mode_encoding - 1, successes, 6)          # mode ids 1..6 = the six insertion orders
        _wandb_log({"score": 0.5 * (success_rate + float(entropy))})
        _wandb_log({"Metrics/successes": success_rate})
        _wandb_log({"Metrics/entropy": float(entropy)})
        _wandb_log({"Metrics/distance": mean_distance.mean().item()})
        print(f"Successrate {success_rate}")
        print(f"entropy {float(entropy)}")
        return successes.cpu(), mode_encoding.cpu(), mean_distance.cpu()
