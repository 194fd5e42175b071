"""``Sorting_Sim`` — drop-in for ``simulation/sorting_sim.py:20-221`` on the batched CUDA env.

Constructor kwargs as in ``configs/sorting_{2,4,6}_config.yaml``.  The reference loads ``<k>_test_contexts.pkl`` and
``<k>_mode_prob.pkl`` from ``environments/dataset/data/sorting/`` — neither file is shipped: contexts
default to the committed draws from ``BlockContextManager``'s boxes (``d3il_b200/data/sorting_<k>_contexts.npy``), and
the mode prior can be passed as ``mode_prob`` ({packed mode: probability}); without it the prior is uniform over the
modes observed in the rollouts.
"""
from __future__ import annotations

import logging
import os

import numpy as np
import torch

from .base_sim import BaseSim, _wandb_log, cartesian_rollout
from .metrics import mode_kl

log = logging.getLogger(__name__)
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data")


class Sorting_Sim(BaseSim):
    def __init__(self, seed: int, device: str, render: bool, n_cores: int = 1, n_contexts: int = 30, n_trajectories_per_context: int = 1,
                 num_box: int = 2, if_vision: bool = False, max_steps_per_episode: int = 500, test_contexts: np.ndarray | None = None,
                 mode_prob: dict | None = None):
        super().__init__(seed, device, render, n_cores, if_vision)
        self.n_contexts = n_contexts
        self.n_trajectories_per_context = n_trajectories_per_context
        self.max_steps_per_episode = max_steps_per_episode
        self.num_box = num_box
        self.test_contexts = np.load(os.path.join(_DATA, f"sorting_{num_box}_contexts.npy")) if test_contexts is None else np.asarray(test_contexts)
        self.modes = mode_prob

    def eval_agent(self, agent, items: np.ndarray):
        """[n, 2] rows (mode, success) of the (context, rollout) pairs in ``items`` (``sorting_sim.py:59-136``)."""
        dev_index = self._cuda_index()
        ctx = torch.tensor(self.test_contexts[items[:, 0]], dtype=torch.float32, device=f"cuda:{dev_index}")
        info = cartesian_rollout(agent, f"sorting_{self.num_box}", ctx, len(items), dev_index, self.seed, 2, self.max_steps_per_episode)
        return torch.stack([info[:, 1], info[:, 0]], 1)

    def test_agent(self, agent):
        log.info("Starting trained model evaluation")
        n_items = self.n_contexts * self.n_trajectories_per_context
        items = np.stack(np.meshgrid(np.arange(self.n_contexts), np.arange(self.n_trajectories_per_context), indexing="ij"), -1).reshape(-1, 2)
        rank, world = self.dist_info()
        lo, hi = self.shard_range(n_items, rank, world)
        rows = self.gather_rows(self.eval_agent(agent, items[lo:hi]), n_items)          # result rows and the metrics below stay on the device; only scalars and the returned tensors come back
        shape = (self.n_contexts, self.n_trajectories_per_context)
        mode_encoding, successes = rows[:, 0].reshape(shape).clone(), rows[:, 1].reshape(shape).clone()
        success_rate = torch.mean(successes).item()
        mode_probs, entropy, KL = mode_kl(mode_encoding, successes, self.modes)
        print(f"p(m|c) {mode_probs}")
        _wandb_log({"score": (success_rate - KL)})
        _wandb_log({"Metrics/successes": success_rate})
        _wandb_log({"Metrics/KL": KL})
        _wandb_log({"Metrics/entropy": entropy})
        print(f"Successrate {success_rate}")
        print(f"entropy {entropy}")
        print(f"KL {KL}")
        return success_rate, mode_encoding.cpu()
