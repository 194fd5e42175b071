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

from ..batched_env import BatchedEnv
from .agent_adapter import predict_batch
from .base_sim import BaseSim, _wandb_log
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
        (mode, success, mean_distance) — the batched ``eval_agent`` of ``pushing_sim.py:43-85``."""
        dev_index = self._cuda_index()
        dev = torch.device(f"cuda:{dev_index}")
        n = len(items)
        test_contexts = load_test_contexts()
        env = BatchedEnv("pushing", n, dev_index)
        torch.manual_seed(self.seed)
        agent.reset()
        ctx = torch.tensor(test_contexts[items[:, 0]], dtype=torch.float32, device=dev)
        obs = env.reset(ctx).clone()                                   # :63
        pred_action = env.robot_state().clone()                        # :68  tcp xyz
        fixed_z = pred_action[:, 2:3].clone()                          # :69
        quat = torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev).repeat(n, 1)
        rows = torch.zeros(n, 3, device=dev)
        active = torch.ones(n, dtype=torch.bool, device=dev)
        des_xy = pred_action[:, :2].clone()
        for _ in range(env.max_steps_per_episode + 1):
            agent_in = torch.cat([des_xy, obs], 1)                     # :74  [des_xy, env_obs]
            delta = predict_batch(agent, agent_in)                     # :76
            des_xy = torch.where(active.unsqueeze(1), delta + agent_in[:, :2], des_xy)   # :77 integrate on the last DESIRED xy
            action = torch.cat([des_xy, fixed_z, quat], 1)             # :79
            obs_t, _, done, info = env.step(action)                    # :81
            obs = obs_t.clone()
            just_done = active & done.bool()
            rows = torch.where(just_done.unsqueeze(1), torch.stack([info[:, 1], info[:, 0], info[:, 2]], 1), rows)   # :83-85
            active = active & ~done.bool()
            if not bool(active.any()):
                break
        env.close()
        return rows

    def test_agent(self, agent):
        log.info("Starting trained model evaluation")
        n_items = self.n_contexts * self.n_trajectories_per_context
        items = np.stack(np.meshgrid(np.arange(self.n_contexts), np.arange(self.n_trajectories_per_context), indexing="ij"), -1).reshape(-1, 2)
        rank, world = self.dist_info()
        lo, hi = self.shard_range(n_items, rank, world)
        rows = self.eval_agent(agent, items[lo:hi])
        rows = self.gather_rows(rows, n_items).cpu()
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
        return successes, mode_encoding, mean_distance
