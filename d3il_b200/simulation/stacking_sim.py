"""``Stacking_Sim`` — drop-in for ``simulation/stacking_sim.py:20-257`` on the batched CUDA env.

Joint-space action (7 joint set-points + gripper command); the agent sees ``[last action (8) || env obs (12)]`` and its
first 7 outputs are deltas integrated on the last desired joint positions (``stacking_sim.py:92-115``).  The mode string of
the reference (``'rgb'`` ...) travels as base-4 digits in ``info[1]`` (r 1, g 2, b 3, first arrival in the lowest digit).
"""
from __future__ import annotations

import logging
import os

import numpy as np
import torch

from ..batched_env import BatchedEnv
from .agent_adapter import graph_safe, predict_batch
from .base_sim import BaseSim, _wandb_log, episode_loop, report_faults
from .metrics import mode_kl, stacking_rows

log = logging.getLogger(__name__)
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data")

# environments/dataset/data/stacking/mode_prob.pkl (prior over the six 3-box orders)
MODE_PROB = {"brg": 0.235, "gbr": 0.287, "rbg": 0.127, "rgb": 0.152, "grb": 0.106, "bgr": 0.093}
MODE_1 = {"r": 0, "g": 1, "b": 2}
MODE_2 = {"rg": 0, "rb": 1, "gr": 2, "gb": 3, "br": 4, "bg": 5}
MODE_3 = {"rgb": 0, "rbg": 1, "grb": 2, "gbr": 3, "brg": 4, "bgr": 5}


def decode_mode(code: int, length: int) -> str:
    return "".join("rgb"[(int(code) // 4 ** k) % 4 - 1] for k in range(int(length)))


class Stacking_Sim(BaseSim):
    def __init__(self, seed: int, device: str, render: bool, n_cores: int = 1, n_contexts: int = 30, n_trajectories_per_context: int = 1,
                 max_steps_per_episode: int = 500, mode_prob: dict | None = None):
        super().__init__(seed, device, render, n_cores)
        self.n_contexts = n_contexts
        self.n_trajectories_per_context = n_trajectories_per_context
        self.max_steps_per_episode = max_steps_per_episode
        self.test_contexts = np.load(os.path.join(_DATA, "stacking_test_contexts.npy"))
        self.modes = dict(MODE_PROB if mode_prob is None else mode_prob)
        enc3 = np.zeros(6)
        for k, v in MODE_3.items():
            enc3[v] = self.modes[k]
        self.mode_encoding_3 = torch.tensor(enc3)
        self.mode_encoding_2 = torch.tensor(enc3.copy())                   # stacking_sim.py:49-54 (sic: the 2-box prior reuses the 3-box one)
        self.mode_encoding_1 = torch.tensor([enc3[i] + enc3[i + 1] for i in (0, 2, 4)])

    @torch.no_grad()
    def eval_agent(self, agent, items: np.ndarray):
        """[n, 6] rows (mode_3, mode_1, mode_2, success, success_1, success_2) of the (context, rollout) pairs in ``items``;
        modes are -1 when the mode string is too short (the reference leaves zeros there, ``stacking_sim.py:122-141``)."""
        dev_index = self._cuda_index()
        dev = torch.device(f"cuda:{dev_index}")
        n = len(items)
        if n == 0:
            return torch.zeros(0, 6, device=dev)
        env = BatchedEnv("stacking", n, dev_index)
        torch.manual_seed(self.seed)
        agent.reset()
        env.reset(torch.tensor(self.test_contexts[items[:, 0]], dtype=torch.float32, device=dev))
        act = env.joint_state().clone()                                        # :89 joint positions + gripper width

        def policy_step(active):
            agent_in = torch.cat([act, env.obs], 1)                             # :99
            out = predict_batch(agent, agent_in)                                # :103
            out[:, :7] = out[:, :7] + agent_in[:, :7]                           # :104
            act.copy_(torch.where(active.unsqueeze(1), out, act))
            return act                                                          # :114

        cap = min(int(self.max_steps_per_episode), env.max_steps_per_episode)
        info_rows, status = episode_loop(env, policy_step, cap, graph=graph_safe(agent))
        report_faults(status, "stacking")
        env.close()
        return stacking_rows(info_rows)

    def cal_KL(self, mode_encoding, successes, prior_encoding, n_mode=6):
        _, entropy, KL = mode_kl(mode_encoding, successes, {k: float(prior_encoding[k]) for k in range(n_mode)})
        return entropy, KL

    def test_agent(self, agent):
        log.info("Starting trained model evaluation")
        n_items = self.n_contexts * self.n_trajectories_per_context
        items = np.stack(np.meshgrid(np.arange(self.n_contexts), np.arange(self.n_trajectories_per_context), indexing="ij"), -1).reshape(-1, 2)
        rank, world = self.dist_info()
        lo, hi = self.shard_range(n_items, rank, world)
        rows = self.gather_rows(self.eval_agent(agent, items[lo:hi]), n_items)          # result rows and the metrics below stay on the device; only scalars and the returned tensors come back
        shape = (self.n_contexts, self.n_trajectories_per_context)
        mode_encoding, mode_1, mode_2, successes, successes_1, successes_2 = (rows[:, k].reshape(shape).clone() for k in range(6))
        box1, box2, success_rate = successes_1.mean().item(), successes_2.mean().item(), successes.mean().item()
        entropy_1, KL_1 = self.cal_KL(mode_1, successes_1, self.mode_encoding_1, n_mode=3)
        entropy_2, KL_2 = self.cal_KL(mode_2, successes_2, self.mode_encoding_2, n_mode=6)
        entropy_3, KL_3 = self.cal_KL(mode_encoding, successes, self.mode_encoding_3, n_mode=6)
        _wandb_log({"score": box1 + box2 + success_rate})
        for k, v in (("successes", success_rate), ("entropy_3", entropy_3), ("KL_3", KL_3), ("successes_1_box", box1), ("entropy_1", entropy_1),
                     ("KL_1", KL_1), ("successes_2_boxes", box2), ("entropy_2", entropy_2), ("KL_2", KL_2)):
            _wandb_log({f"Metrics/{k}": v})
        print(f"Successrate {success_rate}")
        print(f"Successrate_1 {box1}")
        print(f"Successrate_2 {box2}")
        return successes.cpu(), mode_encoding.cpu()
