"""MixedBatch: several task scenes stepped side by side on one GPU (BASELINE.json configs[4], "all 7 tasks mixed batch").

Env instances are grouped by task in contiguous blocks; every task has its own ``BatchedEnv`` handle (its own compiled
scene tables and kernel specialisation) and its own CUDA stream, so the per-task step kernels of one env step overlap on
the device.  There is no cross-task exchange: the reference evaluates each task in its own process as well
(``simulation/*_sim.py``); what is shared is the GPU.
"""
from __future__ import annotations

import torch

from .batched_env import BatchedEnv

SEVEN_CONFIGS = ("avoiding", "aligning", "pushing", "sorting_2", "sorting_4", "sorting_6", "stacking")     # configs/*.yaml (state-based)


class MixedBatch:
    def __init__(self, n_envs_total: int, device: int = 0, tasks=SEVEN_CONFIGS, weights=None):
        self.device = torch.device(f"cuda:{device}")
        self.tasks = tuple(tasks)
        w = [1.0] * len(self.tasks) if weights is None else list(weights)
        counts = [max(1, int(n_envs_total * x / sum(w))) for x in w]
        counts[0] += n_envs_total - sum(counts)
        self.counts = counts
        self.envs = [BatchedEnv(t, c, device) for t, c in zip(self.tasks, counts)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.tasks]
        self.n_envs = sum(counts)

    def close(self):
        for e in self.envs:
            e.close()

    def _fanout(self, fn):
        """Run fn(env, k) for every task on that task's stream; the caller's stream waits for all of them."""
        cur = torch.cuda.current_stream(self.device)
        outs = []
        for k, (e, s) in enumerate(zip(self.envs, self.streams)):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs.append(fn(e, k))
        for s in self.streams:
            cur.wait_stream(s)
        return outs

    def reset(self, contexts: list, masks: list | None = None):
        return self._fanout(lambda e, k: e.reset(contexts[k], None if masks is None else masks[k]))

    def step(self, actions: list):
        """actions[k]: [counts[k], act_dim_k].  Returns the per-task (obs, reward, done, info) tuples."""
        return self._fanout(lambda e, k: e.step(actions[k]))

    @property
    def kernel_launches(self) -> int:
        return sum(e.kernel_launches for e in self.envs)
