"""Tiny driver for ncu captures: a few env steps of the bench workload (Pushing, 4096 envs, random-walk actions)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
ctx = torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda")
env.reset(ctx)
des = torch.cat([env.robot_state().clone(), torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
g = torch.Generator(device="cuda").manual_seed(0)
lo, hi = torch.tensor([0.3, -0.45], device="cuda"), torch.tensor([0.8, 0.45], device="cuda")
tcp0 = env.robot_state().clone()
ids = torch.arange(n, device="cuda")
preroll = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # 400 = bench.py's steady-state mix of episode phases
for k in range(preroll + steps):
    des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + torch.rand(n, 2, generator=g, device="cuda") * 0.02 - 0.01, lo), hi)
    o, r, d, i = env.step(des)
    m = (ids % 400 == k).to(torch.uint8) if k < preroll else d
    env.reset(ctx, m)
    des[:, :3] = torch.where(m.bool().unsqueeze(1), tcp0, des[:, :3])
torch.cuda.synchronize()
print("done", env.kernel_launches)
