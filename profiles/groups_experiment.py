"""Experiment: 4096 Pushing envs as G independent lock-step groups on G CUDA streams (cross-step pipelining: one group's
slow tail overlaps the other groups' bulk).  Prints env-steps/s for several G."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv

N = 4096
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
dev = torch.device("cuda:0")
lo, hi = torch.tensor([0.3, -0.45], device=dev), torch.tensor([0.8, 0.45], device=dev)


class Group:
    def __init__(self, n, off, seed):
        self.n = n
        self.env = BatchedEnv("pushing", n, 0)
        self.stream = torch.cuda.Stream(device=dev)
        self.ctx = torch.tensor(ctxs[(np.arange(n) + off) % 60], dtype=torch.float32, device=dev)
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.ids = torch.arange(n, device=dev) + off
        with torch.cuda.stream(self.stream):
            self.env.reset(self.ctx)
            self.tcp0 = self.env.robot_state().clone()
            self.des = torch.cat([self.tcp0, torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev).repeat(n, 1)], 1)

    def advance(self, k=None):
        with torch.cuda.stream(self.stream):
            self.des[:, :2] = torch.minimum(torch.maximum(self.des[:, :2] + torch.rand(self.n, 2, generator=self.gen, device=dev) * 0.02 - 0.01, lo), hi)
            o, r, d, i = self.env.step(self.des)
            m = d if k is None else (self.ids % 400 == k).to(torch.uint8)
            self.env.reset(self.ctx, m)
            self.des[:, :3] = torch.where(m.bool().unsqueeze(1), self.tcp0, self.des[:, :3])


for G in [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8]:
    groups = [Group(N // G, g * (N // G), 100 + g) for g in range(G)]
    for k in range(400):
        for g in groups:
            g.advance(k)
    for k in range(10):
        for g in groups:
            g.advance()
    torch.cuda.synchronize()
    K = 60
    t0 = time.perf_counter()
    for k in range(K):
        for g in groups:
            g.advance()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"G={G}: {N * K / dt:.0f} env-steps/s  ({1e3 * dt / K:.2f} ms per round of all groups)", flush=True)
    for g in groups:
        g.env.close()
    del groups
