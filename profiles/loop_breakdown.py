import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv
n = 4096
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
ctx = torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda")
env.reset(ctx)
tcp0 = env.robot_state().clone()
des = torch.cat([tcp0, torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
g = torch.Generator(device="cuda").manual_seed(0)
lo, hi = torch.tensor([0.3, -0.45], device="cuda"), torch.tensor([0.8, 0.45], device="cuda")
ids = torch.arange(n, device="cuda")
def step(k, force_stagger):
    des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + torch.rand(n, 2, generator=g, device="cuda") * 0.02 - 0.01, lo), hi)
    obs, rew, done, info = env.step(des)
    m = ((ids % 400 == k % 400) | done.bool()).to(torch.uint8) if force_stagger else done
    return m
for k in range(420):
    m = step(k, True); env.reset(ctx, m); des[:, :3] = torch.where(m.bool().unsqueeze(1), tcp0, des[:, :3])
torch.cuda.synchronize()
N = 60
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(N)]
for k in range(N):
    ev[k][0].record()
    des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + torch.rand(n, 2, generator=g, device="cuda") * 0.02 - 0.01, lo), hi)
    ev[k][1].record()
    obs, rew, done, info = env.step(des)
    ev[k][2].record()
    env.reset(ctx, done)
    des[:, :3] = torch.where(done.bool().unsqueeze(1), tcp0, des[:, :3])
    ev[k][3].record()
torch.cuda.synchronize()
a = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in ev])
tot = ev[0][0].elapsed_time(ev[-1][3]) / N
print("per step ms: action update %.3f | step (k_sched+k_ik+k_env) %.3f | reset+where %.3f | total/step %.3f" % (*a.mean(0), tot))
print("number of resets per step ~", float(sum(env.done.sum().item() for _ in range(1))))
