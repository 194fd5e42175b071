import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv
n = 4096
mode = sys.argv[1] if len(sys.argv) > 1 else "walk"
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
ctx = torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda")
env.reset(ctx)
des = torch.cat([env.robot_state().clone(), torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
g = torch.Generator(device="cuda").manual_seed(0)
lo, hi = torch.tensor([0.3, -0.45], device="cuda"), torch.tensor([0.8, 0.45], device="cuda")
out = []
for blk in range(20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(10):
        if mode == "walk":
            des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + torch.rand(n, 2, generator=g, device="cuda") * 0.02 - 0.01, lo), hi)
        elif mode == "drift":
            des[:, :2] += 0.0005
        obs, rew, done, info = env.step(des)
    e1.record(); torch.cuda.synchronize()
    st = env.get_state(0)
    out.append(f"{e0.elapsed_time(e1)/10:.2f}")
print(mode, "ms/step per 10-step block:", " ".join(out))
print("status flags nonzero:", int((info[:, 3] != 0).sum()), " fingers env0:", env.get_state(0)[7:9])
