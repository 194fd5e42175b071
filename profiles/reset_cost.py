import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv
n = 4096
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
ctx = torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda")
env.reset(ctx)
mask0 = torch.zeros(n, dtype=torch.uint8, device="cuda")
mask1 = torch.zeros(n, dtype=torch.uint8, device="cuda"); mask1[::400] = 1
maskall = torch.ones(n, dtype=torch.uint8, device="cuda")
def t(label, fn, reps=50):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1)/reps*1e3:.1f} us")
t("reset mask=0", lambda: env.reset(ctx, mask0))
t("reset 1/400 envs", lambda: env.reset(ctx, mask1))
t("reset all", lambda: env.reset(ctx, maskall), 10)
