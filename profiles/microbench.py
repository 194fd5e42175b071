"""Steady-state timing of the bare d3il_step call (no resets, no torch-side action updates), CUDA events."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
ctx = torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda")
env.reset(ctx)
des = torch.cat([env.robot_state().clone(), torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
g = torch.Generator(device="cuda").manual_seed(0)
deltas = torch.rand(64, n, 2, generator=g, device="cuda") * 0.02 - 0.01
lo, hi = torch.tensor([0.3, -0.45], device="cuda"), torch.tensor([0.8, 0.45], device="cuda")
def run(k0, k1, with_reset):
    for k in range(k0, k1):
        des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + deltas[k % 64], lo), hi)
        obs, rew, done, info = env.step(des)
        if with_reset:
            env.reset(ctx, done)
run(0, 30, False)
for label, wr in (("step only", False), ("step + masked reset", True)):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(30, 70, wr); e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) / 40:.3f} ms per step  ({n * 40 / e0.elapsed_time(e1) * 1e3:.0f} env-steps/s)")
# same, but only the C call in the loop (fixed action)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(40):
    env.step(des)
e1.record(); torch.cuda.synchronize()
print(f"fixed action: {e0.elapsed_time(e1) / 40:.3f} ms per step")
