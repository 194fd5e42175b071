import os, sys
import numpy as np, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv
n = 4096; nwarm = int(sys.argv[1]) if len(sys.argv) > 1 else 180
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
env.reset(torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda"))
des = torch.cat([env.robot_state().clone(), torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
g = torch.Generator(device='cuda').manual_seed(0)
lo, hi = torch.tensor([0.3, -0.45], device='cuda'), torch.tensor([0.8, 0.45], device='cuda')
for k in range(nwarm):
    des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + torch.rand(n, 2, generator=g, device='cuda') * 0.02 - 0.01, lo), hi)
    env.step(des)
torch.cuda.synchronize()
# raw state rows: misc block sits at a fixed offset; fetch through torch from the device pointer is not exposed -> use get_state for a sample
import numpy as np
from d3il_b200 import lib
rows = []
for e in range(0, n, 4):
    s = env.get_state(e)
    rows.append(s[-8:])       # task_state words: [task0..3, cost_iters, coupled, ncon, 0]
rows = np.array(rows)
it, cp, nc = rows[:, 4], rows[:, 5], rows[:, 6]
print("newton iters per env step: mean %.1f p50 %.0f p90 %.0f p99 %.0f max %.0f" % (it.mean(), *np.percentile(it, [50, 90, 99, 100])))
print("coupled ticks per env step: frac envs >0: %.3f, max %.0f" % ((cp > 0).mean(), cp.max()))
print("max contacts: mean %.1f max %.0f" % (nc.mean(), nc.max()))
worst = np.argsort(-it)[:8]
print("worst envs (iters, coupled, ncon):", [(int(it[i]), int(cp[i]), int(nc[i])) for i in worst])
