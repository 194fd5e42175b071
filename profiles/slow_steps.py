"""Debug (timing build): step through the steady-state mix and describe the env steps whose k_env ran long: which CTAs ended
last, where they sat in the cost order, their Newton statistics.  D3IL_VARIANT=timing python profiles/slow_steps.py [steps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _variant  # noqa: F401
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv
from d3il_b200 import lib
n = 4096
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 80
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
ctx_t = torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda")
env.reset(ctx_t)
tcp0 = env.robot_state().clone()
des = torch.cat([tcp0, torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
lo, hi = torch.tensor([0.3, -0.45], device="cuda"), torch.tensor([0.8, 0.45], device="cuda")
ids = torch.arange(n, device="cuda")
L = lib.lib()
def step(force=None):
    des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + torch.rand(n, 2, device="cuda") * 0.02 - 0.01, lo), hi)
    o, r, d, i = env.step(des)
    m = d if force is None else (force | d)
    env.reset(ctx_t, m)
    des[:, :3] = torch.where(m.bool().unsqueeze(1), tcp0, des[:, :3])
    return d
for k in range(400):
    step((ids % 400 == k).to(torch.uint8))
torch.cuda.synchronize()
ends = []
for k in range(steps):
    L.d3il_debug_cta_stat(None, 1)
    done = step()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (4 * 4096))(); L.d3il_debug_timeline(buf)
    a = np.array(buf, dtype=np.uint64).reshape(4096, 4).astype(np.int64)
    sb = (C.c_uint * (4 * 4096))(); L.d3il_debug_cta_stat(C.cast(sb, C.POINTER(C.c_uint)), 0); stat = np.array(sb, dtype=np.int64).reshape(4096, 4)
    rows = np.nonzero(a[:, 3] == 2)[0]
    t0 = a[a[:, 3] == 3][0, 0]
    end = (a[rows, 1].max() - t0) / 1e6
    ends.append(end)
    if end > 7.0:
        late = rows[np.argsort(-a[rows, 1])[:5]]
        print(f"step {k}: k_env ended at {end:.2f} ms; last CTAs (block, start, end, passes/coupled/own steps/ls evals): " +
              ", ".join(f"({b}, {(a[b, 0] - t0) / 1e6:.2f}, {(a[b, 1] - t0) / 1e6:.2f}, {'/'.join(map(str, stat[b]))})" for b in late))
ends = np.array(ends)
print("k_env end (ms): mean %.2f p50 %.2f p90 %.2f max %.2f; steps over 7 ms: %d of %d" % (ends.mean(), np.median(ends), np.percentile(ends, 90), ends.max(), (ends > 7).sum(), len(ends)))
