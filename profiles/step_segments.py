"""Debug: device-time segments of the free-running bench loop (eager launches, no synchronisation inside the loop):
[action ops | env.step | bookkeeping ops | reset | tail ops], CUDA events read back at the end."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _variant  # noqa: F401
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
s = bench.EnvStream("pushing", 4096, 0, 0)
n = s.n; dev = s.dev
ids = torch.arange(n, device=dev)
for k in range(400):
    s.advance((ids % s.ep_len == k).to(torch.uint8))
torch.cuda.synchronize()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 60
mode = sys.argv[2] if len(sys.argv) > 2 else "full"
names = ["action ops", "env.step", "bookkeeping ops", "env.reset", "tail ops"]
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(K)]
env = s.env
for k in range(K):
    e = evs[k]
    e[0].record()
    delta = torch.rand(n, s.n_act, device=dev) * 0.02 - 0.01
    s.des[:, :s.n_act] = torch.minimum(torch.maximum(s.des[:, :s.n_act] + delta, s.lo), s.hi)
    e[1].record()
    obs, rew, done, info = env.step(s.des)
    e[2].record()
    if mode != "steponly":
        s.last_obs.copy_(obs); s.step_no.add_(1)
        st = info[:, -1].to(torch.int32)
        s.fault_steps.add_((st != 0).sum()); s.bit_counts.add_(((st.unsqueeze(1) & s.bits) != 0).sum(0))
        m = done; mb = m.bool().unsqueeze(1)
        s.last_info.copy_(torch.where(mb, info, s.last_info)); s.episodes.add_(m.sum())
    e[3].record()
    if mode != "steponly":
        env.reset(s.ctx_t, m)
    e[4].record()
    if mode != "steponly":
        s.des.copy_(torch.where(mb, s.start, s.des))
    e[5].record()
torch.cuda.synchronize()
seg = np.array([[evs[k][i].elapsed_time(evs[k][i + 1]) for i in range(5)] for k in range(K)])
per = np.array([evs[k][0].elapsed_time(evs[k + 1][0]) for k in range(K - 1)])
print(f"mode {mode}: period mean {per.mean():.3f} ms (p50 {np.median(per):.3f}, p90 {np.percentile(per, 90):.3f})")
for i, nm in enumerate(names):
    print(f"  {nm:18s} mean {seg[:, i].mean():.3f} ms  p50 {np.median(seg[:, i]):.3f}  p90 {np.percentile(seg[:, i], 90):.3f}  max {seg[:, i].max():.3f}")
