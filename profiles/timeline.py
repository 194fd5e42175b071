"""Debug: block-level timeline of one env step (k_ik blocks + k_env CTAs), from %globaltimer (timing build only)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _variant  # noqa: F401  (D3IL_VARIANT=<name> selects a diagnostic build)
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv
from d3il_b200 import lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
env.reset(torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda"))
des = torch.cat([env.robot_state().clone(), torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
nwarm = int(sys.argv[2]) if len(sys.argv) > 2 else 39
g = torch.Generator(device='cuda').manual_seed(0)
lo, hi = torch.tensor([0.3, -0.45], device='cuda'), torch.tensor([0.8, 0.45], device='cuda')
ctx_t = torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device='cuda')
tcp0 = env.robot_state().clone(); ids = torch.arange(n, device='cuda')
stagger = len(sys.argv) > 3
for k in range(nwarm):
    des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + torch.rand(n, 2, generator=g, device='cuda') * 0.02 - 0.01, lo), hi)
    obs, rew, done, info = env.step(des)
    if stagger:
        force = ((ids % 400 == k % 400) | done.bool()).to(torch.uint8)
        env.reset(ctx_t, force)
        des[:, :3] = torch.where(force.bool().unsqueeze(1), tcp0, des[:, :3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
if hasattr(lib.lib(), 'd3il_debug_cta_stat'): lib.lib().d3il_debug_cta_stat(None, 1)
e0.record(); env.step(des); e1.record(); torch.cuda.synchronize()
stat = None
if hasattr(lib.lib(), 'd3il_debug_cta_stat'):
    sb = (C.c_uint * (4 * 4096))(); lib.lib().d3il_debug_cta_stat(C.cast(sb, C.POINTER(C.c_uint)), 0); stat = np.array(sb, dtype=np.int64).reshape(4096, 4)
print(f'last step by CUDA events: {e0.elapsed_time(e1):.3f} ms')
buf = (C.c_ulonglong * (4 * 4096))()
lib.lib().d3il_debug_timeline(buf)
a = np.array(buf, dtype=np.uint64).reshape(4096, 4).astype(np.int64)
env_b = a[(a[:, 3] == 2)]; ik_b = a[(a[:, 3] == 1)]; sch = a[(a[:, 3] == 3)]
t0 = min(env_b[:, 0].min(), ik_b[:, 0].min()) if len(ik_b) else env_b[:, 0].min()
if len(sch): print(f"k_sched: start {(sch[0,0]-t0)/1e6:.3f} end {(sch[0,1]-t0)/1e6:.3f} ms (relative to the first k_ik block)")
if len(ik_b): print(f"k_ik blocks {len(ik_b)}: start {(ik_b[:,0].min()-t0)/1e6:.3f}..{(ik_b[:,0].max()-t0)/1e6:.3f} ms, end {(ik_b[:,1].min()-t0)/1e6:.3f}..{(ik_b[:,1].max()-t0)/1e6:.3f} ms")
print(f"k_env CTAs {len(env_b)}: start {(env_b[:,0].min()-t0)/1e6:.3f}..{(env_b[:,0].max()-t0)/1e6:.3f} ms, end {(env_b[:,1].min()-t0)/1e6:.3f}..{(env_b[:,1].max()-t0)/1e6:.3f} ms")
dur = (env_b[:, 1] - env_b[:, 0]) / 1e6
st = (env_b[:, 0] - t0) / 1e6
for lo, hi in ((0, 0.5), (0.5, 3), (3, 6), (6, 9), (9, 20)):
    sel = (st >= lo) & (st < hi)
    if sel.any():
        print(f"  CTAs starting in [{lo},{hi}) ms: {sel.sum():4d}  duration mean {dur[sel].mean():.3f} min {dur[sel].min():.3f} max {dur[sel].max():.3f} ms")
print('CTA duration percentiles (ms): ' + ' '.join(f'p{p}={np.percentile(dur, p):.2f}' for p in (5, 25, 50, 75, 90, 95, 99, 100)))
sm_counts = np.bincount(env_b[:, 2].astype(int), minlength=148)
print("CTAs per SM: min", sm_counts.min(), "max", sm_counts.max(), " SMs hosting k_ik blocks:", len(set(ik_b[:, 2].tolist())))
# finer view: start-time histogram and how many k_env CTAs run beside a k_ik block at t = 0.3 ms
edges = np.arange(0, 8.01, 0.25)
hist, _ = np.histogram(st, bins=edges)
print("k_env CTA starts per 0.25 ms:", " ".join(f"{h}" for h in hist))
en = (env_b[:, 1] - t0) / 1e6
ik_sms = set(ik_b[:, 2].tolist())
for t in (0.3, 1.0, 2.0):
    running = (st <= t) & (en > t)
    per_sm = np.bincount(env_b[running, 2].astype(int), minlength=148)
    on_ik = [per_sm[s] for s in range(148) if s in ik_sms]
    off_ik = [per_sm[s] for s in range(148) if s not in ik_sms]
    print(f"t={t} ms: running k_env CTAs {running.sum()}; per SM with k_ik: {np.bincount(on_ik, minlength=3)}; without: {np.bincount(off_ik, minlength=3) if off_ik else []}")

if stat is not None:
    order = np.argsort(-dur)[:12]
    rows = np.nonzero(a[:, 3] == 2)[0]
    print("slowest CTAs of the timeline step (block, start ms, duration ms, passes/coupled/own-steps/ls-evals):", [(int(rows[i]), round(float(st[i]), 2), round(float(dur[i]), 2), *map(int, stat[rows[i]])) for i in order])
    last = np.argsort(-en)[:12]
    print("last CTAs to end (block, start ms, end ms, passes):", [(int(rows[i]), round(float(st[i]), 2), round(float(en[i]), 2), int(stat[rows[i], 0])) for i in last])
    top = np.argsort(-stat[:len(env_b), 0])[:12]
    print("most Newton passes in the stat step: block: passes/coupled/own-steps/ls-evals:", [(int(b), *map(int, stat[b])) for b in top])
    print("Newton passes per CTA: p50 %d p90 %d p99 %d max %d" % tuple(np.percentile(stat[:len(env_b), 0], [50, 90, 99, 100])))
