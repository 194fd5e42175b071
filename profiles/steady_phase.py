"""Debug: per-phase cycle counts of one CTA and all-env cost statistics in the STEADY-STATE episode mix of bench.py
(build with make EXTRA='-DD3IL_PHASE_TIMING -DD3IL_PHASE_BLOCK="(gridDim.x/2)"')."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _variant  # noqa: F401  (D3IL_VARIANT=<name> selects a diagnostic build)
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv
from d3il_b200 import lib
n = 4096
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
ctx_t = torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda")
env.reset(ctx_t)
tcp0 = env.robot_state().clone()
des = torch.cat([tcp0, torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
g = torch.Generator(device="cuda").manual_seed(0)
lo, hi = torch.tensor([0.3, -0.45], device="cuda"), torch.tensor([0.8, 0.45], device="cuda")
ids = torch.arange(n, device="cuda")
def step(force=None):
    des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + torch.rand(n, 2, generator=g, device="cuda") * 0.02 - 0.01, lo), hi)
    o, r, d, i = env.step(des)
    m = d if force is None else force
    env.reset(ctx_t, m)
    des[:, :3] = torch.where(m.bool().unsqueeze(1), tcp0, des[:, :3])
for k in range(400):
    step((ids % 400 == k).to(torch.uint8))
torch.cuda.synchronize()
has_phase = hasattr(lib.lib(), "d3il_debug_phase_cycles")
base = (C.c_ulonglong * 40)()
if has_phase: lib.lib().d3il_debug_phase_cycles(base)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for k in range(steps): step()
ev1.record(); torch.cuda.synchronize()
print("ms/step", ev0.elapsed_time(ev1) / steps)
if has_phase:
    out = (C.c_ulonglong * 40)(); lib.lib().d3il_debug_phase_cycles(out)
    out = [out[i] - base[i] for i in range(40)]
    names = {0: "ctrl+kinematics+tcp", 1: "dynamics", 2: "collision", 3: "make_constraints", 4: "chol(M)+solve", 5: "newton total", 6: "euler+integrate",
             15: "newton: loop top", 16: "wait for IK tick (thread 0)", 8: "newton: jar+eval+grad", 9: "newton: H assembly", 10: "newton: chol(H)", 11: "newton: solve", 12: "newton: line search"}
    ticks = steps * 35
    tot = sum(out[k] for k in range(7))
    for k in sorted(names):
        print(f"{names[k]:28s} {out[k]/ticks:10.0f} cycles/tick  {100*out[k]/tot:5.1f}%")
    print(f"all envs: coupled ticks {out[21]/max(out[23],1):.4f}, mean contacts {out[22]/max(out[23],1):.2f}, mean rows {out[19]/max(out[23],1):.2f}")
    print("total cycles/tick", tot / ticks, "newton iterations/tick (sampled CTA)", out[20] / ticks)
    nit = max(out[20], 1)
    sub = {24: "ls: J p", 25: "ls: M p + reductions", 26: "ls: search", 27: "ls: update", 28: "ls: eval at new iterate", 29: "solve: block solve", 30: "solve: woodbury",
           31: "grad: gradient loop", 32: "H: clear (dense)", 33: "H: in-block assembly", 34: "warm start", 35: "after loop"}
    print("finer (cycles per pass): " + ", ".join(f"{v} {out[k]/nit:.0f}" for k, v in sub.items()))
    print(f"sampled warps per tick: contacts {out[17]/ticks:.2f}, rows {out[7]/ticks:.2f}, ticks with one coupling contact {out[14]/ticks:.3f}, with several {out[13]/ticks:.3f} (per sampled warp: divide by their number)")
    print("per Newton loop pass of the sampled warps (cycles): " + ", ".join(f"{names[k].split(': ')[1]} {out[k]/nit:.0f}" for k in (15, 8, 9, 10, 11, 12)) + f"; line-search evaluations per pass {out[18]/nit:.2f}; passes {nit}")
rows = np.array([env.get_state(e)[-8:] for e in range(0, n, 4)])
it, cp, nc = rows[:, 4], rows[:, 5], rows[:, 6]
print("newton steps per env step: mean %.1f p10 %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f" % (it.mean(), *np.percentile(it, [10, 50, 90, 99, 100])))
print("coupled ticks per env step: frac envs >0: %.3f" % ((cp > 0).mean()))
print("max contacts per env: mean %.1f; hist" % nc.mean(), np.bincount(nc.astype(int)))

if hasattr(lib.lib(), "d3il_debug_iter_hist"):
    h = (C.c_ulonglong * 40)(); lib.lib().d3il_debug_iter_hist(h); h = np.array(list(h), dtype=np.float64)
    tot = h[:16].sum()
    print("Newton steps per tick, all ticks since start (%):", " ".join(f"{100*x/tot:.2f}" for x in h[:16]))
    print("  ... ticks with a coupling contact (% of all ticks):", " ".join(f"{100*x/tot:.3f}" for x in h[16:32]))
    print("  work share by steps/tick (%):", " ".join(f"{100*k*x/max((np.arange(16)*h[:16]).sum(),1):.1f}" for k, x in enumerate(h[:16])))
    print("  ticks with >= 8 steps: mean contacts %.1f" % (h[32] / max(h[33], 1)))
