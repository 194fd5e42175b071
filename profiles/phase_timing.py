"""Debug: per-phase cycle counts of one warp (build with -DD3IL_PHASE_TIMING)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3il_b200.batched_env import BatchedEnv
from d3il_b200 import lib
n = 4096
ctxs = np.load(os.path.join(os.path.dirname(__file__), "..", "d3il_b200", "data", "pushing_test_contexts.npy"))
env = BatchedEnv("pushing", n, 0)
env.reset(torch.tensor(ctxs[np.arange(n) % 60], dtype=torch.float32, device="cuda"))
des = torch.cat([env.robot_state().clone(), torch.tensor([0.0, 1.0, 0.0, 0.0], device="cuda").repeat(n, 1)], 1)
g = torch.Generator(device="cuda").manual_seed(0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
lo, hi = torch.tensor([0.3, -0.45], device="cuda"), torch.tensor([0.8, 0.45], device="cuda")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
base = (C.c_ulonglong * 24)()
for k in range(steps):
    if k == skip:
        torch.cuda.synchronize(); lib.lib().d3il_debug_phase_cycles(base)
    des[:, :2] = torch.minimum(torch.maximum(des[:, :2] + torch.rand(n, 2, generator=g, device="cuda") * 0.02 - 0.01, lo), hi)
    env.step(des)
torch.cuda.synchronize()
out = (C.c_ulonglong * 24)()
lib.lib().d3il_debug_phase_cycles(out)
out = [out[i] - base[i] for i in range(24)]
steps = steps - skip
names = {0: "ctrl+kinematics+tcp", 1: "dynamics", 2: "collision", 3: "make_constraints", 4: "chol(M)+solve", 5: "newton total", 6: "euler+integrate",
         15: "newton: loop top", 16: "wait for IK tick (thread 0)", 8: "newton: jar+eval+grad", 9: "newton: H assembly", 10: "newton: chol(H)", 11: "newton: solve", 12: "newton: line search"}
ticks = steps * 35
tot = sum(out[k] for k in range(7))
for k in sorted(names):
    print(f"{names[k]:28s} {out[k]/ticks:10.0f} cycles/tick  {100*out[k]/tot:5.1f}%")
print(f"all envs: coupled ticks {out[21]/max(out[23],1):.4f}, mean contacts {out[22]/max(out[23],1):.2f}, mean rows {out[19]/max(out[23],1):.2f}")
print("total cycles/tick", tot / ticks, "newton iterations/tick", out[20] / ticks)
