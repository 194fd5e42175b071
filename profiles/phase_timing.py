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
steps = 10
for k in range(steps):
    des[:, :2] += torch.rand(n, 2, generator=g, device="cuda") * 0.02 - 0.01
    env.step(des)
torch.cuda.synchronize()
out = (C.c_ulonglong * 24)()
lib.lib().d3il_debug_phase_cycles(out)
names = {0: "ctrl+kinematics+tcp", 1: "dynamics", 2: "collision", 3: "make_constraints", 4: "chol(M)+solve", 5: "newton total", 6: "euler+integrate",
         15: "newton: loop top", 16: "wait for IK tick (thread 0)", 8: "newton: jar+eval+grad", 9: "newton: H assembly", 10: "newton: chol(H)", 11: "newton: solve", 12: "newton: line search"}
ticks = steps * 35 + 1
tot = sum(out[k] for k in range(7))
for k in sorted(names):
    print(f"{names[k]:28s} {out[k]/ticks:10.0f} cycles/tick  {100*out[k]/tot:5.1f}%")
print("total cycles/tick", tot / ticks, "newton iterations/tick", out[20] / ticks)
