"""Regenerates tests/golden/oracle_rollouts.npz: short scripted rollouts of the fp64 oracle for every compiled scene.

What these fixtures are and are not: the reference itself cannot run anywhere in this project (mujoco / pinocchio are
not installable), so there are no reference-generated vectors to commit.  These are REGRESSION vectors of
the oracle - they pin today's restatement (scene tables + oracle arithmetic + task logic) against accidental drift, and
give the GPU tests a box-independent target.  Inputs that do come from the reference: the evaluation contexts
(`environments/dataset/data/{pushing,aligning,stacking}/test_contexts.pkl`, exported by `d3il_b200/scene/compile.py`).

    python tests/golden/make_golden.py        # rewrites the .npz next to this file
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from d3il_b200.scene.blob import load_scene          # noqa: E402
from oracle.oracle import OracleEnv                  # noqa: E402
from tests.util import scripted_grasp_actions, scripted_push_actions, scripted_task_actions, task_contexts  # noqa: E402

SCENES = ["avoiding", "pushing", "aligning", "sorting_2", "sorting_4", "sorting_6", "stacking", "inserting"]
N_STEPS = {"avoiding": 40, "pushing": 100, "aligning": 50, "sorting_2": 50, "sorting_4": 50, "sorting_6": 40, "stacking": 66, "inserting": 50}


def rollout(task):
    blob, sc = load_scene(task)
    o = OracleEnv(blob, sc.header)
    ctx = None if task == "avoiding" else task_contexts(task)[1]
    obs0 = o.reset(ctx)
    if task == "avoiding":
        des = o.robot_state().copy()
        acts = []
        for k in range(N_STEPS[task]):
            des[:2] += [-0.002, 0.006]
            acts.append(np.concatenate([des, [0, 1, 0, 0]]))
        acts = np.array(acts)
    elif task == "pushing":
        acts = scripted_push_actions(ctx, o.robot_state(), n_steps=N_STEPS[task])
    elif task == "stacking":
        acts = scripted_grasp_actions(sc, ctx, o.robot_state(), o.joint_state()[:7], obs0)
    else:
        acts = scripted_task_actions(task, ctx, o.robot_state(), n_steps=N_STEPS[task])
    states, obs, info = [o.get_state()], [obs0], []
    for a in acts:
        ob, r, d, inf = o.step(a)
        states.append(o.get_state()); obs.append(ob); info.append(np.concatenate([[r, float(d)], inf]))
    return dict(actions=acts, states=np.array(states), obs=np.array(obs), info=np.array(info))


def main():
    out = {}
    for t in SCENES:
        for k, v in rollout(t).items():
            out[f"{t}/{k}"] = v
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_rollouts.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("states")})


if __name__ == "__main__":
    main()
