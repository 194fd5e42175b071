"""CPU checks of the CUDA kernel core (one-lane host build, tests/emu) on the Sorting-k and Aligning scenes against the
fp64 oracle: fp64 build = logic equivalence, fp32 build = the precision the GPU path delivers.  Contact-rich env steps
in which a box is tipping on an edge are ill-conditioned (a 1e-7 perturbation changes the contact set), so the fp32
bound is a quantile bound: >= 85 % of the teacher-forced env steps inside the tolerance box, none blown up."""
import numpy as np
import pytest

from d3il_b200.scene.blob import load_scene
from oracle.oracle import OracleEnv
from tests.emu.emu import EmuEnv
from tests.util import scripted_task_actions, step_errors, task_contexts

TASKS = ["sorting_2", "sorting_4", "sorting_6", "aligning", "inserting"]


@pytest.mark.parametrize("task", TASKS)
def test_new_task_logic_equals_oracle_fp64(task):
    blob, sc = load_scene(task)
    nq, nv, nx = sc.header["nq"], sc.header["nv"], sc.header.get("nextra", 0)
    ctx = task_contexts(task)[1]
    o, e = OracleEnv(blob, sc.header), EmuEnv(blob, sc.header, "f64")
    oo, eo = o.reset(ctx), e.reset(ctx)
    so, se = o.get_state(), e.get_state()
    assert np.abs(so[:nq + 2 * nv + 56] - se[:nq + 2 * nv + 56]).max() < 3e-6      # contexts pass through float32 on the kernel side
    if nx:
        assert np.allclose(so[-nx:], se[-nx:], atol=1e-7)
    assert np.allclose(oo, eo, rtol=1e-4, atol=1e-5)
    touched = False
    for a in scripted_task_actions(task, ctx, o.robot_state(), n_steps=50):
        e.set_state(o.get_state())
        ro, re = o.step(a), e.step(a)
        so, se = o.get_state(), e.get_state()
        assert np.abs(so[:nq] - se[:nq]).max() < 1e-8
        assert np.abs(so[nq:nq + nv] - se[nq:nq + nv]).max() < 1e-6 * (1 + np.abs(so[nq:nq + nv]).max())
        assert np.allclose(ro[0], re[0], rtol=1e-4, atol=1e-5) and ro[2] == re[2] and np.allclose(ro[3], re[3], atol=1e-6)
        assert abs(ro[1] - re[1]) < 1e-6
        con = o.probe("contacts").reshape(-1, 12)
        touched |= bool(len(con)) and bool((con[:, 11] >= 0).any())
    assert touched and so[-nx - 16 + 6] == 0          # status word: no overflow / solver fault


@pytest.mark.parametrize("task", TASKS)
def test_new_task_fp32_env_steps(task):
    blob, sc = load_scene(task)
    nq, nv = sc.header["nq"], sc.header["nv"]
    ctx = task_contexts(task)[2]
    o, e = OracleEnv(blob, sc.header), EmuEnv(blob, sc.header, "f32")
    o.reset(ctx)
    errs = []
    for a in scripted_task_actions(task, ctx, o.robot_state(), n_steps=50):
        e.set_state(o.get_state())
        ro, re = o.step(a), e.step(a)
        errs.append(step_errors(o.get_state(), e.get_state(), nq, nv))
        assert ro[2] == re[2] and np.array_equal(ro[3][:2], re[3][:2]) and re[3][3] == 0
    errs = np.array(errs)
    assert (errs.max(axis=1) <= 1.0).mean() >= 0.85, errs
    assert errs[:, 0].max() < 2e3            # nothing blows up: worst qpos error < 1e-2


def test_sorting_boxes_settle_on_platform():
    """C13: boxes spawn 80 mm inside the platform and are pushed out to rest on its top face (z = 0.10 + 0.03)."""
    blob, sc = load_scene("sorting_4")
    o = OracleEnv(blob, sc.header)
    ctx = task_contexts("sorting_4")[0]
    o.reset(ctx)
    a = np.concatenate([o.robot_state(), [0, 1, 0, 0]])
    for _ in range(12):
        obs, r, done, info = o.step(a)
    s = o.get_state()
    z = s[9 + 2:sc.header["nq"]:7]
    assert np.all(np.abs(z - 0.1299) < 2e-4), z
    assert np.abs(s[sc.header["nq"] + 9:sc.header["nq"] + sc.header["nv"]]).max() < 1e-2          # still creeping into the soft contact
    assert np.allclose(obs[2::3][:4], ctx[:, 0], atol=2e-3) and info[0] == 0 and info[1] == 240 and info[3] == 0


def test_sorting_mode_bookkeeping():
    """check_mode / decode_mode (sorting.py:460-507): teleport boxes into the bins and watch the packed mode word."""
    blob, sc = load_scene("sorting_2")
    o = OracleEnv(blob, sc.header)
    o.reset(task_contexts("sorting_2")[0])
    a = np.concatenate([o.robot_state(), [0, 1, 0, 0]])
    s = o.get_state()
    s[9:12] = [0.62, 0.3, 0.13]         # red box parked in the BLUE bin: nothing is sorted
    s[16:19] = [0.62, 0.35, 0.03]       # blue box in the blue bin (on the table)
    o.set_state(s)
    obs, r, done, info = o.step(a)
    assert info[0] == 0 and info[2] == 1 and info[1] == 0b11000000      # one entry, value 1 (blue) -> packbits keeps 1s
    s = o.get_state()
    s[9:12] = [0.4, 0.3, 0.03]
    o.set_state(s)
    obs, r, done, info = o.step(a)
    assert info[0] == 1 and info[2] == 2 and info[1] == 0b10000000      # second entry red (0)
    obs, r, done, info = o.step(a)
    assert done                                                          # terminated flag seen by the next is_finished


def test_aligning_box_rests_and_reports_target():
    blob, sc = load_scene("aligning")
    o = OracleEnv(blob, sc.header)
    ctx = task_contexts("aligning")[0]
    obs = o.reset(ctx)
    a = np.concatenate([o.robot_state(), [0, 1, 0, 0]])
    for _ in range(10):
        obs, r, done, info = o.step(a)
    assert np.allclose(obs[10:17], ctx[1], atol=1e-6)                    # target pose from model.body_pos/quat
    assert abs(obs[5] - (-0.019 + 0.01)) < 3e-4                          # base plate (half height 0.01) on the table top z = -0.019
    o.forward()
    f = o.probe("efc_force")
    con = o.probe("contacts").reshape(-1, 12)
    normal = sum(f[int(c[11])] for c in con if c[11] >= 0)
    assert abs(normal - 1.004 * 9.81) < 2e-3 * 9.81                      # four plate corners carry m g
    assert np.allclose(con[:, 10], 0.3 / np.sqrt(3))                     # priority-1 plate friction wins over the table's
    pd = np.linalg.norm(obs[3:6] - obs[10:13])
    assert info[1] == 1 and abs(info[2] - 0.5 * (pd + 2 * np.arccos(abs(obs[6:10] @ obs[13:17])) / np.pi)) < 1e-4


def test_stacking_grasp_and_lift_fp64_and_fp32():
    """Joint-space action + gripper (ctrl_kind 1), condim-4 pad / finger-hull contacts, 4 x 4 cone blocks: the scripted
    grasp lifts the red box 14 cm in the oracle; the kernel core follows teacher-forced, fp64 to 1e-8, fp32 inside the
    tolerance box for >= 85 % of the env steps."""
    from tests.util import scripted_grasp_actions
    blob, sc = load_scene("stacking")
    nq, nv = sc.header["nq"], sc.header["nv"]
    ctx = task_contexts("stacking")[1]
    o = OracleEnv(blob, sc.header)
    obs0 = o.reset(ctx)
    acts = scripted_grasp_actions(sc, ctx, o.robot_state(), o.joint_state()[:7], obs0)
    emus = {p: EmuEnv(blob, sc.header, p) for p in ("f64", "f32")}
    for e in emus.values():
        e.reset(ctx)
    assert np.abs(o.get_state()[:nq + 2 * nv + 56] - emus["f64"].get_state()[:nq + 2 * nv + 56]).max() < 1e-6
    errs, max_rows = [], 0
    for a in acts:
        s0 = o.get_state()
        ro = o.step(a)
        max_rows = max(max_rows, o.probe("counts")[1])
        for prec, e in emus.items():
            e.set_state(s0)
            re = e.step(a)
            assert ro[2] == re[2] and np.allclose(ro[3][:4], re[3][:4], atol=1e-4) and re[3][4] == 0
            assert np.allclose(ro[0], re[0], rtol=1e-3, atol=1e-4)
            if prec == "f64":
                assert np.abs(o.get_state()[:nq] - e.get_state()[:nq]).max() < 1e-7          # actions pass through float32 on the kernel side
            else:
                errs.append(step_errors(o.get_state(), e.get_state(), nq, nv))
    errs = np.array(errs)
    assert (errs.max(axis=1) <= 1.0).mean() >= 0.85, errs
    assert ro[0][2] > 0.14 and abs(o.joint_state()[7] - 0.0613) < 2e-3            # red box lifted, gripper closed on a 6 cm box
    assert max_rows >= 80                                                          # the grasp is a >= 80-row problem with condim-4 rows


def test_stacking_mode_and_success_bookkeeping():
    """check_mode / _check_early_termination (stacking.py:395-447) with teleported boxes."""
    blob, sc = load_scene("stacking")
    o = OracleEnv(blob, sc.header)
    ctx = task_contexts("stacking")[0]
    o.reset(ctx)
    hold = np.concatenate([o.joint_state()[:7], [0.08]])
    s = o.get_state()
    s[16:19] = [0.5, 0.2, 0.011]                        # green box on the target
    o.set_state(s)
    obs, r, done, info = o.step(hold)
    assert info[0] == 0 and info[1] == 2 and info[3] == 1              # mode "g"
    s = o.get_state()
    s[9:12] = [0.5, 0.2, 0.071]                         # red stacked on green
    s[23:26] = [0.5, 0.2, 0.131]                        # blue on top
    o.set_state(s)
    obs, r, done, info = o.step(hold)
    assert info[3] == 2 and info[1] in (2 + 4 * 1, 2 + 4 * 3) and info[0] == 1     # second arrival appended; all three within 6 cm, z gaps > 3 cm
    obs, r, done, info = o.step(hold)
    assert done and info[3] == 3


def test_box_off_the_table_lands_on_the_ground():
    """A box released beyond the table's front edge (x = 0.89) falls 0.92 m and comes to rest on the ground plane of base.xml
    (body `ground`, z = -0.94), carrying its weight there - it used to fall for ever.  Oracle, fp64 and fp32 host builds of the
    kernel core agree along the fall, through the impact and at rest."""
    blob, sc = load_scene("pushing")
    nq, nv = sc.header["nq"], sc.header["nv"]
    ctx = task_contexts("pushing")[0]
    o = OracleEnv(blob, sc.header)
    o.reset(ctx)
    s = o.get_state()
    s[9:12] = [1.0, 0.05, 0.0]                    # box 1, 11 cm past the edge
    s[nq:nq + nv] = 0
    o.set_state(s)
    emus = {p: EmuEnv(blob, sc.header, p) for p in ("f64", "f32")}
    for e in emus.values():
        e.reset(ctx); e.set_state(s)
    zs = []
    for k in range(18):                              # 18 x 50 ticks = 0.9 s: free fall takes 0.43 s
        o.substep(50)
        so = o.get_state()
        zs.append(so[11])
        for p, e in emus.items():
            e.substep(50)
            se = e.get_state()
            if p == "f64":
                assert np.abs(so[:nq] - se[:nq]).max() < 1e-7 and np.abs(so[nq:nq + nv] - se[nq:nq + nv]).max() < 1e-5, (k, p)
            e.set_state(so)                          # teacher-forced every 50 ticks
            if p == "f32":
                assert np.abs(so[9:12] - se[9:12]).max() < 2e-4 + 1e-3 * (k in (8, 9)), (k, np.abs(so[9:12] - se[9:12]).max())      # the impact (0.43 s) is the ill-conditioned window
    assert zs[3] < -0.15 and min(zs) > -0.98          # it fell; the impact at 4.2 m/s sinks ~3 cm into the default-solref (20 ms) contact and comes back
    so = o.get_state()
    assert abs(so[11] - (-0.94 + 0.03)) < 5e-4 and np.abs(so[nq + 9:nq + 15]).max() < 2e-3       # at rest on the ground
    con = o.probe("contacts").reshape(-1, 12)
    import json, os
    names = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "d3il_b200", "scenes", "pushing.json")))["geoms"]
    on_ground = con[(con[:, 8] == names.index("ground")) & (con[:, 11] >= 0)]
    assert len(on_ground) == 4
    f = o.probe("efc_force")
    fn = sum(f[int(r)] for r in on_ground[:, 11])
    assert abs(fn - 0.05 * 9.81) < 0.02 * 0.05 * 9.81                               # the four corner contacts carry m g
