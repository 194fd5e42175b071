/*
 * d3il.h — C ABI of libd3il.so: the batched, B200-native replacement for D3IL's env-step / rollout hot path.
 *
 * The reference has no FFI for this path (it is plain Python over the mujoco/pinocchio wheels); these entry points
 * are what a maintainer binds with ctypes from the rollout harness.  Each function names the
 * reference interface it replaces (paths relative to /root/reference/environments/d3il/).
 *
 * Conventions: return 0 = OK, negative = error (message in the thread-local d3il_last_error()); no exceptions cross
 * the ABI; no allocation after create; a handle is not thread-safe (one host thread per handle/device, like one
 * process per core in simulation/pushing_sim.py:114-135).  Pointers marked "dev" are device pointers on the handle's
 * device, "host" are host pointers.  Kernels are enqueued on `cu_stream` (a cudaStream_t, NULL = default stream), so
 * policy inference and env stepping serialise on one stream without host synchronisation.  There is NO CPU execution
 * path in this library: create() fails if no CUDA device is usable.
 */
#ifndef D3IL_H
#define D3IL_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct d3il_env d3il_env; /* opaque: n_envs instances of one compiled task scene on one device */

/* D3SC scene blob layout (d3il_b200/scene/blob.py; produced offline by d3il_b200/scene/compile.py) */
#define D3SC_MAGIC_LE 0x43533344u
enum { D3IL_DIM_OBS = 0, D3IL_DIM_ACT, D3IL_DIM_CTX, D3IL_DIM_INFO, D3IL_DIM_STATE, D3IL_DIM_NENVS, D3IL_DIM_SUBSTEPS, D3IL_DIM_MAXSTEPS, D3IL_NDIMS };

/* Replaces MjFactory.create_scene + MjModel.from_xml_string + env.start() (sims/mj_beta/MjFactory.py:14-29,
 * mj_utils/mj_scene_parser.py:36-53, envs/gym_pushing_env/.../pushing.py:283-333): instantiates n_envs copies of the
 * compiled scene on CUDA device `device`. */
int d3il_create(d3il_env** out, const void* scene_blob, size_t nbytes, int n_envs, int device);
void d3il_destroy(d3il_env* env);
const char* d3il_last_error(void);
int d3il_dims(const d3il_env* env, int32_t out[D3IL_NDIMS]);

/* Replaces Env.reset(random=False, context=c) (pushing.py:461-488, avoiding.py:248-262; MjScene.reset
 * sims/mj_beta/MjScene.py:120-143; RobotBase.beam_to_joint_pos core/Robots.py:580-589) for every env whose mask byte
 * is non-zero (mask NULL = all).  ctx: dev [n_envs, ctx_dim] = (x,y,z,qw,qx,qy,qz) per free object, may be NULL for
 * scenes without objects.  obs (nullable): dev [n_envs, obs_dim], written for the envs that were reset. */
int d3il_reset(d3il_env* env, const float* ctx, const uint8_t* mask, float* obs, void* cu_stream);

/* Replaces GymEnvWrapper.step + task overrides (gyms/gym_env_wrapper.py:45-100, pushing.py:335-339,
 * avoiding.py:168-171): one env step = n_substeps physics ticks of Scene.next_step (core/Scene.py:121-138).
 * action: dev [n_envs, act_dim]: desired tcp xyz + quat wxyz (act_dim 7: Avoiding, Pushing, Sorting, Aligning), or
 * 7 joint set-points + gripper command (act_dim 8: CubeStacking_Env.step, stacking.py:331-393).  Outputs (dev): obs [n_envs, obs_dim] f32,
 * reward [n_envs] f32, done [n_envs] u8 (all three sampled BEFORE the substeps, as the reference does),
 * info [n_envs, info_dim] f32 sampled after them (pushing: success, mode, mean_distance, status;
 * avoiding: success, 9 mode bits, status; sorting: success, packed mode, mode_step, status; aligning: success, mode,
 * mean_distance, status; stacking: success, mode string in base-4 digits, mean_distance, len(mode), status).  status != 0 flags a per-env numerical fault / contact overflow. */
int d3il_step(d3il_env* env, const float* action, float* obs, float* reward, uint8_t* done, float* info, void* cu_stream);

/* Replaces GymEnvWrapper.robot_state() (gym_env_wrapper.py:160-189): tcp position, dev [n_envs, 3]. */
int d3il_robot_state(d3il_env* env, float* tcp, void* cu_stream);

/* Replaces CubeStacking_Env.robot_state() (stacking.py:218-226; MjRobot.receiveState MjRobot.py:140-183): 7 joint
 * positions + gripper width, dev [n_envs, 8]. */
int d3il_joint_state(d3il_env* env, float* j8, void* cu_stream);

/* Replaces the fields of MjRobot.receiveState (sims/mj_beta/MjRobot.py:133-184) the reference's RobotLogger writes into the
 * dataset pickles (core/logger.py; read back by environments/dataset/stacking_dataset.py:92-104): dev [n_envs, 22] =
 * current_c_pos (3), current_c_quat (4, wxyz; both one tick stale like the reference's), current_j_pos (7),
 * current_j_vel (7), gripper_width (1). */
int d3il_robot_kinematics(d3il_env* env, float* out22, void* cu_stream);

/* Replaces Scene.get_obj_pos / get_obj_quat (MjScene._get_obj_pos_and_quat, sims/mj_beta/MjScene.py:233-247) for every free
 * object of the scene: dev [n_envs, n_obj, 7] = (x, y, z, qw, qx, qy, qz) read from qpos (what the reference's ObjectLogger
 * records for the datasets, core/logger.py). */
int d3il_object_poses(d3il_env* env, float* out, void* cu_stream);

/* Host-buffer variants: the same calls with HOST pointers; inputs are staged through pinned memory, copied to the
 * device, the kernels run, outputs are copied back and the call returns after synchronising (this is the
 * reference-facing end-to-end path: one Python call per env step, numpy in / numpy out). */
int d3il_reset_host(d3il_env* env, const float* ctx, const uint8_t* mask, float* obs);
int d3il_step_host(d3il_env* env, const float* action, float* obs, float* reward, uint8_t* done, float* info);
int d3il_robot_state_host(d3il_env* env, float* tcp);
int d3il_joint_state_host(d3il_env* env, float* j8);

/* Parity-test hooks: n physics ticks under the current controller / flat fp64 state of one env (layout shared with
 * oracle/d3il_oracle.c::d3o_get_state: qpos, qvel, qacc_warmstart, qfrc_bias[9], tcp[7], ik_q[7], des pose[7],
 * joint set-point q[7] qd[7], 8 scalars, 8 task words, then the scene's extra words, e.g. Aligning's target pose).  These synchronise the device. */
int d3il_substep(d3il_env* env, int n, void* cu_stream);
int d3il_get_state(d3il_env* env, double* out_host, int env_index);
int d3il_set_state(d3il_env* env, const double* in_host, int env_index);

/* Solver controls (Newton tolerance on the scaled gradient, iteration cap) and launch accounting for bench.py. */
int d3il_set_solver(d3il_env* env, double tolerance, int max_iterations);
long long d3il_kernel_launches(const d3il_env* env);
/* Profiling passes: when on, d3il_step brackets its two kernels (IK reference, env step) with CUDA events on the
 * launching stream and accumulates their device times (ms); d3il_step then synchronises, so keep it off when timing. */
int d3il_set_profiling(d3il_env* env, int on);
int d3il_get_profile(const d3il_env* env, double out_ms[2], long long* n_steps);

#ifdef __cplusplus
}
#endif
#endif
