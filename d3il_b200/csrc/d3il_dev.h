// d3il_dev.h — declarations shared by the two CUDA translation units of libd3il.so:
//   d3il_capi.cu        : C ABI, k_sched, k_ik (fp64 IK reference; compiled with the full register budget)
//   d3il_kernels_env.cu : k_env<3|4>, k_reset, k_robot_state, k_joint_state, k_object_poses.  k_env is capped at 120
//                         registers (__maxnreg__; -maxrregcount covers the rest of the unit): 16 env warps per SM, as one lock-step CTA
//                         of 16 or two of 8 (d3il_create picks per scene; see IK_THREADS below and DESIGN.md §2)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "d3il_model.h"

#ifndef G_LANES
#define G_LANES 32          // lanes cooperating on one env (32 = one warp per env; 16 = two envs per warp)
#endif
#ifndef ENVS_PER_CTA
#define ENVS_PER_CTA 8     // historical default of the diagnostic builds; the product picks the size at d3il_create (ENVS_PER_CTA_MAX)
#endif
#ifndef ENVS_PER_CTA_MAX
#define ENVS_PER_CTA_MAX 16  // one lock-step CTA of 16 envs per SM where the workspace allows (d3il_create picks the size, see there)
#endif
#define CTA_THREADS (G_LANES * ENVS_PER_CTA_MAX)
#ifndef IK_THREADS
#define IK_THREADS 256   // 16 k_ik blocks at 4096 envs.  Registers are allocated to a CTA in units of 4 warps, so ANY k_ik block costs
                         // >= 4 x 32 x 255 registers and evicts a k_env CTA from its SM until it leaves; an 8-warp block takes the
                         // whole register file of its SM (no k_env CTA beside it), a 4-warp block leaves room for one.  Either way 264 of
                         // the 296 k_env slots are free at t = 0; measured: 64 threads -10 %, 128 baseline, 256 +4 % (fewer SMs whose
                         // k_env CTA shares issue slots and L1 with the fp64 IK warps).
#endif
#define IK_FLAG_ENVS 32  // release flags are per k_ik WARP: a warp whose envs need the slow clipped-spectrum path does not hold back the others
struct DevIk {            // SoA views, [field][n]
  double* q;              // [7][n]
  float* des;             // [7][n]  des_pos(3), des_quat(4)
  float* jt;              // [21][n] last set-point: q_hi(7), q_lo(7), qd(7)
  int* valid;             // [n]
};
struct DevCtx {
  const Model* model;     // global copy, staged into shared memory per CTA
  Lay lay;
  float* state;           // [n][row]
  int row, n, ws_stride;
  int model_bytes;        // staged prefix of the Model (d3il_model_bytes), multiple of 128
  int lsg;                // lock-step sub-group size inside a CTA (0 = the whole CTA)
  int fpc;                // envs per FREE-RUNNING CTA (the head of the cost-sorted order, see k_env)
  int epc;                // envs (warps) per CTA: ENVS_PER_CTA unless the scene's workspace needs more shared memory per env
  DevIk ik;
  float* traj;            // [ticks][21][n]
  int* ik_flags;          // [n_ik_flags] ticks published by each k_ik warp (monotonic: launch number * 64 + tick + 1)
  int* launch_no;         // device-resident launch number, advanced by k_sched: no kernel argument changes from step to step, so a
                          // whole env step (and the policy in front of it) can be captured in a CUDA graph and replayed
  int n_ik_flags;
  int* perm;              // [n] env order of this step's k_env groups: most expensive envs (last step's Newton iterations) first
  float tol; int max_iter;
};

// host-side launchers of the kernels that live in d3il_kernels_env.cu
cudaError_t d3il_env_kernels_configure(size_t smem_bytes);
cudaError_t d3il_launch_env(const DevCtx& c, int maxdim, int n_free, int n_ticks, int gym, const float* action, float* obs, float* reward, uint8_t* done, float* info,
                            size_t smem_bytes, cudaStream_t s, bool programmatic);
void d3il_launch_reset(const DevCtx& c, const float* ctx, const uint8_t* mask, float* obs, size_t smem_bytes, cudaStream_t s);
void d3il_launch_robot_state(const DevCtx& c, float* tcp, cudaStream_t s);
void d3il_launch_joint_state(const DevCtx& c, float* j8, cudaStream_t s);
void d3il_launch_robot_kinematics(const DevCtx& c, float* out22, cudaStream_t s);
void d3il_launch_object_poses(const DevCtx& c, int nobj, float* out, cudaStream_t s);
int d3il_env_grid(const DevCtx& c, int n_free);

#ifdef D3IL_PHASE_TIMING
int d3il_debug_timeline_env(unsigned long long* out4096x4);     // k_env CTA records (the k_ik ones live in d3il_capi.cu)
int d3il_debug_phase_cycles_env(unsigned long long* out24);
#endif
