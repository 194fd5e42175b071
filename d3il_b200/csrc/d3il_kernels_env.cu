// d3il_kernels_env.cu — the warp-per-env kernels: k_env (Gym pre-step sampling + n physics ticks + post-step info),
// k_reset, k_robot_state, and their host launchers.  Compiled with -maxrregcount=120 (see d3il_dev.h).
#include <stdio.h>

#include "d3il_dev.h"

// ------------------------------------------------------------------------------------------------ kernels
#ifdef D3IL_PHASE_TIMING
static __device__ unsigned long long g_tl[4 * 4096];     // debug timeline: per block [t0, t1, smid, kind]
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
#define TL_BEGIN(kind, idx) unsigned long long tl_t0 = gtime(); const int tl_i = (idx)
#define TL_END(kind) do { if (threadIdx.x == 0 && tl_i < 4096) { g_tl[4 * tl_i] = tl_t0; g_tl[4 * tl_i + 1] = gtime(); g_tl[4 * tl_i + 2] = smid(); g_tl[4 * tl_i + 3] = kind; } } while (0)
#else
#define TL_BEGIN(kind, idx) ((void)0)
#define TL_END(kind) ((void)0)
#endif
__device__ __forceinline__ void stage_model(Model* sm, const Model* gm, int bytes) {
  const int4* src = (const int4*)gm; int4* dst = (int4*)sm;
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// IK reference generator: one thread per env (a5/a6).  mode: 1 = take the set-point from `action` (env step),
// 0 = keep the stored set-point (d3il_substep).  In joint-PD mode (after reset) the held set-point is replicated.
// Env step: one warp per env.  gym = 1: GymEnvWrapper.step semantics around the ticks; gym = 0: bare ticks (substep).
// 120 registers: 16 env warps (16 x 32 x 120 = 61 440 of the SM's 65 536 registers) as one lock-step CTA or two CTAs of 8.
// (__launch_bounds__ with a min-blocks hint would let ptxas go to 146 and override -maxrregcount; 128 measured the same as 120.)
template <int MD>
#ifndef D3IL_ENV_REGS
#define D3IL_ENV_REGS 120
#endif
__global__ void __maxnreg__(D3IL_ENV_REGS)
k_env(DevCtx c, int n_free, int n_ticks, int gym, const float* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ done, float* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TL_BEGIN(2, blockIdx.x);
  Model* sm = (Model*)smem_raw;
  stage_model(sm, c.model, c.model_bytes);
  const Model& m = *sm;
  const Lay& L = c.lay;
  const int warp = threadIdx.x / G_LANES;
  Cx cx; cx.lane = threadIdx.x % G_LANES;
  cx.mask = G_LANES == 32 ? 0xffffffffu : (((1u << G_LANES) - 1u) << (((threadIdx.x & 31) / G_LANES) * G_LANES));
  // CTA composition: the first n_free CTAs are FREE-RUNNING: they carry the c.fpc most expensive envs of the cost-sorted order
  // each, their warps never meet at a CTA barrier (every env advances at its own pace: its time is the sum of its OWN Newton
  // passes, not of the per-tick maxima over the CTA), and they start first.  The others carry c.epc envs in lock step (phase
  // barriers, CTA-uniform Newton loop: shared instruction fetch).  Spare warps of a free CTA exit here.
  const int flag_base = *(volatile const int*)c.launch_no * 64;      // written by k_sched, which completed before k_ik (and hence this kernel) could start
  const int freecta = (int)blockIdx.x < n_free;
  const int cnt = freecta ? c.fpc : c.epc;
  const int pos0 = freecta ? (int)blockIdx.x * c.fpc : n_free * c.fpc + ((int)blockIdx.x - n_free) * c.epc;
  if (warp >= cnt) return;
  cx.cta_threads = freecta ? G_LANES : cnt * G_LANES;          // cta_sync / cta_any are warp-local when this is one group
  cx.bar_id = 1;
  if (!freecta && c.lsg > 0 && c.lsg < cnt) {                  // lock-step SUB-groups of c.lsg envs, each on its own named barrier
    const int sg = warp / c.lsg, first = sg * c.lsg, members = cnt - first < c.lsg ? cnt - first : c.lsg;
    cx.bar_id = 1 + sg; cx.cta_threads = members * G_LANES;
  }
  // Groups past the end of the batch shadow the last env (same inputs, same control flow, identical outputs) so that
  // every thread of the CTA reaches the phase barriers inside physics_tick.
  const int e_raw = pos0 + warp;                               // position in this step's cost-sorted order
  const int e = c.perm[e_raw < c.n ? e_raw : c.n - 1];
  float* w = (float*)(smem_raw + c.model_bytes) + (size_t)warp * c.ws_stride;
  float* row = c.state + (size_t)e * c.row;
  for (int i = cx.lane; i < L.n_state; i += G_LANES) w[i] = row[i];
  __syncwarp(cx.mask);
  if (gym) env_prestep<G_LANES>(cx, m, L, w, action + (size_t)e * m.act_dim, obs + (size_t)e * m.obs_dim, reward + e, done + e);
  for (int t = 0; t < n_ticks; t++) {
    // acquire tick t of the IK reference (k_ik may still be running: programmatic dependent launch)
    PHASE_T0();
    if (cx.lane == 0) {              // one lane per group waits for the k_ik block that owns its env
      const int want = flag_base + t + 1;
      while (*(volatile int*)(c.ik_flags + e / IK_FLAG_ENVS) < want) __nanosleep(100);
      __threadfence();
    }
    PHASE(16);
    cta_sync<true>(cx);
    const float* tr = c.traj + (size_t)t * 21 * c.n + e;
    for (int k = cx.lane; k < 21; k += G_LANES) w[L.jt + k] = __ldcg(tr + (size_t)k * c.n);
    __syncwarp(cx.mask);
    physics_tick<G_LANES, true, MD>(cx, m, L, w, w + L.jt, w + L.jt + 7, w + L.jt + 14, c.tol, c.max_iter);
  }
  if (gym) env_poststep<G_LANES>(cx, m, L, w, info + (size_t)e * m.info_dim);
  if (e_raw < c.n) for (int i = cx.lane; i < L.n_state; i += G_LANES) row[i] = w[i];
  TL_END(2);
}

__global__ void __launch_bounds__(CTA_THREADS, (CTA_THREADS > 256 ? 1 : 2))
k_reset(DevCtx c, const float* __restrict__ ctx, const uint8_t* __restrict__ mask, float* __restrict__ obs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x / G_LANES;
  const int e = blockIdx.x * c.epc + warp;
  TL_BEGIN(4, 1024 + blockIdx.x);
  // masked reset (auto-reset of the ~1 % of envs that finished): CTAs without a masked env leave before staging the tables
  if (!__syncthreads_or(e < c.n && (!mask || mask[e]))) { TL_END(5); return; }
  Model* sm = (Model*)smem_raw;
  stage_model(sm, c.model, c.model_bytes);
  const Model& m = *sm;
  const Lay& L = c.lay;
  Cx cx; cx.lane = threadIdx.x % G_LANES;
  cx.mask = G_LANES == 32 ? 0xffffffffu : (((1u << G_LANES) - 1u) << (((threadIdx.x & 31) / G_LANES) * G_LANES));
  if (e >= c.n) return;
  cx.cta_threads = c.epc * G_LANES; cx.bar_id = 1;
  if (mask && !mask[e]) return;
  float* w = (float*)(smem_raw + c.model_bytes) + (size_t)warp * c.ws_stride;
  env_reset<G_LANES>(cx, m, L, w, (ctx && m.ctx_dim > 0) ? ctx + (size_t)e * m.ctx_dim : nullptr, c.tol, c.max_iter);
  float* row = c.state + (size_t)e * c.row;
  for (int i = cx.lane; i < L.n_state; i += G_LANES) row[i] = w[i];
  const int n = c.n;
  for (int k = cx.lane; k < 7; k += G_LANES) {
    c.ik.q[k * n + e] = 0; c.ik.des[k * n + e] = 0;
    c.ik.jt[k * n + e] = (float)m.ctrl[D3C_INIT_QPOS + k]; c.ik.jt[(7 + k) * n + e] = 0; c.ik.jt[(14 + k) * n + e] = 0;
  }
  if (cx.lane == 0) {
    c.ik.valid[e] = 0;
    if (obs) task_obs(m, L, w, obs + (size_t)e * m.obs_dim);
  }
  TL_END(4);
}

__global__ void k_robot_state(DevCtx c, float* __restrict__ tcp) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= c.n) return;
  const float* row = c.state + (size_t)e * c.row;
  for (int k = 0; k < 3; k++) tcp[(size_t)e * 3 + k] = row[c.lay.tcp + k];
}


// ------------------------------------------------------------------------------------------------ host launchers
int d3il_env_grid(const DevCtx& c, int n_free) { return n_free + (c.n - n_free * c.fpc + c.epc - 1) / c.epc; }

cudaError_t d3il_env_kernels_configure(size_t smem_bytes) {
  cudaError_t e;
  // The dynamic shared-memory limit is an attribute of the FUNCTION (per device), shared by every handle of the process:
  // several scenes side by side (MixedBatch) need the largest request, so the limit only ever grows.
  static size_t configured[64] = {0};
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if (dev >= 0 && dev < 64) { if (smem_bytes <= configured[dev]) return cudaSuccess; configured[dev] = smem_bytes; }
  if ((e = cudaFuncSetAttribute(k_env<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_env<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_env<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)) != cudaSuccess) return e;
  // an SM keeps the L1/shared carve-out of whatever is resident: every kernel of the step asks for the maximum shared
  // carve-out so k_env CTAs can join SMs that still run a k_ik warp
  if ((e = cudaFuncSetAttribute(k_env<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_reset, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

cudaError_t d3il_launch_env(const DevCtx& c, int maxdim, int n_free, int n_ticks, int gym, const float* action, float* obs, float* reward, uint8_t* done, float* info,
                            size_t smem_bytes, cudaStream_t s, bool programmatic) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(d3il_env_grid(c, n_free)); cfg.blockDim = dim3(c.epc * G_LANES); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = programmatic ? 1 : 0;
  if (maxdim == 4) return cudaLaunchKernelEx(&cfg, k_env<4>, c, n_free, n_ticks, gym, action, obs, reward, done, info);
  return cudaLaunchKernelEx(&cfg, k_env<3>, c, n_free, n_ticks, gym, action, obs, reward, done, info);
}

void d3il_launch_reset(const DevCtx& c, const float* ctx, const uint8_t* mask, float* obs, size_t smem_bytes, cudaStream_t s) {
  k_reset<<<(c.n + c.epc - 1) / c.epc, c.epc * G_LANES, smem_bytes, s>>>(c, ctx, mask, obs);
}

void d3il_launch_robot_state(const DevCtx& c, float* tcp, cudaStream_t s) { k_robot_state<<<(c.n + 127) / 128, 128, 0, s>>>(c, tcp); }

// CubeStacking_Env.robot_state (stacking.py:218-226): 7 joint positions + gripper width (sum of the two finger joints)
__global__ void k_joint_state(DevCtx c, float* __restrict__ j8) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= c.n) return;
  const float* row = c.state + (size_t)e * c.row;
  for (int k = 0; k < 7; k++) j8[(size_t)e * 8 + k] = (float)((double)row[c.lay.qpos + k] + (double)row[c.lay.qlo + k]);
  j8[(size_t)e * 8 + 7] = row[c.lay.qpos + 7] + row[c.lay.qpos + 8];
}
// MjScene._get_obj_pos_and_quat (MjScene.py:233-247) for every free object: qpos words (x, y, z, qw, qx, qy, qz)
__global__ void k_object_poses(DevCtx c, int nobj, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n * nobj * 7) return;
  int e = i / (nobj * 7), k = i - e * nobj * 7;
  out[i] = c.state[(size_t)e * c.row + c.lay.qpos + D3_NROB + k];
}
void d3il_launch_object_poses(const DevCtx& c, int nobj, float* out, cudaStream_t s) {
  if (nobj > 0) k_object_poses<<<(c.n * nobj * 7 + 255) / 256, 256, 0, s>>>(c, nobj, out);
}
__global__ void k_robot_kinematics(DevCtx c, float* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= c.n) return;
  const float* row = c.state + (size_t)e * c.row;
  float* o = out + (size_t)e * 22;
  for (int k = 0; k < 7; k++) o[k] = row[c.lay.tcp + k];
  for (int k = 0; k < 7; k++) { o[7 + k] = (float)((double)row[c.lay.qpos + k] + (double)row[c.lay.qlo + k]); o[14 + k] = row[c.lay.qvel + k]; }
  o[21] = row[c.lay.qpos + 7] + row[c.lay.qpos + 8];
}
void d3il_launch_robot_kinematics(const DevCtx& c, float* out, cudaStream_t s) { k_robot_kinematics<<<(c.n + 127) / 128, 128, 0, s>>>(c, out); }
void d3il_launch_joint_state(const DevCtx& c, float* j8, cudaStream_t s) { k_joint_state<<<(c.n + 127) / 128, 128, 0, s>>>(c, j8); }

#ifdef D3IL_PHASE_TIMING
int d3il_debug_timeline_env(unsigned long long* out) { return cudaMemcpyFromSymbol(out, g_tl, sizeof(unsigned long long) * 4 * 4096) == cudaSuccess ? 0 : -2; }
extern "C" int d3il_debug_iter_hist(unsigned long long* out40) { return cudaMemcpyFromSymbol(out40, g_iter_hist, sizeof(unsigned long long) * 40) == cudaSuccess ? 0 : -2; }
// per-CTA Newton statistics of the launches since the last clear (out == nullptr or clear != 0 zeroes them)
extern "C" int d3il_debug_cta_stat(unsigned* out4096x4, int clear) {
  if (out4096x4 && cudaMemcpyFromSymbol(out4096x4, g_cta_stat, sizeof(unsigned) * 4 * 4096) != cudaSuccess) return -2;
  if (clear || !out4096x4) { static unsigned z[4 * 4096]; if (cudaMemcpyToSymbol(g_cta_stat, z, sizeof(z)) != cudaSuccess) return -2; }
  return 0;
}
int d3il_debug_phase_cycles_env(unsigned long long* out24) { return cudaMemcpyFromSymbol(out24, g_phase_cycles, sizeof(unsigned long long) * 40) == cudaSuccess ? 0 : -2; }
#endif
