// d3il_capi.cu — CUDA kernels (sm_100a) + the C ABI of include/d3il.h.
//
// Data layout in HBM (DESIGN.md "memory"):
//   state  : [n_envs][row] fp32, env-major rows padded to 32 floats (128 B) — one warp owns one env and streams its
//            row in/out with fully coalesced 128 B transactions; between the two touches the state lives in shared memory
//            for all 35 physics ticks of the env step.
//   ik_*   : IK-controller state, field-major SoA [field][n_envs] (the IK kernel is one THREAD per env).
//   traj   : [n_ticks][21][n_envs] fp32 joint set-points (q_hi, q_lo, qd) written by the IK kernel (coalesced along
//            envs), read by the physics kernel.
// Kernels: k_ik (thread per env: the 3 x n_ticks damped-least-squares iterations of the open-loop IK reference, fp64),
//          k_env (warp per env: Gym pre-step sampling, n_ticks physics ticks, post-step info), k_reset (warp per env).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "../../include/d3il.h"
#ifndef IK_THREADS
#define IK_THREADS 256             // see d3il_dev.h
#endif
#define IK_JSTRIDE IK_THREADS      // k_ik keeps the Jacobians of its block in shared memory as J[k][thread]
#include "d3il_dev.h"

static thread_local std::string g_err;
extern "C" const char* d3il_last_error(void) { return g_err.c_str(); }
#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return -2; } \
  } while (0)

struct d3il_env {
  Model m; Lay L;
  DevCtx d;
  int device, n, max_ticks, n_ik_blocks, n_ik_flags, n_free;
  long long launches;
  size_t smem_bytes;
  // pinned + device staging for the *_host calls
  float *h_in, *h_out, *d_in, *d_out; uint8_t *h_mask, *d_mask; size_t in_floats, out_floats;
  cudaStream_t own_stream;
  int profiling; cudaEvent_t ev[3]; double prof_ms[2]; long long prof_n;
};

// ------------------------------------------------------------------------------------------------ kernels (scheduler, IK reference)
#ifdef D3IL_PHASE_TIMING
static __device__ unsigned long long g_tl[4 * 4096];     // debug timeline of the k_ik blocks (k_env's lives in d3il_kernels_env.cu)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
#define TL_BEGIN(kind, idx) unsigned long long tl_t0 = gtime(); const int tl_i = (idx)
#define TL_END(kind) do { if (threadIdx.x == 0 && tl_i < 4096) { if (kind == 3) { g_tl[4 * 4094] = g_tl[4 * tl_i]; g_tl[4 * 4094 + 1] = g_tl[4 * tl_i + 1]; g_tl[4 * 4094 + 3] = 6; } g_tl[4 * tl_i] = tl_t0; g_tl[4 * tl_i + 1] = gtime(); g_tl[4 * tl_i + 2] = smid(); g_tl[4 * tl_i + 3] = kind; } } while (0)
#else
#define TL_BEGIN(kind, idx) ((void)0)
#define TL_END(kind) ((void)0)
#endif
// Cost-aware scheduling.  Env steps differ in cost by ~3x (Newton iterations while a box is being pushed or is
// rocking), CTAs are as slow as their slowest env (CTA-uniform loops, phase barriers) and the kernel as slow as its
// last CTA.  So each step the envs are bucket-sorted by the Newton iterations of their previous step: expensive envs
// share CTAs and are dispatched first.  The order only affects scheduling, never results (envs are independent).
__global__ void __launch_bounds__(1024) k_sched(DevCtx c) {
  TL_BEGIN(3, 4095);
  __shared__ int hist[256], start[256];
  // advance the launch number (the base of this step's IK release flags); when it wraps, the monotonic flags restart from zero
  // — nothing else runs on the stream while k_sched does
  {
    const int id = (*c.launch_no + 1) & 0xffffff;
    __syncthreads();
    if (threadIdx.x == 0) *c.launch_no = id;
    if (id == 0) for (int i = threadIdx.x; i < c.n_ik_flags; i += blockDim.x) c.ik_flags[i] = 0;
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int off = c.lay.misc + ST_COST_ITERS;
  for (int e = threadIdx.x; e < c.n; e += blockDim.x) {
    const float* row = c.state + (size_t)e * c.row;
    int b = (int)(row[off] + 4.f * row[off + 1] + (row[off + 3] > 0.f ? 16.f : 0.f)) >> 1;      // Newton iterations, ticks with a coupling contact, contact imminent
    atomicAdd(&hist[b > 255 ? 255 : b], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) { int acc = 0; for (int b = 255; b >= 0; b--) { start[b] = acc; acc += hist[b]; } }
  __syncthreads();
  for (int e = threadIdx.x; e < c.n; e += blockDim.x) {
    const float* row = c.state + (size_t)e * c.row;
    int b = (int)(row[off] + 4.f * row[off + 1] + (row[off + 3] > 0.f ? 16.f : 0.f)) >> 1;      // Newton iterations, ticks with a coupling contact, contact imminent
    c.perm[atomicAdd(&start[b > 255 ? 255 : b], 1)] = e;
  }
  TL_END(3);
}

__global__ void __launch_bounds__(IK_THREADS) k_ik(DevCtx c, const float* __restrict__ action, int n_ticks, int use_action, int ctrl_kind, int act_dim) {
  const int flag_base = *(volatile const int*)c.launch_no * 64;
  // Programmatic dependent launch: let the env-step kernel (next in the stream) start while this one is still running.
  // It consumes our set-points tick by tick through the release flags below; we never wait on anything, and we are
  // already resident when it is allowed to launch, so the hand-off cannot deadlock.
  asm volatile("griddepcontrol.launch_dependents;");
  TL_BEGIN(1, 2048 + blockIdx.x);
  extern __shared__ __align__(16) double ik_smem[];      // IK_SMEM_BYTES: controller table (as doubles), Jacobians J[k][thread], per-warp cooperative scratch
  double* sctrl = ik_smem;
  double* sJ = sctrl + D3_CTRL_W;
  double* sA = sJ + 42 * IK_THREADS;
  double* coop_all = sA + 27 * IK_THREADS;
  double* coop = coop_all + 160 * (threadIdx.x / 32);
  Cx cx; cx.lane = threadIdx.x & 31; cx.mask = 0xffffffffu; cx.cta_threads = IK_THREADS; cx.bar_id = 1;
  for (int i = threadIdx.x; i < D3_CTRL_W; i += blockDim.x) sctrl[i] = (double)c.model->ctrl[i];
  __syncthreads();
  const int e_raw = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = e_raw < c.n;               // threads past the batch shadow the last env (they must reach the barriers)
  const int e = live ? e_raw : c.n - 1;
  const int n = c.n;
  const float* row = c.state + (size_t)e * c.row;
  IkState s;
  for (int k = 0; k < 7; k++) { s.q[k] = c.ik.q[k * n + e]; s.jt_q[k] = c.ik.jt[k * n + e]; s.jt_qlo[k] = c.ik.jt[(7 + k) * n + e]; s.jt_qd[k] = c.ik.jt[(14 + k) * n + e]; }
  s.valid = c.ik.valid[e];
  int cart = use_action ? 1 : (row[c.lay.misc + ST_CTRL_MODE] == 1.f);
  // an unusable action (NaN / inf / zero quaternion) keeps the previous set-point; k_env's pre-step raises the status bit
  const bool aok = use_action && action_ok(action + (size_t)e * act_dim, act_dim, ctrl_kind);
  if (use_action && ctrl_kind == 1) {
    cart = 0;
    if (aok) {
    // joint-space action (Stacking): the set-point IS the action, held for the whole env step (Controller.py:99-127, zero
    // desired velocity); no IK.  The gripper command action[7] is consumed by k_env's pre-step.
    const float* a = action + (size_t)e * act_dim;
    for (int k = 0; k < 7; k++) { s.jt_q[k] = a[k]; s.jt_qlo[k] = 0; s.jt_qd[k] = 0; }
    }
  } else if (aok) {
    const float* a = action + (size_t)e * act_dim;
    float nq = rsqrtf(a[3] * a[3] + a[4] * a[4] + a[5] * a[5] + a[6] * a[6]);
    s.des_pos[0] = a[0]; s.des_pos[1] = a[1]; s.des_pos[2] = a[2];
    for (int k = 0; k < 4; k++) s.des_quat[k] = a[3 + k] * nq;
    if (live) for (int k = 0; k < 7; k++) c.ik.des[k * n + e] = k < 3 ? s.des_pos[k] : s.des_quat[k - 3];
  } else {
    for (int k = 0; k < 3; k++) s.des_pos[k] = c.ik.des[k * n + e];
    for (int k = 0; k < 4; k++) s.des_quat[k] = c.ik.des[(3 + k) * n + e];
    // held set-point that was never installed (bad action on the first step after a reset): stay under the joint PD hold
    if (use_action && s.des_quat[0] == 0.f && s.des_quat[1] == 0.f && s.des_quat[2] == 0.f && s.des_quat[3] == 0.f) cart = 0;
  }
  if (cart && !s.valid) {                      // IKControllers.py:168-169: old_q NaN -> measured joints
    for (int k = 0; k < 7; k++) s.q[k] = (double)row[c.lay.qpos + k] + (double)row[c.lay.qlo + k];
    s.valid = 1;
  }
  double V[42], sn[7], cs[7]; int vwarm = 0;   // eigenbasis and joint sines/cosines carried across the IK iterations of this launch
  for (int t = 0; t < n_ticks; t++) {
    ik_tick<32>(cx, sctrl, s, cart && live, V, &vwarm, sn, cs, sJ + threadIdx.x, sA + threadIdx.x, coop);
    if (live) {
      float* tr = c.traj + (size_t)t * 21 * n + e;
      for (int k = 0; k < 7; k++) { tr[k * n] = s.jt_q[k]; tr[(7 + k) * n] = s.jt_qlo[k]; tr[(14 + k) * n] = s.jt_qd[k]; }
    }
    // publish tick t of this block's envs
    __threadfence();
    __syncwarp();
    if (cx.lane == 0) *(volatile int*)(c.ik_flags + (blockIdx.x * blockDim.x + threadIdx.x) / IK_FLAG_ENVS) = flag_base + t + 1;
  }
  if (live) {
    for (int k = 0; k < 7; k++) { c.ik.q[k * n + e] = s.q[k]; c.ik.jt[k * n + e] = s.jt_q[k]; c.ik.jt[(7 + k) * n + e] = s.jt_qlo[k]; c.ik.jt[(14 + k) * n + e] = s.jt_qd[k]; }
    c.ik.valid[e] = s.valid;
  }
  TL_END(1);
}

#define IK_SMEM_BYTES ((int)sizeof(double) * (D3_CTRL_W + (42 + 27) * IK_THREADS + 160 * (IK_THREADS / 32)))

// ------------------------------------------------------------------------------------------------ C ABI
static int create_impl(d3il_env* h, const void* blob, size_t nbytes, int n_envs, int device);

extern "C" int d3il_create(d3il_env** out, const void* blob, size_t nbytes, int n_envs, int device) {
  if (!out || !blob || n_envs <= 0) { g_err = "d3il_create: bad arguments"; return -1; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { g_err = "d3il_create: no usable CUDA device (this library has no CPU path)"; return -3; }
  d3il_env* h = new (std::nothrow) d3il_env();
  if (!h) { g_err = "out of memory"; return -1; }
  memset(&h->d, 0, sizeof(h->d));
  h->device = device;
  const int rc = create_impl(h, blob, nbytes, n_envs, device);
  if (rc != 0) { const std::string keep = g_err; d3il_destroy(h); g_err = keep; return rc; }      // every allocation made so far is released
  *out = h;
  return 0;
}

static int create_impl(d3il_env* h, const void* blob, size_t nbytes, int n_envs, int device) {
  std::string err;
  if (!d3il_build_model(blob, nbytes, h->m, h->L, err)) { g_err = "d3il_create: " + err; return -1; }
  if (!((h->m.ctrl_kind == 0 && h->m.act_dim == 7) || (h->m.ctrl_kind == 1 && h->m.act_dim == 8))) { g_err = "d3il_create: unsupported action layout"; return -1; }
  h->device = device; h->n = n_envs; h->launches = 0; h->max_ticks = h->m.n_substeps > 64 ? h->m.n_substeps : 64;
  CK(cudaSetDevice(device));
  DevCtx& d = h->d;
  d.lay = h->L; d.n = n_envs; d.row = (h->L.n_state + 31) & ~31; d.ws_stride = (h->L.total + 3) & ~3;       // 16-byte alignment is all the workspace needs (warps never share an instruction across workspaces)
  d.tol = 1e-6f; d.max_iter = 32;      // the cap is not reached on the benchmark workloads (status bit 8 reports it if it is; 12 was hit by 0.4 % of the Pushing env steps, 24 by a few closing-gripper steps of Stacking); the loop is CTA-uniform with early exit, so a high cap costs nothing
  Model* dm = nullptr;
  CK(cudaMalloc(&dm, sizeof(Model)));
  CK(cudaMemcpy(dm, &h->m, sizeof(Model), cudaMemcpyHostToDevice));
  d.model = dm;
  CK(cudaMalloc(&d.state, (size_t)n_envs * d.row * sizeof(float)));
  CK(cudaMemset(d.state, 0, (size_t)n_envs * d.row * sizeof(float)));
  CK(cudaMalloc(&d.ik.q, (size_t)7 * n_envs * sizeof(double)));
  CK(cudaMalloc(&d.ik.des, (size_t)7 * n_envs * sizeof(float)));
  CK(cudaMalloc(&d.ik.jt, (size_t)21 * n_envs * sizeof(float)));
  CK(cudaMalloc(&d.ik.valid, (size_t)n_envs * sizeof(int)));
  CK(cudaMemset(d.ik.q, 0, (size_t)7 * n_envs * sizeof(double)));
  CK(cudaMemset(d.ik.des, 0, (size_t)7 * n_envs * sizeof(float)));
  CK(cudaMemset(d.ik.jt, 0, (size_t)21 * n_envs * sizeof(float)));
  CK(cudaMemset(d.ik.valid, 0, (size_t)n_envs * sizeof(int)));
  CK(cudaMalloc(&d.traj, (size_t)h->max_ticks * 21 * n_envs * sizeof(float)));
  h->n_ik_blocks = (n_envs + IK_THREADS - 1) / IK_THREADS;
  h->n_ik_flags = h->n_ik_blocks * (IK_THREADS / IK_FLAG_ENVS);
  CK(cudaMalloc(&d.ik_flags, (size_t)h->n_ik_flags * sizeof(int)));
  CK(cudaMalloc(&d.perm, (size_t)n_envs * sizeof(int)));
  CK(cudaMemset(d.ik_flags, 0, (size_t)h->n_ik_flags * sizeof(int)));
  d.n_ik_flags = h->n_ik_flags;
  CK(cudaMalloc(&d.launch_no, sizeof(int)));
  CK(cudaMemset(d.launch_no, 0, sizeof(int)));
  // envs per CTA: ENVS_PER_CTA (two CTAs per SM for the small scenes), fewer when the per-env workspace is large (Sorting-4/6)
  const size_t model_bytes = d3il_model_bytes(h->m), env_bytes = (size_t)d.ws_stride * sizeof(float);
  d.model_bytes = (int)model_bytes;
  // CTA size.  Measured on Pushing (4096 envs): one lock-step CTA of 16 envs per SM 676 k env-steps/s, two CTAs of 8 620 k, lock-step
  // sub-groups of 4 inside them 532 k, one CTA of 12 564 k; Sorting-4: 8 envs 253 k, 9 envs 227 k.  So: (1) as many resident env
  // warps per SM as shared memory (CTA bytes + 1 KB reserve out of 228 KB) and registers (120 per thread, allocated in units of 4
  // warps: two CTAs only up to 8 warps each) allow, counted in whole multiples of 4 - a 9th warp only unbalances the four
  // schedulers; (2) among equals ONE big lock-step group, because every instruction fetch then serves all its warps (the kernel
  // is 490 KB of mostly straight-line code); (3) then the smaller group (less waiting for the slowest env).
  d.epc = 1;
  { int best = -1;
    for (int e = ENVS_PER_CTA_MAX; e >= 1; e--) {
      const size_t bytes = model_bytes + (size_t)e * env_bytes;
      if (bytes > 227 * 1024) continue;
      const int by_smem = (int)((228 * 1024) / (bytes + 1024)), by_regs = ((e + 3) / 4) * 4 <= 8 ? 2 : 1;
      const int ctas = by_smem >= 2 && by_regs >= 2 ? 2 : 1, warps = ctas * e;
      const int score = (warps >= 8 ? (warps & ~3) : warps) * 64 + (ctas == 1 ? 32 : 0) + (ENVS_PER_CTA_MAX - e);
      if (score > best) { best = score; d.epc = e; }
    } }
#ifdef D3IL_DIAG
  if (const char* ev = getenv("D3IL_EPC")) { const int e = atoi(ev); if (e >= 1 && e <= ENVS_PER_CTA_MAX && model_bytes + (size_t)e * env_bytes <= 227 * 1024) d.epc = e; }
#endif
  h->smem_bytes = model_bytes + (size_t)d.epc * env_bytes;
#ifdef D3IL_DIAG
  if (const char* ev = getenv("D3IL_SMEM_PAD_KB")) {      // diagnosis build only: pad the request so fewer CTAs share an SM
    const size_t padded = h->smem_bytes + (size_t)atoi(ev) * 1024;
    if (padded <= 227 * 1024) h->smem_bytes = padded;
  }
#endif
  if (h->smem_bytes > 227 * 1024) { g_err = "d3il_create: scene workspace does not fit in shared memory"; return -1; }
  CK(d3il_env_kernels_configure(h->smem_bytes));
  CK(cudaFuncSetAttribute(k_ik, cudaFuncAttributeMaxDynamicSharedMemorySize, IK_SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_ik, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(k_sched, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // the head of each step's cost-sorted order runs in free-running CTAs (see k_sched / k_env): fpc envs each (one warp per SM
  // sub-partition), enough of them for the ~1.5 % of envs that are in a contact-rich phase at any time, rounded so that
  // the lock-step CTAs behind them are all full
  d.fpc = 4; h->n_free = 0; d.lsg = 0;
  if (G_LANES == 32 && n_envs >= 512) {
    int want = (n_envs / 64 + d.fpc - 1) / d.fpc;                        // ~1.5 % of the envs
    if (want > 64) want = 64;
    h->n_free = want;
    for (int k = 0; k < d.epc; k++) if ((n_envs - (want + k) * d.fpc) % d.epc == 0) { h->n_free = want + k; break; }
  }
#ifdef D3IL_DIAG
  if (const char* ev = getenv("D3IL_LSG")) d.lsg = atoi(ev);
  if (const char* ev = getenv("D3IL_FPC")) { d.fpc = atoi(ev); if (d.fpc < 1 || d.fpc > d.epc) d.fpc = 1; }
  if (const char* ev = getenv("D3IL_N_FREE")) { h->n_free = atoi(ev); if (h->n_free < 0 || h->n_free * d.fpc > n_envs / 2 || G_LANES != 32) h->n_free = 0; }
#endif
  // staging for the host-buffer entry points
  const Model& m = h->m;
  h->in_floats = (size_t)n_envs * (m.act_dim > m.ctx_dim ? m.act_dim : m.ctx_dim);
  h->out_floats = (size_t)n_envs * (m.obs_dim + 1 + m.info_dim + 8) + (n_envs + 3) / 4 + 8;
  CK(cudaMallocHost(&h->h_in, h->in_floats * sizeof(float)));
  CK(cudaMallocHost(&h->h_out, h->out_floats * sizeof(float)));
  CK(cudaMallocHost(&h->h_mask, n_envs));
  CK(cudaMalloc(&h->d_in, h->in_floats * sizeof(float)));
  CK(cudaMalloc(&h->d_out, h->out_floats * sizeof(float)));
  CK(cudaMalloc(&h->d_mask, n_envs));
  CK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->profiling = 0; h->prof_ms[0] = h->prof_ms[1] = 0; h->prof_n = 0;
  for (int i = 0; i < 3; i++) CK(cudaEventCreate(&h->ev[i]));
  return 0;
}

extern "C" void d3il_destroy(d3il_env* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree((void*)h->d.model); cudaFree(h->d.state); cudaFree(h->d.ik.q); cudaFree(h->d.ik.des); cudaFree(h->d.ik.jt); cudaFree(h->d.ik.valid);
  cudaFree(h->d.traj); cudaFree(h->d.ik_flags); cudaFree(h->d.launch_no); cudaFree(h->d.perm); cudaFreeHost(h->h_in); cudaFreeHost(h->h_out); cudaFreeHost(h->h_mask); cudaFree(h->d_in); cudaFree(h->d_out); cudaFree(h->d_mask);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  for (int i = 0; i < 3; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  delete h;
}

extern "C" int d3il_dims(const d3il_env* h, int32_t out[D3IL_NDIMS]) {
  if (!h || !out) { g_err = "d3il_dims: bad arguments"; return -1; }
  out[D3IL_DIM_OBS] = h->m.obs_dim; out[D3IL_DIM_ACT] = h->m.act_dim; out[D3IL_DIM_CTX] = h->m.ctx_dim; out[D3IL_DIM_INFO] = h->m.info_dim;
  out[D3IL_DIM_STATE] = d3il_state_dim(h->m); out[D3IL_DIM_NENVS] = h->n; out[D3IL_DIM_SUBSTEPS] = h->m.n_substeps; out[D3IL_DIM_MAXSTEPS] = h->m.max_steps;
  return 0;
}

extern "C" int d3il_set_solver(d3il_env* h, double tol, int max_iter) {
  if (!h || tol <= 0 || max_iter < 1) { g_err = "d3il_set_solver: bad arguments"; return -1; }
  h->d.tol = (float)tol; h->d.max_iter = max_iter;
  return 0;
}
extern "C" long long d3il_kernel_launches(const d3il_env* h) { return h ? h->launches : 0; }
extern "C" int d3il_set_profiling(d3il_env* h, int on) {
  if (!h) { g_err = "d3il_set_profiling: null handle"; return -1; }
  h->profiling = on; h->prof_ms[0] = h->prof_ms[1] = 0; h->prof_n = 0;
  return 0;
}
extern "C" int d3il_get_profile(const d3il_env* h, double out_ms[2], long long* n_steps) {
  if (!h || !out_ms || !n_steps) { g_err = "d3il_get_profile: bad arguments"; return -1; }
  out_ms[0] = h->prof_ms[0]; out_ms[1] = h->prof_ms[1]; *n_steps = h->prof_n;
  return 0;
}


// Scheduler + IK reference + env step: three launches on one stream, the last one with programmatic stream serialization
// so that it overlaps k_ik (per-tick hand-off through ik_flags).  If the driver serialises them anyway the result is the same.
static cudaError_t launch_step(d3il_env* h, cudaStream_t s, const float* action, int n_ticks, int gym, float* obs, float* reward, uint8_t* done, float* info, cudaEvent_t after_ik = nullptr) {
#ifdef D3IL_DIAG
  static const bool no_pdl = getenv("D3IL_NO_PDL") != nullptr;      // diagnosis build only: serialise the kernels
#else
  const bool no_pdl = false;
#endif
  k_sched<<<1, 1024, 0, s>>>(h->d);
  k_ik<<<h->n_ik_blocks, IK_THREADS, IK_SMEM_BYTES, s>>>(h->d, action, n_ticks, gym, h->m.ctrl_kind, h->m.act_dim);
  h->launches += 3;
  if (after_ik) cudaEventRecord(after_ik, s);        // completes when k_ik has finished (k_env may already be running: PDL)
  return d3il_launch_env(h->d, h->m.maxdim, h->n_free, n_ticks, gym, action, obs, reward, done, info, h->smem_bytes, s, !no_pdl);
}

extern "C" int d3il_reset(d3il_env* h, const float* ctx, const uint8_t* mask, float* obs, void* stream) {
  if (!h) { g_err = "d3il_reset: null handle"; return -1; }
  if (h->m.ctx_dim > 0 && !ctx) { g_err = "d3il_reset: this scene needs a context per env"; return -1; }
  CK(cudaSetDevice(h->device));
  d3il_launch_reset(h->d, ctx, mask, obs, h->smem_bytes, (cudaStream_t)stream);
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int d3il_step(d3il_env* h, const float* action, float* obs, float* reward, uint8_t* done, float* info, void* stream) {
  if (!h || !action || !obs || !reward || !done || !info) { g_err = "d3il_step: null argument"; return -1; }
  CK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (h->profiling) CK(cudaEventRecord(h->ev[0], s));
  CK(launch_step(h, s, action, h->m.n_substeps, 1, obs, reward, done, info, h->profiling ? h->ev[1] : nullptr));
  CK(cudaGetLastError());
  if (h->profiling) {
    // per-kernel device time on the launching stream (bench.py roofline); synchronises, so only for profiling passes
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaEventSynchronize(h->ev[2]));
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, h->ev[0], h->ev[1]));
    CK(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
    h->prof_ms[0] += a; h->prof_ms[1] += b; h->prof_n += 1;
  }
  return 0;
}

extern "C" int d3il_substep(d3il_env* h, int n, void* stream) {
  if (!h || n < 0) { g_err = "d3il_substep: bad arguments"; return -1; }
  CK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  while (n > 0) {
    int k = n < h->max_ticks ? n : h->max_ticks;
    CK(launch_step(h, s, nullptr, k, 0, nullptr, nullptr, nullptr, nullptr));
    n -= k;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int d3il_robot_state(d3il_env* h, float* tcp, void* stream) {
  if (!h || !tcp) { g_err = "d3il_robot_state: null argument"; return -1; }
  CK(cudaSetDevice(h->device));
  d3il_launch_robot_state(h->d, tcp, (cudaStream_t)stream);
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int d3il_joint_state(d3il_env* h, float* j8, void* stream) {
  if (!h || !j8) { g_err = "d3il_joint_state: null argument"; return -1; }
  CK(cudaSetDevice(h->device));
  d3il_launch_joint_state(h->d, j8, (cudaStream_t)stream);
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int d3il_robot_kinematics(d3il_env* h, float* out22, void* stream) {
  if (!h || !out22) { g_err = "d3il_robot_kinematics: null argument"; return -1; }
  CK(cudaSetDevice(h->device));
  d3il_launch_robot_kinematics(h->d, out22, (cudaStream_t)stream);
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int d3il_object_poses(d3il_env* h, float* out, void* stream) {
  if (!h || !out) { g_err = "d3il_object_poses: null argument"; return -1; }
  CK(cudaSetDevice(h->device));
  d3il_launch_object_poses(h->d, h->m.nobj, out, (cudaStream_t)stream);
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// ---- host-buffer variants (end-to-end path: pinned staging, H2D, kernels, D2H, sync)
extern "C" int d3il_reset_host(d3il_env* h, const float* ctx, const uint8_t* mask, float* obs) {
  if (!h) { g_err = "d3il_reset_host: null handle"; return -1; }
  const Model& m = h->m; const size_t n = h->n;
  CK(cudaSetDevice(h->device));
  cudaStream_t s = h->own_stream;
  if (m.ctx_dim > 0) {
    if (!ctx) { g_err = "d3il_reset_host: this scene needs a context per env"; return -1; }
    memcpy(h->h_in, ctx, n * m.ctx_dim * sizeof(float));
    CK(cudaMemcpyAsync(h->d_in, h->h_in, n * m.ctx_dim * sizeof(float), cudaMemcpyHostToDevice, s));
  }
  if (mask) { memcpy(h->h_mask, mask, n); CK(cudaMemcpyAsync(h->d_mask, h->h_mask, n, cudaMemcpyHostToDevice, s)); }
  int rc = d3il_reset(h, m.ctx_dim > 0 ? h->d_in : nullptr, mask ? h->d_mask : nullptr, obs ? h->d_out : nullptr, s);
  if (rc) return rc;
  if (obs) {
    // rows of envs that were not reset keep their previous device-side content; only masked rows are meaningful
    CK(cudaMemcpyAsync(h->h_out, h->d_out, n * m.obs_dim * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (size_t e = 0; e < n; e++) if (!mask || mask[e]) memcpy(obs + e * m.obs_dim, h->h_out + e * m.obs_dim, m.obs_dim * sizeof(float));
  } else CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int d3il_step_host(d3il_env* h, const float* action, float* obs, float* reward, uint8_t* done, float* info) {
  if (!h || !action || !obs || !reward || !done || !info) { g_err = "d3il_step_host: null argument"; return -1; }
  const Model& m = h->m; const size_t n = h->n;
  CK(cudaSetDevice(h->device));
  cudaStream_t s = h->own_stream;
  memcpy(h->h_in, action, n * m.act_dim * sizeof(float));
  CK(cudaMemcpyAsync(h->d_in, h->h_in, n * m.act_dim * sizeof(float), cudaMemcpyHostToDevice, s));
  float* d_obs = h->d_out; float* d_rew = d_obs + n * m.obs_dim; float* d_info = d_rew + n; uint8_t* d_done = (uint8_t*)(d_info + n * m.info_dim);
  int rc = d3il_step(h, h->d_in, d_obs, d_rew, d_done, d_info, s);
  if (rc) return rc;
  size_t bytes = (n * (m.obs_dim + 1 + m.info_dim)) * sizeof(float) + n;
  CK(cudaMemcpyAsync(h->h_out, h->d_out, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  memcpy(obs, h->h_out, n * m.obs_dim * sizeof(float));
  memcpy(reward, h->h_out + n * m.obs_dim, n * sizeof(float));
  memcpy(info, h->h_out + n * (m.obs_dim + 1), n * m.info_dim * sizeof(float));
  memcpy(done, (uint8_t*)(h->h_out + n * (m.obs_dim + 1 + m.info_dim)), n);
  return 0;
}

extern "C" int d3il_robot_state_host(d3il_env* h, float* tcp) {
  if (!h || !tcp) { g_err = "d3il_robot_state_host: null argument"; return -1; }
  CK(cudaSetDevice(h->device));
  int rc = d3il_robot_state(h, h->d_out, h->own_stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_out, h->d_out, (size_t)h->n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->own_stream));
  CK(cudaStreamSynchronize(h->own_stream));
  memcpy(tcp, h->h_out, (size_t)h->n * 3 * sizeof(float));
  return 0;
}

extern "C" int d3il_joint_state_host(d3il_env* h, float* j8) {
  if (!h || !j8) { g_err = "d3il_joint_state_host: null argument"; return -1; }
  CK(cudaSetDevice(h->device));
  int rc = d3il_joint_state(h, h->d_out, h->own_stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_out, h->d_out, (size_t)h->n * 8 * sizeof(float), cudaMemcpyDeviceToHost, h->own_stream));
  CK(cudaStreamSynchronize(h->own_stream));
  memcpy(j8, h->h_out, (size_t)h->n * 8 * sizeof(float));
  return 0;
}

// ---- flat fp64 state of one env (parity tests)
static int fetch_env(d3il_env* h, int e, std::vector<float>& row, IkState& ik) {
  const int n = h->n;
  row.resize(h->d.row);
  CK(cudaMemcpy(row.data(), h->d.state + (size_t)e * h->d.row, h->d.row * sizeof(float), cudaMemcpyDeviceToHost));
  float des[7], jt[21];
  for (int k = 0; k < 7; k++) {
    CK(cudaMemcpy(&ik.q[k], h->d.ik.q + (size_t)k * n + e, sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&des[k], h->d.ik.des + (size_t)k * n + e, sizeof(float), cudaMemcpyDeviceToHost));
  }
  for (int k = 0; k < 21; k++) CK(cudaMemcpy(&jt[k], h->d.ik.jt + (size_t)k * n + e, sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&ik.valid, h->d.ik.valid + e, sizeof(int), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 3; k++) ik.des_pos[k] = des[k];
  for (int k = 0; k < 4; k++) ik.des_quat[k] = des[3 + k];
  for (int k = 0; k < 7; k++) { ik.jt_q[k] = jt[k]; ik.jt_qlo[k] = jt[7 + k]; ik.jt_qd[k] = jt[14 + k]; }
  return 0;
}

extern "C" int d3il_get_state(d3il_env* h, double* out, int e) {
  if (!h || !out || e < 0 || e >= h->n) { g_err = "d3il_get_state: bad arguments"; return -1; }
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  std::vector<float> row; IkState ik;
  int rc = fetch_env(h, e, row, ik);
  if (rc) return rc;
  d3il_pack_state(h->m, h->L, row.data(), ik, out);
  return 0;
}

extern "C" int d3il_set_state(d3il_env* h, const double* in, int e) {
  if (!h || !in || e < 0 || e >= h->n) { g_err = "d3il_set_state: bad arguments"; return -1; }
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  const int n = h->n;
  std::vector<float> row(h->d.row, 0.f); IkState ik;
  d3il_unpack_state(h->m, h->L, row.data(), ik, in);
  CK(cudaMemcpy(h->d.state + (size_t)e * h->d.row, row.data(), h->d.row * sizeof(float), cudaMemcpyHostToDevice));
  float des[7], jt[21];
  for (int k = 0; k < 3; k++) des[k] = ik.des_pos[k];
  for (int k = 0; k < 4; k++) des[3 + k] = ik.des_quat[k];
  for (int k = 0; k < 7; k++) { jt[k] = ik.jt_q[k]; jt[7 + k] = ik.jt_qlo[k]; jt[14 + k] = ik.jt_qd[k]; }
  for (int k = 0; k < 7; k++) {
    CK(cudaMemcpy(h->d.ik.q + (size_t)k * n + e, &ik.q[k], sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d.ik.des + (size_t)k * n + e, &des[k], sizeof(float), cudaMemcpyHostToDevice));
  }
  for (int k = 0; k < 21; k++) CK(cudaMemcpy(h->d.ik.jt + (size_t)k * n + e, &jt[k], sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d.ik.valid + e, &ik.valid, sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

#ifdef D3IL_PHASE_TIMING
extern "C" int d3il_debug_timeline(unsigned long long* out) {
  std::vector<unsigned long long> a(4 * 4096), b(4 * 4096);
  if (cudaMemcpyFromSymbol(a.data(), g_tl, sizeof(unsigned long long) * 4 * 4096) != cudaSuccess) return -2;
  if (d3il_debug_timeline_env(b.data())) return -2;
  for (int i = 0; i < 4096; i++) for (int k = 0; k < 4; k++) out[4 * i + k] = i >= 2048 ? a[4 * i + k] : b[4 * i + k];
  return 0;
}
extern "C" int d3il_debug_phase_cycles(unsigned long long* out24) { return d3il_debug_phase_cycles_env(out24); }
#endif
