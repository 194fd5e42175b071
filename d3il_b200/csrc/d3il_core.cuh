// d3il_core.cuh — lane-cooperative fp32 core of the batched env step (one G-lane group per env).
//
// The hot path of SURVEY.md §8(a) rows a3/a4/a7/a8/a9 (Scene.next_step -> MjRobot.prepare_step -> mj_step ->
// receiveState; reference environments/d3il/d3il_sim/core/Scene.py:121-138, sims/mj_beta/MjRobot.py:125-184,
// mujoco.mj_step [EXT]) re-designed for a warp: every phase is a strided loop over links / dofs / pairs / contacts /
// constraint rows with the group's lanes, operands staged in shared memory, group barriers between phases and
// shuffle reductions for the scalar solver quantities.  No tensor cores: there is no dense contraction here.
//
// The file compiles two ways:
//   * nvcc (device): G = 32 (or 16/8) lanes per env, GSYNC = __syncwarp, reductions = __shfl_xor_sync.
//   * -DD3IL_EMU (host, tests only): G = 1, one "lane" walks every strided loop sequentially.  This is how the
//     kernel logic is checked against the fp64 oracle in the CPU container.  To keep both builds equivalent the code
//     never branches on a particular lane id: cross-lane data always goes through workspace memory + GSYNC, or
//     through the all-reduce helpers below.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef D3IL_REAL
#define D3IL_REAL float
#endif
typedef D3IL_REAL real;
#ifdef D3IL_TAB64
typedef double tab_t;   // test-only: exact tables for logic checks against the fp64 oracle
#else
typedef float tab_t;    // model tables are fp32 in shared memory
#endif

#ifdef D3IL_EMU
#define DEVFN static inline
#define DEVNI static inline
#define HDFN static inline
#define D3_RESTRICT
struct Cx { int lane; unsigned mask; int cta_threads; int bar_id; };   // cta_threads: threads of this CTA that take part in the phase barriers
template <int G> DEVFN void gsync(const Cx&) {}
template <int G> DEVFN real gsum(const Cx&, real x) { return x; }
template <int G> DEVFN real gmaxr(const Cx&, real x) { return x; }
template <int G> DEVFN void gsum2(const Cx&, real&, real&) {}
template <int G> DEVFN int gsumi(const Cx&, int x) { return x; }
template <int G> DEVFN int gori(const Cx&, int x) { return x; }
template <int G, int N> DEVFN void gsumn(const Cx&, real*) {}
template <int G> DEVFN double gsumd(const Cx&, double x) { return x; }
#else
#define DEVFN __device__ __forceinline__
#define HDFN __host__ __device__ __forceinline__
#define DEVNI static __device__ __noinline__       // big, multiply-instantiated routines: the kernel is instruction-fetch bound
#define D3_RESTRICT __restrict__
struct Cx { int lane; unsigned mask; int cta_threads; int bar_id; };   // cta_threads: threads of this CTA that take part in the phase barriers
template <int G> DEVFN void gsync(const Cx& cx) { __syncwarp(cx.mask); }
template <int G> DEVFN real gsum(const Cx& cx, real x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(cx.mask, x, o, G);
  return x;
}
// two all-reduces in one butterfly: the shuffles of the two values interleave, half the latency of two gsum calls
template <int G> DEVFN void gsum2(const Cx& cx, real& x, real& y) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) { const real a = __shfl_xor_sync(cx.mask, x, o, G), b = __shfl_xor_sync(cx.mask, y, o, G); x += a; y += b; }
}
// N all-reduces in one butterfly (the shuffles of the N values pipeline)
template <int G, int N> DEVFN void gsumn(const Cx& cx, real* v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    real t[N];
#pragma unroll
    for (int k = 0; k < N; k++) t[k] = __shfl_xor_sync(cx.mask, v[k], o, G);
#pragma unroll
    for (int k = 0; k < N; k++) v[k] += t[k];
  }
}
template <int G> DEVFN real gmaxr(const Cx& cx, real x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(cx.mask, x, o, G));
  return x;
}
template <int G> DEVFN int gsumi(const Cx& cx, int x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(cx.mask, x, o, G);
  return x;
}
template <int G> DEVFN int gori(const Cx& cx, int x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x |= __shfl_xor_sync(cx.mask, x, o, G);
  return x;
}
template <int G> DEVFN double gsumd(const Cx& cx, double x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(cx.mask, x, o, G);
  return x;
}
#endif

// compiled table sizes (largest scene: Sorting-6, 15 links, nv 45)
#define D3_MAXLINK 16
#define D3_MAXV 48
#define D3_MAXQ 56
#define D3_MAXGEOM 24
#define D3_MAXPAIR 96
#define D3_MAXHE 192       // in-block lower-triangle entries (Sorting-6: 45 + 6 x 21 = 171)

#define LANES(i, n) for (int i = cx.lane; i < (n); i += G)
#ifndef D3IL_GRAD_FLOOR
#define D3IL_GRAD_FLOOR 6e-6  // relative size of the Newton gradient below which fp32 cannot resolve it (see solve_constraints)
#endif
#ifndef D3IL_LS_TOL
#define D3IL_LS_TOL 1e-3   // exact line search: |phi'(alpha)| <= tol |phi'(0)|  (MuJoCo's ls_tolerance default is 1e-2; the solution does not depend on it)
#endif

// The env-step kernel is instruction-fetch bound (tens of KB of straight-line code per phase, one env per group):
// CTA-wide barriers at fixed phase boundaries keep every warp of the CTA inside the same code region, so one
// instruction fetch serves all of them.  CS = false (reset kernel with masked early exits, host emulation): no barriers.
// CTA-wide "does anybody still need another iteration" vote (plain flag without CTA barriers)
template <bool CS> DEVFN int cta_any(const Cx& cx, int pred) {
#if defined(__CUDA_ARCH__)
  if (CS && cx.cta_threads > 32) {        // a free-running CTA (cta_threads = one warp) votes alone: pred is group-uniform
    int out;
    asm volatile("{ .reg .pred p, q; setp.ne.s32 p, %1, 0; bar.red.or.pred q, %3, %2, p; selp.s32 %0, 1, 0, q; }" : "=r"(out) : "r"(pred), "r"(cx.cta_threads), "r"(cx.bar_id) : "memory");
    return out;
  }
#endif
  return pred;
}
// Named barrier 1 over the participating threads only (CTAs that carry fewer envs than warps let the spare warps exit)
template <bool CS> DEVFN void cta_sync(const Cx& cx) {
#if defined(__CUDA_ARCH__)
  if (CS) { if (cx.cta_threads > 32) asm volatile("bar.sync %1, %0;" :: "r"(cx.cta_threads), "r"(cx.bar_id) : "memory"); else __syncwarp(); }
#endif
}

#if defined(D3IL_PHASE_TIMING) && !defined(D3IL_EMU)
// debug build only: per-phase cycle counts of the group that owns env 0 (profiles/phase_timing.py)
static __device__ unsigned long long g_phase_cycles[40];
static __device__ int g_ph_on[64];            // D3IL_PHASE_DENSE: per free-running CTA, the current tick has >= 2 tree-coupling contacts
static __device__ unsigned g_cta_stat[4 * 4096];            // per CTA: [0] CTA-uniform Newton passes, [1] of which warp 0's env was coupled, [2] Newton steps of warp 0's env, [3] line-search evaluations of warp 0
static __device__ unsigned long long g_iter_hist[40];      // [0..15] Newton steps per tick, [16..31] same for ticks with a coupling contact, [32] sum ncon / [33] count of ticks with >= 8 steps
#endif
#if defined(D3IL_PHASE_TIMING) && defined(__CUDA_ARCH__)
#define PHASE_T0() long long t_ph = clock64()
#define PHASE2_T0() long long t_ph2 = clock64()
#ifndef D3IL_PHASE_BLOCK
#define D3IL_PHASE_BLOCK 0      // which CTA of the cost-sorted grid is sampled (e.g. -DD3IL_PHASE_BLOCK="(gridDim.x/2)" = median cost)
#endif
#ifdef D3IL_PHASE_DENSE                                       // first warp of the free-running CTAs, only in ticks with >= 2 tree-coupling contacts
#define D3IL_PHASE_COND (blockIdx.x < 16 && g_ph_on[blockIdx.x])
#define PHASE_MARK_DENSE(on) do { if (threadIdx.x == 0 && blockIdx.x < 16) g_ph_on[blockIdx.x] = (on); } while (0)
#elif defined(D3IL_PHASE_LO)                                          // a range of CTAs (first warp of each): -DD3IL_PHASE_LO=3 -DD3IL_PHASE_HI=8
#define D3IL_PHASE_COND (blockIdx.x >= D3IL_PHASE_LO && blockIdx.x < D3IL_PHASE_HI)
#else
#define D3IL_PHASE_COND (blockIdx.x == D3IL_PHASE_BLOCK)
#endif
#ifndef PHASE_MARK_DENSE
#define PHASE_MARK_DENSE(on) ((void)0)
#endif
__device__ __forceinline__ bool blockIdx_is0() { return D3IL_PHASE_COND && threadIdx.x == 0; }
__device__ __forceinline__ void count_iter() { atomicAdd(&g_phase_cycles[20], 1ull); }
__device__ __forceinline__ void count_stat(int k, int v) { atomicAdd(&g_phase_cycles[k], (unsigned long long)v); }
#define PHASE2(k) do { long long t_now2 = clock64(); if (D3IL_PHASE_COND && threadIdx.x == 0) atomicAdd(&g_phase_cycles[k], (unsigned long long)(t_now2 - t_ph2)); t_ph2 = t_now2; } while (0)
#define PHASE(k) do { long long t_now = clock64(); if (D3IL_PHASE_COND && threadIdx.x == 0) atomicAdd(&g_phase_cycles[k], (unsigned long long)(t_now - t_ph)); t_ph = t_now; } while (0)
#else
#define PHASE_T0() ((void)0)
#define PHASE2_T0() ((void)0)
#define PHASE_MARK_DENSE(on) ((void)0)
#define PHASE(k) ((void)0)
#define PHASE2(k) ((void)0)
DEVFN bool blockIdx_is0() { return false; }
DEVFN void count_iter() {}
DEVFN void count_stat(int, int) {}
#endif

// Register-distributed vectors: element i lives in slot i / G of lane i % G.
#define D3_SLOTS(G) ((D3_MAXV + (G) - 1) / (G))
template <int G, int NS>
DEVFN real elem_bcast(const Cx& cx, const real* xr, int c) {
#ifdef D3IL_EMU
  return xr[c];
#else
  if (NS == 1) return __shfl_sync(cx.mask, xr[0], c, G);
  real v = 0;
#pragma unroll
  for (int sl = 0; sl < NS; sl++) { real t = __shfl_sync(cx.mask, xr[sl], c % G, G); if (sl == c / G) v = t; }
  return v;
#endif
}

// ------------------------------------------------------------------------------------------------ model tables
// Float copy of the D3SC scene blob (d3il_b200/scene/blob.py) + derived index tables, staged into shared memory
// once per CTA.  All arrays are sized for the largest scene we compile (Sorting-6: 15 links, nv 45).
#define D3_NARM 7
#define D3_NROB 9
#define D3_LINK_W 32
#define D3_GEOM_W 24
#define D3_PAIR_W 24
#define D3_CTRL_W 192

enum { D3C_IK_ORIGIN = 0, D3C_IK_EE = 84, D3C_PGAIN_POS = 96, D3C_PGAIN_QUAT = 99, D3C_PGAIN_NULL = 102, D3C_REST = 109,
       D3C_JMIN = 116, D3C_JMAX = 123, D3C_PD_P = 130, D3C_PD_D = 137, D3C_JREG = 144, D3C_SVD_MIN = 145, D3C_SVD_MAX = 146,
       D3C_NUM_ITER = 147, D3C_LRATE = 148, D3C_DT = 149, D3C_INIT_QPOS = 150, D3C_TCP_POS = 157, D3C_TCP_QUAT = 160,
       D3C_GRAVITY = 164, D3C_IMPRATIO = 167, D3C_TOL = 168, D3C_JNT_SOLREF = 169, D3C_JNT_SOLIMP = 171, D3C_MEANINERTIA = 179 };
enum { D3G_CYLINDER = 5, D3G_BOX = 6 };
enum { D3T_AVOIDING = 0, D3T_PUSHING = 1, D3T_ALIGNING = 2, D3T_SORTING = 3, D3T_STACKING = 4, D3T_INSERTING = 5 };

struct Model {
  int task_id, nlink, nobj, nq, nv, ngeom, npair, n_substeps, max_steps, obs_dim, act_dim, ctx_dim, info_dim, ctrl_kind, ntaskp;
  int maxcon, maxrow, nmpair;            // workspace caps and number of structurally non-zero (a>=b) entries of M
  int ws_floats;                          // per-env workspace size
  int nextra;                             // per-env task words outside qpos (Aligning: target pose)
  int maxdim;                             // largest contact dimension of the scene (3, or 4 with torsional friction: gripper pads)
  // per link
  int l_parent[D3_MAXLINK], l_jtype[D3_MAXLINK], l_qadr[D3_MAXLINK], l_dadr[D3_MAXLINK], l_ndof[D3_MAXLINK], l_limited[D3_MAXLINK];
  unsigned l_anc[D3_MAXLINK];             // bit j set: link j is an ancestor-or-self
  unsigned l_desc[D3_MAXLINK];            // bit j set: link j is a descendant-or-self
  int l_ref[D3_MAXLINK];                  // link whose origin is the spatial reference point of this link's kinematic tree
  int d_link[D3_MAXV];
  int d_bs[D3_MAXV], d_be[D3_MAXV];       // [start, end) of the dof's kinematic-tree block (M is block diagonal over these)
  int maxblk;                             // largest block
  int ndamp, damp_first, damp_end;        // trailing damped dofs of the arm block (implicit-damping refactor)
  unsigned char p_rng[D3_MAXPAIR * 4];    // per pair: dof ranges [a0,a1) [b0,b1) touched by its two geoms (static)
  unsigned char p_cpl[D3_MAXPAIR];        // pair joins two different kinematic-tree blocks
  unsigned char g_slab[32];               // geom is a static, axis-aligned box (table top, support): eligible for the slab fast path
  int nblk, blk_s[8], blk_e[8];           // the kinematic-tree blocks as a list
  unsigned char d_blk[D3_MAXV];           // block index of each dof
  short m_row[D3_MAXV];                   // packed mass matrix: M(r, c) = Mbuf[m_row[r] + c] for c inside r's block (blocks stored back to back, row-major squares)
  short m_base[D3_MAXV];                  // offset of the dof's block in the packed buffer
  int m_size;                             // floats of the packed buffer = sum of block sizes squared
  int nhe; unsigned char he_i[D3_MAXHE], he_j[D3_MAXHE];   // all in-block lower-triangle entries (gi >= gj) of the block-diagonal Hessian, as one list
  unsigned diag_blk;                      // bit b: block b of M is diagonal (free body, CoM at its origin, principal axes = body axes)
  int nzp; unsigned char zp_a[8], zp_b[8]; // in-block dof pairs that CRBA never writes (the two fingers): must read as zero
  unsigned char tri_i[300], tri_j[300];   // row-major lower-triangle unranking table for n <= 24
  unsigned char mp_a[D3_MAXV * 8], mp_b[D3_MAXV * 8];   // (a,b) list of related dof pairs, a >= b
  tab_t link[D3_MAXLINK * D3_LINK_W];
  tab_t geom[D3_MAXGEOM * D3_GEOM_W];
  tab_t geomR[D3_MAXGEOM * 9];
  tab_t linkR[D3_MAXLINK * 9];             // constant rotation of each link frame in its parent (quat2mat of the table's quaternion)
  tab_t ctrl[D3_CTRL_W];
  tab_t taskp[32];
  tab_t pair[D3_MAXPAIR * D3_PAIR_W];      // LAST: only the first npair rows are staged into shared memory (d3il_model_bytes)
};
// bytes of the Model a CTA stages into shared memory: everything up to the scene's last candidate pair
static inline size_t d3il_model_bytes(const Model& m) {
  return (((size_t)((const char*)m.pair - (const char*)&m) + sizeof(tab_t) * D3_PAIR_W * (size_t)m.npair) + 127) & ~(size_t)127;
}

// ------------------------------------------------------------------------------------------------ per-env workspace
// Offsets (in reals) into the env's slice of shared memory.  Persistent state first (mirrors the HBM row), then
// scratch.  Everything is a function of the Model's sizes, computed once on the host (d3il_layout).
struct Lay {
  // persistent state (HBM <-> shared at kernel entry/exit)
  int qpos, qlo, qvel, warm, bias_prev, tcp, misc, extra; // qlo: low words of the 9 robot joint angles (two-float qpos);
                                                    // tcp: pos3+quat4 ; misc: 16 scalars (see ST_*)
  int n_state;
  // scratch
  int xpos, xmat, S, I10, Ic, vel, cj, frc, F, M, mdinv, H, hdinv, hpiv, bias, qfrc_smooth, qacc_smooth, qacc, qfrc_c, grad, pvec, Ma, tmpv;
  int wood;           // Woodbury scratch: maxdim dof vectors (view 2)
  int rowc;           // per constraint row >= nlimit: contact id | (row within the contact << 6), bytes
  int blist;          // per kinematic-tree block: [count, contact ids ...] bytes, stride maxcon + 1 (bit 7: the contact couples two blocks)
  int act, jt, con, ncon_pair, limflag, cflag, J, aref, D, jar, frcE, Jp, hd, hb, etype, econ, total;
};
enum { ST_GRIP_SET = 0, ST_GRASP, ST_CTRL_MODE, ST_STEP, ST_TERM, ST_STATUS, ST_OBST, ST_TASK0, ST_TASK1, ST_TASK2, ST_TASK3, ST_COST_ITERS /* Newton iterations of the last env step */, ST_COST_COUPLED /* ticks with a tree-coupling contact */, ST_COST_NCON /* max contacts */, ST_COST_NEAR /* ticks in which a tree-coupling pair passed the broad phase: contact is imminent */, ST_NMISC = 16 };
// per-env fault word (misc[ST_STATUS], last column of `info`; sticky until the env is reset)
enum { D3_STATUS_M_NOT_PD = 1, D3_STATUS_OVERFLOW = 2 /* contact / row budget exceeded: contacts dropped */, D3_STATUS_H_NOT_PD = 4,
       D3_STATUS_ITER_CAP = 8 /* Newton loop left at max_iter without passing the convergence test */, D3_STATUS_BAD_ACTION = 16 /* non-finite action or zero quaternion: last set-point held */ };
#define D3_CON_W 24    // per contact: pos3, frame9, dist, incl, mu, dim, g1, g2, pair, row0, dof ranges a0,a1,b0,b1

#define D3_JW 16       // compact Jacobian row: entries of the contact's two dof ranges (<= 9 + 6), padded to 16

static inline void d3il_layout(const Model& m, Lay& L) {
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  L.qpos = take(m.nq); L.qlo = take(D3_NROB); L.qvel = take(m.nv); L.warm = take(m.nv); L.bias_prev = take(D3_NROB); L.tcp = take(7); L.misc = take(ST_NMISC); L.extra = take(m.nextra);
  L.n_state = o;
  // live for the whole tick
  L.M = take(m.m_size); L.mdinv = take(m.nv);
  L.bias = take(m.nv); L.qfrc_smooth = take(m.nv); L.qacc_smooth = take(m.nv); L.qacc = take(m.nv); L.qfrc_c = take(m.nv);
  L.act = take(D3_NROB); L.jt = take(3 * D3_NARM); L.con = take(D3_CON_W * m.maxcon);
  L.J = take(m.maxrow * D3_JW); L.aref = take(m.maxrow); L.D = take(m.maxrow); L.hd = take(2 * D3_NROB); L.econ = take(2 * D3_NROB); L.blist = take((m.nblk * (m.maxcon + 1) + 3) / 4); L.rowc = take((m.maxrow + 3) / 4);
  // region X, two views that are never live together:
  //   view 1 (kinematics, dynamics, collision, constraint assembly)   view 2 (Newton solver, Euler)
  const int x0 = o;
  L.xpos = take(3 * m.nlink); L.xmat = take(9 * m.nlink); L.S = take(6 * m.nv); L.I10 = take(10 * m.nlink); L.Ic = take(10 * m.nlink);
  L.vel = take(6 * m.nlink); L.cj = take(6 * m.nlink); L.frc = take(6 * m.nlink); L.F = take(6 * m.nv); L.ncon_pair = take(m.npair + 4); L.limflag = take(2 * D3_NROB); L.cflag = take(m.maxcon);
  const int x1 = o;
  o = x0;
  L.H = take(m.nv * m.nv); L.hdinv = take(m.nv); L.hpiv = take(m.nv); L.jar = take(m.maxrow); L.frcE = take(m.maxrow); L.Jp = take(m.maxrow);
  L.hb = take(m.maxdim * m.maxdim * m.maxcon); L.grad = take(m.nv); L.pvec = take(m.nv); L.Ma = take(m.nv); L.tmpv = take(m.nv);
  // Woodbury scratch (maxdim dof vectors): dead before the line search writes Jp, so it shares Jp's storage when it fits
  L.wood = m.maxrow >= m.maxdim * m.nv ? L.Jp : take(m.maxdim * m.nv);
  if (x1 > o) o = x1;
  L.etype = 0;
  L.total = o;
}

// ------------------------------------------------------------------------------------------------ small math
DEVFN real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
DEVFN void cross3(real* o, const real* a, const real* b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
DEVFN real norm3(const real* a) { return sqrt(dot3(a, a)); }
DEVFN void mat_vec3(real* o, const real* R, const real* v) {
  real x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2], z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
DEVFN void matT_vec3(real* o, const real* R, const real* v) {
  real x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2], z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
DEVFN void mat_mul3(real* o, const real* A, const real* B) {
  real t[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) o[i] = t[i];
}
DEVFN void quat2mat(real* R, const real* q) {
  real n = 1 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  real w = q[0] * n, x = q[1] * n, y = q[2] * n, z = q[3] * n;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
DEVFN void mat2quat(real* q, const real* R) {
  real t = R[0] + R[4] + R[8];
  if (t > 0) { real s = sqrt(t + 1) * 2; q[0] = (real)0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { real s = sqrt(1 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = (real)0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { real s = sqrt(1 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = (real)0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { real s = sqrt(1 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = (real)0.25 * s; }
  real n = 1 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
}
// sin/cos for |x| <= ~4 (joint half-angles, per-tick rotation increments): Cody-Waite reduction by pi/2 and
// minimax polynomials on [-pi/4, pi/4]; ~1 ulp in fp32 and a fraction of the code size of sinf/cosf.
DEVFN void sincos_small(real x, real* sn, real* cs) {
#ifdef __CUDA_ARCH__
  float k = rintf(x * 0.636619772f);
  float r = fmaf(k, -1.57079601e+00f, x);
  r = fmaf(k, -3.13916473e-07f, r);
  r = fmaf(k, -5.39030253e-15f, r);
  float r2 = r * r;
  float sp = fmaf(fmaf(fmaf(-1.95152959e-4f, r2, 8.33216087e-3f), r2, -1.66666546e-1f), r2 * r, r);
  float cp = fmaf(fmaf(fmaf(fmaf(2.44331571e-5f, r2, -1.38873163e-3f), r2, 4.16666457e-2f), r2, -0.5f), r2, 1.0f);
  int q = (int)k & 3;
  float s1 = (q & 1) ? cp : sp, c1 = (q & 1) ? sp : cp;
  *sn = (q & 2) ? -s1 : s1;
  *cs = ((q + 1) & 2) ? -c1 : c1;
#else
  *sn = sin(x); *cs = cos(x);
#endif
}
DEVFN real clampr(real x, real lo, real hi) { return x < lo ? lo : (x > hi ? hi : x); }
DEVFN real absr(real x) { return x < 0 ? -x : x; }
DEVFN real maxr(real a, real b) { return a > b ? a : b; }
DEVFN real minr(real a, real b) { return a < b ? a : b; }

// ------------------------------------------------------------------------------------------------ kinematics
// Each lane composes the transforms from the root down to "its" link (the arm is a chain of depth <= 9, free bodies
// depth 1), so no inter-lane communication is needed; world pose, joint motion subspace S (world axes, linear part at
// the world origin) and the 10-parameter spatial inertia about the world origin land in the workspace.
template <int G>
DEVNI void kinematics(const Cx& cx, const Model& m, const Lay& L, real* w) {
  LANES(i, m.nlink) {
    real p[3] = {0, 0, 0}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (m.l_jtype[i] == 2) {
      real* q = w + L.qpos + m.l_qadr[i];
      real n = 1 / sqrt(q[3] * q[3] + q[4] * q[4] + q[5] * q[5] + q[6] * q[6]);
      q[3] *= n; q[4] *= n; q[5] *= n; q[6] *= n;                     // mj_kinematics normalises the quaternion in qpos
      p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
      quat2mat(R, q + 3);
    } else {
      // walk root -> i through the ancestor mask (links are topologically ordered)
      unsigned anc = m.l_anc[i];
      for (int j = 0; j <= i; j++) {
        if (!((anc >> j) & 1u)) continue;
        const tab_t* Lk = m.link + D3_LINK_W * j;
        real off[3], lp[3] = {(real)Lk[2], (real)Lk[3], (real)Lk[4]};
        mat_vec3(off, R, lp);
        p[0] += off[0]; p[1] += off[1]; p[2] += off[2];
        real Rq[9];
        for (int k = 0; k < 9; k++) Rq[k] = (real)m.linkR[9 * j + k];
        mat_mul3(R, R, Rq);
        real q = w[L.qpos + m.l_qadr[j]];
        real ax[3] = {(real)Lk[9], (real)Lk[10], (real)Lk[11]};
        if (m.l_jtype[j] == 0 && ax[0] == 0 && ax[1] == 0 && ax[2] == 1) {
          // hinge about the link's z axis (every Panda joint): R <- R Rz(q) mixes the first two columns
          real s, c; sincos_small(q, &s, &c);
          for (int r = 0; r < 3; r++) { real x0 = R[3 * r], x1 = R[3 * r + 1]; R[3 * r] = c * x0 + s * x1; R[3 * r + 1] = c * x1 - s * x0; }
        } else if (m.l_jtype[j] == 0) {
          real s, c; sincos_small((real)0.5 * q, &s, &c);
          real hq[4] = {c, s * ax[0], s * ax[1], s * ax[2]}, Rj[9];
          quat2mat(Rj, hq); mat_mul3(R, R, Rj);
        } else {
          real a[3]; mat_vec3(a, R, ax);
          p[0] += a[0] * q; p[1] += a[1] * q; p[2] += a[2] * q;
        }
      }
    }
    for (int k = 0; k < 3; k++) w[L.xpos + 3 * i + k] = p[k];
    for (int k = 0; k < 9; k++) w[L.xmat + 9 * i + k] = R[k];
  }
  gsync<G>(cx);
  // Spatial quantities are taken about a per-tree reference point (the origin of link l_ref: the wrist for the arm, the
  // body itself for free objects) instead of the world origin: in fp32 the parallel-axis terms m|c|^2 would otherwise
  // swamp the small distal inertias (MuJoCo uses the subtree CoM for the same reason).
  LANES(i, m.nlink) {
    const real* R = w + L.xmat + 9 * i;
    const real* pr = w + L.xpos + 3 * m.l_ref[i];
    real p[3] = {w[L.xpos + 3 * i] - pr[0], w[L.xpos + 3 * i + 1] - pr[1], w[L.xpos + 3 * i + 2] - pr[2]};
    const tab_t* Lk = m.link + D3_LINK_W * i;
    if (m.l_jtype[i] == 2) {
      for (int k = 0; k < 3; k++) {
        real* Sl = w + L.S + 6 * (m.l_dadr[i] + k); real* Sa = w + L.S + 6 * (m.l_dadr[i] + 3 + k);
        for (int c = 0; c < 6; c++) Sl[c] = 0;
        Sl[3 + k] = 1;
        real a[3] = {R[k], R[3 + k], R[6 + k]};
        Sa[0] = a[0]; Sa[1] = a[1]; Sa[2] = a[2]; cross3(Sa + 3, p, a);
      }
    } else {
      real ax[3] = {(real)Lk[9], (real)Lk[10], (real)Lk[11]}, a[3];
      mat_vec3(a, R, ax);
      real* S = w + L.S + 6 * m.l_dadr[i];
      if (m.l_jtype[i] == 0) { S[0] = a[0]; S[1] = a[1]; S[2] = a[2]; cross3(S + 3, p, a); }
      else { S[0] = S[1] = S[2] = 0; S[3] = a[0]; S[4] = a[1]; S[5] = a[2]; }
    }
    // spatial inertia about the reference point: m, h = m c, IO = R I R^T + m (|c|^2 1 - c c^T)
    real mass = Lk[12], ip[3] = {(real)Lk[13], (real)Lk[14], (real)Lk[15]}, c[3];
    mat_vec3(c, R, ip); c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
    real Ib[9] = {(real)Lk[16], (real)Lk[19], (real)Lk[20], (real)Lk[19], (real)Lk[17], (real)Lk[21], (real)Lk[20], (real)Lk[21], (real)Lk[18]};
    real RI[9], Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]}, Iw[9];
    mat_mul3(RI, R, Ib); mat_mul3(Iw, RI, Rt);
    real cc = dot3(c, c);
    real* I10 = w + L.I10 + 10 * i;
    I10[0] = mass; I10[1] = mass * c[0]; I10[2] = mass * c[1]; I10[3] = mass * c[2];
    I10[4] = Iw[0] + mass * (cc - c[0] * c[0]); I10[5] = Iw[4] + mass * (cc - c[1] * c[1]); I10[6] = Iw[8] + mass * (cc - c[2] * c[2]);
    I10[7] = Iw[1] - mass * c[0] * c[1]; I10[8] = Iw[2] - mass * c[0] * c[2]; I10[9] = Iw[5] - mass * c[1] * c[2];
  }
  gsync<G>(cx);
}

// [n; f] = I10 * [w; v]
DEVFN void inertia_apply(const real* I, const real* sv, real* o) {
  const real* h = I + 1;
  real hv[3], hw[3];
  cross3(hv, h, sv + 3); cross3(hw, h, sv);
  o[0] = I[4] * sv[0] + I[7] * sv[1] + I[8] * sv[2] + hv[0];
  o[1] = I[7] * sv[0] + I[5] * sv[1] + I[9] * sv[2] + hv[1];
  o[2] = I[8] * sv[0] + I[9] * sv[1] + I[6] * sv[2] + hv[2];
  o[3] = I[0] * sv[3] - hw[0]; o[4] = I[0] * sv[4] - hw[1]; o[5] = I[0] * sv[5] - hw[2];
}

// Composite inertias, mass matrix (CRBA) and bias forces (RNE with qacc = 0) — SURVEY App. B.3.
template <int G>
DEVNI void dynamics(const Cx& cx, const Model& m, const Lay& L, real* w) {
  const int nl = m.nlink, nv = m.nv;
  // (1) composite inertia = sum over descendants ; link velocity = sum over ancestor dofs
  LANES(i, nl) {
    real acc[10];
    for (int k = 0; k < 10; k++) acc[k] = 0;
    unsigned desc = m.l_desc[i];
    for (int j = i; j < nl; j++) if ((desc >> j) & 1u) for (int k = 0; k < 10; k++) acc[k] += w[L.I10 + 10 * j + k];
    for (int k = 0; k < 10; k++) w[L.Ic + 10 * i + k] = acc[k];
    real v[6] = {0, 0, 0, 0, 0, 0};
    unsigned anc = m.l_anc[i];
    for (int j = 0; j <= i; j++) if ((anc >> j) & 1u)
      for (int d = m.l_dadr[j]; d < m.l_dadr[j] + m.l_ndof[j]; d++) { real qd = w[L.qvel + d]; for (int k = 0; k < 6; k++) v[k] += w[L.S + 6 * d + k] * qd; }
    for (int k = 0; k < 6; k++) w[L.vel + 6 * i + k] = v[k];
    // joint bias acceleration cJ = d/dt(S) qd
    real vj[6] = {0, 0, 0, 0, 0, 0}, cj[6] = {0, 0, 0, 0, 0, 0};
    for (int d = m.l_dadr[i]; d < m.l_dadr[i] + m.l_ndof[i]; d++) { real qd = w[L.qvel + d]; for (int k = 0; k < 6; k++) vj[k] += w[L.S + 6 * d + k] * qd; }
    if (m.l_jtype[i] == 2) {
      cross3(cj + 3, w + L.qvel + m.l_dadr[i], vj);                     // v_p x omega
    } else {
      real t1[3], t2[3];
      cross3(cj, v, vj); cross3(t1, v, vj + 3); cross3(t2, v + 3, vj);
      cj[3] = t1[0] + t2[0]; cj[4] = t1[1] + t2[1]; cj[5] = t1[2] + t2[2];
    }
    for (int k = 0; k < 6; k++) w[L.cj + 6 * i + k] = cj[k];
  }
  gsync<G>(cx);
  // (2) F_d = Ic[link(d)] S_d ; link force f_i = I_i a_i + v_i x* (I_i v_i)
  LANES(d, nv) { inertia_apply(w + L.Ic + 10 * m.d_link[d], w + L.S + 6 * d, w + L.F + 6 * d); }
  LANES(i, nl) {
    real a[6] = {0, 0, 0, -(real)m.ctrl[D3C_GRAVITY], -(real)m.ctrl[D3C_GRAVITY + 1], -(real)m.ctrl[D3C_GRAVITY + 2]};
    unsigned anc = m.l_anc[i];
    for (int j = 0; j <= i; j++) if ((anc >> j) & 1u) for (int k = 0; k < 6; k++) a[k] += w[L.cj + 6 * j + k];
    real Ia[6], Iv[6], t1[3], t2[3], t3[3];
    const real* v = w + L.vel + 6 * i;
    inertia_apply(w + L.I10 + 10 * i, a, Ia); inertia_apply(w + L.I10 + 10 * i, v, Iv);
    cross3(t1, v, Iv); cross3(t2, v + 3, Iv + 3); cross3(t3, v, Iv + 3);
    real* f = w + L.frc + 6 * i;
    for (int k = 0; k < 3; k++) { f[k] = Ia[k] + t1[k] + t2[k]; f[3 + k] = Ia[3 + k] + t3[k]; }
  }
  gsync<G>(cx);
  // (3) M[a][b] = S_b . F_a over the related pair list ; bias_d = S_d . sum_{desc} f
  LANES(e, m.nmpair) {
    int a = m.mp_a[e], b = m.mp_b[e];
    real s = 0;
    for (int k = 0; k < 6; k++) s += w[L.S + 6 * b + k] * w[L.F + 6 * a + k];
    w[L.M + m.m_row[b] + a] = s;                                      // upper triangle + diagonal hold M (b <= a), blocks packed
  }
  LANES(d, nv) {
    int li = m.d_link[d];
    unsigned desc = m.l_desc[li];
    real f[6] = {0, 0, 0, 0, 0, 0};
    for (int j = li; j < nl; j++) if ((desc >> j) & 1u) for (int k = 0; k < 6; k++) f[k] += w[L.frc + 6 * j + k];
    real s = 0;
    for (int k = 0; k < 6; k++) s += w[L.S + 6 * d + k] * f[k];
    w[L.bias + d] = s;
  }
  gsync<G>(cx);
}

// ------------------------------------------------------------------------------------------------ narrow phase
// Same deterministic rules as the oracle (DESIGN.md "collision"): box-box SAT + face clipping / edge-edge,
// cylinder-box candidate-axis SAT + feature contact point, cylinder-cylinder via axis segments.
struct RawCon { real pos[3], n[3], dist; };

DEVFN int clip_poly(real (*P)[3], int n, real hu, real hv) {
  real Q[12][3];
  for (int plane = 0; plane < 4; plane++) {
    int ax = plane >> 1; real sg = (plane & 1) ? (real)-1 : (real)1, h = ax ? hv : hu;
    int mcount = 0;
    for (int i = 0; i < n; i++) {
      const real* a = P[i]; const real* b = P[(i + 1) % n];
      real da = h - sg * a[ax], db = h - sg * b[ax];
      if (da >= 0) { Q[mcount][0] = a[0]; Q[mcount][1] = a[1]; Q[mcount][2] = a[2]; mcount++; }
      if ((da >= 0) != (db >= 0)) { real t = da / (da - db); for (int k = 0; k < 3; k++) Q[mcount][k] = a[k] + t * (b[k] - a[k]); mcount++; }
    }
    n = mcount;
    for (int i = 0; i < n; i++) { P[i][0] = Q[i][0]; P[i][1] = Q[i][1]; P[i][2] = Q[i][2]; }
    if (n == 0) return 0;
  }
  return n;
}

// Fast path of collide_box_box for a dynamic box B resting on / sunk into the top face of a big static axis-aligned
// box A (table plane, support body): when B's bounding circle lies inside A's top rectangle, B's centre is above A's
// and the overlap along z is shallower than B's bounding radius, the SAT of the general routine always selects A's +z
// face (every other axis overlaps by at least the bounding radius; ties go to A's faces), nothing is clipped, and the
// contacts are the vertices of B's most downward face that lie below the top plane.  Same arithmetic, same order.
// Returns -1 when the preconditions do not hold (caller falls back to the general routine).
DEVFN int collide_slab_box(const real* pA, const real* hA, const real* pB, const real* RB, const real* hB, real rboundB, real margin, RawCon* out) {
  real tx = pB[0] - pA[0], ty = pB[1] - pA[1], tz = pB[2] - pA[2];
  if (!(absr(tx) + rboundB <= hA[0] && absr(ty) + rboundB <= hA[1] && tz > 0)) return -1;
  real rBz = hB[0] * absr(RB[6]) + hB[1] * absr(RB[7]) + hB[2] * absr(RB[8]);
  real sepz = tz - (hA[2] + rBz);
  if (sepz > margin) return 0;
  if (!(sepz > (real)-0.5 * rboundB)) return -1;
  int jx = 0; real jb = -1;
  for (int j = 0; j < 3; j++) { real v = absr(RB[6 + j]); if (v > jb) { jb = v; jx = j; } }
  real isg = RB[6 + jx] > 0 ? (real)-1 : (real)1;
  int k1 = (jx + 1) % 3, k2 = (jx + 2) % 3;
  if (k1 > k2) { int tmp = k1; k1 = k2; k2 = tmp; }
  int cnt = 0;
  for (int c = 0; c < 4; c++) {
    real s1 = (c == 1 || c == 2) ? (real)1 : (real)-1, s2 = (c >= 2) ? (real)1 : (real)-1;
    real ww[3];
    for (int k = 0; k < 3; k++) ww[k] = pB[k] + isg * hB[jx] * RB[3 * k + jx] + s1 * hB[k1] * RB[3 * k + k1] + s2 * hB[k2] * RB[3 * k + k2] - pA[k];
    real dist = ww[2] - hA[2];
    if (dist >= margin) continue;
    out[cnt].pos[0] = pA[0] + ww[0]; out[cnt].pos[1] = pA[1] + ww[1]; out[cnt].pos[2] = pA[2] + (hA[2] + (real)0.5 * dist);
    out[cnt].n[0] = 0; out[cnt].n[1] = 0; out[cnt].n[2] = 1;
    out[cnt].dist = dist; cnt++;
  }
  return cnt;
}

DEVNI int collide_box_box(const real* pA, const real* RA, const real* hA, const real* pB, const real* RB, const real* hB, real margin, RawCon* out) {
  real Rr[9], AbsR[9], t[3], d[3] = {pB[0] - pA[0], pB[1] - pA[1], pB[2] - pA[2]};
  matT_vec3(t, RA, d);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    real s = RA[i] * RB[j] + RA[3 + i] * RB[3 + j] + RA[6 + i] * RB[6 + j];
    Rr[3 * i + j] = s; AbsR[3 * i + j] = absr(s);
  }
  real best = (real)-1e30; int code = -1; real bsign = 1;
  for (int i = 0; i < 3; i++) {
    real sep = absr(t[i]) - (hA[i] + hB[0] * AbsR[3 * i] + hB[1] * AbsR[3 * i + 1] + hB[2] * AbsR[3 * i + 2]);
    if (sep > margin) return 0;
    if (sep > best) { best = sep; code = i; bsign = t[i] >= 0 ? (real)1 : (real)-1; }
  }
  for (int j = 0; j < 3; j++) {
    real tb = t[0] * Rr[j] + t[1] * Rr[3 + j] + t[2] * Rr[6 + j];
    real sep = absr(tb) - (hB[j] + hA[0] * AbsR[j] + hA[1] * AbsR[3 + j] + hA[2] * AbsR[6 + j]);
    if (sep > margin) return 0;
    if (sep > best + (real)1e-6) { best = sep; code = 3 + j; bsign = tb >= 0 ? (real)1 : (real)-1; }
  }
  real ebest = (real)-1e30; int ecode = -1; real en[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    real ai[3] = {RA[i], RA[3 + i], RA[6 + i]}, bj[3] = {RB[j], RB[3 + j], RB[6 + j]}, ax[3];
    cross3(ax, ai, bj);
    real l = norm3(ax);
    if (l < (real)1e-6) continue;
    ax[0] /= l; ax[1] /= l; ax[2] /= l;
    real dist = dot3(ax, d), ra = 0, rb = 0, mx = 0;
    for (int k = 0; k < 3; k++) {
      real ak[3] = {RA[k], RA[3 + k], RA[6 + k]}, bk[3] = {RB[k], RB[3 + k], RB[6 + k]};
      real da = absr(dot3(ax, ak)), db = absr(dot3(ax, bk));
      ra += hA[k] * da; rb += hB[k] * db;
      mx = maxr(mx, maxr(da, db));
    }
    real sep = absr(dist) - (ra + rb);
    if (sep > margin) return 0;
    if (mx > (real)0.9994) continue;      // edge axis within 2 degrees of a face normal: covered by the face test (see oracle)
    if (sep > ebest) { ebest = sep; ecode = 3 * i + j; real sg = dist >= 0 ? (real)1 : (real)-1; en[0] = sg * ax[0]; en[1] = sg * ax[1]; en[2] = sg * ax[2]; }
  }
  if (ecode >= 0 && ebest > best + (real)0.05 * absr(best) + (real)1e-6) {      // edges must beat faces by 5 % + 1 um (sign-symmetric)
    int i = ecode / 3, j = ecode % 3;
    real ca[3] = {pA[0], pA[1], pA[2]}, cb[3] = {pB[0], pB[1], pB[2]};
    for (int k = 0; k < 3; k++) {
      if (k != i) { real ak[3] = {RA[k], RA[3 + k], RA[6 + k]}; real s = dot3(en, ak) >= 0 ? (real)1 : (real)-1; for (int c = 0; c < 3; c++) ca[c] += s * hA[k] * ak[c]; }
      if (k != j) { real bk[3] = {RB[k], RB[3 + k], RB[6 + k]}; real s = dot3(en, bk) >= 0 ? (real)-1 : (real)1; for (int c = 0; c < 3; c++) cb[c] += s * hB[k] * bk[c]; }
    }
    real ua[3] = {RA[i], RA[3 + i], RA[6 + i]}, ub[3] = {RB[j], RB[3 + j], RB[6 + j]}, ww[3] = {ca[0] - cb[0], ca[1] - cb[1], ca[2] - cb[2]};
    real b = dot3(ua, ub), dd = dot3(ua, ww), ee = dot3(ub, ww), den = 1 - b * b;
    real sa = den > (real)1e-12 ? (b * ee - dd) / den : 0, sb = den > (real)1e-12 ? (ee - b * dd) / den : 0;
    sa = clampr(sa, -hA[i], hA[i]); sb = clampr(sb, -hB[j], hB[j]);
    for (int k = 0; k < 3; k++) { out[0].pos[k] = (real)0.5 * (ca[k] + sa * ua[k] + cb[k] + sb * ub[k]); out[0].n[k] = en[k]; }
    out[0].dist = ebest;
    return 1;
  }
  int refIsA = code < 3, ax = refIsA ? code : code - 3;
  const real *pR_ = refIsA ? pA : pB, *RR = refIsA ? RA : RB, *hR = refIsA ? hA : hB;
  const real *pI = refIsA ? pB : pA, *RI = refIsA ? RB : RA, *hI = refIsA ? hB : hA;
  real nref[3], sgn = refIsA ? bsign : -bsign;
  for (int k = 0; k < 3; k++) nref[k] = sgn * RR[3 * k + ax];
  int jx = 0; real jb = -1;
  for (int j = 0; j < 3; j++) { real c[3] = {RI[j], RI[3 + j], RI[6 + j]}; real v = absr(dot3(nref, c)); if (v > jb) { jb = v; jx = j; } }
  real cI[3] = {RI[jx], RI[3 + jx], RI[6 + jx]};
  real isg = dot3(nref, cI) > 0 ? (real)-1 : (real)1;
  int k1 = (jx + 1) % 3, k2 = (jx + 2) % 3;
  if (k1 > k2) { int tmp = k1; k1 = k2; k2 = tmp; }
  int u = (ax + 1) % 3, v = (ax + 2) % 3;
  if (u > v) { int tmp = u; u = v; v = tmp; }
  real P[12][3];
  for (int c = 0; c < 4; c++) {
    real s1 = (c == 1 || c == 2) ? (real)1 : (real)-1, s2 = (c >= 2) ? (real)1 : (real)-1;
    real ww[3], wl[3];
    for (int k = 0; k < 3; k++) ww[k] = pI[k] + isg * hI[jx] * RI[3 * k + jx] + s1 * hI[k1] * RI[3 * k + k1] + s2 * hI[k2] * RI[3 * k + k2] - pR_[k];
    matT_vec3(wl, RR, ww);
    P[c][0] = wl[u]; P[c][1] = wl[v]; P[c][2] = sgn * wl[ax] - hR[ax];
  }
  int n = clip_poly(P, 4, hR[u], hR[v]);
  int cnt = 0;
  for (int c = 0; c < n && cnt < 8; c++) {
    real dist = P[c][2];
    if (dist >= margin) continue;
    real wl[3]; wl[u] = P[c][0]; wl[v] = P[c][1]; wl[ax] = sgn * (hR[ax] + (real)0.5 * dist);
    real ww[3]; mat_vec3(ww, RR, wl);
    for (int k = 0; k < 3; k++) { out[cnt].pos[k] = pR_[k] + ww[k]; out[cnt].n[k] = refIsA ? nref[k] : -nref[k]; }
    out[cnt].dist = dist; cnt++;
  }
  return cnt;
}

DEVFN int zonotope_closest(const real d[2], const real g[3][2], real q[2], real nin[2]) {
  int inside = 1; real bestpen = (real)1e30;
  for (int k = 0; k < 3; k++) {
    real l = sqrt(g[k][0] * g[k][0] + g[k][1] * g[k][1]);
    if (l < (real)1e-12) continue;
    real nk[2] = {-g[k][1] / l, g[k][0] / l};
    real ww = 0;
    for (int j = 0; j < 3; j++) if (j != k) ww += absr(nk[0] * g[j][0] + nk[1] * g[j][1]);
    real s = d[0] * nk[0] + d[1] * nk[1];
    real sep = absr(s) - ww;
    if (sep > 0) inside = 0;
    if (-sep < bestpen) { bestpen = -sep; real sg = s >= 0 ? (real)1 : (real)-1; nin[0] = sg * nk[0]; nin[1] = sg * nk[1]; }
  }
  if (inside) return 0;
  real bd = (real)1e30;
  for (int k = 0; k < 3; k++) {
    real l2 = g[k][0] * g[k][0] + g[k][1] * g[k][1];
    real nk[2] = {-g[k][1], g[k][0]};
    for (int sgi = 0; sgi < 2; sgi++) {
      real sg = sgi ? (real)-1 : (real)1, p0[2] = {d[0], d[1]};
      for (int j = 0; j < 3; j++) if (j != k) { real c = nk[0] * g[j][0] + nk[1] * g[j][1]; real s = (c >= 0 ? (real)1 : (real)-1) * sg; p0[0] += s * g[j][0]; p0[1] += s * g[j][1]; }
      real lam = l2 > (real)1e-24 ? clampr(-(p0[0] * g[k][0] + p0[1] * g[k][1]) / l2, -1, 1) : 0;
      real c[2] = {p0[0] + lam * g[k][0], p0[1] + lam * g[k][1]};
      real dd = c[0] * c[0] + c[1] * c[1];
      if (dd < bd) { bd = dd; q[0] = c[0]; q[1] = c[1]; }
    }
  }
  return 1;
}

DEVNI int collide_cyl_box(const real* c, const real* Rc, const real* sz, const real* b, const real* Rb, const real* e3, real margin, RawCon* out) {
  real r = sz[0], h = sz[1];
  real a[3] = {Rc[2], Rc[5], Rc[8]}, B[3][3], d[3];
  for (int k = 0; k < 3; k++) { B[k][0] = Rb[k]; B[k][1] = Rb[3 + k]; B[k][2] = Rb[6 + k]; d[k] = b[k] - c[k]; }
  real best = (real)1e30, n[3] = {0, 0, 0};
  real cand[5][3]; int nc = 0;
  for (int k = 0; k < 3; k++) { real s = dot3(B[k], d) >= 0 ? (real)1 : (real)-1; cand[nc][0] = s * B[k][0]; cand[nc][1] = s * B[k][1]; cand[nc][2] = s * B[k][2]; nc++; }
  { real s = dot3(a, d) >= 0 ? (real)1 : (real)-1; cand[nc][0] = s * a[0]; cand[nc][1] = s * a[1]; cand[nc][2] = s * a[2]; nc++; }
  {
    real u[3] = {Rc[0], Rc[3], Rc[6]}, ww[3] = {Rc[1], Rc[4], Rc[7]};
    real d2[2] = {dot3(d, u), dot3(d, ww)}, g[3][2], q[2] = {0, 0}, nin[2] = {0, 0};
    for (int k = 0; k < 3; k++) { g[k][0] = e3[k] * dot3(B[k], u); g[k][1] = e3[k] * dot3(B[k], ww); }
    if (zonotope_closest(d2, g, q, nin)) {
      real l = sqrt(q[0] * q[0] + q[1] * q[1]);
      if (l > (real)1e-12) { for (int k = 0; k < 3; k++) cand[nc][k] = (q[0] * u[k] + q[1] * ww[k]) / l; nc++; }
    } else {
      for (int k = 0; k < 3; k++) cand[nc][k] = nin[0] * u[k] + nin[1] * ww[k];
      nc++;
    }
  }
  for (int ci = 0; ci < nc; ci++) {
    const real* nx = cand[ci];
    real na = dot3(nx, a), perp = 1 - na * na;
    real hc = h * absr(na) + r * sqrt(perp > 0 ? perp : 0);
    real hb = e3[0] * absr(dot3(nx, B[0])) + e3[1] * absr(dot3(nx, B[1])) + e3[2] * absr(dot3(nx, B[2]));
    real ov = hc + hb - dot3(nx, d);
    if (ov < -margin) return 0;
    if (ov < best - (real)1e-9) { best = ov; n[0] = nx[0]; n[1] = nx[1]; n[2] = nx[2]; }
  }
  // Contact point, continuous in the relative pose (same rule as the oracle): penetration-weighted centroid along the
  // side line through the supporting rim point, clipped to the box, moved towards the cap centre as the cap flattens.
  real na = dot3(n, a), s = na >= 0 ? (real)1 : (real)-1, pr[3], u[3] = {0, 0, 0}, pc[3];
  for (int k = 0; k < 3; k++) pr[k] = n[k] - na * a[k];
  real l = norm3(pr);
  if (l > (real)1e-12) for (int k = 0; k < 3; k++) u[k] = pr[k] / l;
  real t0 = -h, t1 = h, base[3]; int empty = 0;
  for (int k = 0; k < 3; k++) base[k] = c[k] + r * u[k] - b[k];
  for (int k = 0; k < 3; k++) {
    real x0 = dot3(base, B[k]), dx = dot3(a, B[k]);
    if (absr(dx) < (real)1e-12) { if (absr(x0) > e3[k]) empty = 1; continue; }
    real ta = (-e3[k] - x0) / dx, tb = (e3[k] - x0) / dx;
    if (ta > tb) { real tmp = ta; ta = tb; tb = tmp; }
    if (ta > t0) t0 = ta;
    if (tb < t1) t1 = tb;
  }
  if (t0 > t1) empty = 1;
  real ts;           // default (the side line misses the box: edge / corner contacts): midpoint of the axial overlap
  {
    real ca = dot3(d, a), ha = e3[0] * absr(dot3(a, B[0])) + e3[1] * absr(dot3(a, B[1])) + e3[2] * absr(dot3(a, B[2]));
    real lo = maxr(-h, ca - ha), hi = minr(h, ca + ha);
    ts = lo <= hi ? (real)0.5 * (lo + hi) : clampr(ca, -h, h);
  }
  if (!empty) {
    real da = best - absr(na) * (h - s * t0), db = best - absr(na) * (h - s * t1);
    if (da <= 0 && db <= 0) ts = db > da ? t1 : t0;
    else {
      if (da < 0) { t0 += (t1 - t0) * (-da) / (db - da); da = 0; }
      if (db < 0) { t1 -= (t1 - t0) * (-db) / (da - db); db = 0; }
      ts = t0 + (t1 - t0) * (da + 2 * db) / (3 * (da + db));
    }
  }
  ts = s * h + (ts - s * h) * l * l;        // side contacts (l -> 1) take the centroid, a flat cap (l -> 0) keeps its plane
  real dc = best - r * l, wcap = 1;
  if (dc > 0) { real w = r * l / (4 * dc); if (w < 1) wcap = w; }
  for (int k = 0; k < 3; k++) pc[k] = c[k] + ts * a[k] + r * wcap * u[k];
  for (int k = 0; k < 3; k++) { out->pos[k] = pc[k] - (real)0.5 * best * n[k]; out->n[k] = n[k]; }
  out->dist = -best;
  return out->dist < margin;
}

DEVNI int collide_cyl_cyl(const real* c1, const real* R1, const real* s1, const real* c2, const real* R2, const real* s2, real margin, RawCon* out) {
  real a1[3] = {R1[2], R1[5], R1[8]}, a2[3] = {R2[2], R2[5], R2[8]}, ww[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
  real b = dot3(a1, a2), d = dot3(a1, ww), e = dot3(a2, ww), den = 1 - b * b, t1, t2;
  if (den > (real)1e-8) {
    t1 = clampr((b * e - d) / den, -s1[1], s1[1]);
    t2 = clampr(e + b * t1, -s2[1], s2[1]);
    t1 = clampr(-d + b * t2, -s1[1], s1[1]);
  } else {
    real l2 = (-s2[1] - e) / b, h2 = (s2[1] - e) / b;
    if (l2 > h2) { real tmp = l2; l2 = h2; h2 = tmp; }
    real lo = maxr(-s1[1], l2), hi = minr(s1[1], h2);
    t1 = lo > hi ? clampr((real)0.5 * (lo + hi), -s1[1], s1[1]) : (real)0.5 * (lo + hi);
    t2 = clampr(e + b * t1, -s2[1], s2[1]);
  }
  real p1[3], dv[3];
  for (int k = 0; k < 3; k++) { p1[k] = c1[k] + t1 * a1[k]; dv[k] = c2[k] + t2 * a2[k] - p1[k]; }
  real l = norm3(dv), dist = l - s1[0] - s2[0];
  if (dist >= margin || l < (real)1e-12) return 0;
  for (int k = 0; k < 3; k++) { out->n[k] = dv[k] / l; out->pos[k] = p1[k] + out->n[k] * (s1[0] + (real)0.5 * dist); }
  out->dist = dist;
  return 1;
}

DEVFN void make_frame(real* f) {
  real y[3] = {0, 0, 0};
  if (f[1] < (real)0.5 && f[1] > (real)-0.5) y[1] = 1; else y[2] = 1;
  real t = dot3(f, y);
  y[0] -= t * f[0]; y[1] -= t * f[1]; y[2] -= t * f[2];
  real l = 1 / norm3(y);
  f[3] = y[0] * l; f[4] = y[1] * l; f[5] = y[2] * l;
  cross3(f + 6, f, f + 3);
}

DEVFN void geom_pose(const Model& m, const Lay& L, const real* w, int g, real* p, real* R) {
  const tab_t* gm = m.geom + D3_GEOM_W * g;
  int li = (int)gm[1];
  real gp[3] = {(real)gm[2], (real)gm[3], (real)gm[4]}, gR[9];
  for (int k = 0; k < 9; k++) gR[k] = m.geomR[9 * g + k];
  if (li < 0) { p[0] = gp[0]; p[1] = gp[1]; p[2] = gp[2]; for (int k = 0; k < 9; k++) R[k] = gR[k]; return; }
  real o[3]; mat_vec3(o, w + L.xmat + 9 * li, gp);
  for (int k = 0; k < 3; k++) p[k] = w[L.xpos + 3 * li + k] + o[k];
  mat_mul3(R, w + L.xmat + 9 * li, gR);
}

// One lane per candidate pair; contact slots are assigned in pair order through a count table so the contact list is
// deterministic.  Returns (all lanes) the number of contacts; sets the obstacle flag in misc.
template <int G>
DEVFN int collision(const Cx& cx, const Model& m, const Lay& L, real* w) {
  RawCon rc[8];
  int myn = 0, mypair = -1;
  // NOTE: npair <= G is required for the single-pass scheme below (checked on the host); pairs beyond G use more passes.
  int ntot = 0, obst = 0, near = 0;
  for (int base = 0; base < m.npair; base += G) {
    int ip = base + cx.lane;
    myn = 0; mypair = -1;
    if (ip < m.npair) {
      const tab_t* pr = m.pair + D3_PAIR_W * ip;
      int g1 = (int)pr[0], g2 = (int)pr[1];
      const tab_t *ga = m.geom + D3_GEOM_W * g1, *gb = m.geom + D3_GEOM_W * g2;
      real p1[3], R1[9], p2[3], R2[9];
      geom_pose(m, L, w, g1, p1, R1); geom_pose(m, L, w, g2, p2, R2);
      real dc[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, margin = pr[15];
      if (norm3(dc) <= (real)ga[12] + (real)gb[12] + margin) {
        if (m.p_cpl[ip]) near = 1;
        real s1[3] = {(real)ga[9], (real)ga[10], (real)ga[11]}, s2[3] = {(real)gb[9], (real)gb[10], (real)gb[11]};
        int t1 = (int)ga[0], t2 = (int)gb[0];
        if (t1 == D3G_BOX && t2 == D3G_BOX) {
          myn = m.g_slab[g1] ? collide_slab_box(p1, s1, p2, R2, s2, (real)gb[12], margin, rc) : -1;
          if (myn < 0) myn = collide_box_box(p1, R1, s1, p2, R2, s2, margin, rc);
        }
        else if (t1 == D3G_CYLINDER && t2 == D3G_BOX) myn = collide_cyl_box(p1, R1, s1, p2, R2, s2, margin, rc);
        else if (t1 == D3G_CYLINDER && t2 == D3G_CYLINDER) myn = collide_cyl_cyl(p1, R1, s1, p2, R2, s2, margin, rc);
      }
      mypair = ip;
      w[L.ncon_pair + ip - base] = (real)myn;
    }
    gsync<G>(cx);
    int cnt_here = m.npair - base < G ? m.npair - base : G;
    if (mypair >= 0) {
      int off = ntot;
      for (int j = 0; j < ip - base; j++) off += (int)w[L.ncon_pair + j];
      const tab_t* pr = m.pair + D3_PAIR_W * ip;
      for (int i = 0; i < myn; i++) {
        if (off + i >= m.maxcon) break;
        real* c = w + L.con + D3_CON_W * (off + i);
        c[0] = rc[i].pos[0]; c[1] = rc[i].pos[1]; c[2] = rc[i].pos[2];
        c[3] = rc[i].n[0]; c[4] = rc[i].n[1]; c[5] = rc[i].n[2];
        make_frame(c + 3);
        c[12] = rc[i].dist; c[13] = (real)pr[15] - (real)pr[16]; c[14] = 0; c[15] = pr[2];
        c[16] = pr[0]; c[17] = pr[1]; c[18] = (real)ip; c[19] = -1;
      }
      if (myn > 0 && (((int)pr[17]) & 1)) obst = 1;
    }
    for (int j = 0; j < cnt_here; j++) ntot += (int)w[L.ncon_pair + j];
    gsync<G>(cx);
  }
  obst = gori<G>(cx, obst); near = gori<G>(cx, near);
  LANES(z, 1) {
    w[L.misc + ST_OBST] = (real)obst;
    w[L.misc + ST_COST_NEAR] += (real)near;
    if (ntot > m.maxcon) w[L.misc + ST_STATUS] = (real)(((int)w[L.misc + ST_STATUS]) | 2);
  }
  if (ntot > m.maxcon) ntot = m.maxcon;
  return ntot;
}

// ------------------------------------------------------------------------------------------------ constraints
DEVFN real impedance(const real* solimp, real pos, real margin) {
  real dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (width < (real)1e-15 || dmin == dmax) return (real)0.5 * (dmin + dmax);
  real x = absr(pos - margin) / width;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  real y;
  if (power == 1) y = x;
  else if (power == 2) y = x <= mid ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);
#ifdef __CUDA_ARCH__
  else if (x <= mid) y = __powf(x, power) / __powf(mid, power - 1);      // not reached by the compiled scenes (power == 2)
  else y = 1 - __powf(1 - x, power) / __powf(1 - mid, power - 1);
#else
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
#endif
  return dmin + y * (dmax - dmin);
}

// dof range touched by a geom's link: robot links reach dofs [0, own dof]; free bodies their 6 dofs; static: empty
HDFN void link_range(const Model& m, int li, int* lo, int* hi) {
  if (li < 0) { *lo = 0; *hi = 0; return; }
  if (m.l_jtype[li] == 2) { *lo = m.l_dadr[li]; *hi = m.l_dadr[li] + 6; return; }
  *lo = m.d_bs[m.l_dadr[li]]; *hi = m.l_dadr[li] + 1;
}

// J row . v restricted to the contact's dof ranges
DEVFN real jrow_dot(const real* Jrow, const real* v, int a0, int a1, int b0, int b1) {
  real s = 0;
  for (int d = a0; d < a1; d++) s += Jrow[d - a0] * v[d];
  const real* Jb = Jrow + (a1 - a0) - b0;
  for (int d = b0; d < b1; d++) s += Jb[d] * v[d];
  return s;
}
// compact-row entry of dof d (must lie in one of the ranges)
DEVFN real jrow_at(const real* Jrow, int d, int a0, int a1, int b0) { return d < a1 && d >= a0 ? Jrow[d - a0] : Jrow[(a1 - a0) + d - b0]; }

// Rows: joint limits first (one dof each), then 3 rows per active contact (elliptic, condim 3).  Returns nefc (all
// lanes); *coupled is set when some active contact joins two different kinematic-tree blocks (then H is not block
// diagonal).  Row numbers come from prefix counts over activity flags, so the row order is deterministic.
template <int G, int MD>
DEVFN int make_constraints(const Cx& cx, const Model& m, const Lay& L, real* w, int ncon, int* nlimit, int* coupled) {
  const int nv = m.nv;
  // --- activity flags: joint limit candidates k = (robot link, side), contacts
  LANES(k, 2 * D3_NROB) {
    int i = k >> 1, side = k & 1;
    const tab_t* Lk = m.link + D3_LINK_W * i;
    real q = w[L.qpos + m.l_qadr[i]];
    real dist = side == 0 ? q - (real)Lk[23] : (real)Lk[24] - q;
    w[L.limflag + k] = (m.l_limited[i] && dist < 0) ? (real)1 : (real)0;
  }
  LANES(c, ncon) { const real* cc = w + L.con + D3_CON_W * c; w[L.cflag + c] = cc[12] < cc[13] ? (real)1 : (real)0; }
  gsync<G>(cx);
  int nl = 0;
  for (int k = 0; k < 2 * D3_NROB; k++) nl += (int)w[L.limflag + k];
  if (nl > m.maxrow) nl = m.maxrow;
  LANES(k, 2 * D3_NROB) {
    if (w[L.limflag + k] == 0) continue;
    int ne = 0;
    for (int j = 0; j < k; j++) ne += (int)w[L.limflag + j];
    if (ne >= m.maxrow) continue;
    int i = k >> 1, side = k & 1;
    const tab_t* Lk = m.link + D3_LINK_W * i;
    real q = w[L.qpos + m.l_qadr[i]];
    real dist = side == 0 ? q - (real)Lk[23] : (real)Lk[24] - q;
    real solref[2] = {(real)m.ctrl[D3C_JNT_SOLREF], (real)m.ctrl[D3C_JNT_SOLREF + 1]}, solimp[5];
    for (int kk2 = 0; kk2 < 5; kk2++) solimp[kk2] = m.ctrl[D3C_JNT_SOLIMP + kk2];
    real imp = impedance(solimp, dist, 0);
    real kk = 1 / (solimp[1] * solimp[1] * solref[0] * solref[0] * solref[1] * solref[1]), bb = 2 / (solimp[1] * solref[0]);
    real sg = side == 0 ? (real)1 : (real)-1;
    real Rv = maxr((real)1e-15, (1 - imp) / imp * (real)Lk[27]);
    w[L.D + ne] = 1 / Rv;
    w[L.aref + ne] = -bb * sg * w[L.qvel + m.l_dadr[i]] - kk * imp * dist;
    w[L.econ + ne] = (real)(side == 0 ? m.l_dadr[i] + 1 : -(m.l_dadr[i] + 1));     // signed dof id (+lower / -upper)
  }
  *nlimit = nl;
  int row = nl, cpl = 0, overflow = 0;
  LANES(c, ncon) {
    real* cc = w + L.con + D3_CON_W * c;
    if (w[L.cflag + c] == 0) { cc[19] = -1; continue; }
    int r0 = nl;
    if (MD == 3) { int idx = 0; for (int j = 0; j < c; j++) idx += (int)w[L.cflag + j]; r0 += 3 * idx; }
    else for (int j = 0; j < c; j++) if (w[L.cflag + j] != 0) r0 += (int)w[L.con + D3_CON_W * j + 15];
    if (r0 + (MD == 3 ? 3 : (int)cc[15]) > m.maxrow) { cc[19] = -1; overflow = 1; continue; }
    int ip = (int)cc[18];
    cc[19] = (real)r0; cc[20] = (real)m.p_rng[4 * ip]; cc[21] = (real)m.p_rng[4 * ip + 1]; cc[22] = (real)m.p_rng[4 * ip + 2]; cc[23] = (real)m.p_rng[4 * ip + 3];
    if (m.p_cpl[ip]) cpl = 1;
  }
  for (int c = 0; c < ncon; c++) { int dim = MD == 3 ? 3 : (int)w[L.con + D3_CON_W * c + 15]; if (w[L.cflag + c] != 0 && row + dim <= m.maxrow) row += dim; }
  cpl = gori<G>(cx, cpl); overflow = gori<G>(cx, overflow);
  if (overflow) { LANES(z, 1) w[L.misc + ST_STATUS] = (real)(((int)w[L.misc + ST_STATUS]) | 2); }
  *coupled = cpl;
  gsync<G>(cx);
  const real impratio = m.ctrl[D3C_IMPRATIO];
  // Jacobian rows: one (contact, dof) item per lane step; entries outside the contact's dof ranges are never read
  LANES(item, ncon * D3_JW) {
    int c = item / D3_JW, loc = item - c * D3_JW;                   // one compact-row slot per lane step (D3_JW = 16: shifts)
    const real* cc = w + L.con + D3_CON_W * c;
    int row0 = (int)cc[19];
    if (row0 < 0) continue;
    int a0 = (int)cc[20], a1 = (int)cc[21], b0 = (int)cc[22], b1 = (int)cc[23];
    if (loc >= (a1 - a0) + (b1 - b0)) continue;
    int d = loc < a1 - a0 ? a0 + loc : b0 + loc - (a1 - a0);
    int dim = (int)cc[15];
    int l1 = (int)m.geom[D3_GEOM_W * (int)cc[16] + 1], l2 = (int)m.geom[D3_GEOM_W * (int)cc[17] + 1];
    int dl = m.d_link[d];
    int in1 = l1 >= 0 && ((m.l_anc[l1] >> dl) & 1u), in2 = l2 >= 0 && ((m.l_anc[l2] >> dl) & 1u);
    real jd[3] = {0, 0, 0}, jr[3] = {0, 0, 0};
    if (in1 != in2) {
      const real* S = w + L.S + 6 * d;
      const real* pr = w + L.xpos + 3 * m.l_ref[dl];
      real rr[3] = {cc[0] - pr[0], cc[1] - pr[1], cc[2] - pr[2]}, wxr[3];
      cross3(wxr, S, rr);
      real sg = in2 ? (real)1 : (real)-1;
      jd[0] = sg * (S[3] + wxr[0]); jd[1] = sg * (S[4] + wxr[1]); jd[2] = sg * (S[5] + wxr[2]);
      jr[0] = sg * S[0]; jr[1] = sg * S[1]; jr[2] = sg * S[2];
    }
    for (int r = 0; r < dim; r++) {
      const real* ax = cc + 3 + 3 * (r < 3 ? r : 0);
      w[L.J + (row0 + r) * D3_JW + loc] = r < 3 ? dot3(ax, jd) : dot3(ax, jr);
    }
  }
  gsync<G>(cx);
  // per-contact reference acceleration / regularisation
  LANES(c, ncon) {
    real* cc = w + L.con + D3_CON_W * c;
    int row0 = (int)cc[19];
    if (row0 < 0) continue;
    int dim = (int)cc[15];
    const tab_t* pr = m.pair + D3_PAIR_W * (int)cc[18];
    real solref[2] = {(real)pr[8], (real)pr[9]}, solimp[5] = {(real)pr[10], (real)pr[11], (real)pr[12], (real)pr[13], (real)pr[14]};
    real fr[5] = {(real)pr[3], (real)pr[4], (real)pr[5], (real)pr[6], (real)pr[7]};
    real imp = impedance(solimp, cc[12], cc[13]);
    real kk = 1 / (solimp[1] * solimp[1] * solref[0] * solref[0] * solref[1] * solref[1]), bb = 2 / (solimp[1] * solref[0]);
    real tran = (real)m.geom[D3_GEOM_W * (int)cc[16] + 13] + (real)m.geom[D3_GEOM_W * (int)cc[17] + 13];
    real R0 = maxr((real)1e-15, (1 - imp) / imp * tran), R1 = R0 / impratio;
    cc[14] = fr[0] * sqrt(R1 / R0);
    for (int r = 0; r < dim; r++) {
      real Rv = r == 0 ? R0 : (r == 1 ? R1 : R1 * fr[0] * fr[0] / (fr[r - 1] * fr[r - 1]));
      w[L.D + row0 + r] = 1 / Rv;
      w[L.aref + row0 + r] = -(r == 0 ? kk * imp * (cc[12] - cc[13]) : 0);       // the damping term -b (J qvel) is added per row below
      ((unsigned char*)(w + L.rowc))[row0 + r] = (unsigned char)(c | (r << 6));
    }
  }
  gsync<G>(cx);
  LANES(r, row) {      // one lane per contact row: aref -= b * (J qvel)_r
    if (r < nl) continue;
    const real* cc = w + L.con + D3_CON_W * (((const unsigned char*)(w + L.rowc))[r] & 63);
    const tab_t* pr = m.pair + D3_PAIR_W * (int)cc[18];
    const real bb = 2 / ((real)pr[11] * (real)pr[8]);
    w[L.aref + r] -= bb * jrow_dot(w + L.J + r * D3_JW, w + L.qvel, (int)cc[20], (int)cc[21], (int)cc[22], (int)cc[23]);
  }
  // per-block lists of the active contacts that touch the block (bit 7: the contact also touches another block): the Newton
  // loop's gradient and Hessian assembly walk these instead of scanning every contact for every dof / matrix entry
  LANES(b, m.nblk) {
    unsigned char* bl = (unsigned char*)(w + L.blist) + b * (m.maxcon + 1);
    const int bs = m.blk_s[b], be = m.blk_e[b];
    int n = 0;
    for (int c = 0; c < ncon; c++) {
      const real* cc = w + L.con + D3_CON_W * c;
      if ((int)cc[19] < 0) continue;
      const int a0 = (int)cc[20], b0 = (int)cc[22], b1 = (int)cc[23], cpl = b1 > b0;
      if ((a0 >= bs && a0 < be) || (cpl && b0 >= bs && b0 < be)) bl[1 + n++] = (unsigned char)(c | (cpl ? 0x80 : 0));
    }
    bl[0] = (unsigned char)n;
  }
  gsync<G>(cx);
  return row;
}

// jar = J v - aref for every row (v: workspace offset of a dof vector).  One lane per constraint row.
template <int G>
DEVNI void eval_jar(const Cx& cx, const Model& m, const Lay& L, real* w, int nlimit, int ne, int v_off, int out_off, bool sub_aref) {
  const unsigned char* rowc = (const unsigned char*)(w + L.rowc);
  LANES(r, ne) {
    real s;
    if (r < nlimit) {
      int sd = (int)w[L.econ + r];
      s = sd > 0 ? w[v_off + sd - 1] : -w[v_off - sd - 1];
    } else {
      const real* cc = w + L.con + D3_CON_W * (rowc[r] & 63);
      s = jrow_dot(w + L.J + r * D3_JW, w + v_off, (int)cc[20], (int)cc[21], (int)cc[22], (int)cc[23]);
    }
    w[out_off + r] = sub_aref ? s - w[L.aref + r] : s;
  }
  gsync<G>(cx);
}

// Constraint cost at jar (workspace offset jar_off): forces -> frc_off; HESS: per-limit-row curvature in hd and a 3x3
// block per contact in hb (zero / diagonal / full for the top / bottom / middle zone of the elliptic cone).
// Returns the group-wide cost.  One lane per limit row / per contact.
template <int G, bool HESS, int MD>
DEVNI real constraint_eval(const Cx& cx, const Model& m, const Lay& L, real* w, int nlimit, int ncon, int jar_off, int frc_off) {
  real cost = 0;
  LANES(i, nlimit) {
    real j = w[jar_off + i], Dv = w[L.D + i];
    if (j < 0) { cost += (real)0.5 * Dv * j * j; w[frc_off + i] = -Dv * j; if (HESS) w[L.hd + i] = Dv; }
    else { w[frc_off + i] = 0; if (HESS) w[L.hd + i] = 0; }
  }
  LANES(c, ncon) {
    const real* cc = w + L.con + D3_CON_W * c;
    int i = (int)cc[19];
    if (i < 0) continue;
    const int dim = MD == 3 ? 3 : (int)cc[15];
    const tab_t* pr = m.pair + D3_PAIR_W * (int)cc[18];
    real mu = cc[14], U[MD], fr[MD - 1], T2 = 0;
    U[0] = w[jar_off + i] * mu;
#pragma unroll
    for (int j = 1; j < MD; j++) { fr[j - 1] = (real)pr[2 + j]; U[j] = j < dim ? w[jar_off + i + j] * fr[j - 1] : (real)0; T2 += U[j] * U[j]; }
    real T = sqrt(T2), N = U[0];
    real HB[MD * MD];
#pragma unroll
    for (int k = 0; k < MD * MD; k++) HB[k] = 0;
    if (N >= mu * T || (T <= 0 && N >= 0)) {
#pragma unroll
      for (int j = 0; j < MD; j++) if (j < dim) w[frc_off + i + j] = 0;
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
#pragma unroll
      for (int j = 0; j < MD; j++) if (j < dim) { real Dv = w[L.D + i + j], jj = w[jar_off + i + j]; cost += (real)0.5 * Dv * jj * jj; w[frc_off + i + j] = -Dv * jj; HB[(MD + 1) * j] = Dv; }
    } else {
      real Dm = w[L.D + i] / maxr((real)1e-15, mu * mu * (1 + mu * mu)), NmT = N - mu * T;
      cost += (real)0.5 * Dm * NmT * NmT;
      real f0 = -Dm * NmT * mu;
      w[frc_off + i] = f0;
#pragma unroll
      for (int j = 1; j < MD; j++) if (j < dim) w[frc_off + i + j] = -f0 / T * U[j] * fr[j - 1];
      if (HESS) {
        real g[MD]; g[0] = mu;
#pragma unroll
        for (int j = 1; j < MD; j++) g[j] = -mu * fr[j - 1] * U[j] / T;         // rows >= dim: U = 0
#pragma unroll
        for (int a = 0; a < MD; a++)
#pragma unroll
          for (int b2 = 0; b2 < MD; b2++) {
            if (a >= dim || b2 >= dim) continue;
            real v = g[a] * g[b2];
            if (a > 0 && b2 > 0) v -= mu * NmT / T * fr[a - 1] * fr[b2 - 1] * ((a == b2 ? (real)1 : (real)0) - U[a] * U[b2] / (T * T));
            HB[MD * a + b2] = Dm * v;
          }
      }
    }
    if (HESS) {
#pragma unroll
      for (int k = 0; k < MD * MD; k++) w[L.hb + MD * MD * c + k] = HB[k];
    }
  }
  return gsum<G>(cx, cost);
}

// Cholesky of a symmetric matrix that is block diagonal over the partition cells [ps(i), pe(i)) (all cells advance in
// lock-step, one column per step).  The strict lower triangle of A (leading dimension n) holds the input and receives
// L; the running pivots live in piv[] (initialised by the caller with the diagonal), dinv[] receives 1 / L_kk.
// Row i is owned by lane i % G.  Returns 1 if a pivot was not positive.
template <int G>
DEVNI int chol_factor_part(const Cx& cx, const Model& m, real* A, int n, bool whole, int maxsz, real* piv, real* dinv, unsigned skip = 0u) {
  int bad = 0;
  const int nb = whole ? 1 : m.nblk;
  for (int k = 0; k < maxsz; k++) {
    LANES(i, n) {
      int c = (whole ? 0 : m.d_bs[i]) + k, e = whole ? n : m.d_be[i];
      if (c >= e || i < c) continue;
      real p = piv[c];
      if (!(p > 0)) { bad = 1; p = 1; }
      real inv = 1 / sqrt(p);
      if (i == c) dinv[c] = inv; else A[i * n + c] *= inv;
    }
    gsync<G>(cx);
    // rank-1 update of every block's trailing triangle, one (i, j) entry per lane step (lower-triangle table):
    // the critical path is ceil(entries / G) independent updates instead of a serial walk along the longest row
    for (int b = 0; b < nb; b++) {
      const int c = (whole ? 0 : m.blk_s[b]) + k, e = whole ? n : m.blk_e[b], mb = e - c - 1;
      if (mb <= 0 || ((skip >> b) & 1u)) continue;                   // diagonal blocks have no trailing update
      if (mb <= 24) {
        LANES(t, mb * (mb + 1) / 2) {
          const int i = c + 1 + m.tri_i[t], j = c + 1 + m.tri_j[t];
          const real v = A[i * n + c] * A[j * n + c];
          if (i == j) piv[i] -= v; else A[i * n + j] -= v;
        }
      } else {        // long rows (dense factorisation of a big scene): the unranking table stops at 24
        LANES(t, mb * mb) {
          const int li = t / mb, lj = t - li * mb;
          if (lj > li) continue;
          const int i = c + 1 + li, j = c + 1 + lj;
          const real v = A[i * n + c] * A[j * n + c];
          if (i == j) piv[i] -= v; else A[i * n + j] -= v;
        }
      }
    }
    gsync<G>(cx);
  }
  return gori<G>(cx, bad);
}

// Solve L L^T x = b in place for the same partitioned factor.  x is pulled into registers (element i in lane i % G),
// finished elements are broadcast with shuffles; no shared-memory traffic for x and no barriers inside the sweeps.
template <int G, int NS>
DEVNI void chol_solve_slots(const Cx& cx, const Model& m, const real* A, int n, bool whole, int maxsz, const real* dinv, real* x) {
  real xr[NS];
  int s0[NS], e0[NS];
#pragma unroll
  for (int sl = 0; sl < NS; sl++) {
    int i = sl * G + cx.lane, ii = i < n ? i : n - 1;
    xr[sl] = i < n ? x[i] : (real)0;
    s0[sl] = whole ? 0 : m.d_bs[ii]; e0[sl] = whole ? n : m.d_be[ii];
  }
  for (int k = 0; k < maxsz; k++) {
    real xcs[NS];          // read phase first: every slot sees the pre-step values of x
#pragma unroll
    for (int sl = 0; sl < NS; sl++) {
      int c = s0[sl] + k, cc = c < e0[sl] ? c : e0[sl] - 1;
      xcs[sl] = elem_bcast<G, NS>(cx, xr, cc) * dinv[cc];
    }
#pragma unroll
    for (int sl = 0; sl < NS; sl++) {
      int i = sl * G + cx.lane, c = s0[sl] + k;
      if (i < n && c < e0[sl]) { if (i == c) xr[sl] = xcs[sl]; else if (i > c) xr[sl] -= A[i * n + c] * xcs[sl]; }
    }
  }
  for (int k = maxsz - 1; k >= 0; k--) {
    real xcs[NS];
#pragma unroll
    for (int sl = 0; sl < NS; sl++) {
      int c = s0[sl] + k, cc = c < e0[sl] ? c : e0[sl] - 1;
      xcs[sl] = elem_bcast<G, NS>(cx, xr, cc) * dinv[cc];
    }
#pragma unroll
    for (int sl = 0; sl < NS; sl++) {
      int i = sl * G + cx.lane, c = s0[sl] + k;
      if (i < n && c < e0[sl]) { if (i == c) xr[sl] = xcs[sl]; else if (i < c && i >= s0[sl]) xr[sl] -= A[c * n + i] * xcs[sl]; }
    }
  }
#pragma unroll
  for (int sl = 0; sl < NS; sl++) { int i = sl * G + cx.lane; if (i < n) x[i] = xr[sl]; }
  gsync<G>(cx);
}

// ---- register-resident block Cholesky ------------------------------------------------------------------------------
// The matrices this path factors are block diagonal over the kinematic trees with blocks of at most D3_MAXB rows (the arm:
// 7 joints + 2 fingers; a free body: 6).  Row i of its block lives in the registers of lane i % G (slot i / G): the whole
// right-looking factorisation is shuffles + FMAs on registers, fully unrolled over the D3_MAXB column steps, all blocks
// advancing together — no shared-memory traffic and no barriers inside (the shared-memory version spent two barriers and
// several strided passes per column).  The factor goes back to the strict lower triangle of A (the triangular solves read
// it from there) and 1 / L_kk to dinv.
#define D3_MAXB 9
// value `v[slot]` of the lane/slot that owns row `src_row` (every lane may ask for a different row)
template <int G, int NS>
DEVFN real row_bcast(const Cx& cx, const real* v, int src_row) {
#ifdef D3IL_EMU
  return v[src_row];
#else
  if (NS == 1) return __shfl_sync(cx.mask, v[0], src_row, G);
  real out = 0;
#pragma unroll
  for (int sl = 0; sl < NS; sl++) { const real t = __shfl_sync(cx.mask, v[sl], src_row % G, G); if (sl == src_row / G) out = t; }
  return out;
#endif
}
// upper = true : A is the PACKED mass matrix (Model::m_row) and the input block rows are read from its UPPER triangle
//                (A[j][i], j < i), as CRBA leaves it;
// upper = false: A is a dense n x n buffer, rows from the lower triangle (the assembled Hessian).  diag[] holds the diagonal (nullptr: A's own), diag_add (nullable) is added to it.
template <int G, int NS> DEVFN int chol_reg_core(const Cx& cx, real (*a)[D3_MAXB], const int* bs, const int* be, const int* rr, real* myinv);
template <int G, int NS>
DEVNI int chol_blocks_reg_slots(const Cx& cx, const Model& m, real* A, int n, bool upper, const real* diag, const real* diag_add, real* dinv) {
  real a[NS][D3_MAXB], myinv[NS];
  int bs[NS], rr[NS], be[NS], base[NS], st[NS];
#pragma unroll
  for (int sl = 0; sl < NS; sl++) {
    const int i = sl * G + cx.lane, ii = i < n ? i : n - 1;
    bs[sl] = m.d_bs[ii]; be[sl] = m.d_be[ii]; rr[sl] = i < n ? i - bs[sl] : -1;      // rr < 0: no row in this slot
    // element (bs + r, bs + c) of the lane's block sits at base + r * st + c: the packed mass matrix, or a dense n x n buffer
    st[sl] = upper ? be[sl] - bs[sl] : n; base[sl] = upper ? (int)m.m_base[ii] : bs[sl] * (n + 1);
#pragma unroll
    for (int j = 0; j < D3_MAXB; j++) {
      real v = 0;
      if (j < rr[sl]) v = upper ? A[base[sl] + j * st[sl] + rr[sl]] : A[base[sl] + rr[sl] * st[sl] + j];
      else if (j == rr[sl]) v = (diag ? diag[i] : A[base[sl] + rr[sl] * st[sl] + rr[sl]]) + (diag_add ? diag_add[i] : (real)0);      // diag == nullptr: the buffer's own diagonal (the factor only writes the strict lower triangle)
      a[sl][j] = v;
    }
    myinv[sl] = 1;
  }
  const int bad = chol_reg_core<G, NS>(cx, a, bs, be, rr, myinv);
#pragma unroll
  for (int sl = 0; sl < NS; sl++) {
    const int i = sl * G + cx.lane;
    if (rr[sl] < 0) continue;
#pragma unroll
    for (int j = 0; j < D3_MAXB; j++) if (j < rr[sl]) A[base[sl] + rr[sl] * st[sl] + j] = a[sl][j];
    dinv[i] = myinv[sl];
  }
  gsync<G>(cx);
  return gori<G>(cx, bad);
}
template <int G>
DEVFN int chol_blocks_reg(const Cx& cx, const Model& m, real* A, int n, bool upper, const real* diag, const real* diag_add, real* dinv) {
#ifdef D3IL_EMU
  return chol_blocks_reg_slots<G, D3_SLOTS(G)>(cx, m, A, n, upper, diag, diag_add, dinv);
#else
  if (n <= G) return chol_blocks_reg_slots<G, 1>(cx, m, A, n, upper, diag, diag_add, dinv);
  return chol_blocks_reg_slots<G, D3_SLOTS(G)>(cx, m, A, n, upper, diag, diag_add, dinv);
#endif
}

// Block solve L L^T x = b for the same factor: the lane's row of L (forward sweep) and column of L (backward sweep) are
// loaded into registers up front, x is register-distributed, every step is one shuffle + one FMA.
template <int G, int NS>
DEVNI void chol_blocks_solve_slots(const Cx& cx, const Model& m, const real* A, int n, bool packed, const real* dinv, real* x) {
  real lrow[NS][D3_MAXB], lcol[NS][D3_MAXB], xr[NS], di[NS], mine[NS];
  int bs[NS], rr[NS], be[NS];
#pragma unroll
  for (int sl = 0; sl < NS; sl++) {
    const int i = sl * G + cx.lane, ii = i < n ? i : n - 1;
    bs[sl] = m.d_bs[ii]; be[sl] = m.d_be[ii]; rr[sl] = i < n ? i - bs[sl] : -1;
    xr[sl] = i < n ? x[i] : (real)0; di[sl] = dinv[ii];
    const int st = packed ? be[sl] - bs[sl] : n, base = packed ? (int)m.m_base[ii] : bs[sl] * (n + 1);
#pragma unroll
    for (int j = 0; j < D3_MAXB; j++) {
      lrow[sl][j] = j < rr[sl] ? A[base + rr[sl] * st + j] : (real)0;
      lcol[sl][j] = (rr[sl] >= 0 && j > rr[sl] && bs[sl] + j < be[sl]) ? A[base + j * st + rr[sl]] : (real)0;
    }
  }
#pragma unroll
  for (int k = 0; k < D3_MAXB; k++) {
#pragma unroll
    for (int sl = 0; sl < NS; sl++) mine[sl] = xr[sl] * di[sl];
#pragma unroll
    for (int sl = 0; sl < NS; sl++) {
      const int src = bs[sl] + k < be[sl] ? bs[sl] + k : bs[sl];
      const real xk = row_bcast<G, NS>(cx, mine, src);
      if (rr[sl] == k) xr[sl] = mine[sl]; else if (rr[sl] > k) xr[sl] -= lrow[sl][k] * xk;
    }
  }
#pragma unroll
  for (int k = D3_MAXB - 1; k >= 0; k--) {
#pragma unroll
    for (int sl = 0; sl < NS; sl++) mine[sl] = xr[sl] * di[sl];
#pragma unroll
    for (int sl = 0; sl < NS; sl++) {
      const int src = bs[sl] + k < be[sl] ? bs[sl] + k : bs[sl];
      const real xk = row_bcast<G, NS>(cx, mine, src);
      if (rr[sl] == k) xr[sl] = mine[sl]; else if (rr[sl] >= 0 && rr[sl] < k) xr[sl] -= lcol[sl][k] * xk;      // lcol is zero past the block's end
    }
  }
#pragma unroll
  for (int sl = 0; sl < NS; sl++) { const int i = sl * G + cx.lane; if (i < n) x[i] = xr[sl]; }
  gsync<G>(cx);
}
template <int G>
DEVFN void chol_blocks_solve(const Cx& cx, const Model& m, const real* A, int n, bool packed, const real* dinv, real* x) {
#ifdef D3IL_EMU
  chol_blocks_solve_slots<G, D3_SLOTS(G)>(cx, m, A, n, packed, dinv, x);
#else
  if (n <= G) chol_blocks_solve_slots<G, 1>(cx, m, A, n, packed, dinv, x);
  else chol_blocks_solve_slots<G, D3_SLOTS(G)>(cx, m, A, n, packed, dinv, x);
#endif
}

#ifndef D3IL_EMU
// The same block solve for NR right-hand sides at once (X: NR dof vectors, stride n), systems that fit one group (n <= G):
// the factor's row / column is loaded once and every elimination step is NR independent shuffle + FMA pairs.
template <int G, int NR>
DEVNI void chol_blocks_solve_multi(const Cx& cx, const Model& m, const real* A, int n, bool packed, const real* dinv, real* X, int nr) {
  const int i = cx.lane, ii = i < n ? i : n - 1;
  const int bs = m.d_bs[ii], be = m.d_be[ii], rr = i < n ? i - bs : -1;
  const int st = packed ? be - bs : n, base = packed ? (int)m.m_base[ii] : bs * (n + 1);
  const real di = dinv[ii];
  real lrow[D3_MAXB], lcol[D3_MAXB], xr[NR], mine[NR];
#pragma unroll
  for (int j = 0; j < D3_MAXB; j++) {
    lrow[j] = j < rr ? A[base + rr * st + j] : (real)0;
    lcol[j] = (rr >= 0 && j > rr && bs + j < be) ? A[base + j * st + rr] : (real)0;
  }
#pragma unroll
  for (int r = 0; r < NR; r++) xr[r] = (i < n && r < nr) ? X[r * n + i] : (real)0;
#pragma unroll
  for (int k = 0; k < D3_MAXB; k++) {
    const int src = bs + k < be ? bs + k : bs;
#pragma unroll
    for (int r = 0; r < NR; r++) mine[r] = xr[r] * di;
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const real xk = __shfl_sync(cx.mask, mine[r], src, G);
      if (rr == k) xr[r] = mine[r]; else if (rr > k) xr[r] -= lrow[k] * xk;
    }
  }
#pragma unroll
  for (int k = D3_MAXB - 1; k >= 0; k--) {
    const int src = bs + k < be ? bs + k : bs;
#pragma unroll
    for (int r = 0; r < NR; r++) mine[r] = xr[r] * di;
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const real xk = __shfl_sync(cx.mask, mine[r], src, G);
      if (rr == k) xr[r] = mine[r]; else if (rr >= 0 && rr < k) xr[r] -= lcol[k] * xk;
    }
  }
#pragma unroll
  for (int r = 0; r < NR; r++) if (i < n && r < nr) X[r * n + i] = xr[r];
  gsync<G>(cx);
}
#endif

#ifndef D3IL_EMU
// Dense n x n system (n <= G: one group, row i in lane i) factored AND solved with ROLLED loops on the shared-memory
// matrix: the tree-coupled Newton Hessian of a scene with nv <= 32 (two or more contacts join kinematic trees: a box pushed
// into another box).  Only a handful of envs are in that state at a time and each runs alone in a free-running CTA, where
// instruction fetch is not shared with other warps: 20 KB of unrolled register code cost 34 k cycles per call (one L2 fetch per
// 128-byte line), this loop nest is a few hundred bytes.  A: strict lower triangle (leading dimension n, odd strides are
// bank-conflict free), diag[] its diagonal, x: right-hand side in, solution out.  Returns 1 (all lanes) on a non-positive pivot.
template <int G>
DEVNI int chol_dense_rolled_solve(const Cx& cx, const Model& m, real* A, int n, const real* diag, real* x) {
  const int i = cx.lane;
  const bool mine = i < n;
  real* row = A + (mine ? i : 0) * n;
  real dii = mine ? diag[i] : (real)1, myinv = 1;
  int bad = 0;
  if (n <= 24) {
    // entry-parallel trailing update (the unranking table covers triangles up to 24 rows): per column one round to scale it and
    // ceil(m (m + 1) / 64) rounds over the m x m trailing triangle, instead of m dependent steps per lane
    real* d = const_cast<real*>(diag);
    for (int k = 0; k < n; k++) {
      real p = d[k];
      if (!(p > 0)) { bad = 1; p = 1; }
      const real inv = 1 / sqrt(p);
      if (i == k) myinv = inv;
      if (mine && i > k) row[k] *= inv;
      gsync<G>(cx);
      const int mm = n - k - 1;
      for (int e = i; e < mm * (mm + 1) / 2; e += G) {
        const int r = k + 1 + m.tri_i[e], c = k + 1 + m.tri_j[e];
        const real v = A[r * n + k] * A[c * n + k];
        if (r == c) d[r] -= v; else A[r * n + c] -= v;
      }
      gsync<G>(cx);
    }
  } else
  for (int k = 0; k < n; k++) {
    if (i == k) { real p = dii; if (!(p > 0)) { bad = 1; p = 1; } myinv = 1 / sqrt(p); }
    const real inv = __shfl_sync(cx.mask, myinv, k, G);
    real colk = 0;
    if (mine && i > k) { colk = row[k] * inv; row[k] = colk; }
    // trailing update, column by column: lane i owns row i (entries j <= i), L_jk comes from lane j
    for (int j = k + 1; j < n; j++) {
      const real ljk = __shfl_sync(cx.mask, colk, j, G);
      if (mine && i > j) row[j] -= colk * ljk;
      else if (i == j) dii -= colk * ljk;
    }
  }
  gsync<G>(cx);
  real xr = mine ? x[i] : (real)0;
  for (int k = 0; k < n; k++) {
    const real xk = __shfl_sync(cx.mask, xr * myinv, k, G);
    if (i == k) xr = xk; else if (mine && i > k) xr -= row[k] * xk;
  }
  for (int k = n - 1; k >= 0; k--) {
    const real xk = __shfl_sync(cx.mask, xr * myinv, k, G);
    if (i == k) xr = xk; else if (i < k) xr -= A[k * n + i] * xk;
  }
  if (mine) x[i] = xr;
  gsync<G>(cx);
  return gori<G>(cx, bad);
}
#endif

// NS = register slots per lane for the distributed vector: one when the system fits the group (n <= G)
template <int G>
DEVFN void chol_solve_part(const Cx& cx, const Model& m, const real* A, int n, bool whole, int maxsz, const real* dinv, real* x) {
#ifdef D3IL_EMU
  chol_solve_slots<G, D3_SLOTS(G)>(cx, m, A, n, whole, maxsz, dinv, x);
#else
  if (n <= G) chol_solve_slots<G, 1>(cx, m, A, n, whole, maxsz, dinv, x);
  else chol_solve_slots<G, D3_SLOTS(G)>(cx, m, A, n, whole, maxsz, dinv, x);
#endif
}

// y_d = sum_k M[d][k] v[k] within the dof's block (M stored in the upper triangle + diagonal of the M buffer)
DEVFN real mrow_dot(const Model& m, const real* M, int nv, int d, const real* v, const real* vsub) {
  real s = 0;
  const int rd = m.m_row[d];
  for (int k = m.d_bs[d]; k < m.d_be[d]; k++) {
    real mk = k >= d ? M[rd + k] : M[m.m_row[k] + d];
    s += mk * (vsub ? v[k] - vsub[k] : v[k]);
  }
  return s;
}

// Number of active contacts that couple two kinematic trees (all lanes); *cplc = index of the last one (exact when the
// count is 1, the only case that uses it).
template <int G>
DEVFN int count_coupling(const Cx& cx, const Model& m, const Lay& L, const real* w, int ncon, int* cplc) {
  int nc = 0, last = -1;
  LANES(c, ncon) { const real* cc = w + L.con + D3_CON_W * c; if ((int)cc[19] >= 0 && (int)cc[23] > (int)cc[22]) { nc++; last = c; } }
  nc = gsumi<G>(cx, nc);
  int lm = last + 1; lm = gori<G>(cx, lm);
  *cplc = lm - 1;
  return nc;
}

// core of the register factorisation (shared by chol_blocks_reg_slots and the Newton direction): rows in a[][], in place
template <int G, int NS>
DEVFN int chol_reg_core(const Cx& cx, real (*a)[D3_MAXB], const int* bs, const int* be, const int* rr, real* myinv) {
  real colk[NS];
  int bad = 0;
#pragma unroll
  for (int k = 0; k < D3_MAXB; k++) {
#pragma unroll
    for (int sl = 0; sl < NS; sl++) if (rr[sl] == k) {
      real p = a[sl][k];
      if (!(p > 0)) { bad = 1; p = 1; }
      myinv[sl] = 1 / sqrt(p);
    }
#pragma unroll
    for (int sl = 0; sl < NS; sl++) {
      const int src = bs[sl] + k < be[sl] ? bs[sl] + k : bs[sl];
      const real inv = row_bcast<G, NS>(cx, myinv, src);
      if (rr[sl] > k) a[sl][k] *= inv;
      colk[sl] = a[sl][k];
    }
#pragma unroll
    for (int j = k + 1; j < D3_MAXB; j++) {
#pragma unroll
      for (int sl = 0; sl < NS; sl++) {
        const int src = bs[sl] + j < be[sl] ? bs[sl] + j : bs[sl];
        const real ljk = row_bcast<G, NS>(cx, colk, src);
        if (rr[sl] >= j) a[sl][j] -= a[sl][k] * ljk;
      }
    }
  }
  return bad;
}

// Exact line search of the Newton step: safeguarded 1-D Newton / false position on phi'(alpha), phi = cost along
// qacc + alpha p.  Everything a row contributes is linear in alpha before the cone projection, so each lane loads the data
// of its limit row / contact ONCE into registers (NSL contact slots, NLL limit slots per lane) and an evaluation is
// arithmetic plus one fused all-reduce — no shared-memory traffic inside the loop.  d0 = phi'(0) (= grad . p),
// pMa = p . M (a - a_s), pMp = p . M p.  Rows: jar (at alpha = 0) and Jp = J p in the workspace.
template <int G, int MD, int NSL, int NLL>
DEVNI real line_search(const Cx& cx, const Model& m, const Lay& L, const real* w, int nlimit, int ncon, real d0, real pMa, real pMp) {
  real lj0[NLL], ljp[NLL], lDjp[NLL];
#pragma unroll
  for (int sl = 0; sl < NLL; sl++) {
    const int i = sl * G + cx.lane;
    lj0[sl] = 0; ljp[sl] = 0; lDjp[sl] = 0;
    if (i < nlimit) { lj0[sl] = w[L.jar + i]; ljp[sl] = w[L.Jp + i]; lDjp[sl] = w[L.D + i] * ljp[sl]; }
  }
  real U0[NSL][MD], V[NSL][MD], VV[NSL], A1[NSL], A2[NSL], Dm[NSL], mu[NSL];
  bool act[NSL];
#pragma unroll
  for (int sl = 0; sl < NSL; sl++) {
    const int c = sl * G + cx.lane;
    act[sl] = false; VV[sl] = 0; A1[sl] = 0; A2[sl] = 0; Dm[sl] = 0; mu[sl] = 0;
#pragma unroll
    for (int j = 0; j < MD; j++) { U0[sl][j] = 0; V[sl][j] = 0; }
    if (c < ncon) {
      const real* cc = w + L.con + D3_CON_W * c;
      const int i = (int)cc[19];
      if (i >= 0) {
        act[sl] = true;
        const tab_t* pr = m.pair + D3_PAIR_W * (int)cc[18];
        const int dim = MD == 3 ? 3 : (int)cc[15];
        mu[sl] = cc[14];
#pragma unroll
        for (int j = 0; j < MD; j++) {
          if (j >= dim) continue;
          const real fj = j == 0 ? mu[sl] : (real)pr[2 + j], ja = w[L.jar + i + j], jp = w[L.Jp + i + j], Dv = w[L.D + i + j];
          U0[sl][j] = ja * fj; V[sl][j] = jp * fj;
          if (j > 0) VV[sl] += V[sl][j] * V[sl][j];
          A1[sl] += Dv * ja * jp; A2[sl] += Dv * jp * jp;
        }
        Dm[sl] = w[L.D + i] / maxr((real)1e-15, mu[sl] * mu[sl] * (1 + mu[sl] * mu[sl]));
      }
    }
  }
  real lo = 0, hi = -1, alpha = 1, dlo = d0, dhi = 0;
  for (int ls = 0; ls < 20; ls++) {
    // first and second directional derivatives at jar + alpha Jp
    real d1p = 0, d2p = 0;
#pragma unroll
    for (int sl = 0; sl < NLL; sl++) {
      const real j = lj0[sl] + alpha * ljp[sl];
      if (j < 0) { d1p += lDjp[sl] * j; d2p += lDjp[sl] * ljp[sl]; }
    }
#pragma unroll
    for (int sl = 0; sl < NSL; sl++) {
      if (!act[sl]) continue;
      real U[MD], UV = 0, T2 = 0;
#pragma unroll
      for (int j = 0; j < MD; j++) { U[j] = U0[sl][j] + alpha * V[sl][j]; if (j > 0) { T2 += U[j] * U[j]; UV += U[j] * V[sl][j]; } }
      const real T = sqrt(T2), N = U[0], mus = mu[sl];
      if (N >= mus * T || (T <= 0 && N >= 0)) {
      } else if (mus * N + T <= 0 || (T <= 0 && N < 0)) {
        d1p += A1[sl] + alpha * A2[sl]; d2p += A2[sl];
      } else {
        // s = 0.5 Dm (N - mu T)^2 along the line: dN = V0, dT = (U_t . V_t)/T, d2T = (|V_t|^2 - dT^2)/T
        const real NmT = N - mus * T, dT = UV / T, d2T = (VV[sl] - dT * dT) / T, dn = V[sl][0] - mus * dT;
        d1p += Dm[sl] * NmT * dn;
        d2p += Dm[sl] * (dn * dn - NmT * mus * d2T);
      }
    }
    gsum2<G>(cx, d1p, d2p);
    const real d1 = d1p + pMa + alpha * pMp, d2 = d2p + pMp;
#if defined(D3IL_PHASE_TIMING) && defined(__CUDA_ARCH__)
    if (threadIdx.x == 0 && blockIdx.x < 4096) g_cta_stat[4 * blockIdx.x + 3] += 1;
    if (blockIdx_is0()) atomicAdd(&g_phase_cycles[18], 1ull);
#endif
    if (absr(d1) <= (real)D3IL_LS_TOL * absr(d0)) break;
    if (d1 < 0) { lo = alpha; dlo = d1; } else { hi = alpha; dhi = d1; }
    real next = alpha - d1 / d2;
    if (hi < 0) { if (!(next > lo)) next = 2 * alpha; }
    else {
      // bracketed: Newton step unless it hugs an end point (it can cycle across a kink of phi'), then false position
      const real wd = hi - lo;
      if (!(next > lo + (real)0.05 * wd && next < hi - (real)0.05 * wd)) {
        const real sec = lo + wd * (-dlo) / (dhi - dlo);
        next = clampr(sec, lo + (real)0.05 * wd, hi - (real)0.05 * wd);
      }
      if (wd < (real)1e-6 * (1 + hi)) { alpha = next; break; }
    }
    alpha = next;
  }
  return alpha;
}

// One contact that couples two trees (the rod pushing a box — the slow envs of every batch) makes H = H_blocks + Jc^T B Jc
// a rank-3 (rank-4 with torsional friction) update of the block-diagonal Hessian, so by the Woodbury identity
//     p = y0 - W (I + B S)^-1 B (Jc y0),   y0 = -H_blocks^-1 grad,  W = H_blocks^-1 Jc^T,  S = Jc W :
// dim more block solves and a dim x dim system every lane solves redundantly, instead of a dense nv x nv factorisation in
// shared memory.  Two or more coupling contacts take the dense path (solve_constraints).
// Woodbury correction of the block direction for the single coupling contact `cplc`.  In: pvec =
// y0 = -H_blocks^-1 grad.  Out: pvec = -H^-1 grad.  wood: MD dof vectors of scratch.
template <int G, int MD>
DEVNI void newton_woodbury(const Cx& cx, const Model& m, const Lay& L, real* w, int cplc) {
  const int nv = m.nv;
  const real* J = w + L.J;
  const real* cc = w + L.con + D3_CON_W * cplc;
  const int cdim = MD == 3 ? 3 : (int)cc[15];
  const int r0 = (int)cc[19], a0 = (int)cc[20], a1 = (int)cc[21], b0 = (int)cc[22], b1 = (int)cc[23];
  real* W = w + L.wood;
  // W_p = H_blocks^-1 Jc_p^T
  for (int p = 0; p < cdim; p++) {
    LANES(i, nv) {
      real v = 0;
      if (i >= a0 && i < a1) v = J[(r0 + p) * D3_JW + (i - a0)];
      else if (i >= b0 && i < b1) v = J[(r0 + p) * D3_JW + (a1 - a0) + (i - b0)];
      W[p * nv + i] = v;
    }
  }
  gsync<G>(cx);
#ifndef D3IL_EMU
  if (nv <= G) chol_blocks_solve_multi<G, MD>(cx, m, w + L.H, nv, false, w + L.hdinv, W, cdim);
  else
#endif
  for (int p = 0; p < cdim; p++) chol_blocks_solve<G>(cx, m, w + L.H, nv, false, w + L.hdinv, W + p * nv);
  // S = Jc W (symmetric), t = Jc y0: all partial sums in one pass over the dofs, then ONE interleaved all-reduce
  real S[MD][MD], t[MD], z[MD];
  {
    real acc[MD + MD * (MD + 1) / 2];
#pragma unroll
    for (int k = 0; k < MD + MD * (MD + 1) / 2; k++) acc[k] = 0;
    LANES(i, nv) {
      const int loc = (i >= a0 && i < a1) ? i - a0 : ((i >= b0 && i < b1) ? (a1 - a0) + (i - b0) : -1);
      if (loc < 0) continue;
      const real y0 = w[L.pvec + i];
      real jv[MD], wq[MD];
#pragma unroll
      for (int p = 0; p < MD; p++) { jv[p] = p < cdim ? J[(r0 + p) * D3_JW + loc] : (real)0; wq[p] = p < cdim ? W[p * nv + i] : (real)0; }
      int k = MD;
#pragma unroll
      for (int p = 0; p < MD; p++) {
        acc[p] += jv[p] * y0;
#pragma unroll
        for (int q = p; q < MD; q++) acc[k++] += jv[p] * wq[q];
      }
    }
    gsumn<G, MD + MD * (MD + 1) / 2>(cx, acc);
    int k = MD;
#pragma unroll
    for (int p = 0; p < MD; p++) {
      t[p] = acc[p];
#pragma unroll
      for (int q = p; q < MD; q++) { S[p][q] = acc[k]; S[q][p] = acc[k]; k++; }
    }
  }
  // C z = B t with C = I + B S (B: the contact's cone Hessian block; rows / columns >= cdim are zero -> z = 0 there)
  const real* Bm = w + L.hb + MD * MD * cplc;
  real Cm[MD][MD + 1];
#pragma unroll
  for (int p = 0; p < MD; p++) {
    real bt = 0;
#pragma unroll
    for (int q = 0; q < MD; q++) {
      real v = p == q ? (real)1 : (real)0;
#pragma unroll
      for (int k = 0; k < MD; k++) v += Bm[MD * p + k] * S[k][q];
      Cm[p][q] = v; bt += Bm[MD * p + q] * t[q];
    }
    Cm[p][MD] = bt;
  }
  // Gaussian elimination with partial pivoting (C is not symmetric), unrolled on registers
#pragma unroll
  for (int k = 0; k < MD; k++) {
#pragma unroll
    for (int i2 = k + 1; i2 < MD; i2++) {
      const bool sw = absr(Cm[i2][k]) > absr(Cm[k][k]);
#pragma unroll
      for (int q = 0; q <= MD; q++) { const real u = Cm[k][q], v = Cm[i2][q]; Cm[k][q] = sw ? v : u; Cm[i2][q] = sw ? u : v; }
    }
    const real inv = 1 / Cm[k][k];
#pragma unroll
    for (int i2 = k + 1; i2 < MD; i2++) {
      const real f = Cm[i2][k] * inv;
#pragma unroll
      for (int q = k; q <= MD; q++) Cm[i2][q] -= f * Cm[k][q];
    }
  }
#pragma unroll
  for (int k = MD - 1; k >= 0; k--) {
    real v = Cm[k][MD];
#pragma unroll
    for (int q = k + 1; q < MD; q++) v -= Cm[k][q] * z[q];
    z[k] = v / Cm[k][k];
  }
  LANES(i, nv) {
    real v = w[L.pvec + i];
#pragma unroll
    for (int p = 0; p < MD; p++) if (p < cdim) v -= W[p * nv + i] * z[p];
    w[L.pvec + i] = v;
  }
  gsync<G>(cx);
}

// Newton solver on the primal problem: result in qacc / frcE / qfrc_c.  Returns iterations.
template <int G, bool CS, int MD>
DEVFN int solve_constraints(const Cx& cx, const Model& m, const Lay& L, real* w, int ne, int nlimit, int ncon, int ncpl, int cplc, real tol, int max_iter) {
  PHASE2_T0();
#ifdef D3IL_NO_WOODBURY
  const int coupled = ncpl > 0;
#else
  const int coupled = ncpl > 1;      // dense path: two or more contacts couple kinematic trees (one is handled as a low-rank update of the block path)
#endif
  const int nv = m.nv;
  // Envs without active rows take qacc = qacc_smooth but keep walking the (CTA-uniform) iteration loop below.
  int done = ne == 0;
  if (done) { LANES(d, nv) { w[L.qacc + d] = w[L.qacc_smooth + d]; w[L.qfrc_c + d] = 0; } gsync<G>(cx); }
  const real scale = 1 / ((real)m.ctrl[D3C_MEANINERTIA] * (real)(nv > 1 ? nv : 1));
  const real* M = w + L.M;
  // ---- warm start: cheaper of qacc_warmstart and qacc_smooth.  The evaluation at the warm start (jar, forces, cone
  //      Hessian blocks, M (a - a_s)) is iteration 0's evaluation when the warm start wins - the usual case.
  real cost = 0, oldcost = 0, gn_prev = 0;
  if (!done) {
    eval_jar<G>(cx, m, L, w, nlimit, ne, L.qacc_smooth, L.jar, true);
    real cs = constraint_eval<G, false, MD>(cx, m, L, w, nlimit, ncon, L.jar, L.frcE);
    gsync<G>(cx);
    eval_jar<G>(cx, m, L, w, nlimit, ne, L.warm, L.jar, true);
    real cw = constraint_eval<G, true, MD>(cx, m, L, w, nlimit, ncon, L.jar, L.frcE);
    real part = 0;
    LANES(d, nv) {
      real sM = mrow_dot(m, M, nv, d, w + L.warm, w + L.qacc_smooth);
      w[L.Ma + d] = sM; part += (real)0.5 * (w[L.warm + d] - w[L.qacc_smooth + d]) * sM;
    }
    cw += gsum<G>(cx, part);
    gsync<G>(cx);
    if (cw < cs) { LANES(d, nv) w[L.qacc + d] = w[L.warm + d]; cost = cw; }
    else {
      LANES(d, nv) { w[L.qacc + d] = w[L.qacc_smooth + d]; w[L.Ma + d] = 0; }
      gsync<G>(cx);
      eval_jar<G>(cx, m, L, w, nlimit, ne, L.qacc_smooth, L.jar, true);
      cost = constraint_eval<G, true, MD>(cx, m, L, w, nlimit, ncon, L.jar, L.frcE);
    }
    gsync<G>(cx);
  }
  PHASE2(34);
  int iter = 0, nsteps = 0;      // nsteps: Newton steps this env actually took (iter also counts idle CTA-uniform passes)
  int grad_fresh = 0;            // grad / Ma / frcE all belong to the current iterate (then J^T f = Ma - grad)
  PHASE_T0();
  for (; iter < max_iter; iter++) {
    PHASE(15);
    if (!done) {
    // grad = M (a - a_s) - J^T f : lane per dof, contacts filtered by their dof ranges
    real g2 = 0, ga2 = 0;        // ga2: squared norm of the per-dof sums of |terms| (the scale of the cancellation inside the gradient)
    LANES(d, nv) {
      real s = w[L.Ma + d], sa = absr(s);
      for (int i = 0; i < nlimit; i++) { int sd = (int)w[L.econ + i]; if (sd == d + 1) { s -= w[L.frcE + i]; sa += absr(w[L.frcE + i]); } else if (sd == -(d + 1)) { s += w[L.frcE + i]; sa += absr(w[L.frcE + i]); } }
      const unsigned char* bl = (const unsigned char*)(w + L.blist) + m.d_blk[d] * (m.maxcon + 1);
      const int nbl = bl[0];
      for (int k = 0; k < nbl; k++) {
        const real* cc = w + L.con + D3_CON_W * (bl[1 + k] & 0x7f);
        int i = (int)cc[19];
        int a0 = (int)cc[20], a1 = (int)cc[21], b0 = (int)cc[22], b1 = (int)cc[23];
        if ((d >= a0 && d < a1) || (d >= b0 && d < b1)) {
          const real* Jr = w + L.J + i * D3_JW + (d < a1 && d >= a0 ? d - a0 : (a1 - a0) + d - b0);
          const real t0 = Jr[0] * w[L.frcE + i], t1 = Jr[D3_JW] * w[L.frcE + i + 1], t2 = Jr[2 * D3_JW] * w[L.frcE + i + 2];
          s -= t0 + t1 + t2; sa += absr(t0) + absr(t1) + absr(t2);
          if (MD == 4 && (int)cc[15] == 4) { const real t3 = Jr[3 * D3_JW] * w[L.frcE + i + 3]; s -= t3; sa += absr(t3); }
        }
      }
      w[L.grad + d] = s; g2 += s * s; ga2 += sa * sa;
    }
    gsum2<G>(cx, g2, ga2);
    real gn = sqrt(g2);
    gsync<G>(cx);
    PHASE2(31);
    grad_fresh = 1;
#ifdef D3IL_DEBUG_SOLVER
    printf("  newton it %d cost %.12g gn %.6g rel %.3g\n", iter, (double)cost, (double)gn, (double)(gn / sqrt(ga2 + 1e-30)));
#endif
    PHASE(8);
    if (blockIdx_is0()) count_iter();
    if (scale * gn < tol) done = 1;
    // fp32: the gradient is a difference of terms of size |terms| carried incrementally, whose rounding floor is ~5e-6 |terms|;
    // a gradient at that floor is converged by definition (another step only re-rolls the rounding) - without this test
    // 12 % of the ticks of a RESTING scene took a second Newton step, and a lock-step CTA of 8 envs nearly always did
    if (sizeof(real) == 4 && iter > 0 && absr(cost) < (real)20 && g2 <= (real)(D3IL_GRAD_FLOOR * D3IL_GRAD_FLOOR) * ga2) done = 1;      // quiet scenes only: a grasp or an impact (large cost) keeps the strict test
    // Improvement below what the cost can resolve in this precision: further iterations only chase rounding noise.
    // When the cost is large (impacts, deep spawn penetration, a grasp) its fp32 resolution (2e-6 |cost|) is blind to the
    // light dofs - a box's rotation has inertia 3e-5 kg m^2 - whose accelerations keep converging long after the cost
    // has gone flat: there the loop also waits for the gradient to stall.  Small costs (boxes at rest: ~5) stop on the cost alone.
    if (iter > 0 && (absr(cost) < (real)20 || gn > (real)0.5 * gn_prev) &&
        (scale * (oldcost - cost) < tol * (real)1e-3 || oldcost - cost <= (sizeof(real) == 4 ? (real)2e-6 : (real)1e-14) * absr(oldcost))) done = 1;
    gn_prev = gn;
    }
    // all groups of the CTA iterate together (converged ones idle) so the Newton body stays fetch-shared
    if (!cta_any<CS>(cx, !done)) break;
#if defined(D3IL_PHASE_TIMING) && defined(__CUDA_ARCH__)
    if (threadIdx.x == 0 && blockIdx.x < 4096) { g_cta_stat[4 * blockIdx.x] += 1; if (coupled) g_cta_stat[4 * blockIdx.x + 1] += 1; if (!done) g_cta_stat[4 * blockIdx.x + 2] += 1; }
#endif
    if (!done) {
    nsteps++;
    int hfail;
    // ---- H = M + J^T Hc J (lower triangle).  Block diagonal unless contacts couple two trees.
    // (A) every in-block entry is owned by one lane, which accumulates M, the limit rows and all contacts that live
    //     inside that block in a register and writes H once: no barriers between contacts.
    if (coupled) { LANES(e, nv * nv) w[L.H + e] = 0; gsync<G>(cx); }
    PHASE2(32);
    LANES(e, m.nhe) {
      const int gi = m.he_i[e], gj = m.he_j[e], bs = m.d_bs[gi];
      real acc = M[m.m_row[gj] + gi];
      if (gi == gj) for (int i = 0; i < nlimit; i++) { int sd = (int)w[L.econ + i]; if ((sd > 0 ? sd : -sd) - 1 == gi) acc += w[L.hd + i]; }
      const unsigned char* bl = (const unsigned char*)(w + L.blist) + m.d_blk[gi] * (m.maxcon + 1);
      const int nbl = bl[0];
      for (int k = 0; k < nbl; k++) {
        const int c = bl[1 + k];
        if (c & 0x80) continue;                                                            // couples two blocks: Woodbury / dense path
        const real* cc = w + L.con + D3_CON_W * c;
        const int r0 = (int)cc[19], a1 = (int)cc[21];
        if (gi >= a1) continue;                                                            // gj <= gi < a1
        const real* Hb = w + L.hb + MD * MD * c;
        const real *J0 = w + L.J + r0 * D3_JW, *J1 = J0 + D3_JW, *J2 = J1 + D3_JW;
        const int li = gi - bs, lj = gj - bs;
        real i0 = J0[li], i1 = J1[li], i2 = J2[li];
        if (MD == 3) {
          real t0 = i0 * Hb[0] + i1 * Hb[3] + i2 * Hb[6], t1 = i0 * Hb[1] + i1 * Hb[4] + i2 * Hb[7], t2 = i0 * Hb[2] + i1 * Hb[5] + i2 * Hb[8];
          acc += t0 * J0[lj] + t1 * J1[lj] + t2 * J2[lj];
        } else {           // 4 x 4 cone block; the block's row/column 3 is zero for condim-3 contacts
          const bool d4 = (int)cc[15] == 4;
          real i3 = d4 ? J2[D3_JW + li] : (real)0, j3 = d4 ? J2[D3_JW + lj] : (real)0;
          real t0 = i0 * Hb[0] + i1 * Hb[4] + i2 * Hb[8] + i3 * Hb[12], t1 = i0 * Hb[1] + i1 * Hb[5] + i2 * Hb[9] + i3 * Hb[13];
          real t2 = i0 * Hb[2] + i1 * Hb[6] + i2 * Hb[10] + i3 * Hb[14], t3 = i0 * Hb[3] + i1 * Hb[7] + i2 * Hb[11] + i3 * Hb[15];
          acc += t0 * J0[lj] + t1 * J1[lj] + t2 * J2[lj] + t3 * j3;
        }
      }
      w[L.H + gi * nv + gj] = acc;
    }
    gsync<G>(cx);
    PHASE2(33);
    if (!coupled) {
      // block path: factorisation in registers, block solve; a single coupling contact is a low-rank (Woodbury) update
      PHASE(9);
      LANES(d, nv) { w[L.hpiv + d] = w[L.H + d * nv + d]; w[L.pvec + d] = -w[L.grad + d]; }
      gsync<G>(cx);
      hfail = chol_blocks_reg<G>(cx, m, w + L.H, nv, false, w + L.hpiv, nullptr, w + L.hdinv);
      PHASE(10);
      if (!hfail) {
        chol_blocks_solve<G>(cx, m, w + L.H, nv, false, w + L.hdinv, w + L.pvec);
        PHASE2(29);
        if (ncpl == 1) newton_woodbury<G, MD>(cx, m, L, w, cplc);
        PHASE2(30);
      }
    } else {
    // (B) contacts that couple two blocks (rod-box, box-box): rare, added one after the other
    for (int c = 0; c < ncon; c++) {
      const real* cc = w + L.con + D3_CON_W * c;
      int r0 = (int)cc[19];
      if (r0 < 0 || (int)cc[23] == (int)cc[22]) continue;
      int a0 = (int)cc[20], a1 = (int)cc[21], b0 = (int)cc[22], b1 = (int)cc[23], na = a1 - a0, nn = na + b1 - b0;
      const real* Hb = w + L.hb + MD * MD * c;
      const real *J0 = w + L.J + r0 * D3_JW, *J1 = J0 + D3_JW, *J2 = J1 + D3_JW;
      const bool d4 = MD == 4 && (int)cc[15] == 4;
      LANES(e, nn * (nn + 1) / 2) {
        int li = m.tri_i[e], lj = m.tri_j[e];
        int gi = li < na ? a0 + li : b0 + li - na, gj = lj < na ? a0 + lj : b0 + lj - na;
        real i0 = J0[li], i1 = J1[li], i2 = J2[li];
        if (MD == 3) {
          real t0 = i0 * Hb[0] + i1 * Hb[3] + i2 * Hb[6], t1 = i0 * Hb[1] + i1 * Hb[4] + i2 * Hb[7], t2 = i0 * Hb[2] + i1 * Hb[5] + i2 * Hb[8];
          w[L.H + gi * nv + gj] += t0 * J0[lj] + t1 * J1[lj] + t2 * J2[lj];
        } else {
          real i3 = d4 ? J2[D3_JW + li] : (real)0, j3 = d4 ? J2[D3_JW + lj] : (real)0;
          real t0 = i0 * Hb[0] + i1 * Hb[4] + i2 * Hb[8] + i3 * Hb[12], t1 = i0 * Hb[1] + i1 * Hb[5] + i2 * Hb[9] + i3 * Hb[13];
          real t2 = i0 * Hb[2] + i1 * Hb[6] + i2 * Hb[10] + i3 * Hb[14], t3 = i0 * Hb[3] + i1 * Hb[7] + i2 * Hb[11] + i3 * Hb[15];
          w[L.H + gi * nv + gj] += t0 * J0[lj] + t1 * J1[lj] + t2 * J2[lj] + t3 * j3;
        }
      }
      gsync<G>(cx);
    }
    PHASE(9);
    LANES(d, nv) { w[L.hpiv + d] = w[L.H + d * nv + d]; w[L.pvec + d] = -w[L.grad + d]; }
    gsync<G>(cx);
    // A Hessian that is not positive definite (NaN inputs included) ends this env's solve with status bit 4.  The env must
    // NOT leave the loop on its own: the iteration is CTA-uniform (cta_any above is a barrier every warp of the CTA has to
    // reach), so it turns `done` and idles through the remaining passes like a converged env.
#ifndef D3IL_EMU
    if (nv <= G) { hfail = chol_dense_rolled_solve<G>(cx, m, w + L.H, nv, w + L.hpiv, w + L.pvec); PHASE(10); }
    else
#endif
    {
    hfail = chol_factor_part<G>(cx, m, w + L.H, nv, true, nv, w + L.hpiv, w + L.hdinv);
    PHASE(10);
    if (!hfail) chol_solve_part<G>(cx, m, w + L.H, nv, true, nv, w + L.hdinv, w + L.pvec);
    }
    }
    if (hfail) { LANES(z, 1) w[L.misc + ST_STATUS] = (real)(((int)w[L.misc + ST_STATUS]) | D3_STATUS_H_NOT_PD); done = 1; gsync<G>(cx); }
    if (!hfail) {
    PHASE(11);
    // ---- exact line search (safeguarded 1-D Newton / false position), quantities reduced across lanes
    eval_jar<G>(cx, m, L, w, nlimit, ne, L.pvec, L.Jp, false);
    PHASE2(24);
    real a1s = 0, a2s = 0, a3s = 0;
    LANES(d, nv) {
      real s = mrow_dot(m, M, nv, d, w + L.pvec, nullptr);
      w[L.tmpv + d] = s;                                             // M p (tmpv is free until the Euler stage)
      a1s += w[L.pvec + d] * s; a2s += w[L.pvec + d] * w[L.Ma + d]; a3s += w[L.grad + d] * w[L.pvec + d];
    }
    gsum2<G>(cx, a1s, a2s);
    real pMp = a1s, pMa = a2s, d0 = gsum<G>(cx, a3s);
    PHASE2(25);
#ifdef D3IL_EMU
    const real alpha = line_search<G, MD, 64, 2 * D3_NROB>(cx, m, L, w, nlimit, ncon, d0, pMa, pMp);
#else
    const real alpha = m.maxcon <= G ? line_search<G, MD, 1, 1>(cx, m, L, w, nlimit, ncon, d0, pMa, pMp) : line_search<G, MD, 2, 1>(cx, m, L, w, nlimit, ncon, d0, pMa, pMp);
#endif
    PHASE2(26);
    // step: everything linear in the iterate moves incrementally (jar = J a - aref, Ma = M (a - a_s)); then the
    // evaluation of the NEXT iteration (cost, forces, cone Hessian blocks) at the new iterate
    LANES(d, nv) { w[L.qacc + d] += alpha * w[L.pvec + d]; w[L.Ma + d] += alpha * w[L.tmpv + d]; }
    LANES(i, ne) w[L.jar + i] += alpha * w[L.Jp + i];
    gsync<G>(cx);
    PHASE2(27);
    oldcost = cost;
    cost = constraint_eval<G, true, MD>(cx, m, L, w, nlimit, ncon, L.jar, L.frcE);
    real partc = 0;
    LANES(d, nv) partc += (real)0.5 * (w[L.qacc + d] - w[L.qacc_smooth + d]) * w[L.Ma + d];
    cost += gsum<G>(cx, partc);
    gsync<G>(cx);
    grad_fresh = 0;
    PHASE2(28);
    PHASE(12);
    }
    }
  }
  // iteration cap reached without the convergence test passing: reported, never silent (the oracle flags its own cap the same way)
  if (!done) { LANES(z, 1) w[L.misc + ST_STATUS] = (real)(((int)w[L.misc + ST_STATUS]) | D3_STATUS_ITER_CAP); }
  // qfrc_constraint = J^T f = M (a - a_s) - grad when the gradient belongs to the final iterate (converged exit)
  if (ne > 0 && grad_fresh) { LANES(d, nv) w[L.qfrc_c + d] = w[L.Ma + d] - w[L.grad + d]; }
  else if (ne > 0) LANES(d, nv) {
    real s = 0;
    for (int i = 0; i < nlimit; i++) { int sd = (int)w[L.econ + i]; if (sd == d + 1) s += w[L.frcE + i]; else if (sd == -(d + 1)) s -= w[L.frcE + i]; }
    for (int c = 0; c < ncon; c++) {
      const real* cc = w + L.con + D3_CON_W * c;
      int i = (int)cc[19];
      if (i < 0) continue;
      int a0 = (int)cc[20], a1 = (int)cc[21], b0 = (int)cc[22], b1 = (int)cc[23];
      if ((d >= a0 && d < a1) || (d >= b0 && d < b1)) {
        const real* Jr = w + L.J + i * D3_JW + (d < a1 && d >= a0 ? d - a0 : (a1 - a0) + d - b0);
        s += Jr[0] * w[L.frcE + i] + Jr[D3_JW] * w[L.frcE + i + 1] + Jr[2 * D3_JW] * w[L.frcE + i + 2];
        if (MD == 4 && (int)cc[15] == 4) s += Jr[3 * D3_JW] * w[L.frcE + i + 3];
      }
    }
    w[L.qfrc_c + d] = s;
  }
  gsync<G>(cx);
  PHASE2(35);
  return nsteps;
}

#ifndef D3IL_SKIP_BARS
#define D3IL_SKIP_BARS 0   // diagnostic: bit i drops the i-th phase barrier of the tick (placement sweeps)
#endif
#define PH_SYNC(i) do { if (!((D3IL_SKIP_BARS >> (i)) & 1)) cta_sync<CS>(cx); } while (0)
// ------------------------------------------------------------------------------------------------ one physics tick
// jt_q / jt_qlo / jt_qd: joint set-point for this tick (from the IK reference in Cartesian mode, or the held pose).
template <int G, bool CS, int MD>
DEVFN void physics_tick(const Cx& cx, const Model& m, const Lay& L, real* w, const real* jt_q, const real* jt_qlo, const real* jt_qd, real tol, int max_iter) {
  const int nv = m.nv;
  const real h = m.ctrl[D3C_DT];
  PHASE_T0();
  // --- MjRobot.prepare_step: joint PD + stale-bias gravity compensation, finger law (Robots.py:441-476), actuator clamp
  LANES(k, D3_NARM) {
    // joint angles and their set-points are two-float numbers: the difference of the high words is exact (Sterbenz)
    real perr = (jt_q[k] - w[L.qpos + k]) + (jt_qlo[k] - w[L.qlo + k]);
    real tau = (real)m.ctrl[D3C_PD_P + k] * perr + (real)m.ctrl[D3C_PD_D + k] * (jt_qd[k] - w[L.qvel + k]) + w[L.bias_prev + k];
    real fr = m.link[D3_LINK_W * k + 26];
    w[L.act + k] = clampr(tau, -fr, fr);
  }
  LANES(k, 2) {
    real w0 = w[L.qpos + 7], w1 = w[L.qpos + 8], mean = (real)0.5 * (w0 + w1), wk = k ? w1 : w0, vk = w[L.qvel + 7 + k];
    real set = w[L.misc + ST_GRIP_SET], f = 500 * (mean - wk), g;
    if (mean - set > (real)0.005) g = w[L.misc + ST_GRASP] != 0 ? (real)-20 : 10 * ((real)-0.2 - vk);
    else g = clampr(500 * (set - wk) - 10 * vk, -5, 5);
    real fr = m.link[D3_LINK_W * (7 + k) + 26];
    w[L.act + 7 + k] = clampr(f + g, -fr, fr);
  }
  // --- mj_step: position-dependent stage
  kinematics<G>(cx, m, L, w);
  // stale-by-one-tick tcp pose: body 'tcp' hangs off link 7 (index 6)
  LANES(z, 1) {
    real o[3], tp[3] = {(real)m.ctrl[D3C_TCP_POS], (real)m.ctrl[D3C_TCP_POS + 1], (real)m.ctrl[D3C_TCP_POS + 2]};
    real tq[4] = {(real)m.ctrl[D3C_TCP_QUAT], (real)m.ctrl[D3C_TCP_QUAT + 1], (real)m.ctrl[D3C_TCP_QUAT + 2], (real)m.ctrl[D3C_TCP_QUAT + 3]}, Rt[9], R[9];
    mat_vec3(o, w + L.xmat + 54, tp);
    for (int k = 0; k < 3; k++) w[L.tcp + k] = w[L.xpos + 18 + k] + o[k];
    quat2mat(Rt, tq); mat_mul3(R, w + L.xmat + 54, Rt); mat2quat(w + L.tcp + 3, R);
  }
  PHASE(0);
  PH_SYNC(0);
  LANES(e, m.nzp) { int a = m.zp_a[e], b = m.zp_b[e]; w[L.M + m.m_row[a] + b] = 0; w[L.M + m.m_row[b] + a] = 0; }     // in-block pairs CRBA never writes (the two fingers)
  dynamics<G>(cx, m, L, w);
  PHASE(1);
  PH_SYNC(1);
  int ncon = collision<G>(cx, m, L, w);
  PHASE(2);
  PH_SYNC(2);
  int nlimit = 0, coupled = 0;
  int ne = make_constraints<G, MD>(cx, m, L, w, ncon, &nlimit, &coupled);
  int cplc = -1;
  const int ncpl = count_coupling<G>(cx, m, L, w, ncon, &cplc);
  PHASE(3);
  PHASE_MARK_DENSE(ncpl > 1);
#ifdef D3IL_PHASE_TIMING
  if (cx.lane == 0) { count_stat(21, coupled); count_stat(22, ncon); count_stat(23, 1); count_stat(19, ne); }
  if (blockIdx_is0()) { count_stat(17, ncon); count_stat(14, ncpl == 1); count_stat(13, ncpl > 1); count_stat(7, ne); }
#define D3IL_ITER_HIST 1
#endif
  PH_SYNC(3);
  // --- smooth dynamics: qacc_smooth = M^-1 (passive - bias + actuation); M is block diagonal over the trees
  LANES(d, nv) {
    int li = m.d_link[d];
    real passive = m.l_jtype[li] == 2 ? (real)0 : -(real)m.link[D3_LINK_W * li + 25] * w[L.qvel + d];
    real act = d < D3_NROB ? w[L.act + d] : (real)0;
    real f = passive - w[L.bias + d] + act;
    w[L.qfrc_smooth + d] = f; w[L.qacc_smooth + d] = f;
  }
  gsync<G>(cx);
  // chol(M): rows come from the upper triangle (where CRBA wrote M), the factor goes to the strict lower triangle
  if (chol_blocks_reg<G>(cx, m, w + L.M, nv, true, nullptr, nullptr, w + L.mdinv)) { LANES(z, 1) w[L.misc + ST_STATUS] = (real)(((int)w[L.misc + ST_STATUS]) | D3_STATUS_M_NOT_PD); }
  chol_blocks_solve<G>(cx, m, w + L.M, nv, true, w + L.mdinv, w + L.qacc_smooth);
  PHASE(4);
  PH_SYNC(4);
  int iters = solve_constraints<G, CS, MD>(cx, m, L, w, ne, nlimit, ncon, ncpl, cplc, tol, max_iter);
#if defined(D3IL_ITER_HIST) && defined(__CUDA_ARCH__)
  if (cx.lane == 0) {
    int b = iters > 15 ? 15 : iters;
    atomicAdd(&g_iter_hist[b], 1ull); if (coupled) atomicAdd(&g_iter_hist[16 + b], 1ull);
    if (iters >= 8) { atomicAdd(&g_iter_hist[32], (unsigned long long)ncon); atomicAdd(&g_iter_hist[33], 1ull); }
  }
#endif
  LANES(z, 1) {      // per-env cost counters of the current env step (cost-aware scheduling, diagnostics)
    w[L.misc + ST_COST_ITERS] += (real)iters; w[L.misc + ST_COST_COUPLED] += (real)coupled;
    if ((real)ncon > w[L.misc + ST_COST_NCON]) w[L.misc + ST_COST_NCON] = (real)ncon;
  }
  PHASE(5);
  PH_SYNC(5);
  LANES(d, nv) { w[L.warm + d] = w[L.qacc + d]; w[L.tmpv + d] = w[L.qfrc_smooth + d] + w[L.qfrc_c + d]; }
  LANES(k, D3_NROB) w[L.bias_prev + k] = w[L.bias + k];
  // --- mj_Euler with implicit joint damping: (M + h B) qacc* = qfrc_smooth + qfrc_constraint.  M (upper triangle) and its
  //     diagonal are still intact: the blocks of M + h B are simply factored again in registers (the only damped dofs
  //     are the two fingers, but a register factorisation of every block costs less than a special case for the corner).
  LANES(d, nv) { int li = m.d_link[d]; w[L.hpiv + d] = m.l_jtype[li] == 2 ? (real)0 : h * (real)m.link[D3_LINK_W * li + 25]; }
  gsync<G>(cx);
  chol_blocks_reg<G>(cx, m, w + L.M, nv, true, nullptr, w + L.hpiv, w + L.mdinv);
  chol_blocks_solve<G>(cx, m, w + L.M, nv, true, w + L.mdinv, w + L.tmpv);
  LANES(d, nv) w[L.qvel + d] += h * w[L.tmpv + d];
  gsync<G>(cx);
  LANES(i, m.nlink) {
    if (m.l_jtype[i] != 2) {
      // compensated (Kahan) integration of the robot joint angles; links 0..8 are the robot, qadr == link index
      real hi = w[L.qpos + m.l_qadr[i]], t = h * w[L.qvel + m.l_dadr[i]] + w[L.qlo + i], sres = hi + t;
      w[L.qlo + i] = t - (sres - hi); w[L.qpos + m.l_qadr[i]] = sres;
      continue;
    }
    real* q = w + L.qpos + m.l_qadr[i]; const real* v = w + L.qvel + m.l_dadr[i];
    q[0] += h * v[0]; q[1] += h * v[1]; q[2] += h * v[2];
    real wn = norm3(v + 3), ang = wn * h;
    if (ang > 0) {
      real s, c; sincos_small((real)0.5 * ang, &s, &c);
      s /= wn;
      real a0 = q[3], a1 = q[4], a2 = q[5], a3 = q[6], b1 = s * v[3], b2 = s * v[4], b3 = s * v[5];
      real r0 = a0 * c - a1 * b1 - a2 * b2 - a3 * b3, r1 = a0 * b1 + a1 * c + a2 * b3 - a3 * b2;
      real r2 = a0 * b2 - a1 * b3 + a2 * c + a3 * b1, r3 = a0 * b3 + a1 * b2 - a2 * b1 + a3 * c;
      real n = 1 / sqrt(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3);
      q[3] = r0 * n; q[4] = r1 * n; q[5] = r2 * n; q[6] = r3 * n;
    }
  }
  gsync<G>(cx);
  PHASE(6);
  PH_SYNC(6);
}
