// d3il_env.cuh — env-level logic on top of d3il_core.cuh: the Cartesian-impedance IK reference generator (one thread
// per env), Gym reset/step bookkeeping, observations, termination and mode logic for the compiled tasks.
//
// Reference restated (all under environments/d3il/): controllers/IKControllers.py:163-323 (a5),
// core/Model.py:37-66 (a6), gyms/gym_env_wrapper.py:45-137 (a2), envs/gym_pushing_env/.../pushing.py:255-488,
// envs/gym_avoiding_env/.../avoiding.py:117-262 (a10-a12).  Dual-compilable like the core (see header there).
#pragma once
#include "d3il_core.cuh"

// ------------------------------------------------------------------------------------------------ IK reference (a5/a6)
// Per-env controller state.  The joint reference `q` is integrated in double (increments of 1e-3 * qd are below
// fp32 resolution at |q| ~ 2.5 rad — SURVEY §7 hard part 2); everything else is fp32.
struct IkState {
  double q[D3_NARM];
  real des_pos[3], des_quat[4];
  real jt_q[D3_NARM], jt_qlo[D3_NARM], jt_qd[D3_NARM];    // last joint set-point handed to the PD loop (q as two floats)
  int valid;
};

// The whole IK reference runs in fp64 (B200 has a 1:2 fp64 pipe; this is ~0.6 MFLOP per env step on a path that is
// otherwise fp32): with fp32 forward kinematics the 1e-7 m position noise is amplified by gain/dt into ~1e-3 N m of PD
// torque noise, which is what dominated the kernel-vs-oracle velocity error.
typedef double ikr;
// The controller table is handed to the IK routines as DOUBLES (k_ik converts the fp32 model table once per block into shared
// memory): with a float table the compiler hoists ~100 float->double conversions out of the iteration loops and spills them.
// The 6x7 Jacobian of an env is addressed with a stride: 1 on the host (plain array), the block size in k_ik, where it lives
// in shared memory as J[k][thread] (conflict-free, and 84 registers less per thread: no spills).
#ifndef IK_JSTRIDE
#define IK_JSTRIDE 1
#endif
#define IKJ(J, k) (J)[(k) * IK_JSTRIDE]
DEVFN void ik_cross3(ikr* o, const ikr* a, const ikr* b) {
  ikr x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
DEVFN void ik_mat_vec3(ikr* o, const ikr* R, const ikr* v) {
  ikr x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2], z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
DEVFN void ik_mat_mul3(ikr* o, const ikr* A, const ikr* B) {
  ikr t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  for (int i = 0; i < 9; i++) o[i] = t[i];
}
DEVFN void ik_mat2quat(ikr* q, const ikr* R) {
  ikr t = R[0] + R[4] + R[8];
  if (t > 0) { ikr s = sqrt(t + 1) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { ikr s = sqrt(1 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { ikr s = sqrt(1 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { ikr s = sqrt(1 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
  ikr n = 1 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
}
DEVFN ikr ik_rsqrt(ikr x) {
#ifdef __CUDA_ARCH__
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
DEVFN ikr ik_clamp(ikr x, ikr lo, ikr hi) { return x < lo ? lo : (x > hi ? hi : x); }

// Forward kinematics + geometric Jacobian of the URDF chain (core/Model.py:37-66).  Everything is unrolled and
// statically indexed so that the 42 Jacobian doubles live in registers (k_ik is one thread per env).
DEVFN void ik_fk(const ikr* C, const ikr* sn, const ikr* cs, ikr* pos, ikr* quat, ikr* J) {
  // joint origins and axes go straight into J's storage (origin in the linear rows, axis in the angular rows) and the
  // linear rows are converted in place once the tool position is known: no 42-double org / axs arrays in registers
  ikr p[3] = {0, 0, 0}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
  for (int i = 0; i < 7; i++) {
    const ikr* o = C + D3C_IK_ORIGIN + 12 * i;
    const ikr c = cs[i], s = sn[i];
    ikr Rn[9];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      p[r] += R[3 * r] * (ikr)o[0] + R[3 * r + 1] * (ikr)o[1] + R[3 * r + 2] * (ikr)o[2];
      // R <- R oR Rz(q): first R oR, then mix its first two columns
      const ikr m0 = R[3 * r] * (ikr)o[3] + R[3 * r + 1] * (ikr)o[6] + R[3 * r + 2] * (ikr)o[9];
      const ikr m1 = R[3 * r] * (ikr)o[4] + R[3 * r + 1] * (ikr)o[7] + R[3 * r + 2] * (ikr)o[10];
      const ikr m2 = R[3 * r] * (ikr)o[5] + R[3 * r + 1] * (ikr)o[8] + R[3 * r + 2] * (ikr)o[11];
      Rn[3 * r] = m0 * c + m1 * s; Rn[3 * r + 1] = m1 * c - m0 * s; Rn[3 * r + 2] = m2;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = Rn[k];
    IKJ(J, i) = p[0]; IKJ(J, 7 + i) = p[1]; IKJ(J, 14 + i) = p[2];
    IKJ(J, 21 + i) = R[2]; IKJ(J, 28 + i) = R[5]; IKJ(J, 35 + i) = R[8];
  }
  const ikr* o = C + D3C_IK_EE;
  ikr Re[9];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    pos[r] = p[r] + (R[3 * r] * (ikr)o[0] + R[3 * r + 1] * (ikr)o[1] + R[3 * r + 2] * (ikr)o[2]);
#pragma unroll
    for (int c = 0; c < 3; c++) Re[3 * r + c] = R[3 * r] * (ikr)o[3 + c] + R[3 * r + 1] * (ikr)o[6 + c] + R[3 * r + 2] * (ikr)o[9 + c];
  }
  ik_mat2quat(quat, Re);
#pragma unroll
  for (int i = 0; i < 7; i++) {
    const ikr r0 = pos[0] - IKJ(J, i), r1 = pos[1] - IKJ(J, 7 + i), r2 = pos[2] - IKJ(J, 14 + i);
    const ikr a0 = IKJ(J, 21 + i), a1 = IKJ(J, 28 + i), a2 = IKJ(J, 35 + i);
    IKJ(J, i) = a1 * r2 - a2 * r1; IKJ(J, 7 + i) = a2 * r0 - a0 * r2; IKJ(J, 14 + i) = a0 * r1 - a1 * r0;
  }
}

// Clipped-spectrum solve, fast path.  The controller applies A^-1 with the eigenvalues of A = J J^T + reg I clipped to
// [lo, hi] (IKControllers.py: np.linalg.svd + np.clip).  When the whole spectrum already lies inside the clip range the
// result is simply A^-1 rhs, and that is decidable without an eigen-decomposition: lambda_max <= trace(A) < hi, and
// lambda_min > lo  <=>  A - lo I is positive definite  <=>  its Cholesky factorisation finds positive pivots only.
// Pass 0 factors A - lo I (the test), pass 1 factors A and solves.  Returns 0 (x untouched) when clipping may be active;
// the caller then takes the Jacobi path.  At the boundary both paths agree (the clipped inverse is continuous), so the
// test needs no margin.  Panda poses over the tabletop workspace have lambda in [0.03, 4]: the fast path is the rule.
// Packed lower triangle: element (i, j), j <= i, at i (i + 1) / 2 + j.
#define IK_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))
DEVFN void ik_gram(const ikr* J, ikr diag_add, ikr* A21) {      // A = J J^T + diag_add I
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      ikr s = i == j ? diag_add : (ikr)0;
#pragma unroll
      for (int k = 0; k < 7; k++) s += IKJ(J, i * 7 + k) * IKJ(J, j * 7 + k);
      IKJ(A21, IK_TRI(i, j)) = s;
    }
}
DEVFN int ik_chol21(ikr* A21, ikr* dinv) {      // in place: A21 <- L (strict lower part), dinv <- 1 / L_jj; 0 if a pivot is not positive
#pragma unroll
  for (int j = 0; j < 6; j++) {
    ikr sd = IKJ(A21, IK_TRI(j, j));
#pragma unroll
    for (int k = 0; k < j; k++) sd -= IKJ(A21, IK_TRI(j, k)) * IKJ(A21, IK_TRI(j, k));
    if (!(sd > 1e-14)) return 0;
    const ikr d = ik_rsqrt(sd);
    IKJ(dinv, j) = d;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      ikr so = IKJ(A21, IK_TRI(i, j));
#pragma unroll
      for (int k = 0; k < j; k++) so -= IKJ(A21, IK_TRI(i, k)) * IKJ(A21, IK_TRI(j, k));
      IKJ(A21, IK_TRI(i, j)) = so * d;
    }
  }
  return 1;
}
// Inertia of a symmetric matrix by LDL^T without pivoting (Sylvester): *nneg = number of negative pivots = number of
// negative eigenvalues.  Returns 0 when a pivot is too small to trust its sign.
DEVFN int ik_ldl_inertia(ikr* A21, int* nneg) {
  ikr dd[6];
  int neg = 0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    ikr sd = IKJ(A21, IK_TRI(j, j));
#pragma unroll
    for (int k = 0; k < j; k++) sd -= IKJ(A21, IK_TRI(j, k)) * IKJ(A21, IK_TRI(j, k)) * dd[k];
    if (!(fabs(sd) > 1e-12)) return 0;
    dd[j] = sd; neg += sd < 0;
    const ikr inv = 1 / sd;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      ikr so = IKJ(A21, IK_TRI(i, j));
#pragma unroll
      for (int k = 0; k < j; k++) so -= IKJ(A21, IK_TRI(i, k)) * IKJ(A21, IK_TRI(j, k)) * dd[k];
      IKJ(A21, IK_TRI(i, j)) = so * inv;
    }
  }
  *nneg = neg;
  return 1;
}
DEVFN void ik_chol_solve21(const ikr* L21, const ikr* dinv, const ikr* b, ikr* x) {      // x = (L L^T)^-1 b
  ikr y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { ikr so = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) so -= IKJ(L21, IK_TRI(i, k)) * y[k];
    y[i] = so * IKJ(dinv, i); }
#pragma unroll
  for (int i = 5; i >= 0; i--) { ikr so = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) so -= IKJ(L21, IK_TRI(k, i)) * x[k];
    x[i] = so * IKJ(dinv, i); }
}
// x = V clip(Lambda, lo, hi)^-1 V^T rhs for A = J J^T + reg I WITHOUT an eigen-decomposition, in the two cases that cover the
// tabletop workspace (measured on the random-walk workload: 98 % / 2 % / none):
//   (0) no eigenvalue outside [lo, hi]:  x = A^-1 rhs.  Decided by lambda_max <= trace(A) < hi and a Cholesky factorisation
//       of A - lo I (positive pivots only  <=>  lambda_min > lo).
//   (1) exactly ONE eigenvalue below lo (the arm near the edge of its reach: the elbow singularity; the next eigenvalue
//       is ~0.2, 20x larger):  x = A^-1 rhs + (1/lo - 1/lambda_1) (v_1 . rhs) v_1, with (lambda_1, v_1) from inverse
//       iteration on the Cholesky factor of A (convergence factor lambda_1 / lambda_2 <= 0.05 per step), warm-started from
//       the previous call's v_1.  "Exactly one" is the inertia of A - lo I (LDL^T pivot signs).
// Anything else (two small eigenvalues, an untrustworthy pivot, no convergence) returns 0 and the caller falls back to the
// Jacobi eigen-decomposition.  The clipped inverse is continuous in A, so the case boundaries need no margins.
// Aw: 27 doubles of work space with the same stride as J (packed 6x6 lower triangle + 6 inverse pivots).
DEVFN int ik_solve_spd(const ikr* J, ikr reg, const ikr* rhs, ikr lo, ikr hi, ikr* x, ikr* v1 /*6*/, int* v1_valid, ikr* Aw) {
  ikr* A = Aw; ikr* dinv = Aw + 21 * IK_JSTRIDE;
  ik_gram(J, reg - lo, A);
  if (!(IKJ(A, IK_TRI(0, 0)) + IKJ(A, IK_TRI(1, 1)) + IKJ(A, IK_TRI(2, 2)) + IKJ(A, IK_TRI(3, 3)) + IKJ(A, IK_TRI(4, 4)) + IKJ(A, IK_TRI(5, 5)) + 6 * lo < hi)) return 0;
  const int none_below = ik_chol21(A, dinv);
  if (!none_below) {
    int nneg = 0;
    ik_gram(J, reg - lo, A);
    if (!ik_ldl_inertia(A, &nneg) || nneg != 1) return 0;
  }
  ik_gram(J, reg, A);                                            // A itself (rebuilt: cheaper than keeping a copy in registers)
  if (!ik_chol21(A, dinv)) return 0;
  ik_chol_solve21(A, dinv, rhs, x);
  if (none_below) return 1;
  ikr v[6], u[6];
#pragma unroll
  for (int k = 0; k < 6; k++) v[k] = *v1_valid ? v1[k] : (ikr)0.40824829046386302;
  int conv = 0;
  for (int itn = 0; itn < 48 && !conv; itn++) {
    ik_chol_solve21(A, dinv, v, u);
    ikr n2 = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) n2 += u[k] * u[k];
    const ikr inv = ik_rsqrt(n2);
    ikr d2 = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) { const ikr vn = u[k] * inv; d2 += (vn - v[k]) * (vn - v[k]); v[k] = vn; }
    conv = d2 < 1e-28;
  }
  if (!conv) return 0;
  ik_chol_solve21(A, dinv, v, u);                               // Rayleigh quotient of A^-1: 1 / lambda_1 = v . A^-1 v
  ikr il = 0, vr = 0;
#pragma unroll
  for (int k = 0; k < 6; k++) { il += v[k] * u[k]; vr += v[k] * rhs[k]; }
  const ikr coef = (1 / lo - il) * vr;
  if (il > 1 / lo) {
#pragma unroll
    for (int k = 0; k < 6; k++) x[k] += coef * v[k];
  }
#pragma unroll
  for (int k = 0; k < 6; k++) v1[k] = v[k];
  *v1_valid = 1;
  return 1;
}

// Slow path of the same solve: symmetric eigen-decomposition by parallel-ordered Jacobi rotations, clipped spectrum.
// `V` (persistent within the launch) carries the eigenbasis of the previous call when `warm` is set: A changes by
// O(1e-3) per IK iteration, so rotating into the old basis first leaves a nearly diagonal matrix and two or three sweeps
// finish it.  A sweep = 5 rounds of 3 rotations on disjoint index pairs (round-robin ordering); the three rotations of a
// round commute, so the round is ONE similarity transform A <- G^T A G, V <- V G with two non-zeros per column of G:
// one lane per matrix entry, results into the B1 / B2 buffers, copied back after the barrier.
// In k_ik (one THREAD per env) this is the warp-cooperative service routine: the warp's 32 lanes decompose the matrix of one
// env that needs it (ik_tick's ballot loop) in ~1/6 of the instructions a single lane would spend; the host build runs it with G = 1.
template <int G>
DEVNI void ik_solve_clipped_lanes(const Cx& cx, ikr* A, const ikr* rhs, ikr* V, ikr* B1, ikr* B2, ikr* csb, int warm, ikr lo, ikr hi, ikr* x) {
  if (warm) {
    LANES(e, 36) { const int i = e / 6, j = e - 6 * i; ikr s = 0; for (int k = 0; k < 6; k++) s += A[i * 6 + k] * V[k * 6 + j]; B1[e] = s; }
    gsync<G>(cx);
    LANES(e, 36) {      // V^T (A V); (i, j) and (j, i) evaluate the same expression: exactly symmetric
      const int i = e / 6, j = e - 6 * i, a = i < j ? i : j, b = i < j ? j : i;
      ikr s = 0; for (int k = 0; k < 6; k++) s += V[k * 6 + a] * B1[k * 6 + b];
      A[e] = s;
    }
  } else {
    LANES(e, 36) V[e] = (e / 6 == e % 6) ? 1.0 : 0.0;
  }
  gsync<G>(cx);
  constexpr int PP[5][3] = {{0, 2, 3}, {0, 1, 4}, {0, 2, 1}, {0, 3, 1}, {0, 1, 2}};
  constexpr int QQ[5][3] = {{1, 5, 4}, {2, 3, 5}, {3, 4, 5}, {4, 5, 2}, {5, 4, 3}};
  for (int sweep = 0; sweep < 30; sweep++) {
    ikr po = 0, pd = 0;
    LANES(e, 36) { const int i = e / 6, j = e - 6 * i; const ikr a = A[e]; if (i == j) pd += a * a; else if (j > i) po += a * a; }
    const ikr off = gsumd<G>(cx, po), diag = gsumd<G>(cx, pd);
    if (off <= 1e-22 * diag) break;      // sums of squares: off-diagonal entries below 1e-11 of the diagonal scale (x is then good to ~1e-10)
    for (int rd = 0; rd < 5; rd++) {
      LANES(u, 3) {
        const int p = PP[rd][u], q = QQ[rd][u];
        const ikr apq = A[p * 6 + q], d = A[q * 6 + q] - A[p * 6 + p];
        // tan of the rotation angle: t = sign(d) * 2 apq / (|d| + sqrt(d^2 + 4 apq^2)); (c, s) = (1, t) / sqrt(1 + t^2).
        // ANY t gives an exactly orthogonal rotation once (c, s) are formed in fp64, so t itself is computed in fp32 (its
        // 1e-7 relative error leaves a_pq' ~ 1e-7 a_pq, below what the quadratic convergence delivers anyway) — an fp64
        // sqrt + divide here is a ~1000-cycle dependent chain, five times per sweep.
        float tf = 0.f;
        if (apq != 0) { const float df = (float)d, af = (float)apq; const float hf = sqrtf(df * df + 4.f * af * af); tf = 2.f * af / (df >= 0.f ? df + hf : df - hf); }
        const ikr t = (ikr)tf;
        const ikr c = ik_rsqrt(1 + t * t);
        csb[2 * u] = c; csb[2 * u + 1] = t * c;
      }
      gsync<G>(cx);
      LANES(e, 36) {
        const int i0 = e / 6, j0 = e - 6 * i0, i = i0 < j0 ? i0 : j0, j = i0 < j0 ? j0 : i0;
        // index k of the round: partner pk, new column k = ak * old column k + bk * old column pk  (p: c, -s ; q: c, +s)
        int pi = 0, pj = 0, pj0 = 0; ikr ai = 1, bi = 0, aj = 1, bj = 0, aj0 = 1, bj0 = 0;
        for (int u = 0; u < 3; u++) {
          const int p = PP[rd][u], q = QQ[rd][u]; const ikr c = csb[2 * u], s = csb[2 * u + 1];
          if (i == p) { pi = q; ai = c; bi = -s; } else if (i == q) { pi = p; ai = c; bi = s; }
          if (j == p) { pj = q; aj = c; bj = -s; } else if (j == q) { pj = p; aj = c; bj = s; }
          if (j0 == p) { pj0 = q; aj0 = c; bj0 = -s; } else if (j0 == q) { pj0 = p; aj0 = c; bj0 = s; }
        }
        B1[e] = (ai * aj) * A[i * 6 + j] + (ai * bj) * A[i * 6 + pj] + (bi * aj) * A[pi * 6 + j] + (bi * bj) * A[pi * 6 + pj];
        B2[e] = aj0 * V[i0 * 6 + j0] + bj0 * V[i0 * 6 + pj0];
      }
      gsync<G>(cx);
      LANES(e, 36) { A[e] = B1[e]; V[e] = B2[e]; }
      gsync<G>(cx);
    }
  }
  ikr y[6];
  for (int c = 0; c < 6; c++) { ikr sum = 0; for (int r = 0; r < 6; r++) sum += V[r * 6 + c] * rhs[r]; y[c] = sum / ik_clamp(fabs(A[c * 6 + c]), lo, hi); }
  for (int r = 0; r < 6; r++) { ikr sum = 0; for (int c = 0; c < 6; c++) sum += V[r * 6 + c] * y[c]; x[r] = sum; }
}


// One getControl() call of CartPosQuatImpedenceController: num_iter damped-least-squares iterations on the open-loop
// joint reference; outputs the joint PD set-point (q_des as two floats, qd_des) for this physics tick.
// `active` = 0: the thread only takes part in the warp-cooperative section (its own outputs are discarded).
// coop: 160 doubles of scratch shared by the lane group (k_ik: shared memory per warp; host: a local array).
template <int G>
DEVFN void ik_tick(const Cx& cx, const ikr* C, IkState& s, int active, ikr* V, int* vwarm, ikr* sn, ikr* cs, ikr* J /*42 * IK_JSTRIDE*/, ikr* Aw /*27 * IK_JSTRIDE*/, ikr* coop) {
  ikr q[7], des_quat[4] = {(ikr)s.des_quat[0], (ikr)s.des_quat[1], (ikr)s.des_quat[2], (ikr)s.des_quat[3]};
  for (int k = 0; k < 7; k++) q[k] = s.q[k];
  // *vwarm: bit 0 = sn/cs hold the sines/cosines of q, bit 1 = V[0..35] holds an eigenbasis of an earlier Jacobi call,
  //         bit 2 = V[36..41] holds the smallest eigenvector of an earlier ik_solve_spd call
  if (active && !(*vwarm & 1)) { for (int k = 0; k < 7; k++) { sn[k] = sin(q[k]); cs[k] = cos(q[k]); } *vwarm |= 1; }      // exact once per launch
  const int niter = (int)C[D3C_NUM_ITER];
#pragma unroll 1
  for (int it = 0; it < niter; it++) {
    ikr pos[3], cq[4], rhs[6], qd_null[7], x[6];
    int need = 0;
    if (active) {      // lanes without an env (or in joint-PD mode) only take part in the cooperative section below
    ik_fk(C, sn, cs, pos, cq, J);
    ikr dm = 0, dp = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { dm += (cq[k] - des_quat[k]) * (cq[k] - des_quat[k]); dp += (cq[k] + des_quat[k]) * (cq[k] + des_quat[k]); }
    if (dm > dp) {
#pragma unroll
      for (int k = 0; k < 4; k++) des_quat[k] = -des_quat[k];
    }
    ikr qe[3];
    qe[0] = cq[0] * des_quat[1] - des_quat[0] * cq[1] - cq[3] * des_quat[2] + cq[2] * des_quat[3];
    qe[1] = cq[0] * des_quat[2] - des_quat[0] * cq[2] + cq[3] * des_quat[1] - cq[1] * des_quat[3];
    qe[2] = cq[0] * des_quat[3] - des_quat[0] * cq[3] - cq[2] * des_quat[1] + cq[1] * des_quat[2];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      rhs[k] = (ikr)C[D3C_PGAIN_POS + k] * ik_clamp((ikr)s.des_pos[k] - pos[k], -0.01, 0.01);
      rhs[3 + k] = (ikr)C[D3C_PGAIN_QUAT + k] * ik_clamp(qe[k], -0.1, 0.1);
    }
#pragma unroll
    for (int k = 0; k < 7; k++) qd_null[k] = (ikr)C[D3C_PGAIN_NULL + k] * ik_clamp((ikr)C[D3C_REST + k] - q[k], -0.2, 0.2);
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
      for (int k = 0; k < 7; k++) rhs[r] -= IKJ(J, r * 7 + k) * qd_null[k];
    }
    int v1ok = (*vwarm >> 2) & 1;
    need = !ik_solve_spd(J, (ikr)C[D3C_JREG], rhs, (ikr)C[D3C_SVD_MIN], (ikr)C[D3C_SVD_MAX], x, V + 36, &v1ok, Aw);
    if (v1ok) *vwarm |= 4;
    }
    // Some eigenvalue is (or may be) outside the clip range: eigen-decomposition with the clipped spectrum.  Rare (~2 % of
    // the iterations of the random-walk workload, near the edge of the arm's reach) but 5x the cost of everything else, and in
    // k_ik one needy env would stall the 31 others of its warp: so the WARP serves its needy envs one after the other, all
    // lanes cooperating on one 6x6 matrix in shared memory (coop).  Host build: G = 1, the env serves itself.
#ifdef __CUDA_ARCH__
    unsigned todo = __ballot_sync(cx.mask, need);
#else
    unsigned todo = need ? 1u : 0u;
#endif
    while (todo) {
#ifdef __CUDA_ARCH__
      const int src = __ffs(todo) - 1;
#else
      const int src = cx.lane;
#endif
      todo &= todo - 1;
      ikr *cA = coop, *cV = coop + 36, *cB1 = coop + 72, *cB2 = coop + 108, *cR = coop + 144, *cC = coop + 152;
      const ikr* Js = J + (src - cx.lane);                      // the needy env's Jacobian (J[k][thread] in shared memory)
      const bool mine = cx.lane == src;
      int warm = (*vwarm >> 1) & 1;
#ifdef __CUDA_ARCH__
      warm = __shfl_sync(cx.mask, warm, src);
#endif
      LANES(e, 36) {
        const int i = e / 6, j = e - 6 * i, a2 = i < j ? i : j, b2 = i < j ? j : i;
        ikr sum = a2 == b2 ? (ikr)C[D3C_JREG] : (ikr)0;
        for (int k = 0; k < 7; k++) sum += IKJ(Js, a2 * 7 + k) * IKJ(Js, b2 * 7 + k);
        cA[e] = sum;
      }
      if (mine) { for (int k = 0; k < 6; k++) cR[k] = rhs[k]; if (warm) for (int k = 0; k < 36; k++) cV[k] = V[k]; }
      gsync<G>(cx);
      ikr x2[6];
      ik_solve_clipped_lanes<G>(cx, cA, cR, cV, cB1, cB2, cC, warm, (ikr)C[D3C_SVD_MIN], (ikr)C[D3C_SVD_MAX], x2);
      if (mine) { for (int k = 0; k < 6; k++) x[k] = x2[k]; for (int k = 0; k < 36; k++) V[k] = cV[k]; *vwarm |= 2; }
      gsync<G>(cx);
    }
    if (active) {
    ikr qd[7], nrm = 0;
#pragma unroll
    for (int k = 0; k < 7; k++) {
      ikr sum = qd_null[k];
#pragma unroll
      for (int r = 0; r < 6; r++) sum += IKJ(J, r * 7 + k) * x[r];
      qd[k] = sum; nrm += sum * sum;
    }
    nrm = sqrt(nrm);
    const ikr scl = nrm > 3 ? 3 / nrm : (ikr)1;
#pragma unroll
    for (int k = 0; k < 7; k++) {
      ikr qn = ik_clamp(q[k] + (ikr)C[D3C_LRATE] * (qd[k] * scl), (ikr)C[D3C_JMIN + k], (ikr)C[D3C_JMAX + k]);
      // sin/cos by the angle-addition formulas with a short series in the increment (|dq| <= 3e-3: the dq^6 term is
      // 1e-18), then one Newton step back onto the unit circle: 1e-16 per update instead of a software sincos
      ikr dq = qn - q[k], d2 = dq * dq;
      ikr sd = dq * (1 - d2 * (1.0 / 6.0) * (1 - d2 * 0.05)), cd = 1 - d2 * 0.5 * (1 - d2 * (1.0 / 12.0));
      ikr s1 = sn[k] * cd + cs[k] * sd, c1 = cs[k] * cd - sn[k] * sd;
      ikr nn = 1.5 - 0.5 * (s1 * s1 + c1 * c1);
      sn[k] = s1 * nn; cs[k] = c1 * nn;
      q[k] = qn;
    }
    }
  }
  if (active) for (int k = 0; k < 7; k++) {
    s.jt_q[k] = (real)q[k];
    s.jt_qlo[k] = (real)(q[k] - (double)s.jt_q[k]);
    s.jt_qd[k] = (real)((q[k] - s.q[k]) / (ikr)C[D3C_DT]);
    s.q[k] = q[k];
  }
}

// ------------------------------------------------------------------------------------------------ task logic
DEVFN real dist3(const real* a, const real* b) { real d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; return norm3(d); }

// tan(quat2euler(q)[-1]) — utils/geometric_transformation.py:92-153
DEVFN real tan_yaw(const real* q) {
  real R[9]; quat2mat(R, q);
  real cy = sqrt(R[8] * R[8] + R[5] * R[5]);
  real yaw = cy > (real)8.881784197001252e-16 ? -atan2(R[1], R[0]) : -atan2(-R[3], R[4]);
  return tan(yaw);
}

DEVFN void task_obs(const Model& m, const Lay& L, const real* w, float* obs) {
  if (m.task_id == D3T_INSERTING) {           // gate_insertion.py:288-320 (same layout as Sorting: tcp xy, then xy + tan yaw per box)
    obs[0] = (float)w[L.tcp]; obs[1] = (float)w[L.tcp + 1];
    for (int i = 0; i < 3; i++) {
      const real* b = w + L.qpos + D3_NROB + 7 * i;
      obs[2 + 3 * i] = (float)b[0]; obs[3 + 3 * i] = (float)b[1]; obs[4 + 3 * i] = (float)tan_yaw(b + 3);
    }
  } else if (m.task_id == D3T_PUSHING) {
    const real *b1 = w + L.qpos + 9, *b2 = w + L.qpos + 16;
    obs[0] = (float)w[L.tcp]; obs[1] = (float)w[L.tcp + 1];
    obs[2] = (float)b1[0]; obs[3] = (float)b1[1]; obs[4] = (float)tan_yaw(b1 + 3);
    obs[5] = (float)b2[0]; obs[6] = (float)b2[1]; obs[7] = (float)tan_yaw(b2 + 3);
  } else if (m.task_id == D3T_SORTING) {      // sorting.py:308-390
    obs[0] = (float)w[L.tcp]; obs[1] = (float)w[L.tcp + 1];
    for (int i = 0; i < m.nobj; i++) {
      const real* b = w + L.qpos + D3_NROB + 7 * i;
      obs[2 + 3 * i] = (float)b[0]; obs[3 + 3 * i] = (float)b[1]; obs[4 + 3 * i] = (float)tan_yaw(b + 3);
    }
  } else if (m.task_id == D3T_STACKING) {     // stacking.py:228-277
    for (int i = 0; i < 3; i++) {
      const real* b = w + L.qpos + D3_NROB + 7 * i;
      obs[4 * i] = (float)b[0]; obs[4 * i + 1] = (float)b[1]; obs[4 * i + 2] = (float)b[2]; obs[4 * i + 3] = (float)tan_yaw(b + 3);
    }
  } else if (m.task_id == D3T_ALIGNING) {     // aligning.py:205-235
    const real* b = w + L.qpos + D3_NROB;
    for (int k = 0; k < 3; k++) obs[k] = (float)w[L.tcp + k];
    for (int k = 0; k < 7; k++) { obs[3 + k] = (float)b[k]; obs[10 + k] = (float)w[L.extra + k]; }
  } else {
    obs[0] = (float)w[L.tcp]; obs[1] = (float)w[L.tcp + 1];
  }
}

// ---- Sorting (sorting.py:460-543).  misc TASK0 = mode_step, TASK1 = bit i set <=> mode[i] == 0 (red), TASK2 = min_inds mask.
// Slots are (red_1..3, blue_1..3); slots absent from the scene read the constant pose of the model's last body.
DEVFN void sorting_slot_xy(const Model& m, const Lay& L, const real* w, int slot, real* xy) {
  int half = m.nobj / 2, col = slot / 3, idx = slot - 3 * col;
  if (idx < half) { const real* b = w + L.qpos + D3_NROB + 7 * (col * half + idx); xy[0] = b[0]; xy[1] = b[1]; }
  else { xy[0] = (real)m.taskp[11]; xy[1] = (real)m.taskp[12]; }
}
DEVFN int sorting_in_bin(const Model& m, int col, const real* xy) {
  const tab_t* T = m.taskp;
  return xy[0] > (real)T[4 + 2 * col] && xy[0] < (real)T[5 + 2 * col] && xy[1] > (real)T[8] && xy[1] < (real)T[9];
}
DEVFN int sorting_all_binned(const Model& m, const Lay& L, const real* w) {
  int half = m.nobj / 2;
  for (int i = 0; i < m.nobj; i++) if (!sorting_in_bin(m, i / half, w + L.qpos + D3_NROB + 7 * i)) return 0;
  return 1;
}
DEVFN int sorting_check_mode(const Model& m, const Lay& L, real* w) {
  int step = (int)w[L.misc + ST_TASK0], zero = (int)w[L.misc + ST_TASK1], mins = (int)w[L.misc + ST_TASK2];
  if (step <= 5) {
    real best = 0, bxy[2] = {0, 0}; int bi = -1;
    for (int s = 0; s < 6; s++) {
      real xy[2]; sorting_slot_xy(m, L, w, s, xy);
      real dx = xy[0] - (real)m.taskp[2 * (s / 3)], dy = xy[1] - (real)m.taskp[2 * (s / 3) + 1];
      real d = ((mins >> s) & 1) ? (real)100000 : sqrt(dx * dx + dy * dy);
      if (bi < 0 || d < best) { best = d; bi = s; bxy[0] = xy[0]; bxy[1] = xy[1]; }
    }
    if (sorting_in_bin(m, bi / 3, bxy)) {
      if (bi < 3) zero |= 1 << step;
      step++; mins |= 1 << bi;
    }
    w[L.misc + ST_TASK0] = (real)step; w[L.misc + ST_TASK1] = (real)zero; w[L.misc + ST_TASK2] = (real)mins;
  }
  int code = 0;
  for (int i = 0; i < m.nobj; i++) if (!((zero >> i) & 1)) code |= 1 << (7 - i);
  return code;
}

// ---- Stacking (stacking.py:395-447).  misc TASK0 = len(mode_encoding), TASK1 = mode string in base-4 digits (r 1, g 2, b 3),
// TASK2 = min_inds mask.  taskp: target xy, pos_min_dist, min z gap, gripper-open threshold.
DEVFN real hypot2(real x, real y) { return sqrt(x * x + y * y); }
DEVFN real stacking_check_mode(const Model& m, const Lay& L, real* w) {
  const tab_t* T = m.taskp;
  int len = (int)w[L.misc + ST_TASK0], code = (int)w[L.misc + ST_TASK1], mins = (int)w[L.misc + ST_TASK2];
  real d[3], mean = 0, best = 0; int bi = -1;
  for (int i = 0; i < 3; i++) { const real* b = w + L.qpos + D3_NROB + 7 * i; d[i] = hypot2(b[0] - (real)T[0], b[1] - (real)T[1]); mean += d[i] / 3; }
  for (int i = 0; i < 3; i++) { real di = ((mins >> i) & 1) ? (real)100000 : d[i]; if (bi < 0 || di < best) { best = di; bi = i; } }
  if (best <= (real)T[2]) { int p4 = 1; for (int k = 0; k < len; k++) p4 *= 4; code += (bi + 1) * p4; len++; mins |= 1 << bi; }
  w[L.misc + ST_TASK0] = (real)len; w[L.misc + ST_TASK1] = (real)code; w[L.misc + ST_TASK2] = (real)mins;
  return mean;
}

// ---- Inserting (gate_insertion.py:396-470).  misc TASK0 = len(modes), TASK1 = arrival order in base-4 digits, TASK2 = seen mask.
DEVFN void inserting_dists(const Model& m, const Lay& L, const real* w, real* d3) {
  for (int i = 0; i < 3; i++) { real t[3] = {(real)m.taskp[3 * i], (real)m.taskp[3 * i + 1], (real)m.taskp[3 * i + 2]}; d3[i] = dist3(w + L.qpos + D3_NROB + 7 * i, t); }
}

// ---- Aligning (aligning.py:21-30,295-352)
DEVFN void aligning_dists(const Model& m, const Lay& L, const real* w, real* pd, real* rd) {
  const real *b = w + L.qpos + D3_NROB, *t = w + L.extra;
  *pd = dist3(b, t);
  real d = absr(b[3] * t[3] + b[4] * t[4] + b[5] * t[5] + b[6] * t[6]);
  *rd = 2 * acos(d > 1 ? (real)1 : d) * (real)0.3183098861837907;
}

DEVFN void pushing_dists(const Model& m, const Lay& L, const real* w, real* d4) {
  const real *b1 = w + L.qpos + 9, *b2 = w + L.qpos + 16;
  real g1[3] = {(real)m.taskp[0], (real)m.taskp[1], (real)m.taskp[2]}, g2[3] = {(real)m.taskp[3], (real)m.taskp[4], (real)m.taskp[5]};
  d4[0] = dist3(b1, g1); d4[1] = dist3(b1, g2); d4[2] = dist3(b2, g1); d4[3] = dist3(b2, g2);
}

// _check_early_termination (sets `terminated`); serial helper, call from a single lane
DEVFN int task_early_term(const Model& m, const Lay& L, real* w) {
  if (m.task_id == D3T_PUSHING) {
    real d[4], md = m.taskp[6];
    pushing_dists(m, L, w, d);
    if ((d[0] <= md && d[3] <= md) || (d[1] <= md && d[2] <= md)) { w[L.misc + ST_TERM] = 1; return 1; }
    return 0;
  }
  if (m.task_id == D3T_SORTING) {
    if (sorting_all_binned(m, L, w)) { w[L.misc + ST_TERM] = 1; return 1; }
    return 0;
  }
  if (m.task_id == D3T_INSERTING) {
    real d[3]; inserting_dists(m, L, w, d);
    if (d[0] <= (real)m.taskp[9] && d[1] <= (real)m.taskp[9] && d[2] <= (real)m.taskp[9]) { w[L.misc + ST_TERM] = 1; return 1; }
    return 0;
  }
  if (m.task_id == D3T_ALIGNING) {
    real pd, rd; aligning_dists(m, L, w, &pd, &rd);
    if (pd <= (real)m.taskp[0] && rd <= (real)m.taskp[1]) { w[L.misc + ST_TERM] = 1; return 1; }
    return 0;
  }
  if (m.task_id == D3T_STACKING) {
    const real *r = w + L.qpos + D3_NROB, *g = r + 7, *b = g + 7; const tab_t* T = m.taskp;
    real dz = minr(absr(r[2] - g[2]), minr(absr(r[2] - b[2]), absr(g[2] - b[2])));
    real dr = hypot2(r[0] - (real)T[0], r[1] - (real)T[1]), dg = hypot2(g[0] - (real)T[0], g[1] - (real)T[1]), db = hypot2(b[0] - (real)T[0], b[1] - (real)T[1]);
    if (dr <= (real)T[2] && dg <= (real)T[2] && db <= (real)T[2] && dz > (real)T[3]) { w[L.misc + ST_TERM] = 1; return 1; }
    return 0;
  }
  int success = w[L.tcp + 1] > (real)m.taskp[3];
  if (success || w[L.misc + ST_OBST] != 0) { if (success) w[L.misc + ST_TASK1] = 1; w[L.misc + ST_TERM] = 1; return 1; }
  return 0;
}

DEVFN real task_reward(const Model& m, const Lay& L, const real* w) {
  if (m.task_id == D3T_PUSHING) {
    const real* b1 = w + L.qpos + 9;
    real g1[3] = {(real)m.taskp[0], (real)m.taskp[1], (real)m.taskp[2]};
    real dx = w[L.tcp] - b1[0], dy = w[L.tcp + 1] - b1[1];
    return -(sqrt(dx * dx + dy * dy) + dist3(b1, g1));
  }
  if (m.task_id == D3T_ALIGNING) { real pd, rd; aligning_dists(m, L, w, &pd, &rd); return -rd - (real)3.5 * pd; }
  if (m.task_id == D3T_INSERTING) {
    real d[3], mn = (real)1e30; inserting_dists(m, L, w, d);
    for (int i = 0; i < 3; i++) { const real* b = w + L.qpos + D3_NROB + 7 * i; real dx = w[L.tcp] - b[0], dy = w[L.tcp + 1] - b[1], r = sqrt(dx * dx + dy * dy); if (r < mn) mn = r; }
    return -(mn + d[0] + d[1] + d[2]);
  }
  return 0;
}

// info after the substeps: pushing [success, mode, mean_distance, status]; avoiding [success, 9 mode bits, status]
DEVFN void task_post(const Model& m, const Lay& L, real* w, float* info) {
  for (int k = 0; k < m.info_dim; k++) info[k] = 0;
  if (m.task_id == D3T_PUSHING) {
    int success = task_early_term(m, L, w);
    real d[4], md = m.taskp[6];
    pushing_dists(m, L, w, d);
    int fv = (int)w[L.misc + ST_TASK0], visit = -1, mode = -1;
    if (d[0] <= md && fv != 0) visit = 0; else if (d[1] <= md && fv != 1) visit = 1; else if (d[2] <= md && fv != 2) visit = 2; else if (d[3] <= md && fv != 3) visit = 3;
    if (fv == -1) w[L.misc + ST_TASK0] = (real)visit;
    else {
      if (fv == 0 && visit == 3) mode = 0; else if (fv == 3 && visit == 0) mode = 1; else if (fv == 1 && visit == 2) mode = 2; else if (fv == 2 && visit == 1) mode = 3;
    }
    info[0] = (float)success; info[1] = (float)mode; info[2] = (float)((real)0.5 * (minr(d[0], d[1]) + minr(d[2], d[3]))); info[3] = (float)w[L.misc + ST_STATUS];
  } else if (m.task_id == D3T_SORTING) {      // info: success, packed mode, mode_step, status
    int success = task_early_term(m, L, w);
    int code = sorting_check_mode(m, L, w);
    info[0] = (float)success; info[1] = (float)code; info[2] = (float)w[L.misc + ST_TASK0]; info[3] = (float)w[L.misc + ST_STATUS];
  } else if (m.task_id == D3T_INSERTING) {    // info: success, mode id (mode_dict, 0 unless three boxes arrived), mean_distance, len(modes), status
    int success = task_early_term(m, L, w);
    real d[3]; inserting_dists(m, L, w, d);
    int len = (int)w[L.misc + ST_TASK0], code = (int)w[L.misc + ST_TASK1], seen = (int)w[L.misc + ST_TASK2];
    for (int i = 0; i < 3; i++) if (d[i] <= (real)m.taskp[9] && !((seen >> i) & 1)) {
      int p4 = 1; for (int k = 0; k < len; k++) p4 *= 4;
      code += (i + 1) * p4; len++; seen |= 1 << i;
    }
    w[L.misc + ST_TASK0] = (real)len; w[L.misc + ST_TASK1] = (real)code; w[L.misc + ST_TASK2] = (real)seen;
    int a = code % 4, b = (code / 4) % 4;
    int id = len != 3 ? 0 : (a == 1 ? (b == 2 ? 1 : 2) : a == 2 ? (b == 1 ? 3 : 4) : (b == 1 ? 5 : 6));
    info[0] = (float)success; info[1] = (float)id; info[2] = (float)((d[0] + d[1] + d[2]) / 3); info[3] = (float)len; info[4] = (float)w[L.misc + ST_STATUS];
  } else if (m.task_id == D3T_STACKING) {     // info: success, mode string (base-4 digits), mean_distance, len(mode), status
    int success = task_early_term(m, L, w);
    real md = stacking_check_mode(m, L, w);
    info[0] = (float)success; info[1] = (float)w[L.misc + ST_TASK1]; info[2] = (float)md; info[3] = (float)w[L.misc + ST_TASK0]; info[4] = (float)w[L.misc + ST_STATUS];
  } else if (m.task_id == D3T_ALIGNING) {     // info: success, mode (0 inside / 1 outside push), mean_distance, status
    int success = task_early_term(m, L, w);
    real pd, rd; aligning_dists(m, L, w, &pd, &rd);
    const real* b = w + L.qpos + D3_NROB;
    real dx = b[0] - w[L.tcp], dy = b[1] - w[L.tcp + 1];
    info[0] = (float)success; info[1] = sqrt(dx * dx + dy * dy) < (real)m.taskp[2] ? 0.f : 1.f; info[2] = (float)((real)0.5 * (pd + rd)); info[3] = (float)w[L.misc + ST_STATUS];
  } else {
    real x = w[L.tcp], y = w[L.tcp + 1];
    const tab_t* T = m.taskp;
    int passed = (int)w[L.misc + ST_TASK2], code = (int)w[L.misc + ST_TASK3];
    if (y - (real)0.03 <= (real)T[0] && (real)T[0] <= y + (real)0.03 && !(passed & 1)) { if (x < (real)T[4]) code |= 1; else if (x > (real)T[4]) code |= 2; passed |= 1; }
    if (y - (real)0.03 <= (real)T[1] && (real)T[1] <= y + (real)0.03 && !(passed & 2)) {
      if (x < (real)T[5]) code |= 4; else if ((real)T[5] < x && x < (real)T[6]) code |= 8; else if (x > (real)T[6]) code |= 16;
      passed |= 2;
    }
    if (y >= (real)T[2] && !(passed & 4)) {
      if (x < (real)T[7]) code |= 32;
      if ((real)T[7] < x && x < (real)T[8]) code |= 64; else if ((real)T[8] < x && x < (real)T[9]) code |= 128; else if (x > (real)T[7]) code |= 256;
      passed |= 4;
    }
    w[L.misc + ST_TASK2] = (real)passed; w[L.misc + ST_TASK3] = (real)code;
    info[0] = (float)w[L.misc + ST_TASK1];
    for (int k = 0; k < 9; k++) info[1 + k] = (float)((code >> k) & 1);
    info[10] = (float)w[L.misc + ST_STATUS];
  }
}

// ------------------------------------------------------------------------------------------------ Gym reset / step
// reset (a12): state <- (init_qpos, 0), forward pass for tcp + qfrc_bias, object poses <- context, ONE tick under joint PD.
template <int G>
DEVFN void env_reset(const Cx& cx, const Model& m, const Lay& L, real* w, const float* ctx /*[nobj*7] or null*/, real tol, int max_iter) {
  LANES(d, m.nq) w[L.qpos + d] = d < D3_NARM ? (real)m.ctrl[D3C_INIT_QPOS + d] : (real)0;
  LANES(d, m.nv) { w[L.qvel + d] = 0; w[L.warm + d] = 0; }
  LANES(d, D3_NROB) w[L.qlo + d] = 0;
  LANES(k, ST_NMISC) w[L.misc + k] = 0;
  gsync<G>(cx);
  LANES(i, m.nlink - D3_NROB) {
    const tab_t* Lk = m.link + D3_LINK_W * (D3_NROB + i);
    real* q = w + L.qpos + m.l_qadr[D3_NROB + i];
    for (int k = 0; k < 7; k++) q[k] = Lk[2 + k];
  }
  LANES(z, 1) { w[L.misc + ST_GRIP_SET] = (real)0.001; if (m.task_id == D3T_PUSHING) w[L.misc + ST_TASK0] = -1; }
  LANES(k, m.nextra) w[L.extra + k] = ctx ? (real)ctx[7 * m.nobj + k] : (real)m.taskp[3 + k];      // joint-less target body (Aligning)
  gsync<G>(cx);
  kinematics<G>(cx, m, L, w);
  LANES(z, 1) {
    real o[3], tp[3] = {(real)m.ctrl[D3C_TCP_POS], (real)m.ctrl[D3C_TCP_POS + 1], (real)m.ctrl[D3C_TCP_POS + 2]};
    real tq[4] = {(real)m.ctrl[D3C_TCP_QUAT], (real)m.ctrl[D3C_TCP_QUAT + 1], (real)m.ctrl[D3C_TCP_QUAT + 2], (real)m.ctrl[D3C_TCP_QUAT + 3]}, Rt[9], R[9];
    mat_vec3(o, w + L.xmat + 54, tp);
    for (int k = 0; k < 3; k++) w[L.tcp + k] = w[L.xpos + 18 + k] + o[k];
    quat2mat(Rt, tq); mat_mul3(R, w + L.xmat + 54, Rt); mat2quat(w + L.tcp + 3, R);
  }
  LANES(e, m.m_size) w[L.M + e] = 0;       // whole buffer once per reset: out-of-block entries stay zero afterwards
  gsync<G>(cx);
  dynamics<G>(cx, m, L, w);
  LANES(k, D3_NROB) w[L.bias_prev + k] = w[L.bias + k];
  gsync<G>(cx);
  if (ctx) {
    LANES(d, 7 * m.nobj) w[L.qpos + D3_NROB + d] = ctx[d];
    gsync<G>(cx);
  }
  if (m.task_id == D3T_STACKING) { LANES(z, 1) w[L.misc + ST_GRIP_SET] = (real)0.04; gsync<G>(cx); }      // stacking.py:474 open_fingers() before the reset tick
  real jq[D3_NARM], jql[D3_NARM], jqd[D3_NARM];
  for (int k = 0; k < D3_NARM; k++) { jq[k] = m.ctrl[D3C_INIT_QPOS + k]; jql[k] = 0; jqd[k] = 0; }
  if (m.maxdim == 4) physics_tick<G, false, 4>(cx, m, L, w, jq, jql, jqd, tol, max_iter);
  else physics_tick<G, false, 3>(cx, m, L, w, jq, jql, jqd, tol, max_iter);
  // scheduling hint: the first env steps after a reset resolve the deep spawn penetration (16 contacts, many Newton
  // iterations) and are among the most expensive ones, so a fresh env sorts to the front of the next step's order
  LANES(z, 1) w[L.misc + ST_COST_ITERS] = 400;
  gsync<G>(cx);
}

// An action the controllers cannot take as a set-point: a non-finite component, or (Cartesian scenes) a quaternion of zero
// norm.  The reference would propagate NaN into MuJoCo and warn; here the env keeps its previous set-point for this env step
// and raises D3_STATUS_BAD_ACTION.  Evaluated identically by k_ik (set-point) and env_prestep (status bit, gripper command).
DEVFN bool action_ok(const float* a, int act_dim, int ctrl_kind) {
  bool ok = true;
  for (int k = 0; k < act_dim; k++) ok = ok && (absr((real)a[k]) <= (real)3.0e38);      // false for NaN and +-inf
  if (ctrl_kind == 0) ok = ok && (a[3] * a[3] + a[4] * a[4] + a[5] * a[5] + a[6] * a[6] > 1e-12f);
  return ok;
}

// pre-substep half of GymEnvWrapper.step (gym_env_wrapper.py:67-90): open fingers, Cartesian mode, sample obs /
// reward / done BEFORE the substeps.
template <int G>
DEVFN void env_prestep(const Cx& cx, const Model& m, const Lay& L, real* w, const float* action, float* obs, float* reward, unsigned char* done) {
  LANES(z, 1) {
    const bool aok = action_ok(action, m.act_dim, m.ctrl_kind);
    if (!aok) w[L.misc + ST_STATUS] = (real)(((int)w[L.misc + ST_STATUS]) | D3_STATUS_BAD_ACTION);
    if (m.ctrl_kind == 1) {        // CubeStacking_Env.step (stacking.py:337-346): gripper command in action[7], joint-space set-point
      if (aok) {
        const bool open = action[7] > (float)m.taskp[4];
        w[L.misc + ST_GRIP_SET] = open ? (real)0.04 : (real)0; w[L.misc + ST_GRASP] = open ? (real)0 : (real)1;
      }
      w[L.misc + ST_CTRL_MODE] = 2;
    } else { w[L.misc + ST_GRIP_SET] = (real)0.04; w[L.misc + ST_GRASP] = 0; w[L.misc + ST_CTRL_MODE] = 1; }
    w[L.misc + ST_COST_ITERS] = 0; w[L.misc + ST_COST_COUPLED] = 0; w[L.misc + ST_COST_NCON] = 0; w[L.misc + ST_COST_NEAR] = 0;
    task_obs(m, L, w, obs);
    *reward = (float)task_reward(m, L, w);
    int early = task_early_term(m, L, w);
    *done = (w[L.misc + ST_TERM] != 0 || early || (int)w[L.misc + ST_STEP] >= m.max_steps - 1) ? 1 : 0;
  }
  gsync<G>(cx);
}
template <int G>
DEVFN void env_poststep(const Cx& cx, const Model& m, const Lay& L, real* w, float* info) {
  LANES(z, 1) {
    w[L.misc + ST_STEP] += 1;
    task_post(m, L, w, info);
  }
  gsync<G>(cx);
}
