/*
 * d3il_oracle.c — fp64 single-env CPU restatement of the D3IL env-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under d3il_b200/ may link, import or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs use it (as the checker / the timed CPU baseline).
 *
 * PARITY UNPINNED: the arithmetic of this path lives in mujoco==2.3.2 and pinocchio
 * (reference install.sh:36,57), neither of which is present in /root/reference or
 * installable here, and the reference ships no golden vectors for the path
 * (SURVEY.md §8c).  This file restates (a) the reference's own Python logic, cited
 * per function as file:line under /root/reference/environments/d3il/, and (b) the
 * published MuJoCo 2.3.x algorithm (computation chapter: soft constraints, elliptic
 * cones, Newton solver, semi-implicit Euler) as summarised in SURVEY.md App. B.
 * Deliberate, documented deviations from MuJoCo: narrow-phase contact generation is
 * our own deterministic SAT/feature rule (DESIGN.md "collision"), and the solver
 * warm start after reset is zero.
 *
 * One env per handle, all state in `Env`; dense linear algebra; no allocation after
 * create.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "d3il_oracle.h"

/* ------------------------------------------------------------------ scene blob (d3il_b200/scene/blob.py) */
#define HDR_INTS 32
#define LINK_W 32
#define GEOM_W 24
#define PAIR_W 24
#define CTRL_W 192
#define MAGIC 0x43533344

enum { C_IK_ORIGIN = 0, C_IK_EE = 84, C_PGAIN_POS = 96, C_PGAIN_QUAT = 99, C_PGAIN_NULL = 102, C_REST = 109, C_JMIN = 116,
       C_JMAX = 123, C_PD_P = 130, C_PD_D = 137, C_JREG = 144, C_SVD_MIN = 145, C_SVD_MAX = 146, C_NUM_ITER = 147, C_LRATE = 148,
       C_DT = 149, C_INIT_QPOS = 150, C_TCP_POS = 157, C_TCP_QUAT = 160, C_GRAVITY = 164, C_IMPRATIO = 167, C_TOL = 168,
       C_JNT_SOLREF = 169, C_JNT_SOLIMP = 171, C_MEANINERTIA = 179 };

enum { G_PLANE = 0, G_SPHERE = 2, G_CYLINDER = 5, G_BOX = 6 };
enum { TASK_AVOIDING = 0, TASK_PUSHING = 1, TASK_ALIGNING = 2, TASK_SORTING = 3, TASK_STACKING = 4, TASK_INSERTING = 5 };

#define MAXLINK 16
#define MAXQ 64
#define MAXV 48
#define MAXGEOM 32
#define MAXCON 96
#define MAXEFC 320
#define MAXPAIR 128
#define NARM 7
#define NROB 9

typedef struct {
  int parent, jtype, limited, qadr, dadr, ndof;
  double pos[3], quat[4], axis[3], mass, ipos[3], inertia[9], range[2], damping, forcerange, invw[2];
  uint32_t anc;              /* bitmask of ancestor links incl. self */
} Link;

typedef struct { int type, link, tag; double pos[3], R[9], size[3], rbound, invw[2]; } Geom;
typedef struct { int g1, g2, condim, flags; double friction[5], solref[2], solimp[5], margin, gap; } Pair;

typedef struct {
  double pos[3], frame[9], dist, includemargin, friction[5], solref[2], solimp[5], mu;
  int dim, g1, g2, pair, efc;
} Contact;

typedef struct Env {
  /* model */
  int task_id, nlink, nobj, nq, nv, ngeom, npair, n_substeps, max_steps, obs_dim, act_dim, ctx_dim, info_dim, ctrl_kind, ntaskp, nextra;
  Link link[MAXLINK];
  Geom geom[MAXGEOM];
  Pair pair[MAXPAIR];
  double ctrl[CTRL_W], taskp[32];
  int dof_link[MAXV];
  /* state (see d3o_get_state for the flat layout) */
  double qpos[MAXQ], qvel[MAXV], warm[MAXV];
  double bias_prev[NROB], tcp_pos[3], tcp_quat[4];
  double ik_q[NARM], des_pos[3], des_quat[4], jt_q[NARM], jt_qd[NARM];
  int ik_valid, ctrl_mode, grasp_flag, step_count, terminated, status, obst_contact;
  double grip_set;
  double task_state[8];
  double extra[8];            /* per-env task data outside qpos: Aligning target pose (model.body_pos/quat of `target_box`) */
  /* scratch of the last forward pass (exposed to tests) */
  double xpos[MAXLINK][3], xmat[MAXLINK][9];
  double S[MAXV][6];
  double M[MAXV * MAXV], bias[MAXV], qfrc_smooth[MAXV], qacc_smooth[MAXV], qacc[MAXV], qfrc_constraint[MAXV];
  Contact con[MAXCON];
  int ncon, nefc, solver_iter;
  double J[MAXEFC * MAXV], aref[MAXEFC], D[MAXEFC], Rr[MAXEFC], efc_force[MAXEFC];
  int efc_type[MAXEFC], efc_con[MAXEFC];   /* type 0 limit, 1 cone first row, 2 cone other row */
  double ctrl_out[NROB];
  long flops_dummy;
} Env;

/* ------------------------------------------------------------------ small math */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double* o, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static double norm3(const double* a) { return sqrt(dot3(a, a)); }
static void mat_vec3(double* o, const double* R, const double* v) {   /* o = R v (row-major) */
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2], z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void matT_vec3(double* o, const double* R, const double* v) {  /* o = R^T v */
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2], z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void mat_mul3(double* o, const double* A, const double* B) {
  double t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(o, t, sizeof t);
}
static void quat2mat(double* R, const double* qin) {
  double n = sqrt(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
  double w = qin[0] / n, x = qin[1] / n, y = qin[2] / n, z = qin[3] / n;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void mat2quat(double* q, const double* R) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) { double s = sqrt(t + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= n;
}
static void quat_mul(double* o, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* dense Cholesky A = L L^T in place (lower), returns 0 on success */
static int chol_factor(double* A, int n, int ld) {
  for (int j = 0; j < n; j++) {
    double s = A[j * ld + j];
    for (int k = 0; k < j; k++) s -= A[j * ld + k] * A[j * ld + k];
    if (!(s > 0)) return -1;
    double d = sqrt(s);
    A[j * ld + j] = d;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * ld + j];
      for (int k = 0; k < j; k++) t -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = t / d;
    }
  }
  return 0;
}
static void chol_solve(const double* L, int n, int ld, double* x) {
  for (int i = 0; i < n; i++) { double t = x[i]; for (int k = 0; k < i; k++) t -= L[i * ld + k] * x[k]; x[i] = t / L[i * ld + i]; }
  for (int i = n - 1; i >= 0; i--) { double t = x[i]; for (int k = i + 1; k < n; k++) t -= L[k * ld + i] * x[k]; x[i] = t / L[i * ld + i]; }
}

/* cyclic Jacobi eigen-decomposition of a symmetric n x n (n<=6) matrix: A = V diag(w) V^T */
static void jacobi_eig(double* A, int n, double* w, double* V) {
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) { diag += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j]; }
    if (off <= 1e-40 * diag || off == 0) break;
    for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
      double apq = A[p * n + q];
      if (apq == 0) continue;
      double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
      double c = 1 / sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < n; k++) { double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
      for (int k = 0; k < n; k++) { double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
    }
  }
  for (int i = 0; i < n; i++) w[i] = A[i * n + i];
}

/* ------------------------------------------------------------------ scene loading */
static const char* g_err = "";
const char* d3o_last_error(void) { return g_err; }

Env* d3o_create(const void* blob, size_t nbytes) {
  const int32_t* h = (const int32_t*)blob;
  if (nbytes < 4 * HDR_INTS || h[0] != MAGIC || h[1] != 1) { g_err = "bad scene blob"; return NULL; }
  Env* e = (Env*)calloc(1, sizeof(Env));
  e->task_id = h[2]; e->nlink = h[3]; e->nobj = h[4]; e->nq = h[5]; e->nv = h[6]; e->ngeom = h[7]; e->npair = h[8];
  e->n_substeps = h[9]; e->max_steps = h[10]; e->obs_dim = h[11]; e->act_dim = h[12]; e->ctx_dim = h[13]; e->info_dim = h[14];
  e->ctrl_kind = h[15]; e->ntaskp = h[16]; e->nextra = h[17];
  if (e->nlink > MAXLINK || e->nq > MAXQ || e->nv > MAXV || e->ngeom > MAXGEOM || e->npair > MAXPAIR || e->ntaskp > 32 || e->nextra > 8) { g_err = "scene too large"; free(e); return NULL; }
  size_t need = 4 * HDR_INTS + 8 * ((size_t)e->nlink * LINK_W + (size_t)e->ngeom * GEOM_W + (size_t)e->npair * PAIR_W + CTRL_W + e->ntaskp);
  if (need != nbytes) { g_err = "scene blob size mismatch"; free(e); return NULL; }
  const double* p = (const double*)((const char*)blob + 4 * HDR_INTS);
  for (int i = 0; i < e->nlink; i++, p += LINK_W) {
    Link* L = &e->link[i];
    L->parent = (int)p[0]; L->jtype = (int)p[1];
    memcpy(L->pos, p + 2, 24); memcpy(L->quat, p + 5, 32); memcpy(L->axis, p + 9, 24);
    L->mass = p[12]; memcpy(L->ipos, p + 13, 24);
    double xx = p[16], yy = p[17], zz = p[18], xy = p[19], xz = p[20], yz = p[21];
    double I[9] = {xx, xy, xz, xy, yy, yz, xz, yz, zz};
    memcpy(L->inertia, I, sizeof I);
    L->limited = (int)p[22]; L->range[0] = p[23]; L->range[1] = p[24]; L->damping = p[25]; L->forcerange = p[26];
    L->invw[0] = p[27]; L->invw[1] = p[28]; L->qadr = (int)p[29]; L->dadr = (int)p[30];
    L->ndof = L->jtype == 2 ? 6 : 1;
    L->anc = (1u << i) | (L->parent >= 0 ? e->link[L->parent].anc : 0);
    for (int k = 0; k < L->ndof; k++) e->dof_link[L->dadr + k] = i;
  }
  for (int i = 0; i < e->ngeom; i++, p += GEOM_W) {
    Geom* g = &e->geom[i];
    g->type = (int)p[0]; g->link = (int)p[1]; memcpy(g->pos, p + 2, 24); quat2mat(g->R, p + 5);
    memcpy(g->size, p + 9, 24); g->rbound = p[12]; g->invw[0] = p[13]; g->invw[1] = p[14]; g->tag = (int)p[15];
  }
  for (int i = 0; i < e->npair; i++, p += PAIR_W) {
    Pair* q = &e->pair[i];
    q->g1 = (int)p[0]; q->g2 = (int)p[1]; q->condim = (int)p[2]; memcpy(q->friction, p + 3, 40); memcpy(q->solref, p + 8, 16);
    memcpy(q->solimp, p + 10, 40); q->margin = p[15]; q->gap = p[16]; q->flags = (int)p[17];
  }
  memcpy(e->ctrl, p, 8 * CTRL_W); p += CTRL_W;
  memcpy(e->taskp, p, 8 * e->ntaskp);
  return e;
}
void d3o_destroy(Env* e) { free(e); }

/* ------------------------------------------------------------------ kinematics + spatial quantities
 * Spatial vectors are [angular(3); linear-at-world-origin(3)] in world axes. */
static void kinematics(Env* e) {
  for (int i = 0; i < e->nlink; i++) {
    Link* L = &e->link[i];
    double* p = e->xpos[i]; double* R = e->xmat[i];
    if (L->jtype == 2) {
      double* q = e->qpos + L->qadr;
      double n = sqrt(q[3] * q[3] + q[4] * q[4] + q[5] * q[5] + q[6] * q[6]);
      for (int k = 3; k < 7; k++) q[k] /= n;                       /* mj_kinematics normalises free-joint quats in qpos */
      memcpy(p, q, 24); quat2mat(R, q + 3);
      for (int k = 0; k < 3; k++) {
        double* Sl = e->S[L->dadr + k]; double* Sa = e->S[L->dadr + 3 + k];
        memset(Sl, 0, 48); Sl[3 + k] = 1;
        double a[3] = {R[k], R[3 + k], R[6 + k]};
        memcpy(Sa, a, 24); cross3(Sa + 3, p, a);
      }
      continue;
    }
    double pp[3] = {0, 0, 0}, pR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (L->parent >= 0) { memcpy(pp, e->xpos[L->parent], 24); memcpy(pR, e->xmat[L->parent], 72); }
    double off[3]; mat_vec3(off, pR, L->pos);
    for (int k = 0; k < 3; k++) p[k] = pp[k] + off[k];
    double R0[9], Rq[9]; quat2mat(Rq, L->quat); mat_mul3(R0, pR, Rq);
    double q = e->qpos[L->qadr];
    double* S = e->S[L->dadr];
    if (L->jtype == 0) {
      double hq[4] = {cos(q / 2), sin(q / 2) * L->axis[0], sin(q / 2) * L->axis[1], sin(q / 2) * L->axis[2]};
      double Rj[9]; quat2mat(Rj, hq); mat_mul3(R, R0, Rj);
      double a[3]; mat_vec3(a, R, L->axis);
      memcpy(S, a, 24); cross3(S + 3, p, a);
    } else {
      memcpy(R, R0, 72);
      double a[3]; mat_vec3(a, R, L->axis);
      for (int k = 0; k < 3; k++) p[k] += a[k] * q;
      S[0] = S[1] = S[2] = 0; memcpy(S + 3, a, 24);
    }
  }
}

/* world-frame spatial inertia (about the world origin) of link i */
static void link_spatial_inertia(const Env* e, int i, double* I6) {
  const Link* L = &e->link[i];
  const double* R = e->xmat[i];
  double c[3]; mat_vec3(c, R, L->ipos); for (int k = 0; k < 3; k++) c[k] += e->xpos[i][k];
  double RI[9], Iw[9], Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
  mat_mul3(RI, R, L->inertia); mat_mul3(Iw, RI, Rt);
  double m = L->mass;
  double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0}, cxcx[9];
  mat_mul3(cxcx, cx, cx);
  memset(I6, 0, 36 * 8);
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
    I6[a * 6 + b] = Iw[3 * a + b] - m * cxcx[3 * a + b];
    I6[a * 6 + 3 + b] = m * cx[3 * a + b];
    I6[(3 + a) * 6 + b] = -m * cx[3 * a + b];
  }
  for (int a = 0; a < 3; a++) I6[(3 + a) * 6 + 3 + a] = m;
}
static void mv6(double* o, const double* A, const double* v) {
  for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += A[6 * i + j] * v[j]; o[i] = s; }
}
static double dot6(const double* a, const double* b) { double s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s; }

/* composite rigid body mass matrix */
static void crba(Env* e) {
  static double Ic[MAXLINK][36];
  int n = e->nlink, nv = e->nv;
  for (int i = 0; i < n; i++) link_spatial_inertia(e, i, Ic[i]);
  for (int i = n - 1; i >= 0; i--) if (e->link[i].parent >= 0) for (int k = 0; k < 36; k++) Ic[e->link[i].parent][k] += Ic[i][k];
  memset(e->M, 0, sizeof(double) * nv * nv);
  for (int a = 0; a < nv; a++) {
    int la = e->dof_link[a];
    for (int b = 0; b <= a; b++) {
      int lb = e->dof_link[b], deep;
      if (e->link[la].anc >> lb & 1) deep = la; else if (e->link[lb].anc >> la & 1) deep = lb; else continue;
      double F[6]; mv6(F, Ic[deep], e->S[b]);
      double v = dot6(e->S[a], F);
      e->M[a * nv + b] = e->M[b * nv + a] = v;
    }
  }
}

/* recursive Newton-Euler with qacc = 0: gravity + Coriolis/centrifugal  -> bias (qfrc_bias) */
static void rne_bias(Env* e) {
  static double v[MAXLINK][6], a[MAXLINK][6], f[MAXLINK][6];
  int n = e->nlink;
  const double* g = e->ctrl + C_GRAVITY;
  for (int i = 0; i < n; i++) {
    const Link* L = &e->link[i];
    double vp[6] = {0, 0, 0, 0, 0, 0}, ap[6] = {0, 0, 0, -g[0], -g[1], -g[2]};
    if (L->parent >= 0) { memcpy(vp, v[L->parent], 48); memcpy(ap, a[L->parent], 48); }
    double vj[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < L->ndof; k++) for (int c = 0; c < 6; c++) vj[c] += e->S[L->dadr + k][c] * e->qvel[L->dadr + k];
    for (int c = 0; c < 6; c++) v[i][c] = vp[c] + vj[c];
    double cj[6] = {0, 0, 0, 0, 0, 0};
    if (L->jtype == 2) {
      /* d/dt of the free-joint columns times qvel: [0; v_p x omega] (linear dofs are world-fixed) */
      const double* vl = e->qvel + L->dadr;
      cross3(cj + 3, vl, vj);                                   /* vj[0..2] = omega (world) */
    } else {
      /* v x_m (S qd) */
      double t1[3], t2[3];
      cross3(cj, v[i], vj);
      cross3(t1, v[i], vj + 3); cross3(t2, v[i] + 3, vj);
      for (int c = 0; c < 3; c++) cj[3 + c] = t1[c] + t2[c];
    }
    for (int c = 0; c < 6; c++) a[i][c] = ap[c] + cj[c];
    double I6[36], Ia[6], Iv[6];
    link_spatial_inertia(e, i, I6);
    mv6(Ia, I6, a[i]); mv6(Iv, I6, v[i]);
    /* v x* (I v) = [w x n + vO x f ; w x f] */
    double t1[3], t2[3], t3[3];
    cross3(t1, v[i], Iv); cross3(t2, v[i] + 3, Iv + 3); cross3(t3, v[i], Iv + 3);
    for (int c = 0; c < 3; c++) { f[i][c] = Ia[c] + t1[c] + t2[c]; f[i][3 + c] = Ia[3 + c] + t3[c]; }
  }
  for (int i = n - 1; i >= 0; i--) {
    const Link* L = &e->link[i];
    for (int k = 0; k < L->ndof; k++) e->bias[L->dadr + k] = dot6(e->S[L->dadr + k], f[i]);
    if (L->parent >= 0) for (int c = 0; c < 6; c++) f[L->parent][c] += f[i][c];
  }
}

/* 3 x nv translational Jacobian of a world point attached to link li (li<0: zero) */
static void jac_point(const Env* e, int li, const double* pt, double* Jp /*3*nv*/, double* Jr /*3*nv or NULL*/) {
  int nv = e->nv;
  memset(Jp, 0, sizeof(double) * 3 * nv);
  if (Jr) memset(Jr, 0, sizeof(double) * 3 * nv);
  if (li < 0) return;
  for (int d = 0; d < nv; d++) {
    if (!(e->link[li].anc >> e->dof_link[d] & 1)) continue;
    const double* S = e->S[d];
    double wxr[3]; cross3(wxr, S, pt);
    for (int k = 0; k < 3; k++) { Jp[k * nv + d] = S[3 + k] + wxr[k]; if (Jr) Jr[k * nv + d] = S[k]; }
  }
}

/* ------------------------------------------------------------------ collision (own deterministic narrow phase) */
typedef struct { double pos[3], n[3], dist; } RawContact;

static void geom_pose(const Env* e, const Geom* g, double* p, double* R) {
  if (g->link < 0) { memcpy(p, g->pos, 24); memcpy(R, g->R, 72); return; }
  double o[3]; mat_vec3(o, e->xmat[g->link], g->pos);
  for (int k = 0; k < 3; k++) p[k] = e->xpos[g->link][k] + o[k];
  mat_mul3(R, e->xmat[g->link], g->R);
}

/* Sutherland-Hodgman clip of polygon (u,v,w) against |u|<=hu, |v|<=hv; w carried along linearly */
static int clip_poly(double (*P)[3], int n, double hu, double hv) {
  double Q[16][3];
  for (int plane = 0; plane < 4; plane++) {
    int ax = plane >> 1; double sg = (plane & 1) ? -1.0 : 1.0, h = ax ? hv : hu;
    int m = 0;
    for (int i = 0; i < n; i++) {
      double* a = P[i]; double* b = P[(i + 1) % n];
      double da = h - sg * a[ax], db = h - sg * b[ax];
      if (da >= 0) { memcpy(Q[m++], a, 24); }
      if ((da >= 0) != (db >= 0)) { double t = da / (da - db); for (int k = 0; k < 3; k++) Q[m][k] = a[k] + t * (b[k] - a[k]); m++; }
    }
    n = m; for (int i = 0; i < n; i++) memcpy(P[i], Q[i], 24);
    if (n == 0) return 0;
  }
  return n;
}

/* box-box: SAT over 15 axes, face clipping or edge-edge. normal points from A to B. returns #contacts (<=8) */
static int collide_box_box(const double* pA, const double* RA, const double* hA, const double* pB, const double* RB, const double* hB,
                           double margin, RawContact* out) {
  double Rr[9], AbsR[9], t[3], d[3];
  for (int k = 0; k < 3; k++) d[k] = pB[k] - pA[k];
  matT_vec3(t, RA, d);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double s = 0; for (int k = 0; k < 3; k++) s += RA[3 * k + i] * RB[3 * k + j];
    Rr[3 * i + j] = s; AbsR[3 * i + j] = fabs(s);
  }
  double best = -1e300; int code = -1; double bsign = 1;
  /* A faces */
  for (int i = 0; i < 3; i++) {
    double sep = fabs(t[i]) - (hA[i] + hB[0] * AbsR[3 * i] + hB[1] * AbsR[3 * i + 1] + hB[2] * AbsR[3 * i + 2]);
    if (sep > margin) return 0;
    if (sep > best) { best = sep; code = i; bsign = t[i] >= 0 ? 1 : -1; }
  }
  /* B faces: must beat the best A face by 1 um (tie rule, keeps table-vs-box on the table's face) */
  for (int j = 0; j < 3; j++) {
    double tb = t[0] * Rr[j] + t[1] * Rr[3 + j] + t[2] * Rr[6 + j];
    double sep = fabs(tb) - (hB[j] + hA[0] * AbsR[j] + hA[1] * AbsR[3 + j] + hA[2] * AbsR[6 + j]);
    if (sep > margin) return 0;
    if (sep > best + 1e-6) { best = sep; code = 3 + j; bsign = tb >= 0 ? 1 : -1; }
  }
  /* edge axes A_i x B_j: must be 5% better than the best face (ODE-style preference for faces) */
  double ebest = -1e300; int ecode = -1; double en[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double ai[3] = {RA[i], RA[3 + i], RA[6 + i]}, bj[3] = {RB[j], RB[3 + j], RB[6 + j]}, ax[3];
    cross3(ax, ai, bj);
    double l = norm3(ax);
    if (l < 1e-6) continue;
    for (int k = 0; k < 3; k++) ax[k] /= l;
    double dist = dot3(ax, d), ra = 0, rb = 0, mx = 0;
    for (int k = 0; k < 3; k++) {
      double ak[3] = {RA[k], RA[3 + k], RA[6 + k]}, bk[3] = {RB[k], RB[3 + k], RB[6 + k]};
      double da = fabs(dot3(ax, ak)), db = fabs(dot3(ax, bk));
      ra += hA[k] * da; rb += hB[k] * db;
      if (da > mx) mx = da;
      if (db > mx) mx = db;
    }
    double sep = fabs(dist) - (ra + rb);
    if (sep > margin) return 0;
    /* an edge axis within 2 degrees of a face normal (nearly parallel boxes) is that face contact in disguise: between
       two such axes the SAT would be decided by rounding noise, with 1 contact on one side and 4 on the other */
    if (mx > 0.9994) continue;
    if (sep > ebest) { ebest = sep; ecode = 3 * i + j; double sg = dist >= 0 ? 1 : -1; for (int k = 0; k < 3; k++) en[k] = sg * ax[k]; }
  }
  /* an edge-edge axis must beat the best face axis by 5 % + 1 um (sign-symmetric: margin contacts have positive separations) */
  if (ecode >= 0 && ebest > best + 0.05 * fabs(best) + 1e-6) {
    /* edge-edge: one contact at the closest points of the two supporting edges */
    int i = ecode / 3, j = ecode % 3;
    double ca[3], cb[3];
    memcpy(ca, pA, 24); memcpy(cb, pB, 24);
    for (int k = 0; k < 3; k++) {
      if (k != i) { double ak[3] = {RA[k], RA[3 + k], RA[6 + k]}; double s = dot3(en, ak) >= 0 ? 1 : -1; for (int c = 0; c < 3; c++) ca[c] += s * hA[k] * ak[c]; }
      if (k != j) { double bk[3] = {RB[k], RB[3 + k], RB[6 + k]}; double s = dot3(en, bk) >= 0 ? -1 : 1; for (int c = 0; c < 3; c++) cb[c] += s * hB[k] * bk[c]; }
    }
    double ua[3] = {RA[i], RA[3 + i], RA[6 + i]}, ub[3] = {RB[j], RB[3 + j], RB[6 + j]}, w[3];
    for (int k = 0; k < 3; k++) w[k] = ca[k] - cb[k];
    double b = dot3(ua, ub), dd = dot3(ua, w), ee = dot3(ub, w), den = 1 - b * b;
    double sa = den > 1e-12 ? (b * ee - dd) / den : 0, sb = den > 1e-12 ? (ee - b * dd) / den : 0;
    sa = clampd(sa, -hA[i], hA[i]); sb = clampd(sb, -hB[j], hB[j]);
    for (int k = 0; k < 3; k++) out[0].pos[k] = 0.5 * (ca[k] + sa * ua[k] + cb[k] + sb * ub[k]);
    memcpy(out[0].n, en, 24); out[0].dist = ebest;
    return 1;
  }
  /* face contact: reference box = owner of the axis, incident = the other */
  int refIsA = code < 3, ax = refIsA ? code : code - 3;
  const double *pR_ = refIsA ? pA : pB, *RR = refIsA ? RA : RB, *hR = refIsA ? hA : hB;
  const double *pI = refIsA ? pB : pA, *RI = refIsA ? RB : RA, *hI = refIsA ? hB : hA;
  double nref[3];                                   /* outward normal of the reference face, pointing to the incident box */
  double sgn = refIsA ? bsign : -bsign;
  for (int k = 0; k < 3; k++) nref[k] = sgn * RR[3 * k + ax];
  /* incident face: most anti-parallel to nref */
  int jx = 0; double jb = -1;
  for (int j = 0; j < 3; j++) { double c[3] = {RI[j], RI[3 + j], RI[6 + j]}; double v = fabs(dot3(nref, c)); if (v > jb) { jb = v; jx = j; } }
  double cI[3] = {RI[jx], RI[3 + jx], RI[6 + jx]};
  double isg = dot3(nref, cI) > 0 ? -1 : 1;
  int k1 = (jx + 1) % 3, k2 = (jx + 2) % 3;
  if (k1 > k2) { int tmp = k1; k1 = k2; k2 = tmp; }
  int u = (ax + 1) % 3, v = (ax + 2) % 3;
  if (u > v) { int tmp = u; u = v; v = tmp; }
  static const double s1[4] = {-1, 1, 1, -1}, s2[4] = {-1, -1, 1, 1};
  double P[16][3];
  for (int c = 0; c < 4; c++) {
    double w[3], wl[3];
    for (int k = 0; k < 3; k++) w[k] = pI[k] + isg * hI[jx] * RI[3 * k + jx] + s1[c] * hI[k1] * RI[3 * k + k1] + s2[c] * hI[k2] * RI[3 * k + k2] - pR_[k];
    matT_vec3(wl, RR, w);
    P[c][0] = wl[u]; P[c][1] = wl[v]; P[c][2] = sgn * wl[ax] - hR[ax];   /* signed distance above the reference face */
  }
  int n = clip_poly(P, 4, hR[u], hR[v]);
  int m = 0;
  for (int c = 0; c < n && m < 8; c++) {
    double dist = P[c][2];
    if (dist >= margin) continue;
    double wl[3]; wl[u] = P[c][0]; wl[v] = P[c][1]; wl[ax] = sgn * (hR[ax] + 0.5 * dist);   /* midpoint between vertex and face */
    double w[3]; mat_vec3(w, RR, wl);
    for (int k = 0; k < 3; k++) { out[m].pos[k] = pR_[k] + w[k]; out[m].n[k] = refIsA ? nref[k] : -nref[k]; }
    out[m].dist = dist; m++;
  }
  return m;
}

/* closest point of the zonotope {d + sum l_k g_k, |l_k|<=1} (2-D) to the origin; returns 1 if the origin is outside */
static int zonotope_closest(const double d[2], const double g[3][2], double q[2], double* pen_inside, double n_inside[2]) {
  int inside = 1; double bestpen = 1e300;
  for (int k = 0; k < 3; k++) {
    double l = sqrt(g[k][0] * g[k][0] + g[k][1] * g[k][1]);
    if (l < 1e-12) continue;
    double nk[2] = {-g[k][1] / l, g[k][0] / l};
    double w = 0; for (int j = 0; j < 3; j++) if (j != k) w += fabs(nk[0] * g[j][0] + nk[1] * g[j][1]);
    double s = d[0] * nk[0] + d[1] * nk[1];
    double sep = fabs(s) - w;
    if (sep > 0) inside = 0;
    if (-sep < bestpen) { bestpen = -sep; double sg = s >= 0 ? 1 : -1; n_inside[0] = sg * nk[0]; n_inside[1] = sg * nk[1]; }
  }
  *pen_inside = bestpen;
  if (inside) return 0;
  double bd = 1e300;
  for (int k = 0; k < 3; k++) {
    double l2 = g[k][0] * g[k][0] + g[k][1] * g[k][1];
    double nk[2] = {-g[k][1], g[k][0]};
    for (int sgi = 0; sgi < 2; sgi++) {
      double sg = sgi ? -1 : 1, p0[2] = {d[0], d[1]};
      for (int j = 0; j < 3; j++) if (j != k) { double c = nk[0] * g[j][0] + nk[1] * g[j][1]; double s = (c >= 0 ? 1 : -1) * sg; p0[0] += s * g[j][0]; p0[1] += s * g[j][1]; }
      double lam = l2 > 1e-24 ? clampd(-(p0[0] * g[k][0] + p0[1] * g[k][1]) / l2, -1, 1) : 0;
      double c[2] = {p0[0] + lam * g[k][0], p0[1] + lam * g[k][1]};
      double dd = c[0] * c[0] + c[1] * c[1];
      if (dd < bd) { bd = dd; q[0] = c[0]; q[1] = c[1]; }
    }
  }
  return 1;
}

/* cylinder (geom1) vs box (geom2): candidate-axis SAT on support functions + feature-based contact point. <=1 contact */
static int collide_cyl_box(const double* c, const double* Rc, const double* sz, const double* b, const double* Rb, const double* e3,
                           double margin, RawContact* out) {
  double r = sz[0], h = sz[1];
  double a[3] = {Rc[2], Rc[5], Rc[8]}, B[3][3], d[3];
  for (int k = 0; k < 3; k++) { B[k][0] = Rb[k]; B[k][1] = Rb[3 + k]; B[k][2] = Rb[6 + k]; d[k] = b[k] - c[k]; }
  double best = 1e300, n[3] = {0, 0, 0};
#define TRY_AXIS(nx)                                                                                     \
  do {                                                                                                   \
    double na = dot3(nx, a), perp = 1 - na * na;                                                         \
    double hc = h * fabs(na) + r * sqrt(perp > 0 ? perp : 0);                                            \
    double hb = e3[0] * fabs(dot3(nx, B[0])) + e3[1] * fabs(dot3(nx, B[1])) + e3[2] * fabs(dot3(nx, B[2])); \
    double ov = hc + hb - dot3(nx, d);                                                                   \
    if (ov < -margin) return 0;                                                                          \
    if (ov < best - 1e-9) { best = ov; memcpy(n, nx, 24); }                                              \
  } while (0)
  for (int k = 0; k < 3; k++) { double s = dot3(B[k], d) >= 0 ? 1 : -1; double nx[3] = {s * B[k][0], s * B[k][1], s * B[k][2]}; TRY_AXIS(nx); }
  { double s = dot3(a, d) >= 0 ? 1 : -1; double nx[3] = {s * a[0], s * a[1], s * a[2]}; TRY_AXIS(nx); }
  /* side surface: disc vs projected box in the plane perpendicular to the axis */
  {
    double u[3] = {Rc[0], Rc[3], Rc[6]}, w[3] = {Rc[1], Rc[4], Rc[7]};
    double d2[2] = {dot3(d, u), dot3(d, w)}, g[3][2], q[2] = {0, 0}, pin, nin[2] = {0, 0};
    for (int k = 0; k < 3; k++) { g[k][0] = e3[k] * dot3(B[k], u); g[k][1] = e3[k] * dot3(B[k], w); }
    if (zonotope_closest(d2, g, q, &pin, nin)) {
      double l = sqrt(q[0] * q[0] + q[1] * q[1]);
      if (l > 1e-12) { double nx[3]; for (int k = 0; k < 3; k++) nx[k] = (q[0] * u[k] + q[1] * w[k]) / l; TRY_AXIS(nx); }
    } else {
      double nx[3]; for (int k = 0; k < 3; k++) nx[k] = nin[0] * u[k] + nin[1] * w[k]; TRY_AXIS(nx);
    }
  }
#undef TRY_AXIS
  /* Contact point, continuous in the relative pose (no feature thresholds).  Support point of the cylinder along n:
   * cap rim at t = s h, radial direction u.  The contact patch extends from there along the side line
   * x(t) = c + r u + t a (penetration best - |n.a| (h - s t), clipped to the box) and across the cap (penetration
   * best - r |u-part| at the cap centre): take the penetration-weighted centroid along the line and move from the rim
   * towards the cap centre as the cap flattens against the box.  Flush side contact -> midpoint of the overlap;
   * tilted -> slides continuously towards the deep end; flat cap -> cap centre. */
  double na = dot3(n, a), s = na >= 0 ? 1 : -1, pr[3], u[3] = {0, 0, 0}, pc[3];
  for (int k = 0; k < 3; k++) pr[k] = n[k] - na * a[k];
  double l = norm3(pr);
  if (l > 1e-12) for (int k = 0; k < 3; k++) u[k] = pr[k] / l;
  double t0 = -h, t1 = h, base[3]; int empty = 0;
  for (int k = 0; k < 3; k++) base[k] = c[k] + r * u[k] - b[k];
  for (int k = 0; k < 3; k++) {
    double x0 = dot3(base, B[k]), dx = dot3(a, B[k]);
    if (fabs(dx) < 1e-12) { if (fabs(x0) > e3[k]) empty = 1; continue; }
    double ta = (-e3[k] - x0) / dx, tb = (e3[k] - x0) / dx;
    if (ta > tb) { double tmp = ta; ta = tb; tb = tmp; }
    if (ta > t0) t0 = ta;
    if (tb < t1) t1 = tb;
  }
  if (t0 > t1) empty = 1;
  /* default (the side line misses the box: edge / corner contacts): midpoint of the axial overlap of cylinder and box */
  double ts;
  {
    double ca = dot3(d, a), ha = e3[0] * fabs(dot3(a, B[0])) + e3[1] * fabs(dot3(a, B[1])) + e3[2] * fabs(dot3(a, B[2]));
    double lo = fmax(-h, ca - ha), hi = fmin(h, ca + ha);
    ts = lo <= hi ? 0.5 * (lo + hi) : clampd(ca, -h, h);
  }
  if (!empty) {
    double da = best - fabs(na) * (h - s * t0), db = best - fabs(na) * (h - s * t1);
    if (da <= 0 && db <= 0) ts = db > da ? t1 : t0;
    else {
      if (da < 0) { t0 += (t1 - t0) * (-da) / (db - da); da = 0; }
      if (db < 0) { t1 -= (t1 - t0) * (-db) / (da - db); db = 0; }
      ts = t0 + (t1 - t0) * (da + 2 * db) / (3 * (da + db));
    }
  }
  ts = s * h + (ts - s * h) * l * l;        /* the shift along the axis only means something for side contacts (l -> 1); a flat cap (l -> 0) keeps its plane */
  double dc = best - r * l, wcap = 1;
  if (dc > 0) { double w = r * l / (4 * dc); if (w < 1) wcap = w; }
  for (int k = 0; k < 3; k++) pc[k] = c[k] + ts * a[k] + r * wcap * u[k];
  for (int k = 0; k < 3; k++) { out->pos[k] = pc[k] - 0.5 * best * n[k]; out->n[k] = n[k]; }
  out->dist = -best;
  return out->dist < margin;
}

/* cylinder vs cylinder: side-side via closest points of the axis segments (caps treated as in DESIGN.md). <=1 contact */
static int collide_cyl_cyl(const double* c1, const double* R1, const double* s1, const double* c2, const double* R2, const double* s2,
                           double margin, RawContact* out) {
  double a1[3] = {R1[2], R1[5], R1[8]}, a2[3] = {R2[2], R2[5], R2[8]}, w[3];
  for (int k = 0; k < 3; k++) w[k] = c1[k] - c2[k];
  double b = dot3(a1, a2), d = dot3(a1, w), e = dot3(a2, w), den = 1 - b * b, t1, t2;
  if (den > 1e-8) {
    t1 = clampd((b * e - d) / den, -s1[1], s1[1]);
    t2 = clampd(e + b * t1, -s2[1], s2[1]);
    t1 = clampd(-d + b * t2, -s1[1], s1[1]);
  } else {
    /* parallel axes: midpoint of the axial overlap.  t2 = e + b*t1 in [-h2,h2]  ->  t1 in [(-h2-e)/b, (h2-e)/b] */
    double l2 = (-s2[1] - e) / b, h2 = (s2[1] - e) / b;
    if (l2 > h2) { double tmp = l2; l2 = h2; h2 = tmp; }
    double lo = fmax(-s1[1], l2), hi = fmin(s1[1], h2);
    t1 = lo > hi ? clampd(0.5 * (lo + hi), -s1[1], s1[1]) : 0.5 * (lo + hi);
    t2 = clampd(e + b * t1, -s2[1], s2[1]);
  }
  double p1[3], p2[3], dv[3];
  for (int k = 0; k < 3; k++) { p1[k] = c1[k] + t1 * a1[k]; p2[k] = c2[k] + t2 * a2[k]; dv[k] = p2[k] - p1[k]; }
  double l = norm3(dv), dist = l - s1[0] - s2[0];
  if (dist >= margin || l < 1e-12) return 0;
  for (int k = 0; k < 3; k++) { out->n[k] = dv[k] / l; out->pos[k] = p1[k] + out->n[k] * (s1[0] + 0.5 * dist); }
  out->dist = dist;
  return 1;
}

static void make_frame(double* f) {            /* mju_makeFrame [EXT]: x given, y from (0,1,0) or (0,0,1) */
  double y[3] = {0, 0, 0};
  if (f[1] < 0.5 && f[1] > -0.5) y[1] = 1; else y[2] = 1;
  double t = dot3(f, y);
  for (int k = 0; k < 3; k++) y[k] -= t * f[k];
  double l = norm3(y);
  for (int k = 0; k < 3; k++) f[3 + k] = y[k] / l;
  cross3(f + 6, f, f + 3);
}

static void collision(Env* e) {
  e->ncon = 0; e->obst_contact = 0;
  for (int ip = 0; ip < e->npair; ip++) {
    const Pair* pr = &e->pair[ip];
    const Geom *g1 = &e->geom[pr->g1], *g2 = &e->geom[pr->g2];
    double p1[3], R1[9], p2[3], R2[9], dc[3];
    geom_pose(e, g1, p1, R1); geom_pose(e, g2, p2, R2);
    for (int k = 0; k < 3; k++) dc[k] = p2[k] - p1[k];
    if (norm3(dc) > g1->rbound + g2->rbound + pr->margin) continue;     /* bounding spheres */
    RawContact rc[8]; int n = 0;
    if (g1->type == G_BOX && g2->type == G_BOX) n = collide_box_box(p1, R1, g1->size, p2, R2, g2->size, pr->margin, rc);
    else if (g1->type == G_CYLINDER && g2->type == G_BOX) n = collide_cyl_box(p1, R1, g1->size, p2, R2, g2->size, pr->margin, rc);
    else if (g1->type == G_CYLINDER && g2->type == G_CYLINDER) n = collide_cyl_cyl(p1, R1, g1->size, p2, R2, g2->size, pr->margin, rc);
    for (int i = 0; i < n; i++) {
      if (e->ncon >= MAXCON) { e->status |= 2; break; }
      Contact* c = &e->con[e->ncon++];
      memcpy(c->pos, rc[i].pos, 24); memcpy(c->frame, rc[i].n, 24); make_frame(c->frame);
      c->dist = rc[i].dist; c->includemargin = pr->margin - pr->gap; c->dim = pr->condim;
      memcpy(c->friction, pr->friction, 40); memcpy(c->solref, pr->solref, 16); memcpy(c->solimp, pr->solimp, 40);
      c->g1 = pr->g1; c->g2 = pr->g2; c->pair = ip; c->efc = -1;
      if (pr->flags & 1) e->obst_contact = 1;
    }
  }
}

/* ------------------------------------------------------------------ constraints (SURVEY App. B.6 [EXT]) */
static void impedance(const double* solimp, double pos, double margin, double* imp) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (width < 1e-15 || dmin == dmax) { *imp = 0.5 * (dmin + dmax); return; }
  double x = fabs(pos - margin) / width;
  if (x >= 1) { *imp = dmax; return; }
  if (x <= 0) { *imp = dmin; return; }
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  *imp = dmin + y * (dmax - dmin);
}
static void kbi(const double* solref, const double* solimp, double pos, double margin, double* k, double* b, double* imp) {
  double dmax = solimp[1], tc = solref[0], dr = solref[1];
  *k = 1 / (dmax * dmax * tc * tc * dr * dr);
  *b = 2 / (dmax * tc);
  impedance(solimp, pos, margin, imp);
}

static void make_constraints(Env* e) {
  int nv = e->nv, ne = 0;
  double impratio = e->ctrl[C_IMPRATIO];
  /* joint limits (robot dofs are the only limited joints) */
  for (int i = 0; i < e->nlink; i++) {
    const Link* L = &e->link[i];
    if (!L->limited || L->jtype == 2) continue;
    double q = e->qpos[L->qadr];
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? q - L->range[0] : L->range[1] - q;
      if (dist >= 0) continue;
      double k, b, imp; kbi(e->ctrl + C_JNT_SOLREF, e->ctrl + C_JNT_SOLIMP, dist, 0, &k, &b, &imp);
      double sg = side == 0 ? 1 : -1;
      memset(e->J + ne * nv, 0, 8 * nv);
      e->J[ne * nv + L->dadr] = sg;
      e->Rr[ne] = fmax(1e-15, (1 - imp) / imp * L->invw[0]);
      e->D[ne] = 1 / e->Rr[ne];
      e->aref[ne] = -b * sg * e->qvel[L->dadr] - k * imp * dist;
      e->efc_type[ne] = 0; e->efc_con[ne] = -1; ne++;
    }
  }
  /* elliptic frictional contacts */
  static double Jp1[3 * MAXV], Jr1[3 * MAXV], Jp2[3 * MAXV], Jr2[3 * MAXV];
  for (int ic = 0; ic < e->ncon; ic++) {
    Contact* c = &e->con[ic];
    if (!(c->dist < c->includemargin)) continue;
    if (ne + c->dim > MAXEFC) { e->status |= 2; break; }
    const Geom *g1 = &e->geom[c->g1], *g2 = &e->geom[c->g2];
    jac_point(e, g1->link, c->pos, Jp1, Jr1); jac_point(e, g2->link, c->pos, Jp2, Jr2);
    c->efc = ne;
    for (int r = 0; r < c->dim; r++) {
      const double* ax = c->frame + 3 * (r < 3 ? r : 0);      /* row 3 (torsion) uses the normal on the rotational Jacobian */
      for (int d = 0; d < nv; d++) {
        double s = 0;
        if (r < 3) for (int k = 0; k < 3; k++) s += ax[k] * (Jp2[k * nv + d] - Jp1[k * nv + d]);
        else for (int k = 0; k < 3; k++) s += ax[k] * (Jr2[k * nv + d] - Jr1[k * nv + d]);
        e->J[(ne + r) * nv + d] = s;
      }
    }
    double k, b, imp; kbi(c->solref, c->solimp, c->dist, c->includemargin, &k, &b, &imp);
    double tran = g1->invw[0] + g2->invw[0];
    double R0 = fmax(1e-15, (1 - imp) / imp * tran);
    double R1 = R0 / impratio;
    c->mu = c->friction[0] * sqrt(R1 / R0);
    for (int r = 0; r < c->dim; r++) {
      double Rv = r == 0 ? R0 : (r == 1 ? R1 : R1 * c->friction[0] * c->friction[0] / (c->friction[r - 1] * c->friction[r - 1]));
      double vel = 0; for (int d = 0; d < nv; d++) vel += e->J[(ne + r) * nv + d] * e->qvel[d];
      e->Rr[ne + r] = Rv; e->D[ne + r] = 1 / Rv;
      e->aref[ne + r] = -b * vel - (r == 0 ? k * imp * (c->dist - c->includemargin) : 0);
      e->efc_type[ne + r] = r == 0 ? 1 : 2; e->efc_con[ne + r] = ic;
    }
    ne += c->dim;
  }
  e->nefc = ne;
}

/* cost, force and (optionally) Hessian blocks of the constraint term at jar. Hc: per-row diag in hd[], cone blocks in hb[con][36] */
static double constraint_eval(const Env* e, const double* jar, double* force, double* hd, double (*hb)[36], int* cone_mid) {
  double cost = 0;
  for (int i = 0; i < e->nefc; i++) {
    if (e->efc_type[i] == 0) {
      if (jar[i] < 0) { cost += 0.5 * e->D[i] * jar[i] * jar[i]; force[i] = -e->D[i] * jar[i]; if (hd) hd[i] = e->D[i]; }
      else { force[i] = 0; if (hd) hd[i] = 0; }
    } else if (e->efc_type[i] == 1) {
      const Contact* c = &e->con[e->efc_con[i]];
      int dim = c->dim; double mu = c->mu, U[6], T = 0;
      U[0] = jar[i] * mu;
      for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * c->friction[j - 1]; T += U[j] * U[j]; }
      T = sqrt(T);
      double N = U[0];
      if (cone_mid) cone_mid[e->efc_con[i]] = 0;
      if (N >= mu * T || (T <= 0 && N >= 0)) {                      /* top zone: satisfied */
        for (int j = 0; j < dim; j++) { force[i + j] = 0; if (hd) hd[i + j] = 0; }
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {           /* bottom zone: fully quadratic */
        for (int j = 0; j < dim; j++) { cost += 0.5 * e->D[i + j] * jar[i + j] * jar[i + j]; force[i + j] = -e->D[i + j] * jar[i + j]; if (hd) hd[i + j] = e->D[i + j]; }
      } else {                                                     /* middle zone: on the cone surface */
        double Dm = e->D[i] / fmax(1e-15, mu * mu * (1 + mu * mu)), NmT = N - mu * T;
        cost += 0.5 * Dm * NmT * NmT;
        force[i] = -Dm * NmT * mu;
        for (int j = 1; j < dim; j++) force[i + j] = -force[i] / T * U[j] * c->friction[j - 1];
        if (hb) {
          /* H = Dm [ g g^T - mu*NmT * F (I - uu^T)/T F ],  g = (mu, -mu f_j U_j/T) */
          double g[6]; g[0] = mu; for (int j = 1; j < dim; j++) g[j] = -mu * c->friction[j - 1] * U[j] / T;
          double* H = hb[e->efc_con[i]];
          for (int a = 0; a < dim; a++) for (int b2 = 0; b2 < dim; b2++) {
            double v = g[a] * g[b2];
            if (a > 0 && b2 > 0) v -= mu * NmT / T * c->friction[a - 1] * c->friction[b2 - 1] * ((a == b2) - U[a] * U[b2] / (T * T));
            H[a * 6 + b2] = Dm * v;
          }
          cone_mid[e->efc_con[i]] = 1;
        }
      }
    }
  }
  return cost;
}

/* Newton solver on the primal problem (SURVEY App. B.7 [EXT]); result in e->qacc, e->efc_force */
static void solve_constraints(Env* e) {
  int nv = e->nv, ne = e->nefc;
  if (ne == 0) { memcpy(e->qacc, e->qacc_smooth, 8 * nv); memset(e->qfrc_constraint, 0, 8 * nv); e->solver_iter = 0; return; }
  static double jar[MAXEFC], Jp[MAXEFC], hd[MAXEFC], hb[MAXCON][36], H[MAXV * MAXV], grad[MAXV], p[MAXV], Ma[MAXV], tmp[MAXV], ftmp[MAXEFC];
  static int mid[MAXCON];
  double* a = e->qacc;
  double scale = 1.0 / (e->ctrl[C_MEANINERTIA] * (nv > 1 ? nv : 1)), tol = e->ctrl[C_TOL];
  /* warm start: cheaper of qacc_warmstart and qacc_smooth */
  double cost_w, cost_s;
  {
    for (int i = 0; i < ne; i++) { double s = -e->aref[i]; for (int d = 0; d < nv; d++) s += e->J[i * nv + d] * e->warm[d]; jar[i] = s; }
    cost_w = constraint_eval(e, jar, ftmp, NULL, NULL, NULL);
    for (int d = 0; d < nv; d++) tmp[d] = e->warm[d] - e->qacc_smooth[d];
    for (int d = 0; d < nv; d++) { double s = 0; for (int k = 0; k < nv; k++) s += e->M[d * nv + k] * tmp[k]; cost_w += 0.5 * tmp[d] * s; }
    for (int i = 0; i < ne; i++) { double s = -e->aref[i]; for (int d = 0; d < nv; d++) s += e->J[i * nv + d] * e->qacc_smooth[d]; jar[i] = s; }
    cost_s = constraint_eval(e, jar, ftmp, NULL, NULL, NULL);
    memcpy(a, cost_w < cost_s ? e->warm : e->qacc_smooth, 8 * nv);
  }
  double cost = 0, oldcost;
  int iter;
  for (iter = 0; iter < 100; iter++) {
    for (int i = 0; i < ne; i++) { double s = -e->aref[i]; for (int d = 0; d < nv; d++) s += e->J[i * nv + d] * a[d]; jar[i] = s; }
    oldcost = cost;
    cost = constraint_eval(e, jar, e->efc_force, hd, hb, mid);
    for (int d = 0; d < nv; d++) tmp[d] = a[d] - e->qacc_smooth[d];
    for (int d = 0; d < nv; d++) { double s = 0; for (int k = 0; k < nv; k++) s += e->M[d * nv + k] * tmp[k]; Ma[d] = s; cost += 0.5 * tmp[d] * s; }
    double gn = 0;
    for (int d = 0; d < nv; d++) { double s = Ma[d]; for (int i = 0; i < ne; i++) s -= e->J[i * nv + d] * e->efc_force[i]; grad[d] = s; gn += s * s; }
    gn = sqrt(gn);
    if (iter > 0 && (scale * (oldcost - cost) < tol || scale * gn < tol)) break;
    if (iter == 0 && scale * gn < tol) break;
    /* H = M + J^T Hc J */
    memcpy(H, e->M, 8 * nv * nv);
    for (int i = 0; i < ne; i++) {
      if (e->efc_type[i] == 1 && mid[e->efc_con[i]]) {
        const Contact* c = &e->con[e->efc_con[i]];
        const double* Hb = hb[e->efc_con[i]];
        for (int r = 0; r < c->dim; r++) for (int s2 = 0; s2 < c->dim; s2++) {
          double h = Hb[r * 6 + s2]; if (h == 0) continue;
          const double *Jr = e->J + (i + r) * nv, *Js = e->J + (i + s2) * nv;
          for (int d = 0; d < nv; d++) { if (Jr[d] == 0) continue; double t = h * Jr[d]; for (int k = 0; k < nv; k++) H[d * nv + k] += t * Js[k]; }
        }
        i += c->dim - 1;
      } else if (hd[i] != 0) {
        const double* Jr = e->J + i * nv;
        for (int d = 0; d < nv; d++) { if (Jr[d] == 0) continue; double t = hd[i] * Jr[d]; for (int k = 0; k < nv; k++) H[d * nv + k] += t * Jr[k]; }
      }
    }
    if (chol_factor(H, nv, nv)) { e->status |= 4; break; }
    for (int d = 0; d < nv; d++) p[d] = -grad[d];
    chol_solve(H, nv, nv, p);
    /* exact line search on phi(alpha) = cost(a + alpha p): safeguarded 1-D Newton */
    for (int i = 0; i < ne; i++) { double s = 0; for (int d = 0; d < nv; d++) s += e->J[i * nv + d] * p[d]; Jp[i] = s; }
    double pMp = 0, pMa = 0;
    for (int d = 0; d < nv; d++) { double s = 0; for (int k = 0; k < nv; k++) s += e->M[d * nv + k] * p[k]; pMp += p[d] * s; pMa += p[d] * Ma[d]; }
    double lo = 0, hi = -1, alpha = 1, d0 = 0, dlo, dhi = 0;
    for (int d = 0; d < nv; d++) d0 += grad[d] * p[d];                 /* phi'(0) < 0 */
    dlo = d0;
    static double jal[MAXEFC], hd2[MAXEFC], hb2[MAXCON][36]; static int mid2[MAXCON];
    for (int ls = 0; ls < 50; ls++) {
      for (int i = 0; i < ne; i++) jal[i] = jar[i] + alpha * Jp[i];
      constraint_eval(e, jal, ftmp, hd2, hb2, mid2);
      double d1 = pMa + alpha * pMp, d2 = pMp;
      for (int i = 0; i < ne; i++) d1 -= Jp[i] * ftmp[i];
      for (int i = 0; i < ne; i++) {
        if (e->efc_type[i] == 1 && mid2[e->efc_con[i]]) {
          const Contact* c = &e->con[e->efc_con[i]];
          for (int r = 0; r < c->dim; r++) for (int s2 = 0; s2 < c->dim; s2++) d2 += Jp[i + r] * hb2[e->efc_con[i]][r * 6 + s2] * Jp[i + s2];
          i += c->dim - 1;
        } else d2 += hd2[i] * Jp[i] * Jp[i];
      }
      if (fabs(d1) <= 1e-10 * fabs(d0) + 1e-300) break;
      if (d1 < 0) { lo = alpha; dlo = d1; } else { hi = alpha; dhi = d1; }
      double next = alpha - d1 / d2;
      if (hi < 0) { if (!(next > lo)) next = 2 * alpha + 1e-12; }
      else {
        /* bracketed: Newton step unless it hugs an end point (it can cycle across a kink of phi'), then false position */
        double wd = hi - lo;
        if (!(next > lo + 0.05 * wd && next < hi - 0.05 * wd)) {
          double sec = lo + wd * (-dlo) / (dhi - dlo);
          next = clampd(sec, lo + 0.05 * wd, hi - 0.05 * wd);
        }
        if (wd < 1e-15 * (1 + hi)) { alpha = next; break; }
      }
      alpha = next;
    }
    for (int d = 0; d < nv; d++) a[d] += alpha * p[d];
  }
  if (iter >= 100) e->status |= 4;
  e->solver_iter = iter;
  for (int d = 0; d < nv; d++) { double s = 0; for (int i = 0; i < ne; i++) s += e->J[i * nv + d] * e->efc_force[i]; e->qfrc_constraint[d] = s; }
}

/* ------------------------------------------------------------------ controller side (reference Python restated) */
/* FK + Jacobian of `panda_grasptarget` on the URDF chain — d3il_sim/core/Model.py:37-66 (pinocchio LOCAL_WORLD_ALIGNED) */
static void ik_fk(const Env* e, const double* q, double* pos, double* quat, double* J /*6x7 or NULL*/) {
  double p[3] = {0, 0, 0}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, org[7][3], axs[7][3];
  for (int i = 0; i < 7; i++) {
    const double* o = e->ctrl + C_IK_ORIGIN + 12 * i;
    double t[3]; mat_vec3(t, R, o); for (int k = 0; k < 3; k++) p[k] += t[k];
    mat_mul3(R, R, o + 3);
    double c = cos(q[i]), s = sin(q[i]), Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
    mat_mul3(R, R, Rz);
    memcpy(org[i], p, 24); axs[i][0] = R[2]; axs[i][1] = R[5]; axs[i][2] = R[8];
  }
  const double* o = e->ctrl + C_IK_EE;
  double t[3], Re[9]; mat_vec3(t, R, o); for (int k = 0; k < 3; k++) pos[k] = p[k] + t[k];
  mat_mul3(Re, R, o + 3); mat2quat(quat, Re);
  if (J) for (int i = 0; i < 7; i++) {
    double r[3] = {pos[0] - org[i][0], pos[1] - org[i][1], pos[2] - org[i][2]}, l[3];
    cross3(l, axs[i], r);
    for (int k = 0; k < 3; k++) { J[k * 7 + i] = l[k]; J[(3 + k) * 7 + i] = axs[i][k]; }
  }
}
void d3o_ik_fk(Env* e, const double* q, double* pos, double* quat, double* J) { ik_fk(e, q, pos, quat, J); }

/* utils/geometric_transformation.py:14-46 */
static void quat_error(const double* c, const double* d, double* o) {
  o[0] = c[0] * d[1] - d[0] * c[1] - c[3] * d[2] + c[2] * d[3];
  o[1] = c[0] * d[2] - d[0] * c[2] + c[3] * d[1] - c[1] * d[3];
  o[2] = c[0] * d[3] - d[0] * c[3] - c[2] * d[1] + c[1] * d[2];
}

/* CartPosQuatImpedenceController.getControl — controllers/IKControllers.py:163-323 (+ JointPDController, Controller.py:164-185) */
static void cart_controller(Env* e, double* tau /*7*/) {
  const double* C = e->ctrl;
  if (!e->ik_valid) { memcpy(e->ik_q, e->qpos, 8 * NARM); e->ik_valid = 1; }     /* :168-169 old_q NaN -> measured q */
  double q[7], old_q[7]; memcpy(q, e->ik_q, sizeof q); memcpy(old_q, q, sizeof q);
  double des_quat[4]; memcpy(des_quat, e->des_quat, 32);
  int niter = (int)C[C_NUM_ITER];
  for (int it = 0; it < niter; it++) {
    double pos[3], cq[4], J[42];
    ik_fk(e, q, pos, cq, J);
    double dm = 0, dp = 0;
    for (int k = 0; k < 4; k++) { dm += (cq[k] - des_quat[k]) * (cq[k] - des_quat[k]); dp += (cq[k] + des_quat[k]) * (cq[k] + des_quat[k]); }
    if (sqrt(dm) > sqrt(dp)) for (int k = 0; k < 4; k++) des_quat[k] = -des_quat[k];                       /* :204-207 */
    double qe[3], acc[6]; quat_error(cq, des_quat, qe);
    for (int k = 0; k < 3; k++) {
      acc[k] = C[C_PGAIN_POS + k] * clampd(e->des_pos[k] - pos[k], -0.01, 0.01);
      acc[3 + k] = C[C_PGAIN_QUAT + k] * clampd(qe[k], -0.1, 0.1);
    }
    double A[36], w[6], V[36];
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) { double s = 0; for (int k = 0; k < 7; k++) s += J[r * 7 + k] * J[c * 7 + k]; A[r * 6 + c] = s + (r == c ? C[C_JREG] : 0); }
    jacobi_eig(A, 6, w, V);                                                                               /* svd of an SPD matrix = eig */
    double qd_null[7], rhs[6], y[6], x[6];
    for (int k = 0; k < 7; k++) qd_null[k] = C[C_PGAIN_NULL + k] * clampd(C[C_REST + k] - q[k], -0.2, 0.2);
    for (int r = 0; r < 6; r++) { double s = acc[r]; for (int k = 0; k < 7; k++) s -= J[r * 7 + k] * qd_null[k]; rhs[r] = s; }
    for (int c = 0; c < 6; c++) { double s = 0; for (int r = 0; r < 6; r++) s += V[r * 6 + c] * rhs[r]; y[c] = s / clampd(fabs(w[c]), C[C_SVD_MIN], C[C_SVD_MAX]); }
    for (int r = 0; r < 6; r++) { double s = 0; for (int c = 0; c < 6; c++) s += V[r * 6 + c] * y[c]; x[r] = s; }
    double qd[7], nrm = 0;
    for (int k = 0; k < 7; k++) { double s = qd_null[k]; for (int r = 0; r < 6; r++) s += J[r * 7 + k] * x[r]; qd[k] = s; nrm += s * s; }
    nrm = sqrt(nrm);
    if (nrm > 3) for (int k = 0; k < 7; k++) qd[k] *= 3 / nrm;
    for (int k = 0; k < 7; k++) q[k] = clampd(q[k] + C[C_LRATE] * qd[k], C[C_JMIN + k], C[C_JMAX + k]);
  }
  for (int k = 0; k < 7; k++) { e->jt_q[k] = q[k]; e->jt_qd[k] = (q[k] - old_q[k]) / C[C_DT]; e->ik_q[k] = q[k]; }
  for (int k = 0; k < 7; k++) tau[k] = C[C_PD_P + k] * (e->jt_q[k] - e->qpos[k]) + C[C_PD_D + k] * (e->jt_qd[k] - e->qvel[k]);
}

/* RobotBase.fing_ctrl_step — core/Robots.py:441-476 */
static void finger_ctrl(const Env* e, double* f) {
  double w0 = e->qpos[7], w1 = e->qpos[8], v0 = e->qvel[7], v1 = e->qvel[8], mean = 0.5 * (w0 + w1);
  double f0 = 500 * (mean - w0), f1 = 500 * (mean - w1), g0, g1;
  if (mean - e->grip_set > 0.005) {
    if (e->grasp_flag) { g0 = g1 = -20; } else { g0 = 10 * (-0.2 - v0); g1 = 10 * (-0.2 - v1); }
  } else {
    g0 = clampd(500 * (e->grip_set - w0) - 10 * v0, -5, 5); g1 = clampd(500 * (e->grip_set - w1) - 10 * v1, -5, 5);
  }
  f[0] = f0 + g0; f[1] = f1 + g1;
}

/* ------------------------------------------------------------------ one physics tick: Scene.next_step — core/Scene.py:121-138 */
static void forward_position(Env* e) { kinematics(e); crba(e); collision(e); make_constraints(e); }

static void update_tcp(Env* e) {               /* data.body('tcp_rb0').xpos/xquat of the current forward pass */
  double o[3], Rt[9], R[9];
  mat_vec3(o, e->xmat[6], e->ctrl + C_TCP_POS);
  for (int k = 0; k < 3; k++) e->tcp_pos[k] = e->xpos[6][k] + o[k];
  quat2mat(Rt, e->ctrl + C_TCP_QUAT); mat_mul3(R, e->xmat[6], Rt); mat2quat(e->tcp_quat, R);
}

static void physics_step(Env* e) {
  int nv = e->nv;
  double h = e->ctrl[C_DT];
  /* --- MjRobot.prepare_step (MjRobot.py:125-131): controller -> ctrl, using last step's qfrc_bias */
  double tau[7], fing[2];
  if (e->ctrl_mode == 1) cart_controller(e, tau);
  else for (int k = 0; k < 7; k++) tau[k] = e->ctrl[C_PD_P + k] * (e->jt_q[k] - e->qpos[k]) + e->ctrl[C_PD_D + k] * (e->jt_qd[k] - e->qvel[k]);
  finger_ctrl(e, fing);
  for (int k = 0; k < 7; k++) e->ctrl_out[k] = tau[k] + e->bias_prev[k];
  for (int k = 0; k < 2; k++) e->ctrl_out[7 + k] = fing[k];      /* MjRobot.py:130: gripper ctrl = finger_commands (no gravity comp) */
  /* --- mujoco.mj_step [EXT] */
  forward_position(e);
  update_tcp(e);
  rne_bias(e);
  for (int d = 0; d < nv; d++) {
    int li = e->dof_link[d];
    double passive = e->link[li].jtype == 2 ? 0 : -e->link[li].damping * e->qvel[d];
    double act = 0;
    if (d < NROB) { double fr = e->link[li].forcerange; act = fr > 0 ? clampd(e->ctrl_out[d], -fr, fr) : e->ctrl_out[d]; }
    e->qfrc_smooth[d] = passive - e->bias[d] + act;
  }
  static double L[MAXV * MAXV];
  memcpy(L, e->M, 8 * nv * nv);
  if (chol_factor(L, nv, nv)) { e->status |= 1; return; }
  memcpy(e->qacc_smooth, e->qfrc_smooth, 8 * nv); chol_solve(L, nv, nv, e->qacc_smooth);
  solve_constraints(e);
  memcpy(e->warm, e->qacc, 8 * nv);
  for (int k = 0; k < NROB; k++) e->bias_prev[k] = e->bias[k];
  /* --- mj_Euler with implicit joint damping */
  static double qa[MAXV];
  int damped = 0; for (int i = 0; i < e->nlink; i++) if (e->link[i].jtype != 2 && e->link[i].damping > 0) damped = 1;
  if (damped) {
    memcpy(L, e->M, 8 * nv * nv);
    for (int d = 0; d < nv; d++) { int li = e->dof_link[d]; if (e->link[li].jtype != 2) L[d * nv + d] += h * e->link[li].damping; }
    for (int d = 0; d < nv; d++) qa[d] = e->qfrc_smooth[d] + e->qfrc_constraint[d];
    chol_factor(L, nv, nv); chol_solve(L, nv, nv, qa);
  } else memcpy(qa, e->qacc, 8 * nv);
  for (int d = 0; d < nv; d++) e->qvel[d] += h * qa[d];
  for (int i = 0; i < e->nlink; i++) {
    const Link* Lk = &e->link[i];
    if (Lk->jtype != 2) { e->qpos[Lk->qadr] += h * e->qvel[Lk->dadr]; continue; }
    double* q = e->qpos + Lk->qadr; const double* v = e->qvel + Lk->dadr;
    for (int k = 0; k < 3; k++) q[k] += h * v[k];
    double ang = norm3(v + 3) * h;
    if (ang > 0) {
      double s = sin(ang / 2) / (ang / h), dq[4] = {cos(ang / 2), s * v[3], s * v[4], s * v[5]}, r[4];
      quat_mul(r, q + 3, dq);
      double n = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
      for (int k = 0; k < 4; k++) q[3 + k] = r[k] / n;
    }
  }
  for (int d = 0; d < e->nq; d++) if (!isfinite(e->qpos[d])) e->status |= 1;
}
void d3o_substep(Env* e, int n) { for (int i = 0; i < n; i++) physics_step(e); }

/* ------------------------------------------------------------------ task logic */
static double dist3(const double* a, const double* b) { double d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; return norm3(d); }

/* yaw observation: tan(quat2euler(q)[-1]) — geometric_transformation.py:92-153 (mat2euler: euler[2] = -atan2(R01, R00) when cy > eps) */
static double tan_yaw(const double* q) {
  double R[9]; quat2mat(R, q);
  double cy = sqrt(R[8] * R[8] + R[5] * R[5]);
  double yaw = cy > 4 * 2.220446049250313e-16 ? -atan2(R[1], R[0]) : -atan2(-R[3], R[4]);
  return tan(yaw);
}

static void get_obs(const Env* e, float* obs) {
  if (e->task_id == TASK_PUSHING) {            /* pushing.py:255-280 */
    const double *b1 = e->qpos + 9, *b2 = e->qpos + 16;
    obs[0] = (float)e->tcp_pos[0]; obs[1] = (float)e->tcp_pos[1];
    obs[2] = (float)b1[0]; obs[3] = (float)b1[1]; obs[4] = (float)tan_yaw(b1 + 3);
    obs[5] = (float)b2[0]; obs[6] = (float)b2[1]; obs[7] = (float)tan_yaw(b2 + 3);
  } else if (e->task_id == TASK_AVOIDING) {    /* avoiding.py:117-119 */
    obs[0] = (float)e->tcp_pos[0]; obs[1] = (float)e->tcp_pos[1];
  } else if (e->task_id == TASK_SORTING) {     /* sorting.py:308-390: tcp xy, then (xy, tan yaw) of red_1.., blue_1.. */
    obs[0] = (float)e->tcp_pos[0]; obs[1] = (float)e->tcp_pos[1];
    for (int i = 0; i < e->nobj; i++) {
      const double* b = e->qpos + NROB + 7 * i;
      obs[2 + 3 * i] = (float)b[0]; obs[3 + 3 * i] = (float)b[1]; obs[4 + 3 * i] = (float)tan_yaw(b + 3);
    }
  } else if (e->task_id == TASK_INSERTING) {   /* gate_insertion.py:288-320: tcp xy, then (xy, tan yaw) of push_box1..3 */
    obs[0] = (float)e->tcp_pos[0]; obs[1] = (float)e->tcp_pos[1];
    for (int i = 0; i < 3; i++) {
      const double* b = e->qpos + NROB + 7 * i;
      obs[2 + 3 * i] = (float)b[0]; obs[3 + 3 * i] = (float)b[1]; obs[4 + 3 * i] = (float)tan_yaw(b + 3);
    }
  } else if (e->task_id == TASK_STACKING) {    /* stacking.py:228-277: (pos xyz, tan yaw) of red, green, blue */
    for (int i = 0; i < 3; i++) {
      const double* b = e->qpos + NROB + 7 * i;
      obs[4 * i] = (float)b[0]; obs[4 * i + 1] = (float)b[1]; obs[4 * i + 2] = (float)b[2]; obs[4 * i + 3] = (float)tan_yaw(b + 3);
    }
  } else if (e->task_id == TASK_ALIGNING) {    /* aligning.py:205-235: tcp xyz, box pos + quat (qpos), target pos + quat (model.body_pos/quat) */
    const double* b = e->qpos + NROB;
    for (int k = 0; k < 3; k++) obs[k] = (float)e->tcp_pos[k];
    for (int k = 0; k < 7; k++) { obs[3 + k] = (float)b[k]; obs[10 + k] = (float)e->extra[k]; }
  }
}

/* ---- Sorting (sorting.py:460-543).  task_state: [0] mode_step, [1] bit i set <=> mode[i] == 0 (red), [2] bitmask of min_inds.
 * Box slots are (red_1, red_2, red_3, blue_1, blue_2, blue_3); slots absent from the scene read the constant pose of the
 * model's last body through mj_name2id = -1: taskp[11..12]. */
static void sorting_slot_xy(const Env* e, int slot, double* xy) {
  int half = e->nobj / 2, col = slot / 3, idx = slot % 3;
  if (idx < half) { const double* b = e->qpos + NROB + 7 * (col * half + idx); xy[0] = b[0]; xy[1] = b[1]; }
  else { xy[0] = e->taskp[11]; xy[1] = e->taskp[12]; }
}
static int sorting_in_bin(const Env* e, int col, const double* xy) {
  const double* T = e->taskp;
  return xy[0] > T[4 + 2 * col] && xy[0] < T[5 + 2 * col] && xy[1] > T[8] && xy[1] < T[9];
}
static int sorting_early_term(Env* e) {        /* sorting.py:509-543 */
  int half = e->nobj / 2;
  for (int i = 0; i < e->nobj; i++) if (!sorting_in_bin(e, i / half, e->qpos + NROB + 7 * i)) return 0;
  e->terminated = 1;
  return 1;
}
static int sorting_check_mode(Env* e) {         /* sorting.py:460-507 + decode_mode (np.packbits of mode[:k], -1 and 1 both pack as 1) */
  int step = (int)e->task_state[0], zero = (int)e->task_state[1], mins = (int)e->task_state[2];
  if (step <= 5) {
    double best = 0; int bi = -1; double bxy[2] = {0, 0};
    for (int s2 = 0; s2 < 6; s2++) {
      double xy[2]; sorting_slot_xy(e, s2, xy);
      const double* tg = e->taskp + 2 * (s2 / 3);
      double d = (mins >> s2) & 1 ? 100000.0 : sqrt((xy[0] - tg[0]) * (xy[0] - tg[0]) + (xy[1] - tg[1]) * (xy[1] - tg[1]));
      if (bi < 0 || d < best) { best = d; bi = s2; bxy[0] = xy[0]; bxy[1] = xy[1]; }
    }
    if (sorting_in_bin(e, bi / 3, bxy)) {
      if (bi < 3) zero |= 1 << step;
      step++; mins |= 1 << bi;
    }
    e->task_state[0] = step; e->task_state[1] = zero; e->task_state[2] = mins;
  }
  int code = 0;
  for (int i = 0; i < e->nobj; i++) if (!((zero >> i) & 1)) code |= 1 << (7 - i);
  return code;
}

/* ---- Stacking (stacking.py:395-447).  task_state: [0] len(mode_encoding), [1] mode string as base-4 digits (r 1, g 2, b 3;
 * first arrival in the lowest digit), [2] bitmask of min_inds.  taskp: target xy, pos_min_dist, min z gap, gripper threshold. */
static int stacking_early_term(Env* e) {
  const double *r = e->qpos + NROB, *g = r + 7, *b = g + 7, *T = e->taskp;
  double dz = fmin(fabs(r[2] - g[2]), fmin(fabs(r[2] - b[2]), fabs(g[2] - b[2])));
  double dr = hypot(r[0] - T[0], r[1] - T[1]), dg = hypot(g[0] - T[0], g[1] - T[1]), db = hypot(b[0] - T[0], b[1] - T[1]);
  if (dr <= T[2] && dg <= T[2] && db <= T[2] && dz > T[3]) { e->terminated = 1; return 1; }
  return 0;
}
static double stacking_check_mode(Env* e) {     /* returns mean_distance; appends to the mode string */
  const double* T = e->taskp;
  int len = (int)e->task_state[0], code = (int)e->task_state[1], mins = (int)e->task_state[2];
  double d[3], mean = 0, best = 0; int bi = -1;
  for (int i = 0; i < 3; i++) { const double* b = e->qpos + NROB + 7 * i; d[i] = hypot(b[0] - T[0], b[1] - T[1]); mean += d[i] / 3; }
  for (int i = 0; i < 3; i++) { double di = (mins >> i) & 1 ? 100000.0 : d[i]; if (bi < 0 || di < best) { best = di; bi = i; } }
  if (best <= T[2]) { int p4 = 1; for (int k = 0; k < len; k++) p4 *= 4; code += (bi + 1) * p4; len++; mins |= 1 << bi; }
  e->task_state[0] = len; e->task_state[1] = code; e->task_state[2] = mins;
  return mean;
}

/* ---- Inserting (gate_insertion.py:396-470).  task_state: [0] len(modes), [1] arrival order in base-4 digits (r 1, g 2, b 3),
 * [2] bitmask of boxes already listed.  taskp: three target positions (xyz), target_min_dist.  Distances are 3-D (C11). */
static void inserting_dists(const Env* e, double* d3) {
  for (int i = 0; i < 3; i++) d3[i] = dist3(e->qpos + NROB + 7 * i, e->taskp + 3 * i);
}
static int inserting_early_term(Env* e) {
  double d[3]; inserting_dists(e, d);
  if (d[0] <= e->taskp[9] && d[1] <= e->taskp[9] && d[2] <= e->taskp[9]) { e->terminated = 1; return 1; }
  return 0;
}
static void inserting_check_mode(Env* e) {
  double d[3]; inserting_dists(e, d);
  int len = (int)e->task_state[0], code = (int)e->task_state[1], seen = (int)e->task_state[2];
  for (int i = 0; i < 3; i++) if (d[i] <= e->taskp[9] && !((seen >> i) & 1)) {
    int p4 = 1; for (int k = 0; k < len; k++) p4 *= 4;
    code += (i + 1) * p4; len++; seen |= 1 << i;
  }
  e->task_state[0] = len; e->task_state[1] = code; e->task_state[2] = seen;
}
static int inserting_mode_id(int code, int len) {      /* mode_dict {'rgb':1,'rbg':2,'grb':3,'gbr':4,'brg':5,'bgr':6}, 0 unless all three arrived */
  if (len != 3) return 0;
  int a = code % 4, b = (code / 4) % 4;
  return a == 1 ? (b == 2 ? 1 : 2) : a == 2 ? (b == 1 ? 3 : 4) : (b == 1 ? 5 : 6);
}

/* ---- Aligning (aligning.py:21-30,295-352) */
static double rotation_distance(const double* p, const double* q) {
  double d = fabs(p[0] * q[0] + p[1] * q[1] + p[2] * q[2] + p[3] * q[3]);
  return 2 * acos(d > 1 ? 1 : d);               /* numpy would return NaN for |p.q| = 1 + ulp; clamped here */
}
static void aligning_dists(const Env* e, double* pos_d, double* rot_d) {
  const double* b = e->qpos + NROB;
  *pos_d = dist3(b, e->extra); *rot_d = rotation_distance(b + 3, e->extra + 3) / 3.14159265358979323846;
}
static int aligning_early_term(Env* e) {
  double pd, rd; aligning_dists(e, &pd, &rd);
  if (pd <= e->taskp[0] && rd <= e->taskp[1]) { e->terminated = 1; return 1; }
  return 0;
}

static int pushing_early_term(Env* e) {        /* pushing.py:440-459 */
  const double *b1 = e->qpos + 9, *b2 = e->qpos + 16, *g1 = e->taskp, *g2 = e->taskp + 3;
  double md = e->taskp[6];
  double rr = dist3(b1, g1), rg = dist3(b1, g2), gr = dist3(b2, g1), gg = dist3(b2, g2);
  if ((rr <= md && gg <= md) || (rg <= md && gr <= md)) { e->terminated = 1; return 1; }
  return 0;
}
static void pushing_check_mode(Env* e, int* mode, double* mean_distance) {   /* pushing.py:341-377; task_state[0] = first_visit */
  const double *b1 = e->qpos + 9, *b2 = e->qpos + 16, *g1 = e->taskp, *g2 = e->taskp + 3;
  double md = e->taskp[6];
  double rr = dist3(b1, g1), rg = dist3(b1, g2), gr = dist3(b2, g1), gg = dist3(b2, g2);
  int fv = (int)e->task_state[0], visit = -1; *mode = -1;
  if (rr <= md && fv != 0) visit = 0; else if (rg <= md && fv != 1) visit = 1; else if (gr <= md && fv != 2) visit = 2; else if (gg <= md && fv != 3) visit = 3;
  if (fv == -1) e->task_state[0] = visit;
  else {
    if (fv == 0 && visit == 3) *mode = 0; else if (fv == 3 && visit == 0) *mode = 1; else if (fv == 1 && visit == 2) *mode = 2; else if (fv == 2 && visit == 1) *mode = 3;
  }
  *mean_distance = 0.5 * (fmin(rr, rg) + fmin(gr, gg));
}

static int avoiding_early_term(Env* e) {       /* avoiding.py:204-246; task_state[1] = success */
  int success = e->tcp_pos[1] > e->taskp[3];
  if (success || e->obst_contact) { if (success) e->task_state[1] = 1; e->terminated = 1; return 1; }
  return 0;
}
static void avoiding_check_mode(Env* e) {      /* avoiding.py:173-202; task_state[2] = passed bits, [3] = 9-bit mode code */
  double x = e->tcp_pos[0], y = e->tcp_pos[1]; const double* T = e->taskp;
  int passed = (int)e->task_state[2], code = (int)e->task_state[3];
  if (y - 0.03 <= T[0] && T[0] <= y + 0.03 && !(passed & 1)) { if (x < T[4]) code |= 1 << 0; else if (x > T[4]) code |= 1 << 1; passed |= 1; }
  if (y - 0.03 <= T[1] && T[1] <= y + 0.03 && !(passed & 2)) {
    if (x < T[5]) code |= 1 << 2; else if (T[5] < x && x < T[6]) code |= 1 << 3; else if (x > T[6]) code |= 1 << 4; passed |= 2; }
  if (y >= T[2] && !(passed & 4)) {
    if (x < T[7]) code |= 1 << 5;
    if (T[7] < x && x < T[8]) code |= 1 << 6; else if (T[8] < x && x < T[9]) code |= 1 << 7; else if (x > T[7]) code |= 1 << 8;   /* sic: l3_top_xpos, avoiding.py:200 */
    passed |= 4; }
  e->task_state[2] = passed; e->task_state[3] = code;
}

static double get_reward(const Env* e) {
  if (e->task_id == TASK_PUSHING) {            /* pushing.py:379-404 */
    const double *b1 = e->qpos + 9;
    double dx = e->tcp_pos[0] - b1[0], dy = e->tcp_pos[1] - b1[1];
    return -(sqrt(dx * dx + dy * dy) + dist3(b1, e->taskp));
  }
  if (e->task_id == TASK_INSERTING) {          /* gate_insertion.py:427-449 */
    double d[3], mn = 1e300; inserting_dists(e, d);
    for (int i = 0; i < 3; i++) { const double* b = e->qpos + NROB + 7 * i; double r = hypot(e->tcp_pos[0] - b[0], e->tcp_pos[1] - b[1]); if (r < mn) mn = r; }
    return -(mn + d[0] + d[1] + d[2]);
  }
  if (e->task_id == TASK_ALIGNING) {           /* aligning.py:321-332 */
    double pd, rd; aligning_dists(e, &pd, &rd);
    return -rd - 3.5 * pd;
  }
  return 0;
}

/* ------------------------------------------------------------------ Gym API: reset / step */
/* task reset: pushing.py:461-488 / avoiding.py:248-262 (+ MjScene.reset MjScene.py:120-143, beam_to_joint_pos Robots.py:580-589) */
void d3o_reset(Env* e, const double* ctx) {
  memset(e->qpos, 0, sizeof e->qpos); memset(e->qvel, 0, sizeof e->qvel); memset(e->warm, 0, sizeof e->warm);
  memcpy(e->qpos, e->ctrl + C_INIT_QPOS, 8 * NARM);
  for (int i = NROB; i < e->nlink; i++) { memcpy(e->qpos + e->link[i].qadr, e->link[i].pos, 24); memcpy(e->qpos + e->link[i].qadr + 3, e->link[i].quat, 32); }
  e->ik_valid = 0; e->ctrl_mode = 0; e->grasp_flag = 0; e->grip_set = 0.001;      /* Robots.py:127 */
  e->step_count = 0; e->terminated = 0; e->status = 0; e->obst_contact = 0;
  memset(e->task_state, 0, sizeof e->task_state);
  if (e->task_id == TASK_PUSHING) e->task_state[0] = -1;
  if (e->task_id == TASK_ALIGNING) memcpy(e->extra, e->taskp + 3, 56);           /* XML pose of `target_box` */
  /* set_q -> mj_forward at (init_qpos, 0): fresh tcp pose and qfrc_bias for the first command */
  kinematics(e); update_tcp(e); rne_bias(e);
  for (int k = 0; k < NROB; k++) e->bias_prev[k] = e->bias[k];
  memcpy(e->jt_q, e->qpos, 8 * NARM); memset(e->jt_qd, 0, sizeof e->jt_qd);       /* jointTrackingController.setSetPoint(init_qpos) */
  /* manager.start(context): raw qpos write, no mj_forward */
  if (ctx) for (int i = NROB; i < e->nlink; i++) memcpy(e->qpos + e->link[i].qadr, ctx + 7 * (i - NROB), 56);
  if (e->task_id == TASK_STACKING) e->grip_set = 0.04;                             /* stacking.py:474 robot.open_fingers() before the reset tick */
  if (ctx && e->nextra) memcpy(e->extra, ctx + 7 * e->nobj, 8 * e->nextra);        /* joint-less target body: model.body_pos/quat (MjScene.py:271-285) */
  physics_step(e);                                                                 /* scene.next_step(): exactly one tick */
}

/* GymEnvWrapper.step — gyms/gym_env_wrapper.py:45-100 + task overrides */
void d3o_step(Env* e, const double* action, float* obs, double* reward, int* done, double* info) {
  if (e->ctrl_kind == 1) {
    /* CubeStacking_Env.step (stacking.py:331-393): 8-D action = joint set-point + gripper command; joint PD (zero FF) */
    if (action[7] > e->taskp[4]) { e->grip_set = 0.04; e->grasp_flag = 0; }        /* open_fingers() */
    else { e->grip_set = 0.0; e->grasp_flag = 1; }                                 /* close_fingers(duration=0) Robots.py:430-435 */
    memcpy(e->jt_q, action, 8 * NARM); memset(e->jt_qd, 0, sizeof e->jt_qd);       /* Controller.py:99-127 setSetPoint(q, 0, 0) */
    e->ctrl_mode = 2;
  } else {
    e->grip_set = 0.04; e->grasp_flag = 0;                                         /* :67 robot.open_fingers() */
    double n = sqrt(action[3] * action[3] + action[4] * action[4] + action[5] * action[5] + action[6] * action[6]);
    memcpy(e->des_pos, action, 24); for (int k = 0; k < 4; k++) e->des_quat[k] = action[3 + k] / n;   /* IKControllers.py:346-362 */
    e->ctrl_mode = 1;
  }
  get_obs(e, obs); *reward = get_reward(e);
  int early = e->task_id == TASK_PUSHING ? pushing_early_term(e) : e->task_id == TASK_SORTING ? sorting_early_term(e)
            : e->task_id == TASK_ALIGNING ? aligning_early_term(e) : e->task_id == TASK_STACKING ? stacking_early_term(e) : e->task_id == TASK_INSERTING ? inserting_early_term(e) : avoiding_early_term(e);
  *done = e->terminated || early || e->step_count >= e->max_steps - 1;             /* :124-137 */
  for (int i = 0; i < e->n_substeps; i++) physics_step(e);
  e->step_count++;
  memset(info, 0, 8 * e->info_dim);
  if (e->task_id == TASK_PUSHING) {            /* pushing.py:335-339 */
    int success = pushing_early_term(e), mode; double md;
    pushing_check_mode(e, &mode, &md);
    info[0] = success; info[1] = mode; info[2] = md; info[3] = e->status;
  } else if (e->task_id == TASK_SORTING) {     /* sorting.py:444-458 */
    int success = sorting_early_term(e);
    info[0] = success; info[1] = sorting_check_mode(e); info[2] = e->task_state[0]; info[3] = e->status;
  } else if (e->task_id == TASK_INSERTING) {   /* gate_insertion.py:366-394: success, mode id (0 unless three boxes arrived), mean_distance, len(modes) */
    int success = inserting_early_term(e);
    inserting_check_mode(e);
    double d[3]; inserting_dists(e, d);
    info[0] = success; info[1] = inserting_mode_id((int)e->task_state[1], (int)e->task_state[0]); info[2] = (d[0] + d[1] + d[2]) / 3;
    info[3] = e->task_state[0]; info[4] = e->status;
  } else if (e->task_id == TASK_STACKING) {    /* stacking.py:381-393: success, mode string, mean_distance, len(mode) (success_1/2 = len > 0 / > 1) */
    int success = stacking_early_term(e);
    double md = stacking_check_mode(e);
    info[0] = success; info[1] = e->task_state[1]; info[2] = md; info[3] = e->task_state[0]; info[4] = e->status;
  } else if (e->task_id == TASK_ALIGNING) {    /* aligning.py:289-319 */
    int success = aligning_early_term(e);
    double pd, rd; aligning_dists(e, &pd, &rd);
    const double* b = e->qpos + NROB;
    double rb = sqrt((b[0] - e->tcp_pos[0]) * (b[0] - e->tcp_pos[0]) + (b[1] - e->tcp_pos[1]) * (b[1] - e->tcp_pos[1]));
    info[0] = success; info[1] = rb < e->taskp[2] ? 0 : 1; info[2] = 0.5 * (pd + rd); info[3] = e->status;
  } else {                                     /* avoiding.py:168-171 */
    avoiding_check_mode(e);
    info[0] = e->task_state[1];
    int code = (int)e->task_state[3];
    for (int k = 0; k < 9; k++) info[1 + k] = (code >> k) & 1;
    info[10] = e->status;
  }
}

void d3o_robot_state(const Env* e, double* tcp) { memcpy(tcp, e->tcp_pos, 24); }
/* CubeStacking_Env.robot_state (stacking.py:218-226): joint positions + gripper width (MjRobot.py:176-183: sum of the finger joints) */
void d3o_joint_state(const Env* e, double* j8) { memcpy(j8, e->qpos, 8 * NARM); j8[7] = e->qpos[7] + e->qpos[8]; }
void d3o_get_obs(const Env* e, float* obs) { get_obs(e, obs); }

/* ------------------------------------------------------------------ flat state (shared layout with the CUDA library's get/set_state) */
int d3o_state_dim(const Env* e) { return e->nq + 2 * e->nv + 60 + e->nextra; }
#define PUT(arr, n) do { memcpy(p, arr, 8 * (n)); p += (n); } while (0)
#define GET(arr, n) do { memcpy(arr, p, 8 * (n)); p += (n); } while (0)
void d3o_get_state(const Env* e, double* p) {
  PUT(e->qpos, e->nq); PUT(e->qvel, e->nv); PUT(e->warm, e->nv); PUT(e->bias_prev, 9); PUT(e->tcp_pos, 3); PUT(e->tcp_quat, 4);
  PUT(e->ik_q, 7); PUT(e->des_pos, 3); PUT(e->des_quat, 4); PUT(e->jt_q, 7); PUT(e->jt_qd, 7);
  double s[8] = {e->ik_valid, e->ctrl_mode, e->grip_set, e->grasp_flag, e->step_count, e->terminated, e->status, e->obst_contact};
  PUT(s, 8); PUT(e->task_state, 8); PUT(e->extra, e->nextra);
}
void d3o_set_state(Env* e, const double* p) {
  GET(e->qpos, e->nq); GET(e->qvel, e->nv); GET(e->warm, e->nv); GET(e->bias_prev, 9); GET(e->tcp_pos, 3); GET(e->tcp_quat, 4);
  GET(e->ik_q, 7); GET(e->des_pos, 3); GET(e->des_quat, 4); GET(e->jt_q, 7); GET(e->jt_qd, 7);
  double s[8]; GET(s, 8);
  e->ik_valid = (int)s[0]; e->ctrl_mode = (int)s[1]; e->grip_set = s[2]; e->grasp_flag = (int)s[3]; e->step_count = (int)s[4];
  e->terminated = (int)s[5]; e->status = (int)s[6]; e->obst_contact = (int)s[7];
  GET(e->task_state, 8); GET(e->extra, e->nextra);
}

/* ------------------------------------------------------------------ probes for the invariant tests (tests/ only) */
void d3o_forward(Env* e) {                     /* forward pass at the current state with zero control */
  int nv = e->nv;
  forward_position(e); update_tcp(e); rne_bias(e);
  for (int d = 0; d < nv; d++) e->qfrc_smooth[d] = -e->bias[d];
  static double L[MAXV * MAXV];
  memcpy(L, e->M, 8 * nv * nv); chol_factor(L, nv, nv);
  memcpy(e->qacc_smooth, e->qfrc_smooth, 8 * nv); chol_solve(L, nv, nv, e->qacc_smooth);
  solve_constraints(e);
}
int d3o_probe(const Env* e, const char* what, double* out, int cap) {
  int nv = e->nv, n = 0;
#define OUT(ptr, cnt) do { n = (cnt); if (n > cap) return -n; memcpy(out, ptr, 8 * n); return n; } while (0)
  if (!strcmp(what, "M")) OUT(e->M, nv * nv);
  if (!strcmp(what, "bias")) OUT(e->bias, nv);
  if (!strcmp(what, "qacc")) OUT(e->qacc, nv);
  if (!strcmp(what, "qacc_smooth")) OUT(e->qacc_smooth, nv);
  if (!strcmp(what, "qfrc_constraint")) OUT(e->qfrc_constraint, nv);
  if (!strcmp(what, "efc_force")) OUT(e->efc_force, e->nefc);
  if (!strcmp(what, "efc_aref")) OUT(e->aref, e->nefc);
  if (!strcmp(what, "efc_D")) OUT(e->D, e->nefc);
  if (!strcmp(what, "efc_J")) OUT(e->J, e->nefc * nv);
  if (!strcmp(what, "ctrl")) OUT(e->ctrl_out, NROB);
  if (!strcmp(what, "xpos")) OUT(e->xpos, 3 * e->nlink);
  if (!strcmp(what, "xmat")) OUT(e->xmat, 9 * e->nlink);
  if (!strcmp(what, "counts")) { double c[4] = {e->ncon, e->nefc, e->solver_iter, e->status}; OUT(c, 4); }
  if (!strcmp(what, "contacts")) {             /* per contact: pos3, normal3, dist, dim, g1, g2, mu, efc */
    n = 12 * e->ncon; if (n > cap) return -n;
    for (int i = 0; i < e->ncon; i++) {
      const Contact* c = &e->con[i]; double* o = out + 12 * i;
      memcpy(o, c->pos, 24); memcpy(o + 3, c->frame, 24); o[6] = c->dist; o[7] = c->dim; o[8] = c->g1; o[9] = c->g2; o[10] = c->mu; o[11] = c->efc;
    }
    return n;
  }
  return 0;
}
/* stand-alone narrow-phase probes */
int d3o_collide(int t1, const double* p1, const double* q1, const double* s1, int t2, const double* p2, const double* q2, const double* s2, double* out) {
  double R1[9], R2[9]; quat2mat(R1, q1); quat2mat(R2, q2);
  RawContact rc[8]; int n = 0;
  if (t1 == G_BOX && t2 == G_BOX) n = collide_box_box(p1, R1, s1, p2, R2, s2, 0, rc);
  else if (t1 == G_CYLINDER && t2 == G_BOX) n = collide_cyl_box(p1, R1, s1, p2, R2, s2, 0, rc);
  else if (t1 == G_CYLINDER && t2 == G_CYLINDER) n = collide_cyl_cyl(p1, R1, s1, p2, R2, s2, 0, rc);
  for (int i = 0; i < n; i++) { memcpy(out + 7 * i, rc[i].pos, 24); memcpy(out + 7 * i + 3, rc[i].n, 24); out[7 * i + 6] = rc[i].dist; }
  return n;
}
