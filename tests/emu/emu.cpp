// emu.cpp — TEST-ONLY host build of the CUDA kernel core (d3il_b200/csrc/d3il_core.cuh, d3il_env.cuh) with G = 1.
// Lets the kernel logic be checked against the fp64 oracle in the CPU container; not part of the product path
// (the product library d3il_b200/csrc/libd3il.so has no CPU execution path at all).
#define D3IL_EMU 1
#include <stdlib.h>

#include <vector>

#include "../../d3il_b200/csrc/d3il_model.h"

struct Emu {
  Model m; Lay L; IkState ik; double V[42], sn[7], cs[7]; int vwarm;
  std::vector<real> w;
  real tol; int max_iter;
};
static const Cx CX = {0, 0};

extern "C" {
Emu* emu_create(const void* blob, size_t n) {
  Emu* e = new Emu();
  std::string err;
  if (!d3il_build_model(blob, n, e->m, e->L, err)) { fprintf(stderr, "emu_create: %s\n", err.c_str()); delete e; return nullptr; }
  e->w.assign(e->L.total, 0);
  memset(&e->ik, 0, sizeof(e->ik));
  e->vwarm = 0;
  e->tol = sizeof(real) == 4 ? (real)1e-6 : (real)1e-10;
  e->max_iter = sizeof(real) == 4 ? 32 : 100;
  return e;
}
void emu_destroy(Emu* e) { delete e; }
void emu_set_solver(Emu* e, double tol, int max_iter) { e->tol = (real)tol; e->max_iter = max_iter; }
int emu_ws_floats(Emu* e) { return e->L.total; }
int emu_n_state(Emu* e) { return e->L.n_state; }
int emu_state_dim(Emu* e) { return d3il_state_dim(e->m); }

static void tick(Emu* e) {
  real* w = e->w.data();
  if (w[e->L.misc + ST_CTRL_MODE] == 1) {
    if (!e->ik.valid) { for (int k = 0; k < 7; k++) e->ik.q[k] = (double)w[e->L.qpos + k] + (double)w[e->L.qlo + k]; e->ik.valid = 1; }
    double J[42], Aw[27], Cd[D3_CTRL_W], coop[160];
    for (int k = 0; k < D3_CTRL_W; k++) Cd[k] = (double)e->m.ctrl[k];
    ik_tick<1>(CX, Cd, e->ik, 1, e->V, &e->vwarm, e->sn, e->cs, J, Aw, coop);
  }
  if (e->m.maxdim == 4) physics_tick<1, false, 4>(CX, e->m, e->L, w, e->ik.jt_q, e->ik.jt_qlo, e->ik.jt_qd, e->tol, e->max_iter);
  else physics_tick<1, false, 3>(CX, e->m, e->L, w, e->ik.jt_q, e->ik.jt_qlo, e->ik.jt_qd, e->tol, e->max_iter);
}
void emu_reset(Emu* e, const double* ctx) {
  std::vector<float> c(e->m.ctx_dim > 0 ? e->m.ctx_dim : 1);
  if (ctx) for (int k = 0; k < e->m.ctx_dim; k++) c[k] = (float)ctx[k];
  memset(&e->ik, 0, sizeof(e->ik));
  for (int k = 0; k < 7; k++) { e->ik.jt_q[k] = e->m.ctrl[D3C_INIT_QPOS + k]; e->ik.jt_qlo[k] = 0; e->ik.jt_qd[k] = 0; }
  env_reset<1>(CX, e->m, e->L, e->w.data(), ctx ? c.data() : nullptr, e->tol, e->max_iter);
}
void emu_step(Emu* e, const double* action, float* obs, double* reward, int* done, double* info) {
  real* w = e->w.data();
  float act[8];
  for (int k = 0; k < e->m.act_dim; k++) act[k] = (float)action[k];
  if (e->m.ctrl_kind == 1) {      // joint-space set-point (k_ik's ctrl_kind == 1 branch); float32 like the device action buffer
    for (int k = 0; k < 7; k++) { e->ik.jt_q[k] = (real)act[k]; e->ik.jt_qlo[k] = 0; e->ik.jt_qd[k] = 0; }
  } else {
    double n = sqrt(action[3] * action[3] + action[4] * action[4] + action[5] * action[5] + action[6] * action[6]);
    for (int k = 0; k < 3; k++) e->ik.des_pos[k] = (real)action[k];
    for (int k = 0; k < 4; k++) e->ik.des_quat[k] = (real)(action[3 + k] / n);
  }
  float r; unsigned char d; float inf[16];
  e->vwarm = 0;                 // like a kernel launch: exact sin/cos and a cold eigenbasis at the first IK iteration
  env_prestep<1>(CX, e->m, e->L, w, act, obs, &r, &d);
  for (int i = 0; i < e->m.n_substeps; i++) tick(e);
  env_poststep<1>(CX, e->m, e->L, w, inf);
  *reward = r; *done = d;
  for (int k = 0; k < e->m.info_dim; k++) info[k] = inf[k];
}
void emu_substep(Emu* e, int n) { e->vwarm = 0; for (int i = 0; i < n; i++) tick(e); }
void emu_robot_state(Emu* e, double* tcp) { for (int k = 0; k < 3; k++) tcp[k] = e->w[e->L.tcp + k]; }
void emu_get_obs(Emu* e, float* obs) { task_obs(e->m, e->L, e->w.data(), obs); }
int emu_probe(Emu* e, const char* what, double* out, int cap) {
  const Lay& L = e->L; const Model& m = e->m; int off = -1, n = 0;
  std::string s(what);
  if (s == "M") {          // dense nv x nv view of the packed buffer (upper triangle + diagonal are the mass matrix, the strict lower triangle its factor)
    n = m.nv * m.nv; if (n > cap) return -1;
    for (int r = 0; r < m.nv; r++) for (int c = 0; c < m.nv; c++) out[r * m.nv + c] = (c >= m.d_bs[r] && c < m.d_be[r]) ? (double)e->w[L.M + m.m_row[r] + c] : 0.0;
    return n;
  } else if (s == "bias") { off = L.bias; n = m.nv; } else if (s == "qacc") { off = L.qacc; n = m.nv; }
  else if (s == "qacc_smooth") { off = L.qacc_smooth; n = m.nv; } else if (s == "efc_J") { off = L.J; n = m.maxrow * m.nv; }
  else if (s == "efc_aref") { off = L.aref; n = m.maxrow; } else if (s == "efc_D") { off = L.D; n = m.maxrow; } else if (s == "efc_force") { off = L.frcE; n = m.maxrow; }
  else if (s == "contacts") { off = L.con; n = D3_CON_W * m.maxcon; } else if (s == "qfrc_c") { off = L.qfrc_c; n = m.nv; } else if (s == "act") { off = L.act; n = 9; }
  else if (s == "qfrc_smooth") { off = L.qfrc_smooth; n = m.nv; }
  if (off < 0 || n > cap) return -1;
  for (int k = 0; k < n; k++) out[k] = e->w[off + k];
  return n;
}
void emu_get_state(Emu* e, double* out) { d3il_pack_state(e->m, e->L, e->w.data(), e->ik, out); }
void emu_set_state(Emu* e, const double* in) { d3il_unpack_state(e->m, e->L, e->w.data(), e->ik, in); e->vwarm = 0; }
}

// stand-alone narrow-phase probes of the kernel core (fast path vs general routine)
extern "C" int emu_collide_boxes(const double* pA, const double* hA, const double* pB, const double* qB, const double* hB, int use_fast, double* out) {
  real pa[3], ha[3], pb[3], qb[4], hb[3], RA[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, RB[9];
  for (int k = 0; k < 3; k++) { pa[k] = (real)pA[k]; ha[k] = (real)hA[k]; pb[k] = (real)pB[k]; hb[k] = (real)hB[k]; }
  for (int k = 0; k < 4; k++) qb[k] = (real)qB[k];
  quat2mat(RB, qb);
  RawCon rc[8];
  int n = use_fast ? collide_slab_box(pa, ha, pb, RB, hb, sqrt(hb[0] * hb[0] + hb[1] * hb[1] + hb[2] * hb[2]), 0, rc) : collide_box_box(pa, RA, ha, pb, RB, hb, 0, rc);
  for (int i = 0; i < n; i++) { for (int k = 0; k < 3; k++) { out[7 * i + k] = rc[i].pos[k]; out[7 * i + 3 + k] = rc[i].n[k]; } out[7 * i + 6] = rc[i].dist; }
  return n;
}

// stand-alone probe of the IK's clipped-spectrum solve: method 0 = ik_solve_spd (Cholesky / one-eigenvalue deflation), 1 = Jacobi
// lanes path.  Returns 1 if the method produced x.
extern "C" int emu_ik_clipped_solve(const double* J42, const double* rhs, double reg, double lo, double hi, int method, double* x) {
  if (method == 0) { double v1[6], Aw[27]; int ok = 0; return ik_solve_spd(J42, reg, rhs, lo, hi, x, v1, &ok, Aw); }
  double A[36], V[36], B1[36], B2[36], cs[8], r[6];
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double sum = i == j ? reg : 0; for (int k = 0; k < 7; k++) sum += J42[i * 7 + k] * J42[j * 7 + k]; A[i * 6 + j] = sum; }
  for (int k = 0; k < 6; k++) r[k] = rhs[k];
  ik_solve_clipped_lanes<1>(CX, A, r, V, B1, B2, cs, 0, lo, hi, x);
  return 1;
}
