// d3il_model.h — host-side: D3SC scene blob -> `Model` (fp32 tables + index tables) and flat-state (de)serialisation.
// Shared by the CUDA library (d3il_capi.cu) and the CPU emulation harness used in tests.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>

#include "d3il_env.cuh"

#define D3SC_MAGIC 0x43533344
#define D3SC_HDR_INTS 32

static inline bool d3il_build_model(const void* blob, size_t nbytes, Model& m, Lay& L, std::string& err) {
  const int32_t* h = (const int32_t*)blob;
  if (nbytes < 4 * D3SC_HDR_INTS || h[0] != D3SC_MAGIC || h[1] != 1) { err = "not a D3SC v1 scene blob"; return false; }
  memset(&m, 0, sizeof(m));
  m.task_id = h[2]; m.nlink = h[3]; m.nobj = h[4]; m.nq = h[5]; m.nv = h[6]; m.ngeom = h[7]; m.npair = h[8]; m.n_substeps = h[9];
  m.max_steps = h[10]; m.obs_dim = h[11]; m.act_dim = h[12]; m.ctx_dim = h[13]; m.info_dim = h[14]; m.ctrl_kind = h[15]; m.ntaskp = h[16]; m.nextra = h[17];
  const int hdr_maxcon = h[18];
  if (m.nlink > D3_MAXLINK || m.nq > D3_MAXQ || m.nv > D3_MAXV || m.ngeom > D3_MAXGEOM || m.npair > D3_MAXPAIR || m.ntaskp > 32 || m.nextra < 0 || m.nextra > 8) { err = "scene exceeds compiled table sizes"; return false; }
  if (m.task_id != D3T_AVOIDING && m.task_id != D3T_PUSHING && m.task_id != D3T_ALIGNING && m.task_id != D3T_SORTING && m.task_id != D3T_STACKING && m.task_id != D3T_INSERTING) { err = "task not supported by this build of the CUDA path"; return false; }
  size_t need = 4 * D3SC_HDR_INTS + 8 * ((size_t)m.nlink * D3_LINK_W + (size_t)m.ngeom * D3_GEOM_W + (size_t)m.npair * D3_PAIR_W + D3_CTRL_W + m.ntaskp);
  if (need != nbytes) { err = "scene blob size mismatch"; return false; }
  const double* p = (const double*)((const char*)blob + 4 * D3SC_HDR_INTS);
  for (int i = 0; i < m.nlink; i++, p += D3_LINK_W) {
    for (int k = 0; k < D3_LINK_W; k++) m.link[D3_LINK_W * i + k] = (tab_t)p[k];
    m.l_parent[i] = (int)p[0]; m.l_jtype[i] = (int)p[1]; m.l_limited[i] = (int)p[22]; m.l_qadr[i] = (int)p[29]; m.l_dadr[i] = (int)p[30];
    m.l_ndof[i] = m.l_jtype[i] == 2 ? 6 : 1;
    {
      double n = sqrt(p[5] * p[5] + p[6] * p[6] + p[7] * p[7] + p[8] * p[8]);
      double w = p[5] / n, x = p[6] / n, y = p[7] / n, z = p[8] / n;
      double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
      for (int k = 0; k < 9; k++) m.linkR[9 * i + k] = (tab_t)R[k];
    }
    m.l_anc[i] = (1u << i) | (m.l_parent[i] >= 0 ? m.l_anc[m.l_parent[i]] : 0u);
    for (int k = 0; k < m.l_ndof[i]; k++) m.d_link[m.l_dadr[i] + k] = i;
  }
  for (int i = 0; i < m.nlink; i++) for (int j = 0; j < m.nlink; j++) if ((m.l_anc[j] >> i) & 1u) m.l_desc[i] |= 1u << j;
  for (int i = 0; i < m.nlink; i++) {
    int root = i; while (m.l_parent[root] >= 0) root = m.l_parent[root];
    m.l_ref[i] = (root == 0 && m.nlink > 6) ? 6 : root;     // arm tree: wrist (link 7 origin); free bodies: themselves
  }
  for (int i = 0; i < m.ngeom; i++, p += D3_GEOM_W) {
    for (int k = 0; k < D3_GEOM_W; k++) m.geom[D3_GEOM_W * i + k] = (tab_t)p[k];
    double q[4] = {p[5], p[6], p[7], p[8]};
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                   2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
    for (int k = 0; k < 9; k++) m.geomR[9 * i + k] = (tab_t)R[k];
    m.g_slab[i] = ((int)p[0] == D3G_BOX && (int)p[1] < 0 && w == 1.0 && x == 0.0 && y == 0.0 && z == 0.0) ? 1 : 0;
  }
  int conmax = 0;
  for (int i = 0; i < m.npair; i++, p += D3_PAIR_W) {
    for (int k = 0; k < D3_PAIR_W; k++) m.pair[D3_PAIR_W * i + k] = (tab_t)p[k];
    int t1 = (int)m.geom[D3_GEOM_W * (int)p[0]], t2 = (int)m.geom[D3_GEOM_W * (int)p[1]];
    conmax += (t1 == D3G_BOX && t2 == D3G_BOX) ? 8 : 1;
    if ((int)p[2] != 3 && (int)p[2] != 4) { err = "only condim 3 / 4 contacts are supported by the CUDA path"; return false; }
    if ((int)p[2] > m.maxdim) m.maxdim = (int)p[2];
  }
  for (int k = 0; k < D3_CTRL_W; k++) m.ctrl[k] = (tab_t)p[k];
  p += D3_CTRL_W;
  for (int k = 0; k < m.ntaskp; k++) m.taskp[k] = (tab_t)p[k];
  // structurally non-zero entries of M: dof a >= b with link(b) an ancestor-or-self of link(a)
  m.nmpair = 0;
  for (int a = 0; a < m.nv; a++) for (int b = 0; b <= a; b++)
    if ((m.l_anc[m.d_link[a]] >> m.d_link[b]) & 1u) {
      if (m.nmpair >= D3_MAXV * 8) { err = "mass-matrix pair table overflow"; return false; }
      m.mp_a[m.nmpair] = (unsigned char)a; m.mp_b[m.nmpair] = (unsigned char)b; m.nmpair++;
    }
  // kinematic-tree blocks of the (block-diagonal) mass matrix, implicit-damping corner, triangle unranking table
  m.maxblk = 0;
  for (int d = 0; d < m.nv; d++) {
    int root = m.d_link[d]; while (m.l_parent[root] >= 0) root = m.l_parent[root];
    int lo = m.nv, hi = 0;
    for (int e = 0; e < m.nv; e++) { int r = m.d_link[e]; while (m.l_parent[r] >= 0) r = m.l_parent[r]; if (r == root) { if (e < lo) lo = e; if (e + 1 > hi) hi = e + 1; } }
    m.d_bs[d] = lo; m.d_be[d] = hi;
    if (hi - lo > m.maxblk) m.maxblk = hi - lo;
  }
  if (m.maxblk > D3_MAXB) { err = "a kinematic tree has more dofs than the register-resident block factorisation handles (D3_MAXB)"; return false; }
  m.ndamp = 0; m.damp_first = m.damp_end = 0;
  for (int d = 0; d < m.nv; d++) if (m.l_jtype[m.d_link[d]] != 2 && m.link[D3_LINK_W * m.d_link[d] + 25] > 0) { if (!m.ndamp) m.damp_first = d; m.ndamp++; m.damp_end = d + 1; }
  if (m.ndamp) {
    bool ok = m.ndamp <= 4 && m.damp_end - m.damp_first == m.ndamp && m.damp_end == m.d_be[m.damp_first];
    if (!ok) { err = "damped dofs must be the trailing dofs of their block"; return false; }
  }
  for (int ip = 0; ip < m.npair; ip++) {
    int l1 = (int)m.geom[D3_GEOM_W * (int)m.pair[D3_PAIR_W * ip] + 1], l2 = (int)m.geom[D3_GEOM_W * (int)m.pair[D3_PAIR_W * ip + 1] + 1];
    int a0, a1, b0, b1, cpl = 0;
    link_range(m, l1, &a0, &a1); link_range(m, l2, &b0, &b1);
    if (a1 > a0 && b1 > b0) {
      if (b0 < a0) { int t0 = a0, t1 = a1; a0 = b0; a1 = b1; b0 = t0; b1 = t1; }
      if (b0 < a1) { a1 = a1 > b1 ? a1 : b1; b0 = b1 = 0; }      // same tree: one merged range
      else cpl = 1;                                               // two different blocks
    } else if (a1 == a0) { a0 = b0; a1 = b1; b0 = b1 = 0; }
    if ((a1 - a0) + (b1 - b0) > D3_JW) { err = "contact dof ranges exceed the compact Jacobian row width"; return false; }
    m.p_rng[4 * ip] = (unsigned char)a0; m.p_rng[4 * ip + 1] = (unsigned char)a1; m.p_rng[4 * ip + 2] = (unsigned char)b0; m.p_rng[4 * ip + 3] = (unsigned char)b1;
    m.p_cpl[ip] = (unsigned char)cpl;
  }
  { int e = 0; for (int i = 0; i < 24; i++) for (int j = 0; j <= i; j++) { m.tri_i[e] = (unsigned char)i; m.tri_j[e] = (unsigned char)j; e++; } }
  m.nblk = 0; m.diag_blk = 0; m.nzp = 0;
  for (int d = 0; d < m.nv; d++) if (m.d_bs[d] == d) { if (m.nblk >= 8) { err = "too many kinematic trees"; return false; } m.blk_s[m.nblk] = d; m.blk_e[m.nblk] = m.d_be[d];
    {
      const int li = m.d_link[d]; const tab_t* Lk = m.link + D3_LINK_W * li;
      if (m.l_jtype[li] == 2 && Lk[13] == 0 && Lk[14] == 0 && Lk[15] == 0 && Lk[19] == 0 && Lk[20] == 0 && Lk[21] == 0) m.diag_blk |= 1u << m.nblk;
    }
    m.nblk++; }
  for (int d = 0; d < m.nv; d++) for (int b = 0; b < m.nblk; b++) if (d >= m.blk_s[b] && d < m.blk_e[b]) m.d_blk[d] = (unsigned char)b;
  m.nhe = 0;
  for (int b = 0; b < m.nblk; b++) for (int i = m.blk_s[b]; i < m.blk_e[b]; i++) for (int j = m.blk_s[b]; j <= i; j++) {
    if (m.nhe >= D3_MAXHE) { err = "too many in-block Hessian entries (D3_MAXHE)"; return false; }
    m.he_i[m.nhe] = (unsigned char)i; m.he_j[m.nhe] = (unsigned char)j; m.nhe++;
  }
  // packed mass matrix: the blocks back to back as row-major squares
  m.m_size = 0;
  for (int b = 0; b < m.nblk; b++) {
    const int bs = m.blk_s[b], sz = m.blk_e[b] - bs;
    for (int d = bs; d < bs + sz; d++) { m.m_base[d] = (short)m.m_size; m.m_row[d] = (short)(m.m_size + (d - bs) * sz - bs); }
    m.m_size += sz * sz;
  }
  // in-block (a > b) dof pairs that are not ancestor-related: CRBA never writes them
  for (int a = 0; a < m.nv; a++) for (int b = m.d_bs[a]; b < a; b++)
    if (!((m.l_anc[m.d_link[a]] >> m.d_link[b]) & 1u)) {
      if (m.nzp >= 8) { err = "too many unrelated in-block dof pairs"; return false; }
      m.zp_a[m.nzp] = (unsigned char)a; m.zp_b[m.nzp] = (unsigned char)b; m.nzp++;
    }
  // contact budget: 4 per free body resting on a support + 12 for transients (deep spawn penetration touches two supports,
  // box-box / rod contacts); overflow raises status bit 2, never drops silently
  { int cap = hdr_maxcon > 0 ? hdr_maxcon : 4 * m.nobj + 12; m.maxcon = ((conmax < cap ? conmax : cap) + 3) & ~3; }
  if (m.maxcon > 64) { err = "contact budget above 64 (row / block contact lists keep 6-bit contact ids)"; return false; }
  if (m.maxdim < 3) m.maxdim = 3;
  m.maxrow = m.maxdim * m.maxcon + 4;
  d3il_layout(m, L);
  m.ws_floats = L.total;
  return true;
}

// ---- flat fp64 state, identical layout to oracle/d3il_oracle.c::d3o_get_state (nq + 2 nv + 60)
static inline int d3il_state_dim(const Model& m) { return m.nq + 2 * m.nv + 60 + m.nextra; }

template <class T>
static inline void d3il_pack_state(const Model& m, const Lay& L, const T* w /*workspace/state row*/, const IkState& ik, double* out) {
  double* p = out;
  for (int k = 0; k < m.nq; k++) *p++ = (double)w[L.qpos + k] + (k < D3_NROB ? (double)w[L.qlo + k] : 0.0);
  for (int k = 0; k < m.nv; k++) *p++ = w[L.qvel + k];
  for (int k = 0; k < m.nv; k++) *p++ = w[L.warm + k];
  for (int k = 0; k < 9; k++) *p++ = w[L.bias_prev + k];
  for (int k = 0; k < 7; k++) *p++ = w[L.tcp + k];
  for (int k = 0; k < 7; k++) *p++ = ik.q[k];
  for (int k = 0; k < 3; k++) *p++ = ik.des_pos[k];
  for (int k = 0; k < 4; k++) *p++ = ik.des_quat[k];
  for (int k = 0; k < 7; k++) *p++ = (double)ik.jt_q[k] + (double)ik.jt_qlo[k];
  for (int k = 0; k < 7; k++) *p++ = ik.jt_qd[k];
  *p++ = ik.valid; *p++ = w[L.misc + ST_CTRL_MODE]; *p++ = w[L.misc + ST_GRIP_SET]; *p++ = w[L.misc + ST_GRASP]; *p++ = w[L.misc + ST_STEP];
  *p++ = w[L.misc + ST_TERM]; *p++ = w[L.misc + ST_STATUS]; *p++ = w[L.misc + ST_OBST];
  for (int k = 0; k < 4; k++) *p++ = w[L.misc + ST_TASK0 + k];
  for (int k = 0; k < 3; k++) *p++ = w[L.misc + ST_COST_ITERS + k];     // diagnostics only (the oracle keeps zeros here)
  *p++ = 0;
  for (int k = 0; k < m.nextra; k++) *p++ = w[L.extra + k];
}
template <class T>
static inline void d3il_unpack_state(const Model& m, const Lay& L, T* w, IkState& ik, const double* in) {
  const double* p = in;
  for (int k = 0; k < m.nq; k++) { w[L.qpos + k] = (T)*p; if (k < D3_NROB) w[L.qlo + k] = (T)(*p - (double)w[L.qpos + k]); p++; }
  for (int k = 0; k < m.nv; k++) w[L.qvel + k] = (T)*p++;
  for (int k = 0; k < m.nv; k++) w[L.warm + k] = (T)*p++;
  for (int k = 0; k < 9; k++) w[L.bias_prev + k] = (T)*p++;
  for (int k = 0; k < 7; k++) w[L.tcp + k] = (T)*p++;
  for (int k = 0; k < 7; k++) ik.q[k] = *p++;
  for (int k = 0; k < 3; k++) ik.des_pos[k] = (real)*p++;
  for (int k = 0; k < 4; k++) ik.des_quat[k] = (real)*p++;
  for (int k = 0; k < 7; k++) { ik.jt_q[k] = (real)*p; ik.jt_qlo[k] = (real)(*p - (double)ik.jt_q[k]); p++; }
  for (int k = 0; k < 7; k++) ik.jt_qd[k] = (real)*p++;
  ik.valid = (int)*p++;
  w[L.misc + ST_CTRL_MODE] = (T)*p++; w[L.misc + ST_GRIP_SET] = (T)*p++; w[L.misc + ST_GRASP] = (T)*p++; w[L.misc + ST_STEP] = (T)*p++;
  w[L.misc + ST_TERM] = (T)*p++; w[L.misc + ST_STATUS] = (T)*p++; w[L.misc + ST_OBST] = (T)*p++;
  for (int k = 0; k < 4; k++) w[L.misc + ST_TASK0 + k] = (T)*p++;
  p += 4;
  for (int k = 0; k < m.nextra; k++) w[L.extra + k] = (T)*p++;
}
