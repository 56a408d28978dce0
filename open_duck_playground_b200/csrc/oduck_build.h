// oduck_build.h -- OduckModel (include/oduck.h) -> the device-side tables of oduck_device.cuh.  Pure host code: included by
// oduck_cuda.cu (oduck_create) and by the CPU warp emulation of the kernels' device code (tests/emu, test infrastructure).
#pragma once
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/oduck.h"
#include "oduck_hfcollide.cuh"

static int build_dev_model(const OduckModel& M, DevModel& D, std::string& err) {
  memset(&D, 0, sizeof(D));
  if (M.nv > 32 || M.nbody > 32 || M.nq > 36 || M.nu > 16 || M.foot_nvert > 32) { err = "model exceeds the warp-per-env limits (nv, nbody, nvert <= 32)"; return -1; }
  D.nbody = M.nbody; D.njnt = M.njnt; D.nq = M.nq; D.nv = M.nv; D.nu = M.nu;
  D.nvert = M.foot_nvert; D.enable_ff = M.enable_foot_foot;
  D.iterations = M.iterations; D.ls_iterations = M.ls_iterations;
  D.timestep = (float)M.timestep;
  for (int i = 0; i < 3; i++) D.gravity[i] = (float)M.gravity[i];
  D.tolerance = (float)M.tolerance; D.ls_tolerance = (float)M.ls_tolerance; D.meaninertia = (float)M.meaninertia; D.impratio = (float)M.impratio;
  {
    double timeconst = std::max(M.solref[0], 2 * M.timestep), damp = M.solref[1];
    auto clampd = [](double v, double lo, double hi) { return std::min(std::max(v, lo), hi); };
    double dmin = clampd(M.solimp[0], 1e-4, 0.9999), dmax = clampd(M.solimp[1], 1e-4, 0.9999);
    double k = 1 / (dmax * dmax * timeconst * timeconst * damp * damp), b = 2 / (dmax * timeconst);
    if (M.solref[0] <= 0) k = -M.solref[0] / (dmax * dmax);
    if (M.solref[1] <= 0) b = -M.solref[1] / dmax;
    D.sol_k = (float)k; D.sol_b = (float)b; D.dmin = (float)dmin; D.dmax = (float)dmax;
    D.width = (float)std::max(1e-15, M.solimp[2]); D.mid = (float)clampd(M.solimp[3], 1e-4, 0.9999); D.power = (float)std::max(1.0, M.solimp[4]);
  }
  D.floor_mu = (float)M.floor_friction; D.foot_mu = (float)M.foot_friction;
  if (M.iterations != 1) { err = "only option iterations=1 (the reference scenes) is implemented"; return -1; }
  // bodies
  auto quat2mat = [](const double* q, double* R) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
  };
  int depth[32] = {0};
  D.maxdepth = 0;
  for (int b = 0; b < 32; b++) { D.b_depth[b] = -1; D.b_lastdof[b] = -1; D.b_quat[0][b] = 1.f; }
  for (int b = 0; b < M.nbody; b++) {
    int p = M.body_parentid[b];
    depth[b] = b == 0 ? 0 : depth[p] + 1;
    D.b_depth[b] = b == 0 ? 0 : depth[b];
    D.maxdepth = std::max(D.maxdepth, depth[b]);
    D.b_parent[b] = p;
    for (int i = 0; i < 3; i++) { D.b_pos[i][b] = (float)M.body_pos[b][i]; D.b_ipos[i][b] = (float)M.body_ipos[b][i]; }
    for (int i = 0; i < 4; i++) D.b_quat[i][b] = (float)M.body_quat[b][i];
    D.b_mass[b] = (float)M.body_mass[b];
    D.b_invw0[b] = (float)M.body_invweight0[b][0];
    double Ri[9];
    quat2mat(M.body_iquat[b], Ri);
    double Ib[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Ib[3 * r + c] = Ri[3 * r] * M.body_inertia[b][0] * Ri[3 * c] + Ri[3 * r + 1] * M.body_inertia[b][1] * Ri[3 * c + 1] + Ri[3 * r + 2] * M.body_inertia[b][2] * Ri[3 * c + 2];
    D.b_Ib[0][b] = (float)Ib[0]; D.b_Ib[1][b] = (float)Ib[4]; D.b_Ib[2][b] = (float)Ib[8];
    D.b_Ib[3][b] = (float)Ib[1]; D.b_Ib[4][b] = (float)Ib[2]; D.b_Ib[5][b] = (float)Ib[5];
    D.b_dofadr[b] = M.body_dofadr[b];
    int nj = M.body_jntnum[b], j0 = M.body_jntadr[b];
    D.b_jtype[b] = 0;
    if (nj > 0) {
      if (M.jnt_type[j0] == ODUCK_JNT_FREE) {
        if (nj != 1) { err = "free joint must be alone on its body"; return -1; }
        D.b_jtype[b] = 1; D.b_qadr0[b] = M.jnt_qposadr[j0];
      } else {
        if (nj > 2) { err = "at most two hinge joints per body"; return -1; }
        D.b_jtype[b] = 1 + nj;
        for (int k = 0; k < nj; k++) {
          int j = j0 + k;
          if (M.jnt_type[j] != ODUCK_JNT_HINGE) { err = "unsupported joint type"; return -1; }
          if (M.jnt_pos[j][0] != 0 || M.jnt_pos[j][1] != 0 || M.jnt_pos[j][2] != 0) { err = "hinge joints must sit at the body origin (jnt_pos = 0)"; return -1; }
          for (int i = 0; i < 3; i++) (k == 0 ? D.b_ax0 : D.b_ax1)[i][b] = (float)M.jnt_axis[j][i];
          (k == 0 ? D.b_qadr0 : D.b_qadr1)[b] = M.jnt_qposadr[j];
        }
      }
    }
    // last dof of the nearest ancestor-or-self with dofs
    int a = b;
    while (a > 0 && M.body_dofnum[a] == 0) a = M.body_parentid[a];
    D.b_lastdof[b] = (a > 0) ? M.body_dofadr[a] + M.body_dofnum[a] - 1 : -1;
  }
  for (int b = 1; b < M.nbody; b++) {   // subtree masks
    int a = b;
    while (a > 0) { D.b_submask[a] |= 1 << b; a = M.body_parentid[a]; }
  }
  // dofs
  int maxchain = 1, nfr = 0, nlim = 0;
  for (int d = 0; d < 32; d++) { D.d_parent[d] = -1; D.d_vparent[d] = -1; D.d_act[d] = -1; D.d_frrow[d] = -1; D.d_limrow[d] = -1; }
  for (int d = 0; d < M.nv; d++) {
    int j = M.dof_jntid[d], b = M.dof_bodyid[d];
    D.d_body[d] = b;
    D.d_parent[d] = M.dof_parentid[d];
    D.d_bsubmask[d] = D.b_submask[b];
    int chain = 0;
    for (int a = d; a >= 0; a = M.dof_parentid[a]) { D.d_ancmask[d] |= 1 << a; chain++; }
    maxchain = std::max(maxchain, chain);
    D.d_damping[d] = (float)M.dof_damping[d];
    D.d_invw0[d] = (float)M.dof_invweight0[d];
    D.d_floss[d] = (float)M.dof_frictionloss[d];
    D.d_arm[d] = (float)M.dof_armature[d];
    int k = d - M.jnt_dofadr[j];
    if (M.jnt_type[j] == ODUCK_JNT_FREE) {
      D.d_flags[d] = k < 3 ? DF_TRANS : DF_ROT;
      D.d_qadr[d] = k < 3 ? M.jnt_qposadr[j] + k : M.jnt_qposadr[j] + 3;
      D.d_vparent[d] = k < 3 ? M.dof_parentid[d] : M.jnt_dofadr[j] + 2;
    } else {
      D.d_flags[d] = DF_HINGE;
      D.d_qadr[d] = M.jnt_qposadr[j];
      D.d_vparent[d] = M.dof_parentid[d];
      if (M.jnt_limited[j]) { D.d_flags[d] |= DF_LIMITED; D.d_lo[d] = (float)M.jnt_range[j][0]; D.d_hi[d] = (float)M.jnt_range[j][1]; D.d_limrow[d] = nlim++; }
    }
    if (M.dof_frictionloss[d] > 0) {
      D.d_flags[d] |= DF_FLOSS;
      D.d_frrow[d] = nfr++;
      double imp = D.dmin;
      D.d_Dfric[d] = (float)(1.0 / std::max(M.dof_invweight0[d] * (1 - imp) / imp, 1e-15));
    } else {
      D.d_Dfric[d] = 1.f;
    }
  }
  D.nfr = nfr; D.nlim = nlim;
  for (int d = 0; d < M.nv; d++) {
    int cnt = 0, tmp[32];
    for (int a = M.dof_parentid[d]; a >= 0; a = M.dof_parentid[a]) tmp[cnt++] = a;
    D.d_depth[d] = cnt;
    D.max_dof_depth = std::max(D.max_dof_depth, cnt);
    for (int t = 0; t < cnt; t++) D.anc[d][t] = (unsigned char)tmp[cnt - 1 - t];   // root first
  }
  for (int a = 0, p = 0; a < 32 && p < 512; a++)
    for (int b = 0; b <= a && p < 512; b++, p++) D.pair_ab[p] = (unsigned short)((a << 8) | b);
  {
    int cnt = 0;
    for (int k = 0; k < M.nv; k++) {
      D.chol_ofs[k] = cnt;
      const int dd = D.d_depth[k], rk = k * (k + 1) / 2;
      for (int a = 0; a < dd; a++)
        for (int b = 0; b <= a; b++) {
          const int ia = D.anc[k][a], ib = D.anc[k][b];
          if (cnt >= 1536) { err = "factorisation table overflow"; return -1; }
          D.chol_tab[cnt++] = (unsigned)(ia * (ia + 1) / 2 + ib) | ((unsigned)(rk + ia) << 10) | ((unsigned)(rk + ib) << 20);
        }
    }
    for (int k = M.nv; k <= 32; k++) D.chol_ofs[k] = cnt;
  }
  {
    // chain plan: root chain 0..nb-1 (each the parent of the next), every other dof in a pure chain attached to dof nb-1
    int nb = 0;
    while (nb < M.nv && M.dof_parentid[nb] == nb - 1 && (nb == 0 || true)) { nb++; if (nb < M.nv && M.dof_parentid[nb] != nb - 1) break; }
    // nb = length of the initial run with parent(d) = d-1; the first branch shares that run, so cut it at the free joint's 6 dofs
    D.plan_ok = 0;
    const int NB6 = 6;
    if (M.nv > NB6) {
      bool ok = true;
      for (int d2 = 1; d2 < NB6; d2++) ok = ok && M.dof_parentid[d2] == d2 - 1;
      int starts[8], lens[8], nbr = 0;
      for (int d2 = NB6; d2 < M.nv && ok; ) {
        if (M.dof_parentid[d2] != NB6 - 1 || nbr >= 8) { ok = false; break; }
        int len = 1;
        while (d2 + len < M.nv && M.dof_parentid[d2 + len] == d2 + len - 1) len++;
        starts[nbr] = d2; lens[nbr] = len; nbr++;
        d2 += len;
      }
      if (ok && nbr == 3) {
        // two equal branches (legs) + one single (head)
        int a = -1, b = -1, c2 = -1;
        if (lens[0] == lens[2]) { a = 0; b = 2; c2 = 1; } else if (lens[0] == lens[1]) { a = 0; b = 1; c2 = 2; } else if (lens[1] == lens[2]) { a = 1; b = 2; c2 = 0; }
        if (a >= 0 && lens[c2] == 4 && (lens[a] == 10 || lens[a] == 5)) {
          D.plan_ok = lens[a] == 10 ? 1 : 2;
          D.plan_nbase = NB6; D.plan_pair_len = lens[a]; D.plan_pair_start[0] = starts[a]; D.plan_pair_start[1] = starts[b];
          D.plan_single_len = lens[c2]; D.plan_single_start = starts[c2];
        }
      }
    }
    (void)nb;
  }
  D.n_mpairs = 0;
  for (int i = 0; i < M.nv; i++)
    for (int j = i; j >= 0; j = M.dof_parentid[j]) D.mpair[D.n_mpairs++] = (unsigned short)((i << 8) | j);
  D.body_rounds = 0;
  while ((1 << D.body_rounds) < D.maxdepth) D.body_rounds++;
  {
    // chain-scan plan for subtree sums over the moving bodies 1..nbody-1 (world excluded; children of the world are roots)
    int nchild[32] = {0}, only[32];
    for (int b = 0; b < 32; b++) { only[b] = -1; D.b_next[b] = -1; }
    for (int b = 1; b < M.nbody; b++) { int p = M.body_parentid[b]; if (p > 0) { nchild[p]++; only[p] = b; } }
    int maxchain = 1;
    bool ok = true;
    D.n_branch = 0;
    for (int b = 1; b < M.nbody; b++) if (nchild[b] == 1) D.b_next[b] = only[b];
    for (int b = 1; b < M.nbody; b++) { int len = 1; for (int x = b; D.b_next[x] >= 0; x = D.b_next[x]) len++; maxchain = std::max(maxchain, len); }
    D.scan_rounds = 0;
    while ((1 << D.scan_rounds) < maxchain) D.scan_rounds++;
    // branching bodies, deepest first (body ids grow with depth along a path, so descending id order is a valid order)
    for (int b = M.nbody - 1; b >= 1 && ok; b--) {
      if (nchild[b] < 2) continue;
      if (D.n_branch >= 4 || nchild[b] > 4) { ok = false; break; }
      const int k = D.n_branch++;
      D.br_nchild[k] = 0;
      for (int c = 1; c < M.nbody; c++) if (M.body_parentid[c] == b) D.br_child[k][D.br_nchild[k]++] = c;
      int mask = 1 << b;                                   // the chain that ends in b: b and its single-child ancestors
      for (int x = M.body_parentid[b]; x > 0 && nchild[x] == 1; x = M.body_parentid[x]) mask |= 1 << x;
      D.br_chain[k] = mask;
    }
    D.scan_ok = ok ? 1 : 0;
  }
  for (int b = 0; b < M.nbody; b++) {
    int j0 = M.body_jntadr[b];
    D.b_sameaxis[b] = (M.body_jntnum[b] == 2 && M.jnt_axis[j0][0] == M.jnt_axis[j0 + 1][0] && M.jnt_axis[j0][1] == M.jnt_axis[j0 + 1][1] && M.jnt_axis[j0][2] == M.jnt_axis[j0 + 1][2]);
  }
  D.prefix_rounds = 0;
  while ((1 << D.prefix_rounds) < maxchain) D.prefix_rounds++;
  for (int u = 0; u < M.nu; u++) {
    int j = M.act_jntid[u], d = M.jnt_dofadr[j];
    D.d_act[d] = u;
    D.d_kp[d] = (float)M.act_kp[u]; D.d_kv[d] = (float)M.act_kv[u];
    D.d_clo[d] = (float)M.act_ctrlrange[u][0]; D.d_chi[d] = (float)M.act_ctrlrange[u][1];
    D.d_flo[d] = (float)M.act_forcerange[u][0]; D.d_fhi[d] = (float)M.act_forcerange[u][1];
    D.act_dof[u] = d; D.act_qadr[u] = M.jnt_qposadr[j];
    D.act_bl_qadr[u] = -1;   // base.py:121-125: "<name>_backlash" joint = the next joint on the same body
    if (j + 1 < M.njnt && M.jnt_bodyid[j + 1] == M.jnt_bodyid[j] && M.jnt_type[j + 1] == ODUCK_JNT_HINGE) D.act_bl_qadr[u] = M.jnt_qposadr[j + 1];
    D.key_ctrl[u] = (float)M.key_ctrl[u];
  }
  for (int i = 0; i < M.nq; i++) { D.key_qpos[i] = (float)M.key_qpos[i]; D.qpos0[i] = (float)M.qpos0[i]; }
  // sites
  auto chain_of_body = [&](int b) { int a = b; while (a > 0 && M.body_dofnum[a] == 0) a = M.body_parentid[a]; int mask = 0; if (a > 0) for (int d = M.body_dofadr[a] + M.body_dofnum[a] - 1; d >= 0; d = M.dof_parentid[d]) mask |= 1 << d; return mask; };
  {
    int sidx = M.imu_site;
    D.imu_body = M.site_bodyid[sidx];
    D.imu_chain = chain_of_body(D.imu_body);
    double R[9];
    quat2mat(M.site_quat[sidx], R);
    for (int i = 0; i < 9; i++) D.imu_rot[i] = (float)R[i];
    for (int i = 0; i < 3; i++) D.imu_pos[i] = (float)M.site_pos[sidx][i];
  }
  for (int k = 0; k < 2; k++) {
    int sidx = M.foot_site[k];
    D.foot_site_body[k] = M.site_bodyid[sidx];
    for (int i = 0; i < 3; i++) D.foot_site_pos[k][i] = (float)M.site_pos[sidx][i];
    D.foot_body[k] = M.foot_body[k];
    D.foot_chain[k] = chain_of_body(M.foot_body[k]);
    for (int v = 0; v < M.foot_nvert; v++)
      for (int i = 0; i < 3; i++) D.vert[k][i][v] = (float)M.foot_vert[k][v][i];
  }
  return 0;
}

static void build_dev_ff(const OduckModel& M, DevFF& f) {
  const OduckModel* model = &M;
  memset(&f, 0, sizeof(f));
  f.nplane = model->foot_nplane; f.nedge = model->foot_nedge; f.nvert = model->foot_nvert; f.radius = (float)model->foot_radius;
  for (int q = 0; q < 32; q++) { f.plane_nvert[q] = model->foot_plane_nvert[q]; for (int k = 0; k < 8; k++) f.plane_vert[q][k] = model->foot_plane_vert[q][k]; }
  for (int e2 = 0; e2 < 48; e2++) for (int k = 0; k < 2; k++) { f.edge_vert[e2][k] = model->foot_edge_vert[e2][k]; f.edge_plane[e2][k] = model->foot_edge_plane[e2][k]; }
  for (int k = 0; k < 2; k++) { for (int i = 0; i < 3; i++) { f.center[k][i] = (float)model->foot_center[k][i]; for (int q = 0; q < 32; q++) f.plane_normal[k][i][q] = (float)model->foot_plane_normal[k][q][i]; } }
}
