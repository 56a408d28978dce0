// oduck_physics.cuh -- one mjx.forward (+ optional Euler step) for one env, executed by one warp.
// Phases follow MJX forward(): fwd_position (kinematics, com_pos, crb, factor_m, collision,
// make_constraint), fwd_velocity, fwd_actuation, fwd_acceleration, solver.solve (Newton, 1 iteration,
// parallel line search), sensors, euler.  See oduck_device.cuh for the lane mapping.
#pragma once
#include "oduck_device.cuh"
#include "oduck_ffcollide.cuh"
#include "oduck_hfcollide.cuh"

// per-thread state that persists across the substeps of one launch
struct Lane {
  float qvel, qaccw, qacc, ctrl, kp, floss, arm;   // dof role (ctrl/kp: the dof's actuator, if any)
  float mass;                                     // body role
  float imt;                                      // 1 / total mass (constant over the substeps of a launch)
  V3 ipos;
  int bar_n;                                      // threads that take part in the CTA's phase barriers: 32 x the warps that own an env in this pass
};

struct LSPoint { float alpha, d0, d1; };

// FF = false: no foot-foot code is compiled in; if the feet's bounding spheres overlap the function returns true BEFORE
// touching any persistent state and the caller re-runs the substep with FF = true (rare).
// BAR = true (k_step): CTA barriers at the phase boundaries selected by ODUCK_BARRIERS keep the CTA's warps on the same
// stretch of this ~200 KB instruction stream, so that they share instruction-cache lines instead of each streaming the code
// from L2 (measured: -15 % kernel time with one barrier per substep).  Every warp of the CTA must execute the same number of
// barriers: points before the foot-foot early exit exist only in the FF = false instantiation (the FF = true re-run skips
// them), points after it exist in both.  The barrier is a NAMED barrier with an explicit thread count (bar.sync 1, n; n = 32 x
// the warps that own an env in this pass of the grid-stride loop): warps of the grid tail without an env take no part in it, and
// warps may arrive from different instantiations (compute-sanitizer synccheck rejects a plain __syncthreads() used that way).
#ifndef ODUCK_BARRIERS
#define ODUCK_BARRIERS 0x01
#endif
#ifndef ODUCK_PHASE_MARK
#define ODUCK_PHASE_MARK(bit)      // tests/emu counts warp exchanges per phase through this hook
#endif
// Height-field instantiations: the collider's work varies from warp to warp (a swing foot costs nothing, a foot lying across cell
// borders clips dozens of (triangle, face) pairs), so the CTA's warps leave it at different times and then run the rest of the
// substep -- ~25 k straight-line instructions -- out of phase, each streaming them through the instruction cache on its own (ncu
// r02f: sm__icc hit rate 62 %, 5.8 no-instruction stall cycles per issue).  A second barrier before the Newton phase (bit 3) puts
// them back in step: 3.89 -> 2.38 ms per rollout step at 4096 envs on B200 (profiles/r02g_bench_rough_*.json; masks 0x0b and
// 0x19 measure the same, no barrier at all is slower than one).  ODUCK_HF_BARRIERS selects the barrier points of those instantiations.
#ifndef ODUCK_HF_BARRIERS
#define ODUCK_HF_BARRIERS 0x09
#endif
#define PHASE_MASK (HF ? (ODUCK_HF_BARRIERS) : (ODUCK_BARRIERS))
__device__ __forceinline__ void phase_barrier(const int nthreads) {
#ifdef ODUCK_WARP_EMU
  __syncthreads();
#else
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
#endif
}
#define PHASE_SYNC(bit, pre_exit) { ODUCK_PHASE_MARK(bit) if (BAR && ((PHASE_MASK >> (bit)) & 1) && (!(pre_exit) || !FF)) phase_barrier(L.bar_n); }

// root-to-leaf sweep after chol_rev_tree: table-free on the chain plans, level-parallel with the ancestor table otherwise
__device__ __forceinline__ float back_tree(const DevModel& m, const float* L, int n, int lane, float y) {
  if (m.plan_ok == 1) return chol_rev_back_chain<10>(m, L, n, lane, y);
  if (m.plan_ok == 2) return chol_rev_back_chain<5>(m, L, n, lane, y);
  return chol_rev_back(m, L, n, lane, y, true);
}

template <bool DBG, bool FF, bool BAR, bool HF>
__device__ __forceinline__ bool forward_euler_impl(const DevModel& m, WarpSmem& s, Lane& L, const int lane, const bool last,
                                              const bool integrate, float* __restrict__ out, float* __restrict__ dbg,
                                              const DevFF* __restrict__ ffm, float* __restrict__ ffs,
                                              const DevHF* __restrict__ hfm, float* __restrict__ hfs) {
  const int nv = m.nv, nb = m.nbody;
  PHASE_SYNC(0, true)
  // ------------------------------------------------------------------ kinematics (lane = body)
  // local transform of every body at once (joint rotations do not depend on the parents), then log2(depth) rounds of
  // pointer jumping compose them into world poses: X_world[b] = X_world[parent] o T_b is an associative prefix product.
  const int jt = m.b_jtype[lane];
  V3 xp = v3(m.b_pos[0][lane], m.b_pos[1][lane], m.b_pos[2][lane]);
  Q4 xq; xq.w = m.b_quat[0][lane]; xq.x = m.b_quat[1][lane]; xq.y = m.b_quat[2][lane]; xq.z = m.b_quat[3][lane];
  V3 a0f = v3(m.b_ax0[0][lane], m.b_ax0[1][lane], m.b_ax0[2][lane]);   // joint-0 axis in the final body frame
  const V3 a1 = v3(m.b_ax1[0][lane], m.b_ax1[1][lane], m.b_ax1[2][lane]);
  int up = m.b_parent[lane] > 0 ? m.b_parent[lane] : -1;               // nearest ancestor not yet folded in (-1: world)
  if (jt == 1) {
    const int qa = m.b_qadr0[lane];
    xp = v3(s.qpos[qa], s.qpos[qa + 1], s.qpos[qa + 2]);
    xq.w = s.qpos[qa + 3]; xq.x = s.qpos[qa + 4]; xq.y = s.qpos[qa + 5]; xq.z = s.qpos[qa + 6];
    xq = qnormalize(xq);
    up = -1;
  } else if (jt >= 2) {
    const int qa = m.b_qadr0[lane];
    float ang0 = s.qpos[qa] - s.qpos0[qa];
    if (jt == 3) {
      const int qb = m.b_qadr1[lane];
      const float ang1 = s.qpos[qb] - s.qpos0[qb];
      if (m.b_sameaxis[lane]) {
        ang0 += ang1;                                                    // two hinges about one axis: angles add
        xq = qmul(xq, axis_angle(a0f, ang0));
      } else {
        const Q4 q1 = axis_angle(a1, ang1);
        xq = qmul(qmul(xq, axis_angle(a0f, ang0)), q1);
        Q4 q1c = q1; q1c.x = -q1.x; q1c.y = -q1.y; q1c.z = -q1.z;
        a0f = qrot(q1c, a0f);
      }
    } else {
      xq = qmul(xq, axis_angle(a0f, ang0));
    }
  }
  for (int r = 0; r < m.body_rounds; ++r) {
    const int src = up < 0 ? 0 : up;
    const V3 pp = v3(__shfl_sync(FULLMASK, xp.x, src), __shfl_sync(FULLMASK, xp.y, src), __shfl_sync(FULLMASK, xp.z, src));
    Q4 pq;
    pq.w = __shfl_sync(FULLMASK, xq.w, src); pq.x = __shfl_sync(FULLMASK, xq.x, src);
    pq.y = __shfl_sync(FULLMASK, xq.y, src); pq.z = __shfl_sync(FULLMASK, xq.z, src);
    const int pup = __shfl_sync(FULLMASK, up, src);
    if (up >= 0) { xp = pp + qrot(pq, xp); xq = qmul(pq, xq); up = pup; }
  }
  xq = qnormalize(xq);
  const V3 ax0w = qrot(xq, a0f), ax1w = qrot(xq, a1);
  float R[9];
  {
    const float w = xq.w, x = xq.x, y = xq.y, z = xq.z;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
    R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
    R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
  }
  s.xpos[0][lane] = xp.x; s.xpos[1][lane] = xp.y; s.xpos[2][lane] = xp.z;
#pragma unroll
  for (int k = 0; k < 9; ++k) s.xmat[k][lane] = R[k];
  if (lane < nb) {
    const int da = m.b_dofadr[lane];
    if (jt == 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        *reinterpret_cast<float4*>(&s.cdof[da + k][0]) = make_float4(0.f, 0.f, 0.f, k == 0);
        *reinterpret_cast<float4*>(&s.cdof[da + k][4]) = make_float4(k == 1, k == 2, 0.f, 0.f);
        *reinterpret_cast<float4*>(&s.cdof[da + 3 + k][0]) = make_float4(R[k], R[3 + k], R[6 + k], 0.f);
      }
    } else if (jt >= 2) {
      *reinterpret_cast<float4*>(&s.cdof[da][0]) = make_float4(ax0w.x, ax0w.y, ax0w.z, 0.f);
      if (jt == 3) *reinterpret_cast<float4*>(&s.cdof[da + 1][0]) = make_float4(ax1w.x, ax1w.y, ax1w.z, 0.f);
    }
  }
  __syncwarp();

  // ------------------------------------------------------------------ com_pos (lane = body)
  V3 xi = v3(xp.x + R[0] * L.ipos.x + R[1] * L.ipos.y + R[2] * L.ipos.z,
             xp.y + R[3] * L.ipos.x + R[4] * L.ipos.y + R[5] * L.ipos.z,
             xp.z + R[6] * L.ipos.x + R[7] * L.ipos.y + R[8] * L.ipos.z);
  const float ms = L.mass;
  const float imt = L.imt;
  V3 com;
  {
    float mx[3] = {ms * xi.x, ms * xi.y, ms * xi.z};
    const float t = wfold<3>(mx, lane);
    com = v3(wfold_get(t, 0) * imt, wfold_get(t, 1) * imt, wfold_get(t, 2) * imt);
  }
  I10 cin;
  {
    // body-frame inertia about its COM (constant) rotated into the world, shifted to the common reference point
    const float b0 = m.b_Ib[0][lane], b1 = m.b_Ib[1][lane], b2 = m.b_Ib[2][lane], b3 = m.b_Ib[3][lane], b4 = m.b_Ib[4][lane], b5 = m.b_Ib[5][lane];
    float T[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      T[3 * r + 0] = R[3 * r] * b0 + R[3 * r + 1] * b3 + R[3 * r + 2] * b4;
      T[3 * r + 1] = R[3 * r] * b3 + R[3 * r + 1] * b1 + R[3 * r + 2] * b5;
      T[3 * r + 2] = R[3 * r] * b4 + R[3 * r + 1] * b5 + R[3 * r + 2] * b2;
    }
    const V3 d = xi - com;
    cin.xx = T[0] * R[0] + T[1] * R[1] + T[2] * R[2] + ms * (d.y * d.y + d.z * d.z);
    cin.yy = T[3] * R[3] + T[4] * R[4] + T[5] * R[5] + ms * (d.x * d.x + d.z * d.z);
    cin.zz = T[6] * R[6] + T[7] * R[7] + T[8] * R[8] + ms * (d.x * d.x + d.y * d.y);
    cin.xy = T[0] * R[3] + T[1] * R[4] + T[2] * R[5] - ms * d.x * d.y;
    cin.xz = T[0] * R[6] + T[1] * R[7] + T[2] * R[8] - ms * d.x * d.z;
    cin.yz = T[3] * R[6] + T[4] * R[7] + T[5] * R[8] - ms * d.y * d.z;
    cin.hx = ms * d.x; cin.hy = ms * d.y; cin.hz = ms * d.z; cin.m = ms;
  }

  // ------------------------------------------------------------------ cdof (lane = dof)
  const int flags = m.d_flags[lane];
  const int dbody = m.d_body[lane];
  S6 cd = s6zero();
  if (lane < nv) {
    const float4 c0 = lds4(&s.cdof[lane][0]);
    cd.a0 = c0.x; cd.a1 = c0.y; cd.a2 = c0.z;
    if (flags & DF_TRANS) {
      const float4 c1 = lds4(&s.cdof[lane][4]);
      cd.l0 = c0.w; cd.l1 = c1.x; cd.l2 = c1.y;
    } else {
      V3 off = com - v3(s.xpos[0][dbody], s.xpos[1][dbody], s.xpos[2][dbody]);
      V3 l = cross(v3(cd.a0, cd.a1, cd.a2), off);
      cd.l0 = l.x; cd.l1 = l.y; cd.l2 = l.z;
      s.cdof[lane][3] = l.x;
      *reinterpret_cast<float4*>(&s.cdof[lane][4]) = make_float4(l.y, l.z, 0.f, 0.f);
    }
  }
  __syncwarp();

  // ------------------------------------------------------------------ crb: subtree sums of cinert (lane = body)
  I10 crb = cin;
  if (m.scan_ok) {
    // Subtree sums without visiting every body: (1) suffix sums along the single-child chains by pointer jumping (log2 of the
    // longest chain rounds), (2) every branching body, deepest first, collects its children's totals and hands them to the
    // chain that ends in it.  63 shuffles for the duck (three chains under the trunk) instead of 170.
    int nx = m.b_next[lane];
    for (int r = 0; r < m.scan_rounds; ++r) {
      const int src = nx < 0 ? 0 : nx;
#define ACC(f) { const float v = __shfl_sync(FULLMASK, crb.f, src); if (nx >= 0) crb.f += v; }
      ACC(xx) ACC(yy) ACC(zz) ACC(xy) ACC(xz) ACC(yz) ACC(hx) ACC(hy) ACC(hz) ACC(m)
#undef ACC
      const int nn = __shfl_sync(FULLMASK, nx, src);
      if (nx >= 0) nx = nn;
    }
    for (int k = 0; k < m.n_branch; ++k) {
      I10 x;
      x.xx = x.yy = x.zz = x.xy = x.xz = x.yz = x.hx = x.hy = x.hz = x.m = 0.f;
      for (int c = 0; c < m.br_nchild[k]; ++c) {
        const int ch = m.br_child[k][c];
#define ACC(f) x.f += __shfl_sync(FULLMASK, crb.f, ch);
        ACC(xx) ACC(yy) ACC(zz) ACC(xy) ACC(xz) ACC(yz) ACC(hx) ACC(hy) ACC(hz) ACC(m)
#undef ACC
      }
      if ((m.br_chain[k] >> lane) & 1) {
#define ACC(f) crb.f += x.f;
        ACC(xx) ACC(yy) ACC(zz) ACC(xy) ACC(xz) ACC(yz) ACC(hx) ACC(hy) ACC(hz) ACC(m)
#undef ACC
      }
    }
  } else {
    crb.xx = crb.yy = crb.zz = crb.xy = crb.xz = crb.yz = crb.hx = crb.hy = crb.hz = crb.m = 0.f;
    const int sub = m.b_submask[lane];
    for (int c = 1; c < nb; ++c) {
      const bool in = (sub >> c) & 1;
#define ACC(f) { float v = __shfl_sync(FULLMASK, cin.f, c); if (in) crb.f += v; }
      ACC(xx) ACC(yy) ACC(zz) ACC(xy) ACC(xz) ACC(yz) ACC(hx) ACC(hy) ACC(hz) ACC(m)
#undef ACC
    }
  }
  // ------------------------------------------------------------------ M: one (dof, ancestor-or-self) pair per lane and wave
  {
    I10 cb;
#define GET(f) cb.f = __shfl_sync(FULLMASK, crb.f, dbody);
    GET(xx) GET(yy) GET(zz) GET(xy) GET(xz) GET(yz) GET(hx) GET(hy) GET(hz) GET(m)
#undef GET
    const S6 buf = inert_mul(cb, cd);                                   // crb[body_i] * cdof_i, staged for the pair pass
    float (*bf)[8] = reinterpret_cast<float (*)[8]>(s.H);           // staging records alias H (see WarpSmem)
    *reinterpret_cast<float4*>(&bf[lane][0]) = make_float4(buf.a0, buf.a1, buf.a2, buf.l0);
    *reinterpret_cast<float4*>(&bf[lane][4]) = make_float4(buf.l1, buf.l2, L.arm, 0.f);
    __syncwarp();
    // (unrolling this loop 2x / 4x changes nothing: 7.82 / 7.84 vs 7.85 M env-steps/s, profiles/r02t_bench_n1_mp*.json)
    for (int pp = lane; pp < m.n_mpairs; pp += 32) {
      const unsigned ij = m.mpair[pp];
      const int i = ij >> 8, j = ij & 255;
      const float4 c0 = lds4(&s.cdof[j][0]), c1 = lds4(&s.cdof[j][4]), b0 = lds4(&bf[i][0]), b1 = lds4(&bf[i][4]);
      float v = c0.x * b0.x + c0.y * b0.y + c0.z * b0.z + c0.w * b0.w + c1.x * b1.x + c1.y * b1.y;
      if (i == j) v += b1.z;
      s.A[TRI(i) + j] = v;
    }
  }
  __syncwarp();
  PHASE_SYNC(1, true)
  // ------------------------------------------------------------------ com_vel: prefix sums over the dof tree (lane = dof)
  const int dpar = m.d_parent[lane];
  S6 Sv = lane < nv ? s6scale(cd, L.qvel) : s6zero();
  {
    int P = dpar;
    for (int r = 0; r < m.prefix_rounds; ++r) {
      const int src = P < 0 ? 0 : P;
      S6 o = s6shfl(Sv, src);
      int Pn = __shfl_sync(FULLMASK, P, src);
      if (P >= 0) { Sv = s6add(Sv, o); P = Pn; }
    }
  }
  S6 cdd = s6zero();
  {
    const int vp = m.d_vparent[lane];
    S6 cvb = s6shfl(Sv, vp < 0 ? 0 : vp);
    if (vp < 0) cvb = s6zero();
    if (lane < nv && !(flags & DF_TRANS)) cdd = cross_motion(cvb, cd);
  }
  const int lastd = m.b_lastdof[lane];
  S6 cvel = s6shfl(Sv, lastd < 0 ? 0 : lastd);
  if (lastd < 0) cvel = s6zero();
  // ------------------------------------------------------------------ rne (bias force)
  S6 Sa = lane < nv ? s6scale(cdd, L.qvel) : s6zero();
  {
    int P = dpar;
    for (int r = 0; r < m.prefix_rounds; ++r) {
      const int src = P < 0 ? 0 : P;
      S6 o = s6shfl(Sa, src);
      int Pn = __shfl_sync(FULLMASK, P, src);
      if (P >= 0) { Sa = s6add(Sa, o); P = Pn; }
    }
  }
  float qfrc_bias = 0.f;
  {
    S6 cacc = s6shfl(Sa, lastd < 0 ? 0 : lastd);
    if (lastd < 0) cacc = s6zero();
    cacc.l0 -= m.gravity[0]; cacc.l1 -= m.gravity[1]; cacc.l2 -= m.gravity[2];
    S6 cfrc = s6add(inert_mul(cin, cacc), cross_force(cvel, inert_mul(cin, cvel)));
    if (m.scan_ok) {
      // subtree-summed force per body (same chain scan as the composite inertias), then one fetch per dof lane
      if (lane == 0 || lane >= nb) cfrc = s6zero();
      int nx = m.b_next[lane];
      for (int r = 0; r < m.scan_rounds; ++r) {
        const int src = nx < 0 ? 0 : nx;
        const S6 o = s6shfl(cfrc, src);
        const int nn = __shfl_sync(FULLMASK, nx, src);
        if (nx >= 0) { cfrc = s6add(cfrc, o); nx = nn; }
      }
      for (int k = 0; k < m.n_branch; ++k) {
        S6 x = s6zero();
        for (int c = 0; c < m.br_nchild[k]; ++c) x = s6add(x, s6shfl(cfrc, m.br_child[k][c]));
        if ((m.br_chain[k] >> lane) & 1) cfrc = s6add(cfrc, x);
      }
      const S6 f = s6shfl(cfrc, dbody);
      qfrc_bias = lane < nv ? s6dot(cd, f) : 0.f;
    } else {
      const int bsub = m.d_bsubmask[lane];
      for (int b = 1; b < nb; ++b) {
        S6 f = s6shfl(cfrc, b);
        if ((bsub >> b) & 1) qfrc_bias += s6dot(cd, f);
      }
    }
  }
  // ------------------------------------------------------------------ passive + actuation + smooth acceleration (lane = dof)
  const int qadr = m.d_qadr[lane];
  const float qd = (flags & DF_HINGE) ? s.qpos[qadr] : 0.f;
  float aforce = 0.f;
  if (m.d_act[lane] >= 0) {
    float c = fminf(fmaxf(L.ctrl, m.d_clo[lane]), m.d_chi[lane]);
    float f = L.kp * c - L.kp * qd - m.d_kv[lane] * L.qvel;
    aforce = fminf(fmaxf(f, m.d_flo[lane]), m.d_fhi[lane]);
  }
  const float fs = lane < nv ? (-m.d_damping[lane] * L.qvel - qfrc_bias + aforce) : 0.f;   // qfrc_smooth
  for (int idx = lane; 4 * idx < TRI(nv); idx += 32) reinterpret_cast<float4*>(s.H)[idx] = reinterpret_cast<const float4*>(s.A)[idx];   // float4 copy (both 16-byte aligned, 528 floats)
  s.rhs[lane] = fs;
  __syncwarp();
  chol_rev_tree(m, s.H, s.rhs, nv, lane, s.rowbuf);                                         // factor_m + L^-T qfrc_smooth
  const float as = back_tree(m, s.H, nv, lane, s.rhs[lane]);                                // qacc_smooth

  PHASE_SYNC(2, true)
  // ------------------------------------------------------------------ collision: plane (z = 0) vs convex foot hulls (lane = vertex)
  if constexpr (HF) {
    // height-field floor (HF instantiations only): terrain triangles vs foot faces, oduck_hfcollide.cuh
    hf_collide(m, ffm, hfm, s, lane, 0, hfs);
    hf_collide(m, ffm, hfm, s, lane, 1, hfs);
  } else
  for (int f = 0; f < 2; ++f) {
    const int fb = m.foot_body[f];
    float Rf[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rf[k] = s.xmat[k][fb];
    const V3 p0 = v3(s.xpos[0][fb], s.xpos[1][fb], s.xpos[2][fb]);
    const bool valid = lane < m.nvert;
    const V3 vl = v3(m.vert[f][0][lane], m.vert[f][1][lane], m.vert[f][2][lane]);
    const V3 pw = v3(p0.x + Rf[0] * vl.x + Rf[1] * vl.y + Rf[2] * vl.z, p0.y + Rf[3] * vl.x + Rf[4] * vl.y + Rf[5] * vl.z,
                     p0.z + Rf[6] * vl.x + Rf[7] * vl.y + Rf[8] * vl.z);
    const float support = -pw.z;
    const float ninf = -__int_as_float(0x7f800000);
    const float smax = wmaxf(valid ? support : ninf);
    const float thr = fmaxf(0.f, smax - 1e-3f);
    const bool mask = valid && support > thr;
    const float dm = valid ? (mask ? 0.f : -1e6f) : ninf;
    const V3 nl = v3(Rf[6], Rf[7], Rf[8]);  // plane normal in the hull frame
    int idx[4];
    idx[0] = wargmax(dm, lane);
    const V3 va = v3(__shfl_sync(FULLMASK, vl.x, idx[0]), __shfl_sync(FULLMASK, vl.y, idx[0]), __shfl_sync(FULLMASK, vl.z, idx[0]));
    const V3 ap = va - vl;
    idx[1] = wargmax(dot(ap, ap) + dm, lane);
    const V3 vb = v3(__shfl_sync(FULLMASK, vl.x, idx[1]), __shfl_sync(FULLMASK, vl.y, idx[1]), __shfl_sync(FULLMASK, vl.z, idx[1]));
    const V3 ab = cross(nl, va - vb);
    idx[2] = wargmax(fabsf(dot(ap, ab)) + dm, lane);
    const V3 vc = v3(__shfl_sync(FULLMASK, vl.x, idx[2]), __shfl_sync(FULLMASK, vl.y, idx[2]), __shfl_sync(FULLMASK, vl.z, idx[2]));
    const V3 ac = cross(nl, va - vc), bc = cross(nl, vb - vc), bp = vb - vl;
    int i1, i2;
    const float v1 = wargmax_val(fabsf(dot(bp, bc)) + dm, lane, &i1);
    const float v2 = wargmax_val(fabsf(dot(ap, ac)) + dm, lane, &i2);
    idx[3] = (v1 >= v2) ? i1 : i2;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float sup = __shfl_sync(FULLMASK, support, idx[c]);
      const float px = __shfl_sync(FULLMASK, pw.x, idx[c]), py = __shfl_sync(FULLMASK, pw.y, idx[c]), pz = __shfl_sync(FULLMASK, pw.z, idx[c]);
      bool uniq = true;
#pragma unroll
      for (int p = 0; p < c; ++p) uniq = uniq && (idx[p] != idx[c]);
      const float dist = uniq ? -sup : 1.f;
      if (lane == c) {
        float* cc = s.con[4 * f + c];
        cc[0] = dist; cc[1] = px; cc[2] = py; cc[3] = pz - 0.5f * dist;
      }
    }
  }
  if (lane >= NCON_FLOOR && lane < NCON_ALL) { s.con[lane][0] = 1.f; s.con[lane][1] = s.con[lane][2] = s.con[lane][3] = 0.f; }
  __syncwarp();
  // foot-foot convex-convex pair (rare: warp-uniform early exit on the bounding spheres); its 12 Jacobian rows live in HBM scratch
  float* ffJ = ffs;
  bool ffact = false;
  if (m.enable_ff) {
    if (FF) ffact = ff_collide(m, ffm, s, lane, com, cd, ffJ, ffs + FFJ_SIZE);
    else if (ff_maybe_close(m, ffm, s, lane)) return true;
  }

  // ------------------------------------------------------------------ foot-floor contact Jacobians, implicit
  // A floor contact c of foot f has the point Jacobian jp_i(c) = lin_i + ang_i x (pos_c - com) for the dofs i of the foot's
  // chain (cdof_i = [ang_i; lin_i]), rows (n, t1, t2) = (+z, +y, -x) . jp.  The rows are never stored: J x is the foot's
  // spatial velocity sum_i cdof_i x_i (one folded warp reduction for both feet) evaluated at the contact point, and J^T f /
  // J^T D J are accumulated per dof lane from the contact records (see the gradient + Hessian phase).
  const bool inF0 = (m.foot_chain[0] >> lane) & 1, inF1 = (m.foot_chain[1] >> lane) & 1;
  // ------------------------------------------------------------------ constraint rows
  // dof lane: friction-loss row and joint-limit row; contact lane c < 8: four pyramid rows
  const bool hasf = (flags & DF_FLOSS) != 0;
  const float Df = m.d_Dfric[lane];
  const float floss = L.floss;
  const float rf = hasf ? floss / Df : 0.f;
  const float areff = -m.sol_b * L.qvel;
  bool lact = false;
  float Dl = 0.f, arefl = 0.f, lsign = 0.f;
  if (flags & DF_LIMITED) {
    const float dmin_ = qd - m.d_lo[lane], dmax_ = m.d_hi[lane] - qd;
    const float pos = fminf(dmin_, dmax_);
    if (pos < 0.f) {
      lact = true;
      lsign = dmin_ < dmax_ ? 1.f : -1.f;
      const float imp = impedance(m, pos);
      Dl = 1.f / fmaxf(m.d_invw0[lane] * (1.f - imp) / imp, 1e-15f);
      arefl = -m.sol_b * (lsign * L.qvel) - m.sol_k * imp * pos;
    }
  }
  const int cl = lane < NCON_ALL ? lane : 0;          // contact lane: 0..7 foot-floor, 8..11 foot-foot
  const int cs = cl < NCON_FLOOR ? cl : 0, cf = cl >= NCON_FLOOR ? cl - NCON_FLOOR : 0;
  const float cdist = s.con[cl][0];
  const bool cact = lane < NCON_ALL && cdist < 0.f;
  const float mu = cl < NCON_FLOOR ? m.floor_mu : m.foot_mu;
  const V3 coff = v3(s.con[cs][1], s.con[cs][2], s.con[cs][3]) - com;     // contact lane: its point relative to the com
  // contact frame of this lane's floor contact: the flat floor has the fixed (n, t1, t2) = (+z, +y, -x) = make_frame(+z); a
  // height-field contact carries its triangle normal in the contact record
  V3 frn = v3(0.f, 0.f, 1.f), frt1 = v3(0.f, 1.f, 0.f), frt2 = v3(-1.f, 0.f, 0.f);
  if constexpr (HF) {
    float fr_[9];
    ff_make_frame(v3(s.con[cs][13], s.con[cs][14], s.con[cs][15]), fr_);
    frn = v3(fr_[0], fr_[1], fr_[2]); frt1 = v3(fr_[3], fr_[4], fr_[5]); frt2 = v3(fr_[6], fr_[7], fr_[8]);
  }
  // rows (n, t1, t2) of this lane's floor contact applied to the foot's spatial vector [a; l]
#define POINT_ROWS(a_, l_, pn_, p1_, p2_)                                                                               \
  {                                                                                                                    \
    const V3 jp_ = (l_) + cross((a_), coff);                                                                           \
    if constexpr (HF) { pn_ = dot(frn, jp_); p1_ = dot(frt1, jp_); p2_ = dot(frt2, jp_); }                              \
    else { pn_ = jp_.z; p1_ = jp_.y; p2_ = -jp_.x; }                                                                   \
  }
  float Dc = 0.f, arefc[4] = {0.f, 0.f, 0.f, 0.f};
  {
    // J qvel = velocity of the contact point: the foot body's cvel (com_vel prefix sums) at the point
    float pn, pt1, pt2;
    {
      const int fb = m.foot_body[cs >> 2];
      const S6 fv = s6shfl(cvel, fb);
      POINT_ROWS(v3(fv.a0, fv.a1, fv.a2), v3(fv.l0, fv.l1, fv.l2), pn, pt1, pt2)
    }
    if (FF && ffact) {
      const float pf = ffdot(ffJ, nv, lane, L.qvel);
      const float fn = __shfl_sync(FULLMASK, pf, 3 * cf), f1 = __shfl_sync(FULLMASK, pf, 3 * cf + 1), f2 = __shfl_sync(FULLMASK, pf, 3 * cf + 2);
      if (cl >= NCON_FLOOR) { pn = fn; pt1 = f1; pt2 = f2; }
    }
    if (cact) {
      const float imp = impedance(m, cdist);
      const float t = cl < NCON_FLOOR ? m.b_invw0[m.foot_body[cl >> 2]] : m.b_invw0[m.foot_body[0]] + m.b_invw0[m.foot_body[1]];
      const float invw = (t + mu * mu * t) * 2.f * mu * mu / m.impratio;
      Dc = 1.f / fmaxf(invw * (1.f - imp) / imp, 1e-15f);
      const float kp_ = m.sol_k * imp * cdist;
      arefc[0] = -m.sol_b * (pn + mu * pt1) - kp_;
      arefc[1] = -m.sol_b * (pn - mu * pt1) - kp_;
      arefc[2] = -m.sol_b * (pn + mu * pt2) - kp_;
      arefc[3] = -m.sol_b * (pn - mu * pt2) - kp_;
    }
  }
  // J x for the four pyramid rows of this lane's contact
#define CONTACT_PRODUCTS(x, o)                                                                                         \
  {                                                                                                                    \
    float sv_[12];                                                                                                     \
    {                                                                                                                  \
      const float x0_ = inF0 ? (x) : 0.f, x1_ = inF1 ? (x) : 0.f;                                                      \
      sv_[0] = cd.a0 * x0_; sv_[1] = cd.a1 * x0_; sv_[2] = cd.a2 * x0_; sv_[3] = cd.l0 * x0_; sv_[4] = cd.l1 * x0_; sv_[5] = cd.l2 * x0_; \
      sv_[6] = cd.a0 * x1_; sv_[7] = cd.a1 * x1_; sv_[8] = cd.a2 * x1_; sv_[9] = cd.l0 * x1_; sv_[10] = cd.l1 * x1_; sv_[11] = cd.l2 * x1_; \
    }                                                                                                                  \
    const float st_ = wfold<12>(sv_, lane);                                                                            \
    const int fo_ = 6 * (cs >> 2);                                                                                     \
    const V3 fa_ = v3(wfold_get(st_, fo_), wfold_get(st_, fo_ + 1), wfold_get(st_, fo_ + 2));                           \
    const V3 fl_ = v3(wfold_get(st_, fo_ + 3), wfold_get(st_, fo_ + 4), wfold_get(st_, fo_ + 5));                       \
    float pn_, p1_, p2_;                                                                                               \
    POINT_ROWS(fa_, fl_, pn_, p1_, p2_)                                                                                \
    if (FF && ffact) {                                                                                                 \
      const float pf_ = ffdot(ffJ, nv, lane, (x));                                                                     \
      const float fn_ = __shfl_sync(FULLMASK, pf_, 3 * cf), f1_ = __shfl_sync(FULLMASK, pf_, 3 * cf + 1), f2_ = __shfl_sync(FULLMASK, pf_, 3 * cf + 2); \
      if (cl >= NCON_FLOOR) { pn_ = fn_; p1_ = f1_; p2_ = f2_; }                                                       \
    }                                                                                                                  \
    o[0] = pn_ + mu * p1_; o[1] = pn_ - mu * p1_; o[2] = pn_ + mu * p2_; o[3] = pn_ - mu * p2_;                         \
  }
  // lane-local constraint cost for given Jaref values
#define ROW_COST(Jf_, Jl_, Jc_, acc)                                                                                   \
  {                                                                                                                    \
    if (hasf) {                                                                                                        \
      if ((Jf_) <= -rf) acc += floss * (-0.5f * rf - (Jf_));                                                           \
      else if ((Jf_) >= rf) acc += floss * (-0.5f * rf + (Jf_));                                                       \
      else acc += 0.5f * Df * (Jf_) * (Jf_);                                                                           \
    }                                                                                                                  \
    if (lact && (Jl_) < 0.f) acc += 0.5f * Dl * (Jl_) * (Jl_);                                                         \
    if (cact) {                                                                                                        \
      _Pragma("unroll") for (int r_ = 0; r_ < 4; ++r_) if (Jc_[r_] < 0.f) acc += 0.5f * Dc * Jc_[r_] * Jc_[r_];        \
    }                                                                                                                  \
  }

  // ------------------------------------------------------------------ solver.solve: warm start selection
  float qacc, Ma, Jf, Jl, Jc[4], gauss;
  {
    const float qw = lane < nv ? L.qaccw : 0.f;
    const float Maw = symv(s.A, nv, lane, qw, s.rowbuf);
    float Jcw[4], Jcs[4];
    CONTACT_PRODUCTS(qw, Jcw)
    CONTACT_PRODUCTS(as, Jcs)
#pragma unroll
    for (int r = 0; r < 4; ++r) { Jcw[r] -= arefc[r]; Jcs[r] -= arefc[r]; }
    const float Jfw = qw - areff, Jfs = as - areff;
    const float Jlw = lsign * qw - arefl, Jls = lsign * as - arefl;
    float gwl = 0.5f * (Maw - fs) * (qw - as), cw = 0.f, cs = 0.f;
    ROW_COST(Jfw, Jlw, Jcw, cw)
    ROW_COST(Jfs, Jls, Jcs, cs)
    float c3[3] = {gwl, cw, cs};
    const float ct = wfold<3>(c3, lane);
    const float gw = wfold_get(ct, 0), costw = gw + wfold_get(ct, 1), costs = wfold_get(ct, 2);
    const bool usew = costw < costs;
    qacc = usew ? qw : as;
    Ma = usew ? Maw : fs;
    Jf = usew ? Jfw : Jfs;
    Jl = usew ? Jlw : Jls;
#pragma unroll
    for (int r = 0; r < 4; ++r) Jc[r] = usew ? Jcw[r] : Jcs[r];
    gauss = usew ? gw : 0.f;
    if (DBG && lane == 0) { dbg[2536] = costw; dbg[2537] = costs; }
  }
  PHASE_SYNC(3, false)
  // ------------------------------------------------------------------ gradient + Hessian + Newton direction
  float search, grad;
  {
    // constraint forces at the start point (update_constraint)
    float ff = 0.f; bool fquad = false;
    if (hasf) {
      if (Jf <= -rf) ff = floss; else if (Jf >= rf) ff = -floss; else { ff = -Df * Jf; fquad = true; }
    }
    const bool lon = lact && Jl < 0.f;
    const float fl = lon ? -Dl * Jl : 0.f;
    float fc[4]; bool ca[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { ca[r] = cact && Jc[r] < 0.f; fc[r] = ca[r] ? -Dc * Jc[r] : 0.f; }
    {
      // contact record: force and Hessian weights of this lane's contact in its (n, t1, t2) frame
      const float a0 = ca[0], a1 = ca[1], a2 = ca[2], a3 = ca[3];
      if (lane < NCON_ALL) {
        float* cc = s.con[lane];
        cc[4] = fc[0] + fc[1] + fc[2] + fc[3];
        cc[5] = mu * (fc[0] - fc[1]);
        cc[6] = mu * (fc[2] - fc[3]);
        cc[8] = Dc * (a0 + a1 + a2 + a3);
        cc[9] = Dc * mu * (a0 - a1);
        cc[10] = Dc * mu * (a2 - a3);
        cc[11] = Dc * mu * mu * (a0 + a1);
        cc[12] = Dc * mu * mu * (a2 + a3);
      }
    }
    __syncwarp();
    // J^T f and z_i = sum_c T_c^T W_c jp_i(c) (a wrench per dof lane): H_ij += z_i . cdof_j for j ancestor-or-self of i
    float qfc = ff + lsign * fl;
    S6 z = s6zero();
    bool anyc = false;
    {
      const V3 ang = v3(cd.a0, cd.a1, cd.a2), lin = v3(cd.l0, cd.l1, cd.l2);
#pragma unroll 1
      for (int c = 0; c < NCON_FLOOR; ++c) {
        const float4 r0 = lds4(&s.con[c][0]);
        if (!(r0.x < 0.f)) continue;                                        // warp-uniform
        anyc = true;
        const float4 r1 = lds4(&s.con[c][4]), r2 = lds4(&s.con[c][8]);
        const float w22 = s.con[c][12];
        const V3 off = v3(r0.y, r0.z, r0.w) - com;
        V3 jp = lin + cross(ang, off);
        if (!(c < 4 ? inF0 : inF1)) jp = v3(0.f, 0.f, 0.f);
        float jn = jp.z, j1 = jp.y, j2 = -jp.x;
        V3 cn = v3(0.f, 0.f, 1.f), ct1 = v3(0.f, 1.f, 0.f), ct2 = v3(-1.f, 0.f, 0.f);
        if constexpr (HF) {
          float fr_[9];
          ff_make_frame(v3(s.con[c][13], s.con[c][14], s.con[c][15]), fr_);
          cn = v3(fr_[0], fr_[1], fr_[2]); ct1 = v3(fr_[3], fr_[4], fr_[5]); ct2 = v3(fr_[6], fr_[7], fr_[8]);
          jn = dot(cn, jp); j1 = dot(ct1, jp); j2 = dot(ct2, jp);
        }
        qfc += jn * r1.x + j1 * r1.y + j2 * r1.z;
        const float un = r2.x * jn + r2.y * j1 + r2.z * j2, u1 = r2.y * jn + r2.w * j1, u2 = r2.z * jn + w22 * j2;
        V3 U = v3(-u2, u1, un);                                             // back to world axes
        if constexpr (HF) U = un * cn + u1 * ct1 + u2 * ct2;
        const V3 oxU = cross(off, U);
        z.a0 += oxU.x; z.a1 += oxU.y; z.a2 += oxU.z; z.l0 += U.x; z.l1 += U.y; z.l2 += U.z;
      }
    }
    if (FF && ffact) {
      for (int c = 0; c < 4; ++c)
        if (s.con[NCON_FLOOR + c][0] < 0.f)
          qfc += ffJ[(3 * c) * 32 + lane] * s.con[NCON_FLOOR + c][4] + ffJ[(3 * c + 1) * 32 + lane] * s.con[NCON_FLOOR + c][5] + ffJ[(3 * c + 2) * 32 + lane] * s.con[NCON_FLOOR + c][6];
    }
    grad = lane < nv ? Ma - fs - qfc : 0.f;
    // H = M + J^T D J
    for (int idx = lane; 4 * idx < TRI(nv); idx += 32) reinterpret_cast<float4*>(s.H)[idx] = reinterpret_cast<const float4*>(s.A)[idx];   // float4 copy (both 16-byte aligned, 528 floats)
    __syncwarp();
    if (lane < nv) s.H[TRI(lane) + lane] += (fquad ? Df : 0.f) + (lon ? Dl : 0.f);
    if (anyc) {
      const int dep = lane < nv ? m.d_depth[lane] : -1;
      float* Hi = s.H + TRI(lane);
      if (m.plan_ok) {
        // chain plan (see chol_rev_back_chain): ancestor at level lev = lev on the root chain, chain start + lev - CH_NB below it; no
        // table, root-chain axes are warp-wide broadcasts, and the lane's own term uses the axis it already holds in registers
        const int cs = lane - (dep - CH_NB);
        const int nlev = m.max_dof_depth;                                 // levels of proper ancestors: 0 .. max_dof_depth - 1
        if (dep >= 0) Hi[lane] += z.a0 * cd.a0 + z.a1 * cd.a1 + z.a2 * cd.a2 + z.l0 * cd.l0 + z.l1 * cd.l1 + z.l2 * cd.l2;
#pragma unroll
        for (int lev = 0; lev < CH_NB + 10; ++lev) {
          if (lev >= nlev) break;                                        // warp-uniform
          if (lev < dep) {
            const int j = lev < CH_NB ? lev : cs + (lev - CH_NB);
            const float4 c0 = lds4(&s.cdof[j][0]), c1 = lds4(&s.cdof[j][4]);
            Hi[j] += z.a0 * c0.x + z.a1 * c0.y + z.a2 * c0.z + z.l0 * c0.w + z.l1 * c1.x + z.l2 * c1.y;
          }
        }
      } else {
        const unsigned char* al = m.anc[lane];
#pragma unroll 1
        for (int lev = 0; lev <= m.max_dof_depth; ++lev) {
          if (lev <= dep) {
            const int j = lev < dep ? al[lev] : lane;
            const float4 c0 = lds4(&s.cdof[j][0]), c1 = lds4(&s.cdof[j][4]);
            Hi[j] += z.a0 * c0.x + z.a1 * c0.y + z.a2 * c0.z + z.l0 * c0.w + z.l1 * c1.x + z.l2 * c1.y;
          }
        }
      }
    }
    if (FF && ffact) {
      // foot-foot rows couple the two legs: dense update (every i >= j), dense factorisation below
      float Zn[4], Z1[4], Z2[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float* cc = s.con[NCON_FLOOR + c];
        const float wnn = cc[8], wn1 = cc[9], wn2 = cc[10], w11 = cc[11], w22 = cc[12];
        const float jn = ffJ[(3 * c) * 32 + lane], j1 = ffJ[(3 * c + 1) * 32 + lane], j2 = ffJ[(3 * c + 2) * 32 + lane];
        Zn[c] = wnn * jn + wn1 * j1 + wn2 * j2;
        Z1[c] = wn1 * jn + w11 * j1;
        Z2[c] = wn2 * jn + w22 * j2;
      }
      for (int j = 0; j < nv; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc += Zn[c] * ffJ[(3 * c) * 32 + j] + Z1[c] * ffJ[(3 * c + 1) * 32 + j] + Z2[c] * ffJ[(3 * c + 2) * 32 + j];
        if (j <= lane && lane < nv) s.H[TRI(lane) + j] += acc;
      }
    }
    __syncwarp();
    if (DBG) {
      for (int j = 0; j <= lane && lane < nv; ++j) { dbg[4096 + lane * 32 + j] = s.H[TRI(lane) + j]; dbg[4096 + j * 32 + lane] = s.H[TRI(lane) + j]; }
    }
    s.rhs[lane] = grad;
    __syncwarp();
    if (FF && ffact) chol_rev(m, s.H, s.rhs, nv, lane, false);
    else chol_rev_tree(m, s.H, s.rhs, nv, lane, s.rowbuf);
    search = -((FF && ffact) ? chol_rev_back(m, s.H, nv, lane, s.rhs[lane], false) : back_tree(m, s.H, nv, lane, s.rhs[lane]));
    if (lane >= nv) search = 0.f;
  }
  PHASE_SYNC(4, false)
  // ------------------------------------------------------------------ line search (solver.py _linesearch)
  // Along qacc + alpha search the cost is piecewise quadratic; a trial point needs its slope d0 and curvature d1 (sums over
  // the rows that are active AT alpha).  The three trial points of an iteration are evaluated lane-locally and their six
  // sums folded in one butterfly (wfold); costs are only needed to choose between the final bracket ends and the start.
  {
    const float Mv = symv(s.A, nv, lane, search, s.rowbuf);
    float jvc[4];
    CONTACT_PRODUCTS(search, jvc)
    const float jvf = search, jvl = lsign * search;
    float snorm, qg1, qg2;
    {
      float g[3] = {search * search, search * (Ma - fs), search * Mv};
      const float t = wfold<3>(g, lane);
      snorm = sqrtf(wfold_get(t, 0)); qg1 = wfold_get(t, 1); qg2 = 0.5f * wfold_get(t, 2);
    }
    const float gtol = m.tolerance * m.ls_tolerance * snorm * m.meaninertia * (float)max(1, nv);
    const float qg0 = gauss;
    // per-row polynomial pieces (zero for rows this lane does not have)
    const float Dfe = hasf ? Df : 0.f;
    const float fq1 = Dfe * jvf * Jf, fq2 = Dfe * jvf * jvf, flin = hasf ? floss * jvf : 0.f;
    const float lq1 = Dl * jvl * Jl, lq2 = Dl * jvl * jvl;
    // Contact rows of the search.  The four pyramid rows of contact c live in lane c < 8 (12 with the foot-foot contacts), but a
    // trial point costs the WARP the same whether 8 lanes or 32 evaluate a row, so without foot-foot contacts (FF = false: 8
    // contacts = 32 rows) the rows are dealt out one per lane for the search: row r of contact c goes to lane 4 c + r (four
    // shuffles per quantity, once per substep), and every slope / cost evaluation handles one contact row instead of four
    // (17 slope evaluations per substep).  The foot-foot instantiation (rare path) keeps four rows in its contact lanes.
    constexpr int NCR = FF ? 4 : 1;
    float tJ[NCR], tjv[NCR], cq1[NCR], cq2[NCR], tD;
    if constexpr (FF) {
      tD = Dc;
#pragma unroll
      for (int r = 0; r < 4; ++r) { tJ[r] = Jc[r]; tjv[r] = jvc[r]; }
    } else {
      const int sc = lane >> 2, sr = lane & 3;
      const float j0 = __shfl_sync(FULLMASK, Jc[0], sc), j1 = __shfl_sync(FULLMASK, Jc[1], sc), j2 = __shfl_sync(FULLMASK, Jc[2], sc), j3 = __shfl_sync(FULLMASK, Jc[3], sc);
      const float v0 = __shfl_sync(FULLMASK, jvc[0], sc), v1 = __shfl_sync(FULLMASK, jvc[1], sc), v2 = __shfl_sync(FULLMASK, jvc[2], sc), v3_ = __shfl_sync(FULLMASK, jvc[3], sc);
      tJ[0] = sr == 0 ? j0 : (sr == 1 ? j1 : (sr == 2 ? j2 : j3));
      tjv[0] = sr == 0 ? v0 : (sr == 1 ? v1 : (sr == 2 ? v2 : v3_));
      tD = __shfl_sync(FULLMASK, Dc, sc);
    }
#pragma unroll
    for (int r = 0; r < NCR; ++r) { cq1[r] = tD * tjv[r] * tJ[r]; cq2[r] = tD * tjv[r] * tjv[r]; }
    // lane-local slope and curvature at alpha
    auto slope = [&](const float alpha, float& d0, float& d1) {
      const float x = Jf + alpha * jvf;
      const bool quad = hasf && x > -rf && x < rf;
      d0 = quad ? fmaf(alpha, fq2, fq1) : (x >= rf ? flin : -flin);
      d1 = quad ? fq2 : 0.f;
      if (Jl + alpha * jvl < 0.f) { d0 += fmaf(alpha, lq2, lq1); d1 += lq2; }
#pragma unroll
      for (int r = 0; r < NCR; ++r)
        if (tJ[r] + alpha * tjv[r] < 0.f) { d0 += fmaf(alpha, cq2[r], cq1[r]); d1 += cq2[r]; }
    };
    // lane-local cost at alpha
    auto cost = [&](const float alpha) {
      float q0 = 0.f, q1 = 0.f, q2 = 0.f;
      if (hasf) {
        const float x = Jf + alpha * jvf;
        if (x <= -rf) { q0 = floss * (-0.5f * rf - Jf); q1 = -flin; }
        else if (x >= rf) { q0 = floss * (-0.5f * rf + Jf); q1 = flin; }
        else { q0 = 0.5f * Df * Jf * Jf; q1 = fq1; q2 = 0.5f * fq2; }
      }
      if (Jl + alpha * jvl < 0.f) { q0 += 0.5f * Dl * Jl * Jl; q1 += lq1; q2 += 0.5f * lq2; }
#pragma unroll
      for (int r = 0; r < NCR; ++r)
        if (tJ[r] + alpha * tjv[r] < 0.f) { q0 += 0.5f * tD * tJ[r] * tJ[r]; q1 += cq1[r]; q2 += 0.5f * cq2[r]; }
      return fmaf(alpha, fmaf(alpha, q2, q1), q0);
    };
#define LS_FINISH(pt, sd0, sd1)                                                                                        \
  {                                                                                                                    \
    (pt).d0 = (sd0) + qg1 + 2.f * (pt).alpha * qg2;                                                                    \
    const float c_ = (sd1) + 2.f * qg2;                                                                                \
    (pt).d1 = c_ + (c_ == 0.f ? 1e-15f : 0.f);                                                                         \
  }
    LSPoint p0, lo, hi;
    {
      LSPoint ini[2];
      float a_ = 0.f;
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        float v[2];
        slope(a_, v[0], v[1]);
        const float t = wfold<2>(v, lane);
        ini[q].alpha = a_;
        LS_FINISH(ini[q], wfold_get(t, 0), wfold_get(t, 1))
        a_ = ini[0].alpha - ini[0].d0 / ini[0].d1;
      }
      p0 = ini[0]; lo = ini[1];
    }
    if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
    bool swap = true;
    int it = 0;
    while (true) {
      bool done = it >= m.ls_iterations;
      done |= !swap;
      done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
      done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
      if (done) break;
      LSPoint lo_next, mid, hi_next;
      lo_next.alpha = lo.alpha - lo.d0 / lo.d1; mid.alpha = 0.5f * (lo.alpha + hi.alpha); hi_next.alpha = hi.alpha - hi.d0 / hi.d1;
      {
        float v[6];
        slope(lo_next.alpha, v[0], v[1]);
        slope(mid.alpha, v[2], v[3]);
        slope(hi_next.alpha, v[4], v[5]);
        const float t = wfold<6>(v, lane);
        LS_FINISH(lo_next, wfold_get(t, 0), wfold_get(t, 1))
        LS_FINISH(mid, wfold_get(t, 2), wfold_get(t, 3))
        LS_FINISH(hi_next, wfold_get(t, 4), wfold_get(t, 5))
      }
      // Bracket update (oracle/oduck_oracle.cpp linesearch has the rationale): a candidate becomes the new lo if its slope is
      // negative and (lo sits on the wrong side of the root or the candidate is closer to it); symmetrically for hi.
      swap = false;
#define TRY_LO(c) if ((c).d0 < 0.f && (lo.d0 > 0.f || (c).d0 > lo.d0)) { lo = (c); swap = true; }
#define TRY_HI(c) if ((c).d0 >= 0.f && (hi.d0 < 0.f || (c).d0 < hi.d0)) { hi = (c); swap = true; }
      TRY_LO(lo_next) TRY_LO(mid) TRY_LO(hi_next)
      TRY_HI(hi_next) TRY_HI(mid) TRY_HI(lo_next)
#undef TRY_LO
#undef TRY_HI
      ++it;
    }
#undef LS_FINISH
    float alpha;
    {
      float v[3] = {cost(lo.alpha), cost(hi.alpha), cost(0.f)};
      const float t = wfold<3>(v, lane);
      const float clo = wfold_get(t, 0) + fmaf(lo.alpha, fmaf(lo.alpha, qg2, qg1), qg0);
      const float chi = wfold_get(t, 1) + fmaf(hi.alpha, fmaf(hi.alpha, qg2, qg1), qg0);
      const float c00 = wfold_get(t, 2) + qg0;
      const bool improved = (clo < c00) || (chi < c00);
      alpha = improved ? (clo < chi ? lo.alpha : hi.alpha) : 0.f;
    }
    qacc += alpha * search;
    Ma += alpha * Mv;
    Jf += alpha * jvf;
    Jl += alpha * jvl;
#pragma unroll
    for (int r = 0; r < 4; ++r) Jc[r] += alpha * jvc[r];
    if (DBG && lane == 0) { dbg[2538] = alpha; dbg[2539] = (float)it; }
  }
  if (lane >= nv) qacc = 0.f;
  L.qacc = qacc;
  L.qaccw = qacc;

  PHASE_SYNC(5, false)
  // ------------------------------------------------------------------ outputs of the last forward: forces, sensors
  if (last) {
    if (lane < nv) out[OUT_QACC + lane] = qacc;
    // efc_force rows: friction | limits | contacts x 4   (update_constraint after the line search)
    if (hasf) {
      float ff;
      if (Jf <= -rf) ff = floss; else if (Jf >= rf) ff = -floss; else ff = -Df * Jf;
      out[OUT_EFC + m.d_frrow[lane]] = ff;
    }
    if (flags & DF_LIMITED) out[OUT_EFC + m.nfr + m.d_limrow[lane]] = (lact && Jl < 0.f) ? -Dl * Jl : 0.f;
    if (lane < NCON_ALL) {
#pragma unroll
      for (int r = 0; r < 4; ++r) out[OUT_EFC + m.nfr + m.nlim + 4 * lane + r] = (cact && Jc[r] < 0.f) ? -Dc * Jc[r] : 0.f;
      out[OUT_CDIST + lane] = s.con[lane][0];
    }
    if (m.d_act[lane] >= 0) out[OUT_AFRC + m.d_act[lane]] = aforce;
    // sensors (sensor.py): all site-attached; cacc of the imu body needs the post-constraint qacc
    {
      const int ch = m.imu_chain;
      const bool in = lane < nv && ((ch >> lane) & 1);
      S6 t = s6add(s6scale(cdd, L.qvel), s6scale(cd, qacc));
      S6 ca;
      ca.a0 = wsum(in ? t.a0 : 0.f); ca.a1 = wsum(in ? t.a1 : 0.f); ca.a2 = wsum(in ? t.a2 : 0.f);
      ca.l0 = wsum(in ? t.l0 : 0.f) - m.gravity[0]; ca.l1 = wsum(in ? t.l1 : 0.f) - m.gravity[1]; ca.l2 = wsum(in ? t.l2 : 0.f) - m.gravity[2];
      const int ib = m.imu_body;
      const S6 cv = s6shfl(cvel, ib);
      float Rb[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) Rb[k] = s.xmat[k][ib];
      float Rs[9];  // site_xmat = R_body * R_site_local
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Rs[3 * r + c] = Rb[3 * r] * m.imu_rot[c] + Rb[3 * r + 1] * m.imu_rot[3 + c] + Rb[3 * r + 2] * m.imu_rot[6 + c];
      const V3 sp = v3(s.xpos[0][ib] + Rb[0] * m.imu_pos[0] + Rb[1] * m.imu_pos[1] + Rb[2] * m.imu_pos[2],
                       s.xpos[1][ib] + Rb[3] * m.imu_pos[0] + Rb[4] * m.imu_pos[1] + Rb[5] * m.imu_pos[2],
                       s.xpos[2][ib] + Rb[6] * m.imu_pos[0] + Rb[7] * m.imu_pos[1] + Rb[8] * m.imu_pos[2]);
      const V3 diff = sp - com;
      const V3 ang = v3(cv.a0, cv.a1, cv.a2);
      const V3 lin = v3(cv.l0, cv.l1, cv.l2) - cross(diff, ang);
      const V3 acc = v3(ca.l0, ca.l1, ca.l2) - cross(diff, v3(ca.a0, ca.a1, ca.a2));
      auto toLocal = [&](V3 v) { return v3(Rs[0] * v.x + Rs[3] * v.y + Rs[6] * v.z, Rs[1] * v.x + Rs[4] * v.y + Rs[7] * v.z, Rs[2] * v.x + Rs[5] * v.y + Rs[8] * v.z); };
      const V3 angl = toLocal(ang), linl = toLocal(lin), accl = toLocal(acc);
      const V3 corr = cross(angl, linl);
      if (lane == 0) {
        float* sd = out + OUT_SENS;
        sd[0] = angl.x; sd[1] = angl.y; sd[2] = angl.z;
        sd[3] = linl.x; sd[4] = linl.y; sd[5] = linl.z;
        sd[6] = accl.x + corr.x; sd[7] = accl.y + corr.y; sd[8] = accl.z + corr.z;
        sd[9] = Rs[2]; sd[10] = Rs[5]; sd[11] = Rs[8];
        sd[12] = ang.x; sd[13] = ang.y; sd[14] = ang.z;
        sd[21] = 0.f; sd[22] = 0.f; sd[23] = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) out[OUT_IMUMAT + k] = Rs[k];
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int fb = m.foot_site_body[k];
        const S6 fv = s6shfl(cvel, fb);
        const V3 fp = v3(s.xpos[0][fb] + s.xmat[0][fb] * m.foot_site_pos[k][0] + s.xmat[1][fb] * m.foot_site_pos[k][1] + s.xmat[2][fb] * m.foot_site_pos[k][2],
                         s.xpos[1][fb] + s.xmat[3][fb] * m.foot_site_pos[k][0] + s.xmat[4][fb] * m.foot_site_pos[k][1] + s.xmat[5][fb] * m.foot_site_pos[k][2],
                         s.xpos[2][fb] + s.xmat[6][fb] * m.foot_site_pos[k][0] + s.xmat[7][fb] * m.foot_site_pos[k][1] + s.xmat[8][fb] * m.foot_site_pos[k][2]);
        const V3 fl = v3(fv.l0, fv.l1, fv.l2) - cross(fp - com, v3(fv.a0, fv.a1, fv.a2));
        if (lane == 0) {
          out[OUT_SENS + 15 + 3 * k] = fl.x; out[OUT_SENS + 16 + 3 * k] = fl.y; out[OUT_SENS + 17 + 3 * k] = fl.z;
          out[OUT_FEET + 3 * k] = fp.x; out[OUT_FEET + 3 * k + 1] = fp.y; out[OUT_FEET + 3 * k + 2] = fp.z;
        }
      }
    }
  }
  if (DBG) {
    // dense M, smooth dynamics, contacts, rows, Newton direction (layout: tests/test_parity_gpu.py DBG_*)
    for (int j = 0; j <= lane && lane < nv; ++j) { dbg[lane * 32 + j] = s.A[TRI(lane) + j]; dbg[j * 32 + lane] = s.A[TRI(lane) + j]; }
    if (lane < nv) {
      dbg[1024 + lane] = qfrc_bias; dbg[1056 + lane] = fs; dbg[1088 + lane] = as;
      dbg[1184 + lane] = hasf ? Df : 0.f; dbg[1216 + lane] = Dl;
      dbg[1264 + lane] = hasf ? areff : 0.f; dbg[1296 + lane] = arefl;
      dbg[1376 + lane] = search; dbg[1408 + lane] = grad; dbg[1736 + lane] = qacc;
      dbg[1540 + lane * 6 + 0] = cd.a0; dbg[1540 + lane * 6 + 1] = cd.a1; dbg[1540 + lane * 6 + 2] = cd.a2;
      dbg[1540 + lane * 6 + 3] = cd.l0; dbg[1540 + lane * 6 + 4] = cd.l1; dbg[1540 + lane * 6 + 5] = cd.l2;
    }
    if (lane < NCON_ALL) {
      dbg[1120 + lane] = s.con[lane][0];
      dbg[1136 + 3 * lane] = s.con[lane][1]; dbg[1137 + 3 * lane] = s.con[lane][2]; dbg[1138 + 3 * lane] = s.con[lane][3];
      dbg[1248 + lane] = Dc;
      dbg[2560 + 3 * lane] = lane < NCON_FLOOR ? 0.f : ((FF && ffact) ? s.misc[0] : 1.f); dbg[2561 + 3 * lane] = lane < NCON_FLOOR ? 0.f : ((FF && ffact) ? s.misc[1] : 0.f); dbg[2562 + 3 * lane] = lane < NCON_FLOOR ? 1.f : ((FF && ffact) ? s.misc[2] : 0.f);
      if constexpr (HF) { if (lane < NCON_FLOOR) { dbg[2560 + 3 * lane] = s.con[lane][13]; dbg[2561 + 3 * lane] = s.con[lane][14]; dbg[2562 + 3 * lane] = s.con[lane][15]; } }
      for (int r = 0; r < 4; ++r) dbg[1328 + 4 * lane + r] = arefc[r];
    }
    if (lane < nb) { dbg[1440 + 3 * lane] = xp.x; dbg[1441 + 3 * lane] = xp.y; dbg[1442 + 3 * lane] = xp.z; }
    if (lane == 0) { dbg[1536] = com.x; dbg[1537] = com.y; dbg[1538] = com.z; }
    for (int c = 0; c < NCON_FLOOR; ++c) {
      V3 jp = v3(0.f, 0.f, 0.f);
      if (s.con[c][0] < 0.f && (c < 4 ? inF0 : inF1)) jp = v3(cd.l0, cd.l1, cd.l2) + cross(v3(cd.a0, cd.a1, cd.a2), v3(s.con[c][1], s.con[c][2], s.con[c][3]) - com);
      if constexpr (HF) {
        float fr_[9];
        ff_make_frame(v3(s.con[c][13], s.con[c][14], s.con[c][15]), fr_);
        dbg[1768 + (3 * c) * 32 + lane] = dot(v3(fr_[0], fr_[1], fr_[2]), jp); dbg[1768 + (3 * c + 1) * 32 + lane] = dot(v3(fr_[3], fr_[4], fr_[5]), jp);
        dbg[1768 + (3 * c + 2) * 32 + lane] = dot(v3(fr_[6], fr_[7], fr_[8]), jp);
      } else {
        dbg[1768 + (3 * c) * 32 + lane] = jp.z; dbg[1768 + (3 * c + 1) * 32 + lane] = jp.y; dbg[1768 + (3 * c + 2) * 32 + lane] = -jp.x;
      }
    }
  }
  __syncwarp();

  // ------------------------------------------------------------------ euler (eulerdamp disabled)
  if (integrate) {
    const float dt = m.timestep;
    if (lane < nv) L.qvel += dt * qacc;
    if (flags & DF_HINGE) s.qpos[qadr] = qd + dt * L.qvel;
    if (flags & DF_TRANS) s.qpos[qadr] += dt * L.qvel;   // qadr of a free-joint translation dof = its coordinate
    // free-joint quaternion: q <- normalize(q * exp(dt * w_local)); rotation dofs carry qadr = address of qw
    const int rl = __ffs(__ballot_sync(FULLMASK, (flags & DF_ROT) != 0)) - 1;  // first rotation dof lane
    if (rl >= 0) {
      const float wx = __shfl_sync(FULLMASK, L.qvel, rl), wy = __shfl_sync(FULLMASK, L.qvel, rl + 1), wz = __shfl_sync(FULLMASK, L.qvel, rl + 2);
      const int qa = __shfl_sync(FULLMASK, qadr, rl);
      __syncwarp();
      if (lane == 0) {
        const float nrm = sqrtf(wx * wx + wy * wy + wz * wz);
        V3 ax = v3(0.f, 0.f, 0.f);
        if (nrm > 0.f) ax = v3(wx / nrm, wy / nrm, wz / nrm);
        Q4 q; q.w = s.qpos[qa]; q.x = s.qpos[qa + 1]; q.y = s.qpos[qa + 2]; q.z = s.qpos[qa + 3];
        Q4 qn = qnormalize(qmul(q, axis_angle(ax, dt * nrm)));
        s.qpos[qa] = qn.w; s.qpos[qa + 1] = qn.x; s.qpos[qa + 2] = qn.y; s.qpos[qa + 3] = qn.z;
      }
    }
    __syncwarp();
  }
#undef CONTACT_PRODUCTS
#undef ROW_COST
  return false;
}

template <bool DBG, bool BAR = false, bool HF = false>
__device__ __forceinline__ void forward_euler(const DevModel& m, WarpSmem& s, Lane& L, const int lane, const bool last, const bool integrate,
                                              float* __restrict__ out, float* __restrict__ dbg, const DevFF* __restrict__ ffm, float* __restrict__ ffs,
                                              const DevHF* __restrict__ hfm = nullptr, float* __restrict__ hfs = nullptr) {
  if (forward_euler_impl<DBG, false, BAR, HF>(m, s, L, lane, last, integrate, out, dbg, ffm, ffs, hfm, hfs))
    forward_euler_impl<DBG, true, BAR, HF>(m, s, L, lane, last, integrate, out, dbg, ffm, ffs, hfm, hfs);
}

// per-env state record (HBM) <-> the warp's shared memory / registers
__device__ __forceinline__ void load_env(const DevModel& m, WarpSmem& s, Lane& L, int lane, const float* __restrict__ ph, const float* __restrict__ dr) {
  if (lane < m.nq) { s.qpos[lane] = ph[lane]; s.qpos0[lane] = dr[DR_QPOS0 + lane]; }
  if (lane + 32 < m.nq) { s.qpos[lane + 32] = ph[lane + 32]; s.qpos0[lane + 32] = dr[DR_QPOS0 + lane + 32]; }
  const bool isd = lane < m.nv;
  L.qvel = isd ? ph[PHYS_QVEL + lane] : 0.f;
  L.qaccw = isd ? ph[PHYS_QACCW + lane] : 0.f;
  L.qacc = 0.f;
  const int a = m.d_act[lane];
  L.ctrl = a >= 0 ? ph[PHYS_CTRL + a] : 0.f;
  L.kp = a >= 0 ? dr[DR_KP + a] : 0.f;
  L.floss = isd ? dr[DR_FLOSS + lane] : 0.f;
  L.arm = isd ? dr[DR_ARM + lane] : 0.f;
  L.mass = lane < m.nbody ? dr[lane] : 0.f;
  L.imt = 1.f / fmaxf(wsum(L.mass), 1e-15f);
  for (int i = lane; i < 528; i += 32) s.A[i] = 0.f;   // structural zeros of M (only ancestor pairs are rewritten each substep)
  L.ipos = v3(m.b_ipos[0][lane], m.b_ipos[1][lane], m.b_ipos[2][lane]);
  if (lane == 1) L.ipos = v3(dr[DR_IPOS1], dr[DR_IPOS1 + 1], dr[DR_IPOS1 + 2]);   // TORSO_BODY_ID = 1 (randomize.py:23)
  __syncwarp();
}
__device__ __forceinline__ void store_phys(const DevModel& m, const WarpSmem& s, const Lane& L, int lane, float* __restrict__ ph) {
  if (lane < m.nq) ph[lane] = s.qpos[lane];
  if (lane + 32 < m.nq) ph[lane + 32] = s.qpos[lane + 32];
  if (lane < m.nv) { ph[PHYS_QVEL + lane] = L.qvel; ph[PHYS_QACCW + lane] = L.qaccw; }
  const int a = m.d_act[lane];
  if (a >= 0) ph[PHYS_CTRL + a] = L.ctrl;
}
__device__ __forceinline__ void store_out(const WarpSmem& s, int lane, float* __restrict__ o) {
  for (int i = lane; i < OUT_STRIDE; i += 32) o[i] = s.outrec[i];
}
