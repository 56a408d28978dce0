// oduck_ffcollide.cuh -- convex-convex narrow phase for the two feet (rare path: their bounding spheres are apart in every
// nominal pose).  Same scheme as oracle/oduck_oracle.cpp convex_convex: SAT over polygon-face normals and Gauss-map-pruned
// edge pairs, reference-face clipping, <= 4 manifold points.  Warp-cooperative: lanes are vertices / planes / edge pairs
// (operands gathered with shuffles), the clipping of one <= 8-gon against one <= 8-gon is done by lane 0.
#pragma once
#include "oduck_device.cuh"

struct DevFF {   // stays in global memory (read only on the rare path)
  int nplane, nedge, nvert;
  int plane_nvert[32];
  int plane_vert[32][8];
  int edge_vert[48][2];
  int edge_plane[48][2];
  float plane_normal[2][3][NLANE];
  float center[2][3];
  float radius;
};

#define FFV_VA 0
#define FFV_VB 96
#define FFV_NA 192
#define FFV_NB 288
#define FFV_SIZE 384
#define FFJ_SIZE (12 * 32)

__device__ __forceinline__ V3 shfl3(V3 v, int src) { return v3(__shfl_sync(FULLMASK, v.x, src), __shfl_sync(FULLMASK, v.y, src), __shfl_sync(FULLMASK, v.z, src)); }
__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }

__device__ __forceinline__ void ff_make_frame(V3 n, float* f) {   // MJX math.make_frame
  const float nn = sqrtf(dot(n, n));
  const V3 a = (1.f / nn) * n;
  V3 b0 = (-0.5f < a.y && a.y < 0.5f) ? v3(0.f, 1.f, 0.f) : v3(0.f, 0.f, 1.f);
  V3 b = b0 - dot(a, b0) * a;
  b = (1.f / sqrtf(dot(b, b))) * b;
  const V3 c = cross(a, b);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = b.x; f[4] = b.y; f[5] = b.z; f[6] = c.x; f[7] = c.y; f[8] = c.z;
}

// _manifold_points on <= 18 points held by ONE thread (oracle manifold_points, scores rounded like fp32 MJX)
__device__ inline void ff_manifold(int n, const float (*poly)[3], const bool* mask, V3 nrm, int* idx) {
  auto dm = [&](int i) { return mask[i] ? 0.f : -1e6f; };
  auto P = [&](int i) { return v3(poly[i][0], poly[i][1], poly[i][2]); };
  int a = 0; float best = -__int_as_float(0x7f800000);
  for (int i = 0; i < n; i++) if (dm(i) > best) { best = dm(i); a = i; }
  int b = 0; best = -__int_as_float(0x7f800000);
  for (int i = 0; i < n; i++) { V3 d = P(a) - P(i); float v = dot(d, d) + dm(i); if (v > best) { best = v; b = i; } }
  const V3 ab = cross(nrm, P(a) - P(b));
  int c = 0; best = -__int_as_float(0x7f800000);
  for (int i = 0; i < n; i++) { float v = fabsf(dot(P(a) - P(i), ab)) + dm(i); if (v > best) { best = v; c = i; } }
  const V3 ac = cross(nrm, P(a) - P(c)), bc = cross(nrm, P(b) - P(c));
  int d = 0; best = -__int_as_float(0x7f800000);
  for (int i = 0; i < n; i++) { float v = fabsf(dot(P(b) - P(i), bc)) + dm(i); if (v > best) { best = v; d = i; } }
  for (int i = 0; i < n; i++) { float v = fabsf(dot(P(a) - P(i), ac)) + dm(i); if (v > best) { best = v; d = i; } }
  idx[0] = a; idx[1] = b; idx[2] = c; idx[3] = d;
}

// Broad phase of the common-path instantiation.  Conservative: returns false only when a separating axis PROVES that the
// feet do not touch -- first the bounding spheres, then the exact support gap of the two hulls along the centre-to-centre
// direction (lane = vertex, two warp reductions).  Side-by-side feet are rejected here; only real near-contacts restart.
__device__ __forceinline__ bool ff_maybe_close(const DevModel& m, const DevFF* __restrict__ ff, const WarpSmem& s, int lane) {
  V3 c[2], p[2];
  float R[2][9];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int b = m.foot_body[k];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[k][i] = s.xmat[i][b];
    p[k] = v3(s.xpos[0][b], s.xpos[1][b], s.xpos[2][b]);
    const V3 l = v3(ff->center[k][0], ff->center[k][1], ff->center[k][2]);
    c[k] = p[k] + v3(R[k][0] * l.x + R[k][1] * l.y + R[k][2] * l.z, R[k][3] * l.x + R[k][4] * l.y + R[k][5] * l.z, R[k][6] * l.x + R[k][7] * l.y + R[k][8] * l.z);
  }
  const V3 d = c[1] - c[0];
  const float dist = sqrtf(dot(d, d));
  if (dist > 2.f * ff->radius) return false;
  const V3 n = (1.f / fmaxf(dist, 1e-9f)) * d;
  const bool valid = lane < m.nvert;
  const int vl = valid ? lane : 0;
  float pr[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const V3 v = v3(m.vert[k][0][vl], m.vert[k][1][vl], m.vert[k][2][vl]);
    const V3 w = p[k] + v3(R[k][0] * v.x + R[k][1] * v.y + R[k][2] * v.z, R[k][3] * v.x + R[k][4] * v.y + R[k][5] * v.z, R[k][6] * v.x + R[k][7] * v.y + R[k][8] * v.z);
    pr[k] = dot(n, w);
  }
  const float maxA = wmaxf(valid ? pr[0] : -__int_as_float(0x7f800000));
  const float minB = -wmaxf(valid ? -pr[1] : -__int_as_float(0x7f800000));
  return !(minB - maxA > 0.f);
}

// Returns true iff one of the four foot-foot contacts (slots 8..11 of s.con) is active; then s.misc[0..8] = contact frame and
// ffJ[12][32] = the frame-rotated relative point Jacobians (row 3c+a, column = dof).
static __device__ __noinline__ bool ff_collide(const DevModel& m, const DevFF* __restrict__ ff, WarpSmem& s, const int lane, const V3 com, const S6 cd,
                                               float* __restrict__ ffJ, float* __restrict__ ffv) {
  const int bA = m.foot_body[0], bB = m.foot_body[1];
  float RA[9], RB[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { RA[k] = s.xmat[k][bA]; RB[k] = s.xmat[k][bB]; }
  const V3 pA = v3(s.xpos[0][bA], s.xpos[1][bA], s.xpos[2][bA]), pB = v3(s.xpos[0][bB], s.xpos[1][bB], s.xpos[2][bB]);
  auto rotA = [&](V3 v) { return v3(RA[0] * v.x + RA[1] * v.y + RA[2] * v.z, RA[3] * v.x + RA[4] * v.y + RA[5] * v.z, RA[6] * v.x + RA[7] * v.y + RA[8] * v.z); };
  auto rotB = [&](V3 v) { return v3(RB[0] * v.x + RB[1] * v.y + RB[2] * v.z, RB[3] * v.x + RB[4] * v.y + RB[5] * v.z, RB[6] * v.x + RB[7] * v.y + RB[8] * v.z); };
  const V3 cA = pA + rotA(v3(ff->center[0][0], ff->center[0][1], ff->center[0][2])), cB = pB + rotB(v3(ff->center[1][0], ff->center[1][1], ff->center[1][2]));
  {
    const V3 dc = cB - cA;
    if (sqrtf(dot(dc, dc)) > 2.f * ff->radius) return false;                       // bounding spheres apart (the common case)
  }
  const int nvt = ff->nvert, npl = ff->nplane, ne = ff->nedge;
  const float ninf = -__int_as_float(0x7f800000);
  const int vl = lane < nvt ? lane : 0, ql = lane < npl ? lane : 0;
  const V3 VA = pA + rotA(v3(m.vert[0][0][vl], m.vert[0][1][vl], m.vert[0][2][vl])), VB = pB + rotB(v3(m.vert[1][0][vl], m.vert[1][1][vl], m.vert[1][2][vl]));
  const V3 nA = rotA(v3(ff->plane_normal[0][0][ql], ff->plane_normal[0][1][ql], ff->plane_normal[0][2][ql]));
  const V3 nB = rotB(v3(ff->plane_normal[1][0][ql], ff->plane_normal[1][1][ql], ff->plane_normal[1][2][ql]));
  // ---- face axes (lane = polygon face)
  float sepA, sepB;
  {
    const int pv0 = ff->plane_vert[ql][0];
    const float offA = dot(nA, shfl3(VA, pv0)), offB = dot(nB, shfl3(VB, pv0));
    float mnA = __int_as_float(0x7f800000), mnB = mnA;
    for (int j = 0; j < nvt; ++j) { mnA = fminf(mnA, dot(nA, shfl3(VB, j))); mnB = fminf(mnB, dot(nB, shfl3(VA, j))); }
    sepA = lane < npl ? mnA - offA : ninf;
    sepB = lane < npl ? mnB - offB : ninf;
  }
  int iA, iB;
  const float vA = wargmax_val(sepA, lane, &iA), vB = wargmax_val(sepB, lane, &iB);
  const int face_hull = vB > vA ? 1 : 0, face_idx = face_hull ? iB : iA;
  const float face_sep = face_hull ? vB : vA;
  if (face_sep > 0.f) return false;
  // ---- edge axes (lane = edge pair, ne * ne of them)
  float esep = ninf; int epair = 0x7fffffff; V3 en = v3(0.f, 0.f, 1.f);
  const int npairs = ne * ne;
  for (int p0 = 0; p0 < npairs; p0 += 32) {
    const int p = p0 + lane;
    const bool valid = p < npairs;
    const int ea = valid ? p / ne : 0, eb = valid ? p % ne : 0;
    const V3 a = shfl3(nA, ff->edge_plane[ea][0]), b = shfl3(nA, ff->edge_plane[ea][1]);
    const V3 c = -1.f * shfl3(nB, ff->edge_plane[eb][0]), d = -1.f * shfl3(nB, ff->edge_plane[eb][1]);
    const V3 pa0 = shfl3(VA, ff->edge_vert[ea][0]), pa1 = shfl3(VA, ff->edge_vert[ea][1]);
    const V3 pb0 = shfl3(VB, ff->edge_vert[eb][0]), pb1 = shfl3(VB, ff->edge_vert[eb][1]);
    const V3 bxa = cross(b, a), dxc = cross(d, c), dA = pa1 - pa0, dB = pb1 - pb0;
    const float cba = dot(c, bxa), dba = dot(d, bxa), adc = dot(a, dxc), bdc = dot(b, dxc);
    V3 n = cross(dA, dB);
    const float len2 = dot(n, n);
    const bool ok = valid && (cba * dba < 0.f && adc * bdc < 0.f && cba * bdc > 0.f) && !(len2 < 1e-10f * dot(dA, dA) * dot(dB, dB));
    n = rsqrtf(fmaxf(len2, 1e-30f)) * n;
    if (dot(n, pa0 - cA) < 0.f) n = -1.f * n;
    const float sep = dot(n, pb0 - pa0);
    if (ok && sep > esep) { esep = sep; epair = p; en = n; }
  }
  {
    // warp argmax with first-pair tie break
    float v = esep; int pi = epair, src = lane;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(FULLMASK, v, o);
      const int opi = __shfl_xor_sync(FULLMASK, pi, o), osrc = __shfl_xor_sync(FULLMASK, src, o);
      const bool take = (ov > v) || (ov == v && opi < pi);
      v = take ? ov : v; pi = take ? opi : pi; src = take ? osrc : src;
    }
    esep = v; epair = pi; en = shfl3(en, src);
  }
  if (esep > 0.f) return false;
  // ---- manifold (lane 0), operands through the scratch
  if (lane < nvt) { ffv[FFV_VA + 3 * lane] = VA.x; ffv[FFV_VA + 3 * lane + 1] = VA.y; ffv[FFV_VA + 3 * lane + 2] = VA.z;
                    ffv[FFV_VB + 3 * lane] = VB.x; ffv[FFV_VB + 3 * lane + 1] = VB.y; ffv[FFV_VB + 3 * lane + 2] = VB.z; }
  if (lane < npl) { ffv[FFV_NA + 3 * lane] = nA.x; ffv[FFV_NA + 3 * lane + 1] = nA.y; ffv[FFV_NA + 3 * lane + 2] = nA.z;
                    ffv[FFV_NB + 3 * lane] = nB.x; ffv[FFV_NB + 3 * lane + 1] = nB.y; ffv[FFV_NB + 3 * lane + 2] = nB.z; }
  __syncwarp();
  if (lane == 0) {
    for (int c = 8; c < 12; ++c) { s.con[c][0] = 1.f; s.con[c][1] = s.con[c][2] = s.con[c][3] = 0.f; }
    if (epair != 0x7fffffff && esep > face_sep + 1e-4f) {
      const int ea = epair / ne, eb = epair % ne;
      const V3 p1 = ld3(ffv + FFV_VA + 3 * ff->edge_vert[ea][0]), q1 = ld3(ffv + FFV_VA + 3 * ff->edge_vert[ea][1]);
      const V3 p2 = ld3(ffv + FFV_VB + 3 * ff->edge_vert[eb][0]), q2 = ld3(ffv + FFV_VB + 3 * ff->edge_vert[eb][1]);
      const V3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
      const float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2), den = a * e - b * b;
      float sA = den > 1e-18f ? fminf(fmaxf((b * f - c * e) / den, 0.f), 1.f) : 0.f;
      const float tB = fminf(fmaxf((b * sA + f) / e, 0.f), 1.f);
      sA = fminf(fmaxf((b * tB - c) / a, 0.f), 1.f);
      const V3 pos = 0.5f * ((p1 + sA * d1) + (p2 + tB * d2));
      s.con[8][0] = esep; s.con[8][1] = pos.x; s.con[8][2] = pos.y; s.con[8][3] = pos.z;
      ff_make_frame(en, s.misc);
    } else {
      const float* Vr = ffv + (face_hull ? FFV_VB : FFV_VA);
      const float* Vi = ffv + (face_hull ? FFV_VA : FFV_VB);
      const float* Ni = ffv + (face_hull ? FFV_NA : FFV_NB);
      const V3 nref = ld3(ffv + (face_hull ? FFV_NB : FFV_NA) + 3 * face_idx);
      int inc = 0; float best = __int_as_float(0x7f800000);
      for (int q = 0; q < npl; ++q) { const float v = dot(ld3(Ni + 3 * q), nref); if (v < best) { best = v; inc = q; } }
      float poly[2][18][3];
      int cur = 0, cnt = ff->plane_nvert[inc];
      for (int k = 0; k < cnt; ++k) { const V3 x = ld3(Vi + 3 * ff->plane_vert[inc][k]); poly[0][k][0] = x.x; poly[0][k][1] = x.y; poly[0][k][2] = x.z; }
      const int nr = ff->plane_nvert[face_idx];
      for (int k = 0; k < nr && cnt > 0; ++k) {
        const V3 r0 = ld3(Vr + 3 * ff->plane_vert[face_idx][k]), r1 = ld3(Vr + 3 * ff->plane_vert[face_idx][(k + 1) % nr]);
        const V3 sd = cross(r1 - r0, nref);
        int no = 0;
        for (int v = 0; v < cnt; ++v) {
          const V3 x0 = v3(poly[cur][v][0], poly[cur][v][1], poly[cur][v][2]);
          const int v1 = (v + 1) % cnt;
          const V3 x1 = v3(poly[cur][v1][0], poly[cur][v1][1], poly[cur][v1][2]);
          const float d0 = dot(sd, x0 - r0), d1 = dot(sd, x1 - r0);
          if (d0 <= 0.f) { poly[1 - cur][no][0] = x0.x; poly[1 - cur][no][1] = x0.y; poly[1 - cur][no][2] = x0.z; no++; }
          if ((d0 <= 0.f) != (d1 <= 0.f)) { const float t = d0 / (d0 - d1); const V3 xi = x0 + t * (x1 - x0); poly[1 - cur][no][0] = xi.x; poly[1 - cur][no][1] = xi.y; poly[1 - cur][no][2] = xi.z; no++; }
        }
        cur = 1 - cur; cnt = no;
      }
      if (cnt > 0) {
        float dist[18]; bool mask[18];
        const V3 r0 = ld3(Vr + 3 * ff->plane_vert[face_idx][0]);
        for (int v = 0; v < cnt; ++v) { dist[v] = dot(nref, v3(poly[cur][v][0], poly[cur][v][1], poly[cur][v][2]) - r0); mask[v] = dist[v] < 0.f; }
        int idx[4] = {0, 1, 2, 3};
        if (cnt > 4) ff_manifold(cnt, poly[cur], mask, nref, idx);
        ff_make_frame(face_hull ? -1.f * nref : nref, s.misc);
        for (int c = 0; c < 4; ++c) {
          const int v = idx[c];
          bool ok = v < cnt;
          for (int p2 = 0; p2 < c && ok; ++p2) ok = idx[p2] != v;
          if (!ok) continue;
          s.con[8 + c][0] = dist[v];
          s.con[8 + c][1] = poly[cur][v][0] - 0.5f * dist[v] * nref.x;
          s.con[8 + c][2] = poly[cur][v][1] - 0.5f * dist[v] * nref.y;
          s.con[8 + c][3] = poly[cur][v][2] - 0.5f * dist[v] * nref.z;
        }
      }
    }
  }
  __syncwarp();
  const bool any = s.con[8][0] < 0.f || s.con[9][0] < 0.f || s.con[10][0] < 0.f || s.con[11][0] < 0.f;
  if (!any) return false;
  // ---- relative point Jacobians (lane = dof): jacp(right foot) - jacp(left foot), rotated into the contact frame
  const float sg = (float)((m.foot_chain[1] >> lane) & 1) - (float)((m.foot_chain[0] >> lane) & 1);
  for (int c = 0; c < 4; ++c) {
    float j0 = 0.f, j1 = 0.f, j2 = 0.f;
    if (s.con[8 + c][0] < 0.f && sg != 0.f) {
      const V3 off = v3(s.con[8 + c][1], s.con[8 + c][2], s.con[8 + c][3]) - com;
      const V3 jp = sg * (v3(cd.l0, cd.l1, cd.l2) + cross(v3(cd.a0, cd.a1, cd.a2), off));
      j0 = s.misc[0] * jp.x + s.misc[1] * jp.y + s.misc[2] * jp.z;
      j1 = s.misc[3] * jp.x + s.misc[4] * jp.y + s.misc[5] * jp.z;
      j2 = s.misc[6] * jp.x + s.misc[7] * jp.y + s.misc[8] * jp.z;
    }
    ffJ[(3 * c) * 32 + lane] = j0; ffJ[(3 * c + 1) * 32 + lane] = j1; ffJ[(3 * c + 2) * 32 + lane] = j2;
  }
  __syncwarp();
  return true;
}
// lane r < 12 gets ffJ[r] . x
static __device__ __noinline__ float ffdot(const float* __restrict__ ffJ, int n, int lane, float x) {
  float acc = 0.f;
  const float* row = ffJ + (lane < 12 ? lane : 0) * 32;
  for (int d = 0; d < n; ++d) acc = fmaf(row[d], __shfl_sync(FULLMASK, x, d), acc);
  return lane < 12 ? acc : 0.f;
}
