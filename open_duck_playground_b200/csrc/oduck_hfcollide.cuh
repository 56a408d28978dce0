// oduck_hfcollide.cuh -- height-field floor vs one convex foot hull (rough_terrain scenes; only the HF = true kernel
// instantiations contain this code, the flat-floor instantiations are unchanged).  Same scheme as
// oracle/oduck_oracle.cpp hfield_convex: for every terrain triangle under the foot's bounding sphere, every foot face that
// looks down onto it is clipped to the triangle's vertical column; clipped points below the triangle plane are candidates
// (dist along the triangle normal, pos midway, normal = triangle normal); 4 of the candidates within 1 mm of the deepest one
// are kept by _manifold_points with the mean normal.  Warp mapping: lane = hull face for the clipping (each lane clips its
// own <= 8-gon in local memory), candidates are appended in the oracle's order (triangle, face, polygon vertex) to a per-env
// HBM scratch list with a warp prefix sum, and the manifold selection runs over that list 32 candidates at a time with the
// same first-index arg-max as the plane collider.
#pragma once
#include "oduck_ffcollide.cuh"

struct DevHF {   // global memory
  int nrow, ncol;
  float sx, sy, sz, dx, dy;     // radii, elevation scale, grid pitch
  const float* data;            // [nrow][ncol] elevation in [0, 1]
};

#define HF_CAP 512             // candidates kept per foot (the oracle keeps all; a resting foot has a few dozen)
#define HF_REC 8               // dist, pos[3], normal[3], -
#ifdef ODUCK_HF_PAIRS
#define HF_MAXT 32             // triangles per batch of the pair list
#define HF_MAXPAIR 256         // (triangle, face) pairs per batch
#define HF_SCRATCH (HF_CAP * HF_REC + 96 + HF_MAXT * 12 + HF_MAXPAIR)   // + world vertices, triangle table, pair list
#else
#define HF_SCRATCH (HF_CAP * HF_REC)
#endif
#define HF_MAXP 12             // 8-gon clipped by 3 planes: at most 11 vertices

// Sutherland-Hodgman: keep the part of the polygon on the inner side (d <= 0) of the vertical plane through r0 -> r1
__device__ __forceinline__ int hf_clip(const float (*in)[3], int cnt, float (*out)[3], const float r0x, const float r0y, const float sdx, const float sdy) {
  int no = 0;
  for (int v = 0; v < cnt; ++v) {
    const int w = v + 1 == cnt ? 0 : v + 1;
    const float d0 = sdx * (in[v][0] - r0x) + sdy * (in[v][1] - r0y), d1 = sdx * (in[w][0] - r0x) + sdy * (in[w][1] - r0y);
    if (d0 <= 0.f) { out[no][0] = in[v][0]; out[no][1] = in[v][1]; out[no][2] = in[v][2]; ++no; }
    if ((d0 <= 0.f) != (d1 <= 0.f)) {
      const float t = d0 / (d0 - d1);
      out[no][0] = in[v][0] + t * (in[w][0] - in[v][0]); out[no][1] = in[v][1] + t * (in[w][1] - in[v][1]); out[no][2] = in[v][2] + t * (in[w][2] - in[v][2]);
      ++no;
    }
  }
  return no;
}

// Writes the four contact records of foot f: s.con[4 f + c][0] = dist (1: inactive), [1..3] = pos, [13..15] = normal.
static __device__ __noinline__ void hf_collide(const DevModel& m, const DevFF* __restrict__ ff, const DevHF* __restrict__ hf, WarpSmem& s,
                                               const int lane, const int f, float* __restrict__ cand) {
  const int fb = m.foot_body[f];
  float R[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = s.xmat[k][fb];
  const V3 p0 = v3(s.xpos[0][fb], s.xpos[1][fb], s.xpos[2][fb]);
  auto rot = [&](V3 v) { return v3(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z); };
  if (lane < 4) {
    float* cc = s.con[4 * f + lane];
    cc[0] = 1.f; cc[1] = cc[2] = cc[3] = 0.f; cc[13] = 0.f; cc[14] = 0.f; cc[15] = 1.f;
  }
  __syncwarp();
  const V3 C = p0 + rot(v3(ff->center[f][0], ff->center[f][1], ff->center[f][2]));
  const float rb = ff->radius;
  const int nrow = hf->nrow, ncol = hf->ncol;
  const float sx = hf->sx, sy = hf->sy, sz = hf->sz, dx = hf->dx, dy = hf->dy;
  const float* __restrict__ data = hf->data;
  int cmin = (int)floorf((C.x - rb + sx) / dx), cmax = (int)floorf((C.x + rb + sx) / dx);
  int rmin = (int)floorf((C.y - rb + sy) / dy), rmax = (int)floorf((C.y + rb + sy) / dy);
  cmin = max(cmin, 0); rmin = max(rmin, 0); cmax = min(cmax, ncol - 2); rmax = min(rmax, nrow - 2);
  // this lane's face: world normal and world polygon
  const bool has = lane < ff->nplane;
  const int q = has ? lane : 0;
  const V3 Nw = rot(v3(ff->plane_normal[f][0][q], ff->plane_normal[f][1][q], ff->plane_normal[f][2][q]));
  const int cnt0 = has ? ff->plane_nvert[q] : 0;
  float P0[8][3];
  for (int v = 0; v < cnt0; ++v) {
    const int vid = ff->plane_vert[q][v];
    const V3 w = p0 + rot(v3(m.vert[f][0][vid], m.vert[f][1][vid], m.vert[f][2][vid]));
    P0[v][0] = w.x; P0[v][1] = w.y; P0[v][2] = w.z;
  }
#ifdef ODUCK_HF_CULL
  // Result-preserving culls (measured next round; off by default so that the verified build is unchanged): the world box of the
  // hull (lane = vertex, redux min / max) rejects cells it does not overlap and triangles that lie wholly below its lowest
  // vertex (a candidate needs a hull point BELOW the triangle plane, whose height never exceeds the triangle's top); the box
  // of this lane's face rejects the clipping.  1e-6 m margins keep the culls conservative under fp32 rounding of the clipped points.
  float bx0, bx1, by0, by1, bz0;
  {
    const bool vv = lane < m.nvert;
    const int vl = vv ? lane : 0;
    const V3 w = p0 + rot(v3(m.vert[f][0][vl], m.vert[f][1][vl], m.vert[f][2][vl]));
    const float inf = __int_as_float(0x7f800000);
    bx1 = wmaxf(vv ? w.x : -inf) + 1e-6f; bx0 = -wmaxf(vv ? -w.x : -inf) - 1e-6f;
    by1 = wmaxf(vv ? w.y : -inf) + 1e-6f; by0 = -wmaxf(vv ? -w.y : -inf) - 1e-6f;
    bz0 = -wmaxf(vv ? -w.z : -inf) - 1e-6f;
  }
  float fx0 = 0.f, fx1 = 0.f, fy0 = 0.f, fy1 = 0.f;
  if (cnt0 > 0) {
    fx0 = fx1 = P0[0][0]; fy0 = fy1 = P0[0][1];
    for (int v = 1; v < cnt0; ++v) { fx0 = fminf(fx0, P0[v][0]); fx1 = fmaxf(fx1, P0[v][0]); fy0 = fminf(fy0, P0[v][1]); fy1 = fmaxf(fy1, P0[v][1]); }
    fx0 -= 1e-6f; fx1 += 1e-6f; fy0 -= 1e-6f; fy1 += 1e-6f;
  }
#endif
  int nc = 0;                      // candidates so far (warp-uniform)
  V3 nsum = v3(0.f, 0.f, 0.f);     // sum of the candidates' normals (warp-uniform)
  float deep = 0.f;                // lane-local deepest candidate
#ifdef ODUCK_HF_PAIRS
  // Variant for an A/B run (tools/variants.py): instead of every lane clipping ITS face against every triangle in turn (a
  // handful of lanes busy, ~18 sequential rounds), the (triangle, face) pairs that survive the box culls are listed first and
  // then clipped 32 pairs at a time, lane = pair.  The candidate order (triangle, face, polygon vertex) is unchanged.
  float* wv = cand + HF_CAP * HF_REC;                                   // world vertices [32][3]
  float* wt = wv + 96;                                                   // triangles [HF_MAXT][12]: T0, T1, T2, n
  int* wp = reinterpret_cast<int*>(wt + HF_MAXT * 12);                   // pairs [HF_MAXPAIR]: triangle << 5 | face
  float hx0, hx1, hy0, hy1, hz0;                                          // box of the hull, 1e-6 m margins
  {
    const bool vv = lane < m.nvert;
    const int vl = vv ? lane : 0;
    const V3 w = p0 + rot(v3(m.vert[f][0][vl], m.vert[f][1][vl], m.vert[f][2][vl]));
    if (vv) { wv[3 * lane] = w.x; wv[3 * lane + 1] = w.y; wv[3 * lane + 2] = w.z; }
    const float inf = __int_as_float(0x7f800000);
    hx1 = wmaxf(vv ? w.x : -inf) + 1e-6f; hx0 = -wmaxf(vv ? -w.x : -inf) - 1e-6f;
    hy1 = wmaxf(vv ? w.y : -inf) + 1e-6f; hy0 = -wmaxf(vv ? -w.y : -inf) - 1e-6f;
    hz0 = -wmaxf(vv ? -w.z : -inf) - 1e-6f;
  }
  __syncwarp();
  float gx0 = 0.f, gx1 = 0.f, gy0 = 0.f, gy1 = 0.f;                       // box of this lane's face
  if (cnt0 > 0) {
    gx0 = gx1 = P0[0][0]; gy0 = gy1 = P0[0][1];
    for (int v = 1; v < cnt0; ++v) { gx0 = fminf(gx0, P0[v][0]); gx1 = fmaxf(gx1, P0[v][0]); gy0 = fminf(gy0, P0[v][1]); gy1 = fmaxf(gy1, P0[v][1]); }
    gx0 -= 1e-6f; gx1 += 1e-6f; gy0 -= 1e-6f; gy1 += 1e-6f;
  }
  int nt = 0, np = 0;                                                    // triangles / pairs in the current batch (warp-uniform)
  auto flush = [&]() {
    __syncwarp();
    for (int base = 0; base < np; base += 32) {
      const int i = base + lane;
      float A[HF_MAXP][3], B[HF_MAXP][3];
      int cnt = 0;
      V3 T0 = v3(0.f, 0.f, 0.f), n = v3(0.f, 0.f, 1.f);
      if (i < np) {
        const int pr = wp[i];
        const float* tr = wt + (pr >> 5) * 12;
        const int qq = pr & 31;
        T0 = v3(tr[0], tr[1], tr[2]);
        const V3 T1 = v3(tr[3], tr[4], tr[5]), T2 = v3(tr[6], tr[7], tr[8]);
        n = v3(tr[9], tr[10], tr[11]);
        const int c0 = ff->plane_nvert[qq];
        float P[8][3];
        for (int v = 0; v < c0; ++v) {
          const int vid = ff->plane_vert[qq][v];
          P[v][0] = wv[3 * vid]; P[v][1] = wv[3 * vid + 1]; P[v][2] = wv[3 * vid + 2];
        }
        cnt = hf_clip(P, c0, A, T0.x, T0.y, T1.y - T0.y, -(T1.x - T0.x));
        if (cnt > 0) cnt = hf_clip(A, cnt, B, T1.x, T1.y, T2.y - T1.y, -(T2.x - T1.x));
        if (cnt > 0) cnt = hf_clip(B, cnt, A, T2.x, T2.y, T0.y - T2.y, -(T0.x - T2.x));
      }
      int k = 0;
      for (int v = 0; v < cnt; ++v) {
        const float dist = n.x * (A[v][0] - T0.x) + n.y * (A[v][1] - T0.y) + n.z * (A[v][2] - T0.z);
        if (dist < 0.f) ++k;
      }
      int incl = k;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += t; }
      const int total = __shfl_sync(FULLMASK, incl, 31);
      if (total == 0) continue;                                         // warp-uniform
      int at = nc + incl - k;
      for (int v = 0; v < cnt; ++v) {
        const float dist = n.x * (A[v][0] - T0.x) + n.y * (A[v][1] - T0.y) + n.z * (A[v][2] - T0.z);
        if (!(dist < 0.f)) continue;
        if (at < HF_CAP) {
          float* rec = cand + at * HF_REC;
          rec[0] = dist; rec[1] = A[v][0] - 0.5f * dist * n.x; rec[2] = A[v][1] - 0.5f * dist * n.y; rec[3] = A[v][2] - 0.5f * dist * n.z;
          rec[4] = n.x; rec[5] = n.y; rec[6] = n.z;
          deep = fminf(deep, dist);
        }
        ++at;
      }
      const float kf = (float)k;
      nsum = nsum + v3(wsum(kf * n.x), wsum(kf * n.y), wsum(kf * n.z));
      nc += total;
    }
    __syncwarp();
    nt = 0; np = 0;
  };
  for (int r = rmin; r <= rmax; ++r)
    for (int c = cmin; c <= cmax; ++c) {
      const float x0 = c * dx - sx, x1 = (c + 1) * dx - sx, y0 = r * dy - sy, y1 = (r + 1) * dy - sy;
      if (x1 < hx0 || x0 > hx1 || y1 < hy0 || y0 > hy1) continue;      // warp-uniform: the cell misses the hull's box
      const float h00 = data[(size_t)r * ncol + c] * sz, h01 = data[(size_t)r * ncol + c + 1] * sz;
      const float h10 = data[(size_t)(r + 1) * ncol + c] * sz, h11 = data[(size_t)(r + 1) * ncol + c + 1] * sz;
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        const V3 T0 = i == 0 ? v3(x0, y1, h10) : v3(x0, y0, h00);
        const V3 T1 = i == 0 ? v3(x0, y0, h00) : v3(x1, y0, h01);
        const V3 T2 = v3(x1, y1, h11);
        const float top = fmaxf(T0.z, fmaxf(T1.z, T2.z));
        if (C.z - rb > top || hz0 > top) continue;                      // warp-uniform
        V3 n = cross(T1 - T0, T2 - T0);
        n = (1.f / sqrtf(dot(n, n))) * n;
        const bool act = has && dot(Nw, n) < 0.f && !(x1 < gx0 || x0 > gx1 || y1 < gy0 || y0 > gy1);
        const unsigned bm = __ballot_sync(FULLMASK, act);
        if (!bm) continue;                                              // warp-uniform
        if (nt == HF_MAXT || np + 32 > HF_MAXPAIR) flush();
        if (lane == 0) {
          float* tr = wt + nt * 12;
          tr[0] = T0.x; tr[1] = T0.y; tr[2] = T0.z; tr[3] = T1.x; tr[4] = T1.y; tr[5] = T1.z; tr[6] = T2.x; tr[7] = T2.y; tr[8] = T2.z;
          tr[9] = n.x; tr[10] = n.y; tr[11] = n.z;
        }
        if (act) wp[np + __popc(bm & ((1u << lane) - 1u))] = (nt << 5) | lane;
        np += __popc(bm);
        ++nt;
      }
    }
  flush();
#else
  for (int r = rmin; r <= rmax; ++r)
    for (int c = cmin; c <= cmax; ++c) {
      const float x0 = c * dx - sx, x1 = (c + 1) * dx - sx, y0 = r * dy - sy, y1 = (r + 1) * dy - sy;
#ifdef ODUCK_HF_CULL
      if (x1 < bx0 || x0 > bx1 || y1 < by0 || y0 > by1) continue;      // warp-uniform: the cell misses the hull's box
#endif
      const float h00 = data[(size_t)r * ncol + c] * sz, h01 = data[(size_t)r * ncol + c + 1] * sz;
      const float h10 = data[(size_t)(r + 1) * ncol + c] * sz, h11 = data[(size_t)(r + 1) * ncol + c + 1] * sz;
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        // counter-clockwise seen from above; the cell is split along (c, r) - (c + 1, r + 1)
        const V3 T0 = i == 0 ? v3(x0, y1, h10) : v3(x0, y0, h00);
        const V3 T1 = i == 0 ? v3(x0, y0, h00) : v3(x1, y0, h01);
        const V3 T2 = v3(x1, y1, h11);
        const float top = fmaxf(T0.z, fmaxf(T1.z, T2.z));
        if (C.z - rb > top) continue;                                   // warp-uniform
#ifdef ODUCK_HF_CULL
        if (bz0 > top) continue;                                        // warp-uniform: the whole hull is above this triangle
#endif
        V3 n = cross(T1 - T0, T2 - T0);
        n = (1.f / sqrtf(dot(n, n))) * n;
        float A[HF_MAXP][3], B[HF_MAXP][3];
        int cnt = 0;
#ifdef ODUCK_HF_CULL
        if (has && dot(Nw, n) < 0.f && !(x1 < fx0 || x0 > fx1 || y1 < fy0 || y0 > fy1)) {
#else
        if (has && dot(Nw, n) < 0.f) {                                  // only the faces that look down onto the triangle
#endif
          cnt = hf_clip(P0, cnt0, A, T0.x, T0.y, T1.y - T0.y, -(T1.x - T0.x));
          if (cnt > 0) cnt = hf_clip(A, cnt, B, T1.x, T1.y, T2.y - T1.y, -(T2.x - T1.x));
          if (cnt > 0) cnt = hf_clip(B, cnt, A, T2.x, T2.y, T0.y - T2.y, -(T0.x - T2.x));
        }
        int k = 0;
        for (int v = 0; v < cnt; ++v) {
          const float dist = n.x * (A[v][0] - T0.x) + n.y * (A[v][1] - T0.y) + n.z * (A[v][2] - T0.z);
          if (dist < 0.f) ++k;
        }
        int incl = k;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += t; }
        const int total = __shfl_sync(FULLMASK, incl, 31);
        if (total == 0) continue;                                       // warp-uniform
        int at = nc + incl - k;
        for (int v = 0; v < cnt; ++v) {
          const float dist = n.x * (A[v][0] - T0.x) + n.y * (A[v][1] - T0.y) + n.z * (A[v][2] - T0.z);
          if (!(dist < 0.f)) continue;
          if (at < HF_CAP) {
            float* rec = cand + at * HF_REC;
            rec[0] = dist; rec[1] = A[v][0] - 0.5f * dist * n.x; rec[2] = A[v][1] - 0.5f * dist * n.y; rec[3] = A[v][2] - 0.5f * dist * n.z;
            rec[4] = n.x; rec[5] = n.y; rec[6] = n.z;
            deep = fminf(deep, dist);
          }
          ++at;
        }
        nsum = nsum + (float)total * n;
        nc += total;
      }
    }
#endif
  if (nc == 0) return;
  nc = min(nc, HF_CAP);
  __syncwarp();
  const float deepest = -wmaxf(-deep);
  const float thr = fminf(0.f, deepest + 1e-3f);                        // plane_convex's rule: within 1 mm of the deepest
  const V3 nm = (1.f / sqrtf(dot(nsum, nsum))) * nsum;                  // every triangle normal has n_z > 0
  const float ninf = -__int_as_float(0x7f800000);
  // point i of the list as seen by this lane in the chunk starting at `base`
#define HF_LOAD(base, P_, dm_)                                                                                         \
  V3 P_ = v3(0.f, 0.f, 0.f); float dm_ = ninf;                                                                          \
  {                                                                                                                    \
    const int i_ = (base) + lane;                                                                                      \
    if (i_ < nc) { const float* rec = cand + i_ * HF_REC; P_ = v3(rec[1], rec[2], rec[3]); dm_ = rec[0] < thr ? 0.f : -1e6f; }   \
  }
  auto point = [&](int i) { const float* rec = cand + i * HF_REC; return v3(rec[1], rec[2], rec[3]); };
  int idx[4] = {0, 0, 0, 0};
  for (int base = 0; base < nc; base += 32) {                           // a: the first candidate inside the threshold
    HF_LOAD(base, P, dm)
    const unsigned bm = __ballot_sync(FULLMASK, dm == 0.f);
    if (bm) { idx[0] = base + __ffs(bm) - 1; break; }
  }
  const V3 pa = point(idx[0]);
  {
    float best = ninf;                                                   // b: farthest from a
    for (int base = 0; base < nc; base += 32) {
      HF_LOAD(base, P, dm)
      const V3 ap = pa - P;
      int li; const float v = wargmax_val(dot(ap, ap) + dm, lane, &li);
      if (v > best) { best = v; idx[1] = base + li; }
    }
  }
  const V3 pb = point(idx[1]);
  const V3 ab = cross(nm, pa - pb);
  {
    float best = ninf;                                                   // c: farthest from the line a b
    for (int base = 0; base < nc; base += 32) {
      HF_LOAD(base, P, dm)
      int li; const float v = wargmax_val(fabsf(dot(pa - P, ab)) + dm, lane, &li);
      if (v > best) { best = v; idx[2] = base + li; }
    }
  }
  const V3 pc = point(idx[2]);
  const V3 ac = cross(nm, pa - pc), bc = cross(nm, pb - pc);
  {
    float b1 = ninf, b2 = ninf; int i1 = 0, i2 = 0;                     // d: farthest from the edges b c and a c
    for (int base = 0; base < nc; base += 32) {
      HF_LOAD(base, P, dm)
      int li; float v = wargmax_val(fabsf(dot(pb - P, bc)) + dm, lane, &li);
      if (v > b1) { b1 = v; i1 = base + li; }
      v = wargmax_val(fabsf(dot(pa - P, ac)) + dm, lane, &li);
      if (v > b2) { b2 = v; i2 = base + li; }
    }
    idx[3] = b1 >= b2 ? i1 : i2;
  }
#undef HF_LOAD
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    bool uniq = true;
#pragma unroll
    for (int p = 0; p < c; ++p) uniq = uniq && (idx[p] != idx[c]);
    if (uniq && lane == c) {
      const float* rec = cand + idx[c] * HF_REC;
      float* cc = s.con[4 * f + c];
      cc[0] = rec[0]; cc[1] = rec[1]; cc[2] = rec[2]; cc[3] = rec[3]; cc[13] = rec[4]; cc[14] = rec[5]; cc[15] = rec[6];
    }
  }
  __syncwarp();
}
