// oduck_hfcollide.cuh -- height-field floor vs one convex foot hull (rough_terrain scenes; only the HF = true kernel
// instantiations contain this code, the flat-floor instantiations are unchanged).  Same scheme as
// oracle/oduck_oracle.cpp hfield_convex: for every terrain triangle under the foot's bounding sphere, every foot face that
// looks down onto it is clipped to the triangle's vertical column; clipped points below the triangle plane are candidates
// (dist along the triangle normal, pos midway, normal = triangle normal); 4 of the candidates within 1 mm of the deepest one
// are kept by _manifold_points with the mean normal (copies of one point on a shared triangle edge masked first).  Warp mapping:
// see hf_collide.
#pragma once
#include "oduck_ffcollide.cuh"

struct DevHF {   // global memory
  int nrow, ncol;
  float sx, sy, sz, dx, dy;     // radii, elevation scale, grid pitch
  const float* data;            // [nrow][ncol] elevation in [0, 1]
};

#define HF_CAP 512             // candidates kept per foot (the oracle keeps as many; a resting foot has a few dozen)
#define HF_REC 8               // dist, pos[3], normal[3], twin flag
#define HF_SCRATCH (HF_CAP * HF_REC)
#define HF_MARGIN 2e-5f        // box-cull slack: 20 ulp of a 10 m coordinate (the field is 20 m wide); a 1e-6 slack culled faces that the
                               // fp64 oracle clips to a sliver at |x| ~ 9 m
#define HF_MAXP 11             // 8-gon clipped by 3 half-planes: at most 11 vertices
#define HF_LANES 13            // (triangle, face) pairs clipped per round: one 33-float polygon per lane in shared memory
#define HF_MAXPAIR 64          // pair list of a batch (the rhs + rowbuf rows of WarpSmem)
#ifdef ODUCK_HF_NO_TWIN
#define HF_TWIN -1.f
#else
#define HF_TWIN 1e-5f
#endif
//         // clipped points closer than this (max-norm) are copies of one point on a shared triangle edge

// Sutherland-Hodgman IN PLACE: keep the part of the polygon P[cnt][3] on the inner side (d <= 0) of the vertical plane through
// r0 with outward normal (sdx, sdy).  One buffer suffices because an output slot never overtakes the input: the vertex after
// the current one is held in registers (and the first vertex for the wrap-around), and before iteration v at most v + 1 points
// have been written -- v + 1 only when vertex v itself is outside (a convex polygon crosses the plane at most twice).  The same
// arithmetic, in the same order, as the two-buffer form of oracle/oduck_oracle.cpp hfield_convex.  The polygons live in shared
// memory: with the kernel's shared-memory carve-out the L1 that would back per-lane local arrays is ~20 KB per SM, and the
// first version of this collider spent 60 % of k_step<HF> waiting for local-memory loads (ncu r02a, profiles/).
__device__ __forceinline__ int hf_clip(float* __restrict__ P, const int cnt, const float r0x, const float r0y, const float sdx, const float sdy) {
  if (cnt <= 0) return 0;
  const float fx = P[0], fy = P[1], fz = P[2];
  float cx = fx, cy = fy, cz = fz;
  const float dfirst = sdx * (cx - r0x) + sdy * (cy - r0y);
  float d0 = dfirst;
  int no = 0;
  for (int v = 0; v < cnt; ++v) {
    float nx, ny, nz, d1;
    if (v + 1 == cnt) { nx = fx; ny = fy; nz = fz; d1 = dfirst; }
    else { nx = P[3 * v + 3]; ny = P[3 * v + 4]; nz = P[3 * v + 5]; d1 = sdx * (nx - r0x) + sdy * (ny - r0y); }
    if (d0 <= 0.f && no < HF_MAXP) { P[3 * no] = cx; P[3 * no + 1] = cy; P[3 * no + 2] = cz; ++no; }
    if ((d0 <= 0.f) != (d1 <= 0.f) && no < HF_MAXP && no <= v + 1) {
      const float t = d0 / (d0 - d1);
      P[3 * no] = cx + t * (nx - cx); P[3 * no + 1] = cy + t * (ny - cy); P[3 * no + 2] = cz + t * (nz - cz);
      ++no;
    }
    cx = nx; cy = ny; cz = nz; d0 = d1;
  }
  return no;
}

// Writes the four contact records of foot f: s.con[4 f + c][0] = dist (1: inactive), [1..3] = pos, [13..15] = normal.
// Three stages per foot.  (1) The terrain triangles under the hull's box are enumerated by the whole warp (warp-uniform loops);
// lane = hull face decides whether its face looks down onto the triangle and overlaps its cell, and the surviving (triangle,
// face) pairs are appended to a list in shared memory -- cells the hull's box misses and triangles wholly below its lowest
// vertex are skipped (a candidate needs a hull point BELOW the triangle plane, which never rises above the triangle's top;
// HF_MARGIN keeps the culls conservative under fp32 rounding).  (2) The pairs are clipped HF_LANES at a time, lane =
// pair, each lane in its own shared-memory polygon; clipped points below the triangle plane are appended to the env's
// candidate list in the oracle's order (triangle, face, polygon vertex) with a warp prefix sum.  (3) Twins are masked and the
// manifold is selected over the list 32 candidates at a time with the plane collider's first-index arg-max.
// Shared memory borrowed from WarpSmem while the Hessian does not exist: H (world vertices + polygons), rhs + rowbuf (pairs).
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
  const V3 C = p0 + rot(v3(ff->center[f][0], ff->center[f][1], ff->center[f][2]));
  const float rb = ff->radius;
  const int nrow = hf->nrow, ncol = hf->ncol;
  const float sx = hf->sx, sy = hf->sy, sz = hf->sz, dx = hf->dx, dy = hf->dy;
  const float* __restrict__ data = hf->data;
  int cmin = (int)floorf((C.x - rb + sx) / dx), cmax = (int)floorf((C.x + rb + sx) / dx);
  int rmin = (int)floorf((C.y - rb + sy) / dy), rmax = (int)floorf((C.y + rb + sy) / dy);
  cmin = max(cmin, 0); rmin = max(rmin, 0); cmax = min(cmax, ncol - 2); rmax = min(rmax, nrow - 2);
  float* wv = s.H;                                                       // world hull vertices [32][3]
  float* poly = s.H + 96 + 3 * HF_MAXP * (lane < HF_LANES ? lane : 0);    // this lane's polygon
  int* wp = reinterpret_cast<int*>(s.rhs);                               // pair list: (cell * 2 + half) << 5 | face
  float hx0, hx1, hy0, hy1, hz0;                                          // box of the hull
  const bool vv = lane < m.nvert;                                         // lane = hull vertex (world position wl kept for the plane-side cull)
  V3 wl;
  {
    const int vl = vv ? lane : 0;
    const V3 w = p0 + rot(v3(m.vert[f][0][vl], m.vert[f][1][vl], m.vert[f][2][vl]));
    wl = w;
    __syncwarp();                                                         // (H's previous readers are done)
    if (vv) { wv[3 * lane] = w.x; wv[3 * lane + 1] = w.y; wv[3 * lane + 2] = w.z; }
    const float inf = __int_as_float(0x7f800000);
    hx1 = wmaxf(vv ? w.x : -inf) + HF_MARGIN; hx0 = -wmaxf(vv ? -w.x : -inf) - HF_MARGIN;
    hy1 = wmaxf(vv ? w.y : -inf) + HF_MARGIN; hy0 = -wmaxf(vv ? -w.y : -inf) - HF_MARGIN;
    hz0 = -wmaxf(vv ? -w.z : -inf) - HF_MARGIN;
  }
  __syncwarp();
  // this lane's face: world normal and xy box
  const bool has = lane < ff->nplane;
  const int q = has ? lane : 0;
  const V3 Nw = rot(v3(ff->plane_normal[f][0][q], ff->plane_normal[f][1][q], ff->plane_normal[f][2][q]));
  float gx0 = 0.f, gx1 = 0.f, gy0 = 0.f, gy1 = 0.f;
  unsigned fmask = 0u;                                                    // this face's vertices as a bit mask over the hull vertices
  if (has) {
    const int cnt0 = ff->plane_nvert[q];
    for (int v = 0; v < cnt0; ++v) {
      const int vid = ff->plane_vert[q][v];
      fmask |= 1u << vid;
      const float x = wv[3 * vid], y = wv[3 * vid + 1];
      if (v == 0) { gx0 = gx1 = x; gy0 = gy1 = y; }
      else { gx0 = fminf(gx0, x); gx1 = fmaxf(gx1, x); gy0 = fminf(gy0, y); gy1 = fmaxf(gy1, y); }
    }
    gx0 -= HF_MARGIN; gx1 += HF_MARGIN; gy0 -= HF_MARGIN; gy1 += HF_MARGIN;
  }
  int nc = 0;                      // candidates so far (warp-uniform)
  V3 nsum = v3(0.f, 0.f, 0.f);     // sum of the candidates' normals (warp-uniform)
  float deep = 0.f;                // lane-local deepest candidate
  int np = 0;                      // pairs in the list (warp-uniform)
  auto flush = [&]() {
    __syncwarp();
    for (int base = 0; base < np; base += HF_LANES) {
      const int i = base + lane;
      int cnt = 0;
      V3 T0 = v3(0.f, 0.f, 0.f), n = v3(0.f, 0.f, 1.f);
      if (lane < HF_LANES && i < np) {
        const int pr = wp[i];
        const int qq = pr & 31, half = (pr >> 5) & 1, cell = pr >> 6;
        const int r = cell / ncol, c = cell - r * ncol;
        // the triangle, by the expressions of the enumeration below (the same floats)
        const float x0 = c * dx - sx, x1 = (c + 1) * dx - sx, y0 = r * dy - sy, y1 = (r + 1) * dy - sy;
        const float h00 = data[(size_t)r * ncol + c] * sz, h01 = data[(size_t)r * ncol + c + 1] * sz;
        const float h10 = data[(size_t)(r + 1) * ncol + c] * sz, h11 = data[(size_t)(r + 1) * ncol + c + 1] * sz;
        T0 = half == 0 ? v3(x0, y1, h10) : v3(x0, y0, h00);
        const V3 T1 = half == 0 ? v3(x0, y0, h00) : v3(x1, y0, h01);
        const V3 T2 = v3(x1, y1, h11);
        n = cross(T1 - T0, T2 - T0);
        n = (1.f / sqrtf(dot(n, n))) * n;
        const int c0 = ff->plane_nvert[qq];
        for (int v = 0; v < c0; ++v) {
          const int vid = ff->plane_vert[qq][v];
          poly[3 * v] = wv[3 * vid]; poly[3 * v + 1] = wv[3 * vid + 1]; poly[3 * v + 2] = wv[3 * vid + 2];
        }
        cnt = hf_clip(poly, c0, T0.x, T0.y, T1.y - T0.y, -(T1.x - T0.x));
        cnt = hf_clip(poly, cnt, T1.x, T1.y, T2.y - T1.y, -(T2.x - T1.x));
        cnt = hf_clip(poly, cnt, T2.x, T2.y, T0.y - T2.y, -(T0.x - T2.x));
      }
      int k = 0;
      for (int v = 0; v < cnt; ++v) {
        const float dist = n.x * (poly[3 * v] - T0.x) + n.y * (poly[3 * v + 1] - T0.y) + n.z * (poly[3 * v + 2] - T0.z);
        if (dist < 0.f) ++k;
      }
      int incl = k;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += t; }
      const int total = __shfl_sync(FULLMASK, incl, 31);
      if (total == 0) continue;                                         // warp-uniform
      int at = nc + incl - k;
      for (int v = 0; v < cnt; ++v) {
        const float px = poly[3 * v], py = poly[3 * v + 1], pz = poly[3 * v + 2];
        const float dist = n.x * (px - T0.x) + n.y * (py - T0.y) + n.z * (pz - T0.z);
        if (!(dist < 0.f)) continue;
        if (at < HF_CAP) {
          float* rec = cand + at * HF_REC;
          rec[0] = dist; rec[1] = px - 0.5f * dist * n.x; rec[2] = py - 0.5f * dist * n.y; rec[3] = pz - 0.5f * dist * n.z;
          rec[4] = n.x; rec[5] = n.y; rec[6] = n.z; rec[7] = 0.f;
          deep = fminf(deep, dist);
        }
        ++at;
      }
      const float kf = (float)k;
      float ns[3] = {kf * n.x, kf * n.y, kf * n.z};
      const float tt = wfold<3>(ns, lane);
      nsum = nsum + v3(wfold_get(tt, 0), wfold_get(tt, 1), wfold_get(tt, 2));
      nc += total;
    }
    __syncwarp();
    np = 0;
  };
  for (int r = rmin; r <= rmax; ++r)
    for (int c = cmin; c <= cmax; ++c) {
      const float x0 = c * dx - sx, x1 = (c + 1) * dx - sx, y0 = r * dy - sy, y1 = (r + 1) * dy - sy;
      if (x1 < hx0 || x0 > hx1 || y1 < hy0 || y0 > hy1) continue;      // warp-uniform: the cell misses the hull's box
      const float h00 = data[(size_t)r * ncol + c] * sz, h01 = data[(size_t)r * ncol + c + 1] * sz;
      const float h10 = data[(size_t)(r + 1) * ncol + c] * sz, h11 = data[(size_t)(r + 1) * ncol + c + 1] * sz;
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        // counter-clockwise seen from above; the cell is split along (c, r) - (c + 1, r + 1)
        const V3 T0 = i == 0 ? v3(x0, y1, h10) : v3(x0, y0, h00);
        const V3 T1 = i == 0 ? v3(x0, y0, h00) : v3(x1, y0, h01);
        const V3 T2 = v3(x1, y1, h11);
        const float top = fmaxf(T0.z, fmaxf(T1.z, T2.z));
        if (C.z - rb > top || hz0 > top) continue;                      // warp-uniform: the hull is above this triangle
        V3 n = cross(T1 - T0, T2 - T0);
        n = (1.f / sqrtf(dot(n, n))) * n;
        // plane-side cull: a clipped point is a convex combination of its face's vertices, so a face none of whose vertices lies
        // below the triangle's plane (HF_MARGIN slack) yields no candidate.  Lane = hull vertex tests ITS vertex, one ballot hands
        // every face lane the set; a swing foot drops out here entirely, a resting one keeps the faces around its sole.
        const unsigned below = __ballot_sync(FULLMASK, vv && dot(n, wl - T0) < HF_MARGIN);
        if (!below) continue;                                           // warp-uniform
        const bool act = has && (fmask & below) != 0u && dot(Nw, n) < 0.f && !(x1 < gx0 || x0 > gx1 || y1 < gy0 || y0 > gy1);   // the face looks down onto the triangle
        const unsigned bm = __ballot_sync(FULLMASK, act);
        if (!bm) continue;                                              // warp-uniform
        if (np + __popc(bm) > HF_MAXPAIR) flush();
        if (act) wp[np + __popc(bm & ((1u << lane) - 1u))] = ((((r * ncol + c) << 1) | i) << 5) | lane;
        np += __popc(bm);
      }
    }
  flush();
  if (nc == 0) return;
  nc = min(nc, HF_CAP);
  __syncwarp();
  const float deepest = -wmaxf(-deep);
  const float thr = fminf(0.f, deepest + 1e-3f);                        // plane_convex's rule: within 1 mm of the deepest
  const V3 nm = (1.f / sqrtf(dot(nsum, nsum))) * nsum;                  // every triangle normal has n_z > 0
  const float ninf = -__int_as_float(0x7f800000);
  // Twins (oracle hfield_convex): an in-threshold candidate whose clipped point lies within HF_TWIN of an EARLIER in-threshold
  // candidate is a copy of the same point seen through a neighbouring triangle; it is masked out, so the arg-max passes below
  // never choose between copies that differ only by rounding and by the triangle normal they carry.
  for (int base = 0; base < nc; base += 32) {
    const int j = base + lane;
    V3 pj = v3(0.f, 0.f, 0.f);
    bool inj = false;
    if (j < nc) {
      const float* rec = cand + j * HF_REC;
      inj = rec[0] < thr;
      pj = v3(rec[1] + 0.5f * rec[0] * rec[4], rec[2] + 0.5f * rec[0] * rec[5], rec[3] + 0.5f * rec[0] * rec[6]);   // the clipped point itself
    }
    bool tw = false;
    for (int ib = 0; ib <= base; ib += 32) {
      V3 po = pj;
      bool ino = inj;
      if (ib != base) {
        const float* rec = cand + (ib + lane) * HF_REC;                  // earlier chunks are full
        ino = rec[0] < thr;
        po = v3(rec[1] + 0.5f * rec[0] * rec[4], rec[2] + 0.5f * rec[0] * rec[5], rec[3] + 0.5f * rec[0] * rec[6]);
      }
      unsigned bm = __ballot_sync(FULLMASK, ino);
      while (bm) {                                                       // warp-uniform
        const int l = __ffs(bm) - 1;
        bm &= bm - 1;
        const float ox = __shfl_sync(FULLMASK, po.x, l), oy = __shfl_sync(FULLMASK, po.y, l), oz = __shfl_sync(FULLMASK, po.z, l);
        if (inj && ib + l < j && fmaxf(fabsf(pj.x - ox), fmaxf(fabsf(pj.y - oy), fabsf(pj.z - oz))) < HF_TWIN) tw = true;
      }
    }
    if (j < nc && tw) cand[j * HF_REC + 7] = 1.f;
  }
  __syncwarp();
  // point i of the list as seen by this lane in the chunk starting at `base`
#define HF_LOAD(base, P_, dm_)                                                                                         \
  V3 P_ = v3(0.f, 0.f, 0.f); float dm_ = ninf;                                                                          \
  {                                                                                                                    \
    const int i_ = (base) + lane;                                                                                      \
    if (i_ < nc) { const float* rec = cand + i_ * HF_REC; P_ = v3(rec[1], rec[2], rec[3]); dm_ = (rec[0] < thr && rec[7] == 0.f) ? 0.f : -1e6f; }   \
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
