// oduck_hfcollide.cuh -- height-field floor vs one convex foot hull (rough_terrain scenes; only the HF = true kernel
// instantiations contain this code, the flat-floor instantiations are unchanged).  Same scheme as
// oracle/oduck_oracle.cpp hfield_convex: for every terrain triangle under the foot's bounding sphere, every foot face that
// looks down onto it is clipped to the triangle's vertical column; clipped points below the triangle plane are candidates
// (dist along the triangle normal, pos midway, normal = triangle normal); 4 of the candidates within 1 mm of the deepest one
// are kept by _manifold_points with the mean normal (copies of one point on a shared triangle edge masked first).  Warp mapping:
// see hf_collide; the three versions of this collider and their measurements: DESIGN.md 3e.
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
#define HF_MAXPAIR 64          // pair list of a batch (the rhs + rowbuf rows of WarpSmem)
#define HF_NIN 96              // in-threshold candidates the shared-memory selection handles (3 per lane); more: hf_select_generic
#define HF_MAXTRI 28           // triangle table of a foot in shared memory (12 floats each): 96 + 96 + 28 * 12 = 528 floats = WarpSmem::H
#ifndef ODUCK_HF_STAT
#define ODUCK_HF_STAT(what, n)  // tests/emu counts triangles / pairs / candidates through this hook
#endif
// clipped points closer than this (max-norm) are copies of one point seen through a neighbouring triangle or face ("twins")
#ifdef ODUCK_HF_NO_TWIN        // (A/B switch of the oracle and the kernel: no twin masking)
#define HF_TWIN -1.f
#else
#define HF_TWIN 1e-5f
#endif

// One Sutherland-Hodgman pass, one polygon per group of G consecutive lanes (G = 8 or 16, lane j of the group holds vertex j in
// registers): keep the part of the polygon on the inner side (d <= 0) of the vertical plane through r0 with outward normal
// (sdx, sdy).  Every lane classifies its own vertex and the edge to the next one (shuffle), two ballots give it the output slot
// of what it emits (the vertex if inside, then the crossing if the edge changes side), the emitted points go through 3 G floats
// of shared memory and come back as vertex j of the clipped polygon.  The same arithmetic on the same
// points, in the same output order, as the serial loop of oracle/oduck_oracle.cpp hfield_convex; no lane-divergent loop.
// (The first two versions clipped one polygon per lane serially -- local memory, then shared memory: 13 busy lanes, divergent
// trip counts, 11 k warp-instructions per foot and substep in the clipping alone, ncu r02e in profiles/.)
__device__ __forceinline__ void hf_clip_pass(V3& P, int& cnt, const int j, const int gb, const unsigned gmask, const unsigned glt, const int G,
                                             float* __restrict__ sc, const float r0x, const float r0y, const float sdx, const float sdy) {
  const float d0 = sdx * (P.x - r0x) + sdy * (P.y - r0y);
  // no vertex of any of the warp's polygons outside this plane: the pass returns every polygon as it is (cells are 8 cm wide, most
  // faces 1 - 3 cm: a face usually lies inside two of its triangle's three half-planes)
  if (__ballot_sync(FULLMASK, j < cnt && !(d0 <= 0.f)) == 0u) return;   // warp-uniform
  const int src = gb + ((j + 1 >= cnt) ? 0 : j + 1);                    // the next vertex, the first one after the last
  const V3 N = v3(__shfl_sync(FULLMASK, P.x, src), __shfl_sync(FULLMASK, P.y, src), __shfl_sync(FULLMASK, P.z, src));
  const float d1 = __shfl_sync(FULLMASK, d0, src);
  const bool act = j < cnt;
  const bool in0 = act && d0 <= 0.f, cr = act && ((d0 <= 0.f) != (d1 <= 0.f));
  // output slot = what the group's earlier lanes emit (each a vertex and / or a crossing): two ballots instead of a prefix sum
  const unsigned bi = __ballot_sync(FULLMASK, in0), bc = __ballot_sync(FULLMASK, cr);
  int off = __popc(bi & glt) + __popc(bc & glt);
  const int total = __popc(bi & gmask) + __popc(bc & gmask);
  __syncwarp();                                                         // (the previous pass has read its points)
  if (in0 && off < G) { sc[3 * off] = P.x; sc[3 * off + 1] = P.y; sc[3 * off + 2] = P.z; ++off; }
  if (cr && off < G) {
    const float t = d0 / (d0 - d1);
    sc[3 * off] = P.x + t * (N.x - P.x); sc[3 * off + 1] = P.y + t * (N.y - P.y); sc[3 * off + 2] = P.z + t * (N.z - P.z);
  }
  __syncwarp();
  cnt = min(total, G);
  if (j < cnt) P = v3(sc[3 * j], sc[3 * j + 1], sc[3 * j + 2]);
}

// Selection over the candidate list in HBM, any length (the path for more than HF_NIN in-threshold candidates: a foot pressed
// flat into the terrain over many cells): twins masked in the records, then the manifold arg-max passes over the list 32
// candidates at a time.
static __device__ __noinline__ void hf_select_generic(WarpSmem& s, const int lane, const int f, float* __restrict__ cand, const int nc,
                                                      const float thr, const V3 nm) {
  const float ninf = -__int_as_float(0x7f800000);
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

// Writes the four contact records of foot f: s.con[4 f + c][0] = dist (1: inactive), [1..3] = pos, [13..15] = normal.
// Three stages per foot.  (1) The terrain triangles under the hull's box are enumerated by the whole warp (warp-uniform loops);
// lane = hull face decides whether its face looks down onto the triangle, overlaps its cell and has a vertex below its plane,
// and the surviving (triangle, face) pairs are appended to a list in shared memory -- cells the hull's box misses and
// triangles wholly below its lowest vertex are skipped (a candidate needs a hull point BELOW the triangle plane, which never
// rises above the triangle's top; HF_MARGIN keeps the culls conservative under fp32 rounding).  (2) The pairs are clipped 32 / G
// at a time, one polygon per group of G lanes (hf_clip_pass; G = 8 when no face has more than 5 vertices, as on the duck's
// foot, else 16); clipped points below the triangle plane are appended to the env's candidate list in the oracle's order
// (triangle, face, polygon vertex) = lane order, positions from a ballot.  (3) The candidates within 1 mm of the deepest one are
// compacted (3 per lane, their clipped points in shared memory), twins are masked with one broadcast read per earlier
// candidate, and the manifold is selected with the plane collider's first-index arg-max.
// Shared memory borrowed from WarpSmem while the Hessian does not exist: H (world vertices, clip scratch / compacted points and
// indices), rhs + rowbuf (pairs).
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
  float* scr = s.H + 96;                                                 // stage 2: clip scratch, 3 floats per lane; stage 3: compacted points [HF_NIN][3]
  float4* tri = reinterpret_cast<float4*>(s.H + 192);                    // stages 1 - 2: triangles that own pairs, 3 float4 each: T0 | n.x, T1 | n.y, T2 | n.z
                                                                         // (at most HF_MAXTRI: the bounding sphere spans 3 x 3 cells; stage 3 reuses the space)
  int ntri = 0;                                                          // warp-uniform
  int* cidx = reinterpret_cast<int*>(s.H + 96 + 3 * HF_NIN);             // stage 3: list index of compacted candidate i
  int* wp = reinterpret_cast<int*>(s.rhs);                               // pair list: triangle slot << 5 | face
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
  const int cnt0 = has ? ff->plane_nvert[q] : 0;
  if (has) {
    for (int v = 0; v < cnt0; ++v) {
      const int vid = ff->plane_vert[q][v];
      fmask |= 1u << vid;
      const float x = wv[3 * vid], y = wv[3 * vid + 1];
      if (v == 0) { gx0 = gx1 = x; gy0 = gy1 = y; }
      else { gx0 = fminf(gx0, x); gx1 = fmaxf(gx1, x); gy0 = fminf(gy0, y); gy1 = fmaxf(gy1, y); }
    }
    gx0 -= HF_MARGIN; gx1 += HF_MARGIN; gy0 -= HF_MARGIN; gy1 += HF_MARGIN;
  }
  // lanes per polygon: a face of v vertices has at most v + 3 after three clips
  const int lg = (__ballot_sync(FULLMASK, cnt0 > 5) == 0u) ? 3 : 4;
  const int G = 1 << lg, gi = lane >> lg, j = lane & (G - 1), gb = lane & ~(G - 1);
  float* sc = scr + 3 * gb;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned gmask = (G == 8 ? 0xffu : 0xffffu) << gb, glt = gmask & lt;   // this lane's group, and its earlier lanes
  int nc = 0;                      // candidates so far (warp-uniform)
  V3 nsum = v3(0.f, 0.f, 0.f);     // lane-local sum of this lane's candidates' normals
  float deep = 0.f;                // lane-local deepest candidate
  int np = 0;                      // pairs in the list (warp-uniform)
  auto flush = [&]() {
    __syncwarp();
    for (int base = 0; base < np; base += 32 >> lg) {
      const int i = base + gi;
      const bool valid = i < np;
      const int pr = wp[valid ? i : 0];
      const int qq = pr & 31;
      // the triangle as the enumeration below computed it (the same floats): three float4 broadcasts
      const float4 t0 = tri[3 * (pr >> 5)], t1 = tri[3 * (pr >> 5) + 1], t2 = tri[3 * (pr >> 5) + 2];
      const V3 T0 = v3(t0.x, t0.y, t0.z), T1 = v3(t1.x, t1.y, t1.z), T2 = v3(t2.x, t2.y, t2.z), n = v3(t0.w, t1.w, t2.w);
      // (the face's vertex list through two shuffles from lane = face instead of these cached global loads: 2.67 vs 2.71 M
      // env-steps/s at 16384 envs, profiles/r02ae_bench_rough_*.json -- not taken)
      int cnt = valid ? ff->plane_nvert[qq] : 0;
      V3 P = v3(0.f, 0.f, 0.f);
      if (j < cnt) { const int vid = ff->plane_vert[qq][j]; P = v3(wv[3 * vid], wv[3 * vid + 1], wv[3 * vid + 2]); }
      hf_clip_pass(P, cnt, j, gb, gmask, glt, G, sc, T0.x, T0.y, T1.y - T0.y, -(T1.x - T0.x));
      hf_clip_pass(P, cnt, j, gb, gmask, glt, G, sc, T1.x, T1.y, T2.y - T1.y, -(T2.x - T1.x));
      hf_clip_pass(P, cnt, j, gb, gmask, glt, G, sc, T2.x, T2.y, T0.y - T2.y, -(T0.x - T2.x));
      const float dist = n.x * (P.x - T0.x) + n.y * (P.y - T0.y) + n.z * (P.z - T0.z);
      const bool isc = j < cnt && dist < 0.f;
      const unsigned bm = __ballot_sync(FULLMASK, isc);
      if (!bm) continue;                                                // warp-uniform
      if (isc) {
        const int at = nc + __popc(bm & lt);
        if (at < HF_CAP) {
          float4* rec = reinterpret_cast<float4*>(cand + at * HF_REC);
          rec[0] = make_float4(dist, P.x - 0.5f * dist * n.x, P.y - 0.5f * dist * n.y, P.z - 0.5f * dist * n.z);
          rec[1] = make_float4(n.x, n.y, n.z, 0.f);
          deep = fminf(deep, dist);
        }
        nsum = nsum + n;
      }
      nc += __popc(bm);
    }
    __syncwarp();
    np = 0;
  };
  // (lane = triangle for the hull-independent part -- vertices, normal, box / height culls -- and a visit of the survivors through the
  // table was measured slower than these warp-uniform loops with their early exits: 2.61 vs 2.71 M env-steps/s at 16384 envs,
  // profiles/r02af_bench_rough_*.json)
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
        ODUCK_HF_STAT(0, 1)
        V3 n = cross(T1 - T0, T2 - T0);
        n = (1.f / sqrtf(dot(n, n))) * n;
        // plane-side cull: a clipped point is a convex combination of its face's vertices, so a face none of whose vertices lies
        // below the triangle's plane (HF_MARGIN slack) yields no candidate.  Lane = hull vertex tests ITS vertex, one ballot hands
        // every face lane the set; a swing foot drops out here entirely, a resting one keeps the faces around its sole.
        const unsigned below = __ballot_sync(FULLMASK, vv && dot(n, wl - T0) < HF_MARGIN);
        if (!below) continue;                                           // warp-uniform
        ODUCK_HF_STAT(1, 1)
        const bool act = has && (fmask & below) != 0u && dot(Nw, n) < 0.f && !(x1 < gx0 || x0 > gx1 || y1 < gy0 || y0 > gy1);   // the face looks down onto the triangle
        const unsigned bm = __ballot_sync(FULLMASK, act);
        if (!bm) continue;                                              // warp-uniform
        if (np + __popc(bm) > HF_MAXPAIR || ntri == HF_MAXTRI) { flush(); ntri = 0; }
        if (lane < 3) {
          const V3 Tl = lane == 0 ? T0 : (lane == 1 ? T1 : T2);
          tri[3 * ntri + lane] = make_float4(Tl.x, Tl.y, Tl.z, lane == 0 ? n.x : (lane == 1 ? n.y : n.z));
        }
        if (act) wp[np + __popc(bm & lt)] = (ntri << 5) | lane;       // pair = (triangle slot, face)
        ++ntri;
        np += __popc(bm);
        ODUCK_HF_STAT(2, __popc(bm))
      }
    }
  flush();
  ODUCK_HF_STAT(3, nc)
  if (nc == 0) return;
  nc = min(nc, HF_CAP);
  __syncwarp();
  const float deepest = -wmaxf(-deep);
  const float thr = fminf(0.f, deepest + 1e-3f);                        // plane_convex's rule: within 1 mm of the deepest
  {
    float ns[3] = {nsum.x, nsum.y, nsum.z};
    const float tt = wfold<3>(ns, lane);
    nsum = v3(wfold_get(tt, 0), wfold_get(tt, 1), wfold_get(tt, 2));
  }
  const V3 nm = (1.f / sqrtf(dot(nsum, nsum))) * nsum;                  // every triangle normal has n_z > 0
  const float ninf = -__int_as_float(0x7f800000);
  // compact the in-threshold candidates (only they can be picked: a masked-out score is x - 1e6 < 0 <= any in-threshold score)
  int nin = 0;
  for (int base = 0; base < nc; base += 32) {
    const int i = base + lane;
    const bool in = i < nc && cand[i * HF_REC] < thr;
    const unsigned bm = __ballot_sync(FULLMASK, in);
    const int slot = nin + __popc(bm & lt);
    if (in && slot < HF_NIN) cidx[slot] = i;
    nin += __popc(bm);
  }
  ODUCK_HF_STAT(4, nin)
  if (nin > HF_NIN) { hf_select_generic(s, lane, f, cand, nc, thr, nm); return; }   // warp-uniform
  __syncwarp();
  // slot lane + 32 c of the compacted list lives in this lane: midway point Pm (what the manifold rule looks at), clipped point Pc
  V3 Pm[HF_NIN / 32], Pc[HF_NIN / 32];
  bool ok[HF_NIN / 32];
#pragma unroll
  for (int c = 0; c < HF_NIN / 32; ++c) {
    const int sl = lane + 32 * c;
    ok[c] = sl < nin;
    Pm[c] = Pc[c] = v3(0.f, 0.f, 0.f);
    if (ok[c]) {
      const float4* rec = reinterpret_cast<const float4*>(cand + cidx[sl] * HF_REC);
      const float4 a = rec[0], b = rec[1];
      Pm[c] = v3(a.y, a.z, a.w);
      Pc[c] = v3(a.y + 0.5f * a.x * b.x, a.z + 0.5f * a.x * b.y, a.w + 0.5f * a.x * b.z);   // the clipped point itself
      scr[3 * sl] = Pc[c].x; scr[3 * sl + 1] = Pc[c].y; scr[3 * sl + 2] = Pc[c].z;
    }
  }
  __syncwarp();
  // Twins (oracle hfield_convex): an in-threshold candidate whose clipped point lies within HF_TWIN of an EARLIER in-threshold
  // candidate is a copy of the same point seen through a neighbouring triangle or face; it is masked out, so the arg-max passes
  // below never choose between copies that differ only by rounding and by the triangle normal they carry.
  const int nch = (nin + 31) >> 5;                                       // chunks in use (warp-uniform)
  float dm[HF_NIN / 32];
  {
    bool tw[HF_NIN / 32];
#pragma unroll
    for (int c = 0; c < HF_NIN / 32; ++c) tw[c] = false;
#pragma unroll 4
    for (int i = 0; i < nin; ++i) {                                      // warp-uniform
      const float ox = scr[3 * i], oy = scr[3 * i + 1], oz = scr[3 * i + 2];
#pragma unroll
      for (int c = 0; c < HF_NIN / 32; ++c)
        if (c < nch && i < lane + 32 * c && fmaxf(fabsf(Pc[c].x - ox), fmaxf(fabsf(Pc[c].y - oy), fabsf(Pc[c].z - oz))) < HF_TWIN) tw[c] = true;
    }
#pragma unroll
    for (int c = 0; c < HF_NIN / 32; ++c) dm[c] = (ok[c] && !tw[c]) ? 0.f : ninf;
  }
  auto point = [&](int sl) {                                             // midway point of slot sl, from the lane that holds it
    const int c = sl >> 5;
    const V3 v = c == 0 ? Pm[0] : (c == 1 ? Pm[1] : Pm[2]);
    return shfl3(v, sl & 31);
  };
  int idx[4] = {0, 0, 0, 0};
#pragma unroll
  for (int c = HF_NIN / 32 - 1; c >= 0; --c) {                           // a: the first candidate left
    const unsigned bm = __ballot_sync(FULLMASK, dm[c] == 0.f);
    if (bm) idx[0] = 32 * c + __ffs(bm) - 1;
  }
  const V3 pa = point(idx[0]);
  {
    float best = ninf;                                                   // b: farthest from a
#pragma unroll
    for (int c = 0; c < HF_NIN / 32; ++c) if (c < nch) {
      const V3 ap = pa - Pm[c];
      int li; const float v = wargmax_val(dot(ap, ap) + dm[c], lane, &li);
      if (v > best) { best = v; idx[1] = 32 * c + li; }
    }
  }
  const V3 pb = point(idx[1]);
  const V3 ab = cross(nm, pa - pb);
  {
    float best = ninf;                                                   // c: farthest from the line a b
#pragma unroll
    for (int c = 0; c < HF_NIN / 32; ++c) if (c < nch) {
      int li; const float v = wargmax_val(fabsf(dot(pa - Pm[c], ab)) + dm[c], lane, &li);
      if (v > best) { best = v; idx[2] = 32 * c + li; }
    }
  }
  const V3 pc = point(idx[2]);
  const V3 ac = cross(nm, pa - pc), bc = cross(nm, pb - pc);
  {
    float b1 = ninf, b2 = ninf; int i1 = 0, i2 = 0;                     // d: farthest from the edges b c and a c
#pragma unroll
    for (int c = 0; c < HF_NIN / 32; ++c) if (c < nch) {
      int li; float v = wargmax_val(fabsf(dot(pb - Pm[c], bc)) + dm[c], lane, &li);
      if (v > b1) { b1 = v; i1 = 32 * c + li; }
      v = wargmax_val(fabsf(dot(pa - Pm[c], ac)) + dm[c], lane, &li);
      if (v > b2) { b2 = v; i2 = 32 * c + li; }
    }
    idx[3] = b1 >= b2 ? i1 : i2;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    bool uniq = true;
#pragma unroll
    for (int p = 0; p < c; ++p) uniq = uniq && (idx[p] != idx[c]);
    if (uniq && lane == c) {
      const float4* rec = reinterpret_cast<const float4*>(cand + cidx[idx[c]] * HF_REC);
      const float4 a = rec[0], b = rec[1];
      float* cc = s.con[4 * f + c];
      cc[0] = a.x; cc[1] = a.y; cc[2] = a.z; cc[3] = a.w; cc[13] = b.x; cc[14] = b.y; cc[15] = b.z;
    }
  }
  __syncwarp();
}
