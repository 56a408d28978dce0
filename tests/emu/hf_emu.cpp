// CPU emulation of csrc/oduck_hfcollide.cuh hf_collide (TEST INFRASTRUCTURE, tests/test_hf_emu.py): the device routine is
// compiled for the host against tests/emu/cuda_runtime.h and run by 32 threads, one per lane.  It checks the LOGIC of the
// warp-level collider (and of its -DODUCK_HF_CULL / -DODUCK_HF_PAIRS variants) against the oracle without a GPU.
#include <cuda_runtime.h>

#include <thread>
#include <vector>

#ifdef ODUCK_HF_STATS   // tools/hf_stats.py: triangles after the height cull / after the plane-side cull, pairs, candidates per call
static long long g_hf_stat[8];
#define ODUCK_HF_STAT(what, n) { if (lane == 0) g_hf_stat[what] += (n); }
extern "C" long long* emu_hf_stats(void) { return g_hf_stat; }
static float g_hf_last_cand[1 << 14];
extern "C" float* emu_hf_last_candidates(void) { return g_hf_last_cand; }   // the candidate list of the last call
#endif
#include "../../open_duck_playground_b200/csrc/oduck_hfcollide.cuh"

extern "C" int emu_hf_scratch_floats(void) { return HF_SCRATCH; }

// One foot.  vert [nvert][3] (body frame), plane_nvert [nplane], plane_vert [nplane][8], plane_normal [nplane][3] (body frame),
// xpos [3], xmat [9] (row-major) of the foot body, hfield data [nrow][ncol], size = (sx, sy, sz).  out [4][8]: dist, pos[3], normal[3], -
extern "C" int emu_hf_collide(const float* xpos, const float* xmat, const float* vert, int nvert, int nplane, const int* plane_nvert,
                              const int* plane_vert, const float* plane_normal, const float* center, float radius, int nrow, int ncol,
                              const float* size, const float* data, float* out) {
  static DevModel m;
  static WarpSmem s;
  static DevFF ff;
  static DevHF hf;
  std::memset(&m, 0, sizeof(m)); std::memset(&s, 0, sizeof(s)); std::memset(&ff, 0, sizeof(ff));
  const int f = 0, body = 1;
  m.nvert = nvert; m.foot_body[0] = body; m.foot_body[1] = body;
  for (int v = 0; v < nvert; v++) for (int i = 0; i < 3; i++) m.vert[f][i][v] = vert[3 * v + i];
  for (int i = 0; i < 3; i++) s.xpos[i][body] = xpos[i];
  for (int i = 0; i < 9; i++) s.xmat[i][body] = xmat[i];
  ff.nplane = nplane; ff.nvert = nvert; ff.radius = radius;
  for (int q = 0; q < nplane; q++) {
    ff.plane_nvert[q] = plane_nvert[q];
    for (int k = 0; k < 8; k++) ff.plane_vert[q][k] = plane_vert[8 * q + k];
    for (int i = 0; i < 3; i++) ff.plane_normal[f][i][q] = plane_normal[3 * q + i];
  }
  for (int i = 0; i < 3; i++) ff.center[f][i] = center[i];
  hf.nrow = nrow; hf.ncol = ncol; hf.sx = size[0]; hf.sy = size[1]; hf.sz = size[2];
  hf.dx = (float)(2.0 * size[0] / (ncol - 1)); hf.dy = (float)(2.0 * size[1] / (nrow - 1));
  hf.data = data;
  std::vector<float> cand(HF_SCRATCH, 0.f);
  std::barrier<> bar(32);
  warp_emu::bar = &bar;
  std::vector<std::thread> th;
  for (int l = 0; l < 32; l++)
    th.emplace_back([&, l]() {
      warp_emu::lane = l;
      hf_collide(m, &ff, &hf, s, l, f, cand.data());
    });
  for (auto& t : th) t.join();
#ifdef ODUCK_HF_STATS
  std::memcpy(g_hf_last_cand, cand.data(), sizeof(float) * HF_SCRATCH);
#endif
  for (int c = 0; c < 4; c++) {
    const float* cc = s.con[c];
    float* o = out + 8 * c;
    o[0] = cc[0]; o[1] = cc[1]; o[2] = cc[2]; o[3] = cc[3]; o[4] = cc[13]; o[5] = cc[14]; o[6] = cc[15]; o[7] = 0.f;
  }
  return 0;
}
