// CPU emulation of the physics substep of liboduck_cuda (TEST INFRASTRUCTURE, tests/test_step_emu.py): forward_euler of
// csrc/oduck_physics.cuh -- the device code k_physics / k_step run -- compiled for the host against tests/emu/cuda_runtime.h and
// executed by 32 threads, one per lane.  Same body as k_physics<false, HF> for ONE env with the nominal (un-randomised) model.
#include <cuda_runtime.h>

#include <thread>
#include <vector>

#include "../../open_duck_playground_b200/csrc/oduck_build.h"
#include "../../open_duck_playground_b200/csrc/oduck_physics.cuh"

extern "C" int emu_exchange_counts(long long* out, int reset) { for (int i = 0; i < 16; i++) { out[i] = warp_emu::exchanges[i]; if (reset) warp_emu::exchanges[i] = 0; } return 0; }
extern "C" int emu_strides(int* out) { out[0] = PHYS_STRIDE; out[1] = DR_STRIDE; out[2] = OUT_STRIDE; out[3] = PHYS_QVEL; out[4] = PHYS_QACCW; out[5] = PHYS_CTRL; out[6] = OUT_QACC; out[7] = OUT_SENS; out[8] = OUT_EFC; out[9] = OUT_CDIST; out[10] = OUT_AFRC; return 0; }

// phys [PHYS_STRIDE] in/out (qpos | qvel | qacc_warm | ctrl), ctrl [nu] or null, out [OUT_STRIDE].  Returns 0, or -1 with a bad model.
extern "C" int emu_physics(const OduckModel* model, int nsub, int integrate, float* phys, const float* ctrl, float* out) {
  static DevModel m;
  static DevFF ff;
  static DevHF hf;
  static WarpSmem s;
  std::string err;
  if (build_dev_model(*model, m, err) != 0) return -1;
  build_dev_ff(*model, ff);
  std::memset(&s, 0, sizeof(s));
  const bool HFIELD = model->floor_is_hfield != 0;
  if (HFIELD) {
    hf.nrow = model->hfield_nrow; hf.ncol = model->hfield_ncol;
    hf.sx = (float)model->hfield_size[0]; hf.sy = (float)model->hfield_size[1]; hf.sz = (float)model->hfield_size[2];
    hf.dx = (float)(2.0 * model->hfield_size[0] / (model->hfield_ncol - 1)); hf.dy = (float)(2.0 * model->hfield_size[1] / (model->hfield_nrow - 1));
    hf.data = model->hfield_data;
  }
  // nominal per-env model record, as oduck_create fills it
  std::vector<float> dr(DR_STRIDE, 0.f);
  for (int i = 0; i < model->nq; i++) dr[DR_QPOS0 + i] = (float)model->qpos0[i];
  for (int u = 0; u < model->nu; u++) dr[DR_KP + u] = (float)model->act_kp[u];
  for (int b = 0; b < model->nbody; b++) dr[b] = (float)model->body_mass[b];
  for (int i = 0; i < 3; i++) dr[DR_IPOS1 + i] = (float)model->body_ipos[1][i];
  dr[DR_FRIC0] = 1.f;
  for (int d = 0; d < model->nv; d++) { dr[DR_FLOSS + d] = (float)model->dof_frictionloss[d]; dr[DR_ARM + d] = (float)model->dof_armature[d]; }
  std::vector<float> ffs(FFJ_SIZE + FFV_SIZE, 0.f), hfs(HF_SCRATCH, 0.f);
  std::barrier<> bar(32);
  warp_emu::bar = &bar;
  std::vector<std::thread> th;
  for (int l = 0; l < 32; l++)
    th.emplace_back([&, l]() {
      warp_emu::lane = l;
      const int lane = l;
      Lane L;
      load_env(m, s, L, lane, phys, dr.data());
      if (ctrl) { const int a = m.d_act[lane]; if (a >= 0) L.ctrl = ctrl[a]; }
      for (int i = lane; i < OUT_STRIDE; i += 32) s.outrec[i] = 0.f;
      __syncwarp();
      for (int k = 0; k < nsub; ++k) {
        if (HFIELD) forward_euler<false, false, true>(m, s, L, lane, k == nsub - 1, integrate != 0, s.outrec, nullptr, &ff, ffs.data(), &hf, hfs.data());
        else forward_euler<false, false, false>(m, s, L, lane, k == nsub - 1, integrate != 0, s.outrec, nullptr, &ff, ffs.data(), nullptr, nullptr);
      }
      __syncwarp();
      store_phys(m, s, L, lane, phys);
      store_out(s, lane, out);
      __syncwarp();
    });
  for (auto& t : th) t.join();
  return 0;
}
