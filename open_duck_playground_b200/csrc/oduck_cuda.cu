// oduck_cuda.cu -- liboduck_cuda.so: the B200 (sm_100a) implementation of include/oduck.h.
// One warp per env; all per-timestep work of the reference's Joystick.step (joystick.py:323-481), including the ten
// mjx.step substeps it calls at joystick.py:420, runs inside ONE kernel launch per env.step.  No CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/oduck.h"
#include "oduck_env.cuh"
#ifndef ODUCK_WARP_EMU
#include "oduck_policy_tc.cuh"   // mbarrier helpers
#endif

#ifndef WPB
// warps (= envs in flight) per CTA.  16 = ONE CTA per SM: the substep is ~9 k straight-line instructions (140 KB) that stream through
// the instruction cache once per substep and convoy, and a CTA's warps are kept in one convoy by the substep barrier; two 8-warp
// CTAs per SM (rounds 1 and 2a) are two convoys.  Measured on B200 (profiles/r02i_bench_*.json): flat 7.24 -> 7.53 M env-steps/s,
// rough terrain 8.28 -> 6.58 ms per 16384-env step.
#define WPB 16
#endif
#ifndef CTAS_PER_SM
#define CTAS_PER_SM (WPB <= 8 ? 2 : 1)
#endif
// Thread ids behind an opaque move: ptxas otherwise rematerialises lane / warp / the warp's smem base (S2R + shifts + IMAD)
// dozens of times per substep instead of keeping them in registers.
__device__ __forceinline__ int opaque(int x) {
#ifdef ODUCK_OPAQUE_IDS
  asm volatile("mov.b32 %0, %0;" : "+r"(x));
#endif
  return x;
}

// dynamic shared memory of a kernel (tests/emu runs the kernels on CPU threads and hands them a heap block instead)
#ifdef ODUCK_WARP_EMU
#define ODUCK_SMEM_RAW(name) unsigned char* name = warp_emu::smem
#else
#define ODUCK_SMEM_RAW(name) extern __shared__ __align__(16) unsigned char name[]
#endif

static thread_local std::string g_err;
int oduck_fail(int code, const std::string& msg) { g_err = msg; return code; }   // shared with oduck_policy.cu
static int fail(int code, const std::string& msg) { return oduck_fail(code, msg); }
#define CUDA_TRY(x)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (x);                                                                             \
    if (e_ != cudaSuccess) return fail(ODUCK_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

#include "oduck_handle.cuh"

// ------------------------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ size_t smem_model_bytes() { return (sizeof(DevModel) + 15) & ~(size_t)15; }
__device__ __forceinline__ size_t smem_cfg_bytes() { return (sizeof(DevEnvCfg) + 15) & ~(size_t)15; }

// Batch-shared tables (DevModel ~18 KB, DevEnvCfg) -> shared memory: two TMA bulk copies issued by one thread, completion on
// an mbarrier (cp.async.bulk + expect_tx; SASS UBLKCP) instead of an 18-trip load/store loop in all 256 threads.
__device__ __forceinline__ void block_load_tables(const Params& p, unsigned char* raw, DevModel*& m, DevEnvCfg*& c, WarpSmem*& ws) {
  m = reinterpret_cast<DevModel*>(raw);
  c = reinterpret_cast<DevEnvCfg*>(raw + smem_model_bytes());
  ws = reinterpret_cast<WarpSmem*>(raw + smem_model_bytes() + smem_cfg_bytes());
#ifdef ODUCK_WARP_EMU
  if (threadIdx.x == 0) { memcpy(m, p.model, sizeof(DevModel)); memcpy(c, p.cfg, sizeof(DevEnvCfg)); }
  __syncwarp();
#else
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"((uint32_t)(sizeof(DevModel) + sizeof(DevEnvCfg))) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(m)), "l"(p.model), "r"((uint32_t)sizeof(DevModel)), "r"(smem_u32(&bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(c)), "l"(p.cfg), "r"((uint32_t)sizeof(DevEnvCfg)), "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();                                     // the barrier is initialised before anyone polls it
  mbar_wait(&bar, 0u);
#endif
}

// ------------------------------------------------------------------------------------------------- kernels
// A8 / A9 of SURVEY 8a: n x (forward + euler), or one forward without integration (mjx_env.init's forward).
template <bool DBG, bool HF>
__global__ void __launch_bounds__(WPB * 32, CTAS_PER_SM) k_physics(Params p) {
  ODUCK_SMEM_RAW(raw);
  DevModel* mp; DevEnvCfg* cp; WarpSmem* ws;
  block_load_tables(p, raw, mp, cp, ws);
  const DevModel& m = *mp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmem& s = ws[warp];
  for (int env0 = blockIdx.x * WPB; env0 < p.N; env0 += gridDim.x * WPB) {
    const int env = env0 + warp;
    if (env >= p.N) continue;                                  // grid tail: the CTA's other warps synchronise among themselves (L.bar_n)
    Lane L;
    L.bar_n = 32 * min(WPB, p.N - env0);
    float* ph = p.phys + (size_t)env * PHYS_STRIDE;
    load_env(m, s, L, lane, ph, p.dr + (size_t)env * DR_STRIDE);
    if (p.action) { const int a = m.d_act[lane]; if (a >= 0) L.ctrl = p.action[(size_t)env * m.nu + a]; }
    for (int i = lane; i < OUT_STRIDE; i += 32) s.outrec[i] = 0.f;
    __syncwarp();
    float* dbg = DBG ? p.dbg + (size_t)env * DBG_STRIDE : nullptr;
    // like k_step: one CTA barrier per substep keeps the CTA's warps on the same stretch of the instruction stream
    for (int k = 0; k < p.nsub; ++k) forward_euler<DBG, !DBG, HF>(m, s, L, lane, k == p.nsub - 1, p.integrate != 0, s.outrec, dbg, p.ffmodel, p.ffscratch + (size_t)env * (FFJ_SIZE + FFV_SIZE), p.hfmodel, HF ? p.hfscratch + (size_t)env * HF_SCRATCH : nullptr);
    __syncwarp();
    store_phys(m, s, L, lane, ph);
    store_out(s, lane, p.out + (size_t)env * OUT_STRIDE);
    __syncwarp();
  }
}

__device__ __forceinline__ void load_info(const DevModel& m, const DevEnvCfg& c, const float* __restrict__ inf, EnvRegs& er, int lane) {
  const int* ii = reinterpret_cast<const int*>(inf);
  er.rng.a = (uint32_t)ii[INFO_RNG]; er.rng.b = (uint32_t)ii[INFO_RNG + 1];
  er.step = ii[INFO_STEP]; er.steps = ii[INFO_STEPS]; er.push_step = ii[INFO_PUSH_STEP]; er.push_interval = ii[INFO_PUSH_INT];
  er.imitation_i = ii[INFO_IMIT_I];
  er.cmd = lane < 7 ? inf[INFO_CMD + lane] : 0.f;
  const bool u = lane < m.nu;
#pragma unroll
  for (int k = 0; k < 3; ++k) er.last_act[k] = u ? inf[INFO_LAST_ACT + 16 * k + lane] : 0.f;
  er.targets = u ? inf[INFO_TARGETS + lane] : 0.f;
#pragma unroll
  for (int k = 0; k < MAX_DELAY; ++k) er.hist[k] = (u && k < c.act_max_delay) ? inf[INFO_AHIST + k * m.nu + lane] : 0.f;
  er.air = lane < 2 ? inf[INFO_AIR + lane] : 0.f;
  er.lastc = lane < 2 ? inf[INFO_LASTC + lane] : 0.f;
  er.swing = lane < 2 ? inf[INFO_SWING + lane] : 0.f;
  er.ref_lo = inf[INFO_REF + lane];
  er.ref_hi = inf[INFO_REF + 32 + (lane & 7)];
  er.phase = lane < 2 ? inf[INFO_PHASE + lane] : 0.f;
}
__device__ __forceinline__ void store_info(const DevModel& m, const DevEnvCfg& c, float* __restrict__ inf, const EnvRegs& er, float push, int lane) {
  int* ii = reinterpret_cast<int*>(inf);
  if (lane == 0) {
    ii[INFO_RNG] = (int)er.rng.a; ii[INFO_RNG + 1] = (int)er.rng.b;
    ii[INFO_STEP] = er.step; ii[INFO_STEPS] = er.steps; ii[INFO_PUSH_STEP] = er.push_step; ii[INFO_PUSH_INT] = er.push_interval;
    ii[INFO_IMIT_I] = er.imitation_i;
  }
  if (lane < 7) inf[INFO_CMD + lane] = er.cmd;
  if (lane < m.nu) {
#pragma unroll
    for (int k = 0; k < 3; ++k) inf[INFO_LAST_ACT + 16 * k + lane] = er.last_act[k];
    inf[INFO_TARGETS + lane] = er.targets;
#pragma unroll
    for (int k = 0; k < MAX_DELAY; ++k) if (k < c.act_max_delay) inf[INFO_AHIST + k * m.nu + lane] = er.hist[k];
  }
  if (lane < 2) {
    inf[INFO_AIR + lane] = er.air; inf[INFO_LASTC + lane] = er.lastc; inf[INFO_SWING + lane] = er.swing;
    inf[INFO_PUSH + lane] = push; inf[INFO_PHASE + lane] = er.phase;
  }
  inf[INFO_REF + lane] = er.ref_lo;
  if (lane < 8) inf[INFO_REF + 32 + lane] = er.ref_hi;
}
__device__ __forceinline__ float feet_contact(const WarpSmem& s, int lane) {   // geoms_colliding, lane k < 2
  const int k = lane & 1;
  const float* d = s.outrec + OUT_CDIST + 4 * k;
  return fminf(fminf(d[0], d[1]), fminf(d[2], d[3])) < 0.f ? 1.f : 0.f;
}

// A2 (+A9, first_state store): Joystick.reset for the masked envs.
template <bool HF>
__global__ void __launch_bounds__(WPB * 32, CTAS_PER_SM) k_reset(Params p) {
  ODUCK_SMEM_RAW(raw);
  DevModel* mp; DevEnvCfg* cp; WarpSmem* ws;
  block_load_tables(p, raw, mp, cp, ws);
  const DevModel& m = *mp;
  const DevEnvCfg& c = *cp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmem& s = ws[warp];
  for (int env = blockIdx.x * WPB + warp; env < p.N; env += gridDim.x * WPB) {
    if (p.mask && !p.mask[env]) continue;
    Lane L;
    float* ph = p.phys + (size_t)env * PHYS_STRIDE;
    load_env(m, s, L, lane, ph, p.dr + (size_t)env * DR_STRIDE);
    // six chained splits (joystick.py:213-263): dxy, yaw, joints, base qvel, command, push interval
    RKey rng; rng.a = p.keys[2 * env]; rng.b = p.keys[2 * env + 1];
    RKey ks[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { RKey o = rblock(rng, lane & 1); ks[k] = kshfl(o, 1); rng = kshfl(o, 0); }
    // one round of draws: lanes 0-1 dxy | 2 yaw | 3 push interval | 4..9 base qvel | 10..10+nu joints
    RKey kk; uint32_t cnt;
    if (lane < 2) { kk = ks[0]; cnt = lane; }
    else if (lane == 2) { kk = ks[1]; cnt = 0; }
    else if (lane == 3) { kk = ks[5]; cnt = 0; }
    else if (lane < 10) { kk = ks[3]; cnt = lane - 4; }
    else { kk = ks[2]; cnt = lane - 10; }
    RKey b = rblock(kk, cnt);
    const uint32_t bits = b.a ^ b.b;
    if (lane < m.nq) s.qpos[lane] = m.key_qpos[lane];
    if (lane + 32 < m.nq) s.qpos[lane + 32] = m.key_qpos[lane + 32];
    __syncwarp();
    if (lane < 2) s.qpos[lane] = m.key_qpos[lane] + bits_uniform(bits, -0.05f, 0.05f);
    if (lane == 2) {
      const float yaw = bits_uniform(bits, -3.14f, 3.14f);
      Q4 q0; q0.w = m.key_qpos[3]; q0.x = m.key_qpos[4]; q0.y = m.key_qpos[5]; q0.z = m.key_qpos[6];
      Q4 qn = qmul(q0, axis_angle(v3(0.f, 0.f, 1.f), yaw));
      s.qpos[3] = qn.w; s.qpos[4] = qn.x; s.qpos[5] = qn.y; s.qpos[6] = qn.z;
    }
    if (lane >= 10 && lane < 10 + m.nu) { const int qa = m.act_qadr[lane - 10]; s.qpos[qa] = m.key_qpos[qa] * bits_uniform(bits, 0.5f, 1.5f); }
    const float bq = bits_uniform(bits, -c.reset_qvel_noise, c.reset_qvel_noise);
    L.qvel = 0.f;
    {
      const float v = __shfl_sync(FULLMASK, bq, 4 + (lane < 6 ? lane : 0));
      if (lane < 6) L.qvel = v;
    }
    const float push_iv = __shfl_sync(FULLMASK, bits_uniform(bits, c.push_interval[0], c.push_interval[1]), 3);
    L.qaccw = 0.f;
    __syncwarp();
    { const int a = m.d_act[lane]; L.ctrl = a >= 0 ? s.qpos[m.act_qadr[a]] : 0.f; }
    for (int i = lane; i < OUT_STRIDE; i += 32) s.outrec[i] = 0.f;
    __syncwarp();
    forward_euler<false, false, HF>(m, s, L, lane, true, false, s.outrec, nullptr, p.ffmodel, p.ffscratch + (size_t)env * (FFJ_SIZE + FFV_SIZE), p.hfmodel, HF ? p.hfscratch + (size_t)env * HF_SCRATCH : nullptr);   // mjx_env.init
    __syncwarp();
    EnvRegs er;
    er.rng = rng;
    er.step = 0; er.steps = 0; er.push_step = 0; er.imitation_i = 0;
    er.push_interval = (int)rintf(push_iv / c.ctrl_dt);
    er.cmd = sample_command(c, ks[4], lane);
    if (lane >= 7) er.cmd = 0.f;
    er.last_act[0] = er.last_act[1] = er.last_act[2] = 0.f;
    er.targets = (lane < m.nu && c.task != ODUCK_TASK_STANDING) ? m.key_ctrl[lane] : 0.f;   // standing.py:279 starts them at zero
#pragma unroll
    for (int k = 0; k < MAX_DELAY; ++k) er.hist[k] = 0.f;
    er.air = er.lastc = er.swing = 0.f;
    er.phase = 0.f;
    er.ref_lo = er.ref_hi = 0.f;
    if (c.use_imitation) {
      const float c0 = __shfl_sync(FULLMASK, er.cmd, 0), c1 = __shfl_sync(FULLMASK, er.cmd, 1), c2 = __shfl_sync(FULLMASK, er.cmd, 2);
      reference_motion(c, p.poly, c0, c1, c2, 0, lane, er.ref_lo, er.ref_hi);
    }
    float* inf = p.info + (size_t)env * INFO_STRIDE;
    if (lane < c.imu_max_delay * 3) inf[INFO_IMUHIST + lane] = 0.f;
    __syncwarp();
    const float contact = feet_contact(s, lane);
    float* ost = p.obs_state + (size_t)env * ODUCK_OBS_STATE;
    float* opr = p.obs_priv + (size_t)env * ODUCK_OBS_PRIV;
    write_obs(m, c, s, s.outrec, L, er, contact, lane, ost, opr, inf + INFO_IMUHIST);
    store_info(m, c, inf, er, 0.f, lane);
    if (lane == 0) { p.reward[env] = 0.f; p.done[env] = 0.f; p.trunc[env] = 0.f; }
    if (lane < ODUCK_NMETRIC) p.metrics[(size_t)env * ODUCK_NMETRIC + lane] = 0.f;
    store_phys(m, s, L, lane, ph);
    store_out(s, lane, p.out + (size_t)env * OUT_STRIDE);
    __syncwarp();
    // BraxAutoResetWrapper.reset: first_state, first_obs
    float* fp = p.first_phys + (size_t)env * PHYS_STRIDE;
    for (int i = lane; i < PHYS_STRIDE; i += 32) fp[i] = ph[i];
    float* fs_ = p.first_obs_state + (size_t)env * ODUCK_OBS_STATE;
    for (int i = lane; i < ODUCK_OBS_STATE; i += 32) fs_[i] = ost[i];
    float* fpv = p.first_obs_priv + (size_t)env * ODUCK_OBS_PRIV;
    for (int i = lane; i < ODUCK_OBS_PRIV; i += 32) fpv[i] = opr[i];
    __syncwarp();
  }
}

// A1 + A16: Joystick.step fused with EpisodeWrapper + AutoResetWrapper.  HF = height-field floor (rough_terrain scenes).
// RL = reward-library terms switched on (OduckEnvConfig.lib; the shipped tasks use none).
template <bool HF, bool RL>
__global__ void __launch_bounds__(WPB * 32, CTAS_PER_SM) k_step(Params p) {
  ODUCK_SMEM_RAW(raw);
  DevModel* mp; DevEnvCfg* cp; WarpSmem* ws;
  block_load_tables(p, raw, mp, cp, ws);
  const DevModel& m = *mp;
  const DevEnvCfg& c = *cp;
  const int warp = opaque(threadIdx.x >> 5), lane = opaque(threadIdx.x & 31);
  WarpSmem& s = ws[warp];
  const float PI = 3.14159265358979323846f;
  for (int env0 = blockIdx.x * WPB; env0 < p.N; env0 += gridDim.x * WPB) {
    const int env = env0 + warp;
    if (env >= p.N) continue;                                  // grid tail: the CTA's other warps synchronise among themselves (L.bar_n)
    Lane L;
    L.bar_n = 32 * min(WPB, p.N - env0);
    float* ph = p.phys + (size_t)env * PHYS_STRIDE;
    load_env(m, s, L, lane, ph, p.dr + (size_t)env * DR_STRIDE);
    float* inf = p.info + (size_t)env * INFO_STRIDE;
    EnvRegs er;
    load_info(m, c, inf, er, lane);
    const int nu = m.nu;
    const float act = lane < nu ? p.action[(size_t)env * nu + lane] : 0.f;
    if (p.done[env] != 0.f) er.steps = 0;                       // AutoResetWrapper.step prologue
    const float dt = c.ctrl_dt;
    const float c0 = __shfl_sync(FULLMASK, er.cmd, 0), c1 = __shfl_sync(FULLMASK, er.cmd, 1), c2 = __shfl_sync(FULLMASK, er.cmd, 2);
    if (c.use_imitation) {
      er.imitation_i = (er.imitation_i + 1) % c.nb_steps;
      const float phs = ((float)er.imitation_i / (float)c.nb_steps) * 2.f * PI;
      er.phase = lane == 0 ? cosf(phs) : (lane == 1 ? sinf(phs) : 0.f);
      reference_motion(c, p.poly, c0, c1, c2, er.imitation_i, lane, er.ref_lo, er.ref_hi);
    } else {
      er.imitation_i = 0;
    }
    // rng, push1, push2, delay = split(rng, 4)
    RKey push1, push2, delay;
    { RKey o = rblock(er.rng, lane & 3); push1 = kshfl(o, 1); push2 = kshfl(o, 2); delay = kshfl(o, 3); er.rng = kshfl(o, 0); }
    // action delay (joystick.py:361-376)
#pragma unroll
    for (int k = MAX_DELAY - 1; k >= 1; --k) er.hist[k] = er.hist[k - 1];
    er.hist[0] = act;
    const int aidx = warp_randint(delay, c.act_min_delay, c.act_max_delay, lane);
    float act_d = er.hist[0];
#pragma unroll
    for (int k = 1; k < MAX_DELAY; ++k) if (aidx == k) act_d = er.hist[k];
    // push (joystick.py:381-400)
    float pushv = 0.f;
    {
      RKey b = rblock(lane == 0 ? push1 : push2, 0u);
      const uint32_t bits = b.a ^ b.b;
      const float theta = bits_uniform(__shfl_sync(FULLMASK, bits, 0), 0.f, 2.f * PI);
      const float mag = bits_uniform(__shfl_sync(FULLMASK, bits, 1), c.push_magnitude[0], c.push_magnitude[1]);
      const bool fire = ((er.push_step + 1) % er.push_interval) == 0 && c.push_enable;
      if (lane < 2) { pushv = fire ? (lane == 0 ? cosf(theta) : sinf(theta)) : 0.f; L.qvel += pushv * mag; }
    }
    // motor targets with speed limit (joystick.py:404-417)
    float tgt = (lane < nu ? m.key_ctrl[lane] : 0.f) + act_d * c.action_scale;
    if (c.use_speed_limits) { const float lim = c.max_motor_velocity * dt; tgt = fminf(fmaxf(tgt, er.targets - lim), er.targets + lim); }
    { const int a = m.d_act[lane]; const float t = __shfl_sync(FULLMASK, tgt, a < 0 ? 0 : a); if (a >= 0) L.ctrl = t; }
    // physics: n_substeps x mjx.step (joystick.py:420)
    for (int k = 0; k < c.n_substeps; ++k)
      forward_euler<false, true, HF>(m, s, L, lane, k == c.n_substeps - 1, true, s.outrec, nullptr, p.ffmodel, p.ffscratch + (size_t)env * (FFJ_SIZE + FFV_SIZE), p.hfmodel, HF ? p.hfscratch + (size_t)env * HF_SCRATCH : nullptr);
    __syncwarp();
    er.targets = tgt;
    // contacts / air time / swing peak (joystick.py:424-435)
    const float contact = feet_contact(s, lane);
    float first_contact = 0.f;                                       // lane k < 2 (joystick.py:430-431); only the library terms read it
    if constexpr (RL) first_contact = (er.air > 0.f && (contact != 0.f || er.lastc != 0.f)) ? 1.f : 0.f;
    er.air += dt;
    if (lane < 2) er.swing = fmaxf(er.swing, s.outrec[OUT_FEET + 3 * lane + 2]);
    float* ost = p.obs_state + (size_t)env * ODUCK_OBS_STATE;
    float* opr = p.obs_priv + (size_t)env * ODUCK_OBS_PRIV;
    write_obs(m, c, s, s.outrec, L, er, contact, lane, ost, opr, inf + INFO_IMUHIST);
    // termination (joystick.py:483-485)
    bool bad = (lane < m.nq && isnan(s.qpos[lane])) || (lane + 32 < m.nq && isnan(s.qpos[lane + 32])) || isnan(L.qvel);
    const bool done = (s.outrec[OUT_SENS + 11] < 0.f) || (__ballot_sync(FULLMASK, bad) != 0u);
    // rewards (joystick.py:622-669)
    const float* sd = s.outrec + OUT_SENS;
    float q = 0.f, afrc = 0.f, dflt = 0.f;
    if (lane < nu) { q = s.qpos[m.act_qadr[lane]]; afrc = s.outrec[OUT_AFRC + lane]; dflt = m.key_ctrl[lane]; }
    const float qdv = __shfl_sync(FULLMASK, L.qvel, lane < nu ? m.act_dof[lane] : 0);
    const float qd = lane < nu ? qdv : 0.f;
    const float ex = (c0 - sd[3]) * (c0 - sd[3]);
    const float ey = fmaxf(fabsf(sd[4] - c1) - 0.1f, 0.f);
    const float r_lin = nan_to_num(expf(-(ex + ey * ey) / c.tracking_sigma));
    const float r_ang = nan_to_num(expf(-((c2 - sd[2]) * (c2 - sd[2])) / c.tracking_sigma));
    const float c_torque = nan_to_num(wsum(afrc * afrc));
    const float dact = act - er.last_act[0];
    const float c_rate = nan_to_num(wsum(lane < nu ? dact * dact : 0.f));
    const float cmd_norm = sqrtf(c0 * c0 + c1 * c1 + c2 * c2);
    const float c_still = nan_to_num(wsum(lane < nu ? fabsf(q - dflt) + fabsf(qd) : 0.f)) * (cmd_norm < 0.01f ? 1.f : 0.f);
    float r_imit = 0.f;
    if (c.use_imitation) {
      // base velocity error vs ref[34:40]; lanes 0..5
      const float bv = L.qvel;                                        // lane d < 6: floating-base qvel (post-integration)
      const float rv = __shfl_sync(FULLMASK, er.ref_hi, 2 + (lane < 6 ? lane : 0));
      const float e = lane < 6 ? (bv - rv) * (bv - rv) : 0.f;
      const float lxy = __shfl_sync(FULLMASK, e, 0) + __shfl_sync(FULLMASK, e, 1), lz = __shfl_sync(FULLMASK, e, 2);
      const float axy = __shfl_sync(FULLMASK, e, 3) + __shfl_sync(FULLMASK, e, 4), az = __shfl_sync(FULLMASK, e, 5);
      // leg joints: actuator lanes 0..4 and 9..13 vs ref[rr], ref[16 + rr]
      const bool leg = lane < 5 || (lane >= 9 && lane < 14);
      const int rr = lane < 5 ? lane : lane + 2;
      const float rp = __shfl_sync(FULLMASK, er.ref_lo, leg ? rr : 0), rvv = __shfl_sync(FULLMASK, er.ref_lo, leg ? 16 + rr : 0);
      const float jp_ = wsum(leg ? (q - rp) * (q - rp) : 0.f), jv_ = wsum(leg ? (qd - rvv) * (qd - rvv) : 0.f);
      const float rc = er.ref_hi > 0.5f ? 1.f : 0.f;                  // lane k < 2: ref[32 + k]
      const float cm = (lane < 2 && contact == rc) ? 1.f : 0.f;
      const float crew = __shfl_sync(FULLMASK, cm, 0) + __shfl_sync(FULLMASK, cm, 1);
      float rew = expf(-8.f * lxy) + expf(-8.f * lz) + 0.5f * expf(-2.f * axy) + 0.5f * expf(-2.f * az) - 15.f * jp_ - 1e-3f * jv_ + crew;
      rew *= cmd_norm > 0.01f ? 1.f : 0.f;
      r_imit = nan_to_num(rew);
    }
    float scs[7], sg[7], total;
    if (c.task == ODUCK_TASK_STANDING) {
      // standing.py:573-606: orientation, torques, action_rate, alive, stand_still (legs only), head_pos
      const float c_or = nan_to_num(sd[9] * sd[9] + sd[10] * sd[10]);
      const bool legl = lane < 5 || (lane >= 9 && lane < nu);
      const float c_still_legs = nan_to_num(wsum(legl ? fabsf(q - dflt) + fabsf(qd) : 0.f)) * (cmd_norm < 0.01f ? 1.f : 0.f);
      const float hc = __shfl_sync(FULLMASK, er.cmd, (lane >= 5 && lane < 9) ? lane - 2 : 0);       // head joint u <- command[3 + u - 5]
      const float c_head = nan_to_num(wsum((lane >= 5 && lane < 9) ? (q - hc) * (q - hc) : 0.f)) * (cmd_norm > 0.01f ? 1.f : 0.f);
      const float s_or = c_or * c.sc_orient, s_tq = c_torque * c.sc_torques, s_ar = c_rate * c.sc_rate, s_al = 1.f * c.sc_alive;
      const float s_ss = c_still_legs * c.sc_still, s_hp = c_head * c.sc_head;
      total = s_or + s_tq + s_ar + s_al + s_ss + s_hp;               // dict order of standing.py:585-604
      scs[0] = s_or; scs[1] = s_tq; scs[2] = s_ar; scs[3] = s_ss; scs[4] = s_al; scs[5] = s_hp; scs[6] = 0.f;
      sg[0] = c.sc_orient; sg[1] = c.sc_torques; sg[2] = c.sc_rate; sg[3] = c.sc_still; sg[4] = c.sc_alive; sg[5] = c.sc_head; sg[6] = 0.f;
    } else {
      const float sc0 = r_lin * c.sc_lin, sc1 = r_ang * c.sc_ang, sc2 = c_torque * c.sc_torques, sc3 = c_rate * c.sc_rate;
      const float sc4 = c_still * c.sc_still, sc5 = 1.f * c.sc_alive, sc6 = r_imit * c.sc_imit;
      total = sc0 + sc1 + sc2 + sc3 + sc5 + sc6 + sc4;               // dict order of joystick.py:634-667
      scs[0] = sc0; scs[1] = sc1; scs[2] = sc2; scs[3] = sc3; scs[4] = sc4; scs[5] = sc5; scs[6] = sc6;
      sg[0] = c.sc_lin; sg[1] = c.sc_ang; sg[2] = c.sc_torques; sg[3] = c.sc_rate; sg[4] = c.sc_still; sg[5] = c.sc_alive; sg[6] = c.sc_imit;
    }
    if constexpr (RL) {
      // reward library (include/oduck.h OduckLibTerm; oracle reward_library_sum): scaled terms added in enum order
      const DevRewardLib& R = *p.rlib;
      const float* Rs = s.outrec + OUT_IMUMAT;                       // imu site_xmat of the last forward: global_linvel = Rs * local_linvel
      const float gx = Rs[0] * sd[3] + Rs[1] * sd[4] + Rs[2] * sd[5], gy = Rs[3] * sd[3] + Rs[4] * sd[4] + Rs[5] * sd[5], gz = Rs[6] * sd[3] + Rs[7] * sd[4] + Rs[8] * sd[5];
      const bool u = lane < nu;
      const float lo = u ? R.soft_lo[lane] : 0.f, hi = u ? R.soft_hi[lane] : 0.f, pw = u ? R.pose_w[lane] : 0.f;
      float t[ODUCK_NLIBTERM];
      t[ODUCK_LIB_ORIENTATION] = nan_to_num(sd[9] * sd[9] + sd[10] * sd[10]);
      t[ODUCK_LIB_LIN_VEL_Z] = nan_to_num(gz * gz);
      t[ODUCK_LIB_ANG_VEL_XY] = nan_to_num(sd[12] * sd[12] + sd[13] * sd[13]);
      { const float dh = s.qpos[2] - R.base_height_target; t[ODUCK_LIB_BASE_HEIGHT] = nan_to_num(dh * dh); }
      t[ODUCK_LIB_ENERGY] = nan_to_num(wsum(u ? fabsf(qd) * fabsf(afrc) : 0.f));
      t[ODUCK_LIB_JOINT_POS_LIMITS] = nan_to_num(wsum(u ? -fminf(q - lo, 0.f) + fmaxf(q - hi, 0.f) : 0.f));
      t[ODUCK_LIB_TERMINATION] = done ? 1.f : 0.f;
      t[ODUCK_LIB_JOINT_DEVIATION_HIP] = nan_to_num(wsum((u && ((R.hip_mask >> lane) & 1u)) ? fabsf(q - dflt) : 0.f) * (fabsf(c1) > 0.1f ? 1.f : 0.f));
      t[ODUCK_LIB_JOINT_DEVIATION_KNEE] = nan_to_num(wsum((u && ((R.knee_mask >> lane) & 1u)) ? fabsf(q - dflt) : 0.f));
      t[ODUCK_LIB_POSE] = nan_to_num(wsum(u ? (q - dflt) * (q - dflt) * pw : 0.f));
      {
        const float v = sqrtf(gx * gx + gy * gy);
        t[ODUCK_LIB_FEET_SLIP] = nan_to_num(v * __shfl_sync(FULLMASK, contact, 0) + v * __shfl_sync(FULLMASK, contact, 1));
      }
      {
        const int f = lane & 1;                                       // lane k < 2: foot k
        const float vx = sd[15 + 3 * f], vy = sd[16 + 3 * f], fz = s.outrec[OUT_FEET + 3 * f + 2];
        const float clr = fabsf(fz - R.max_foot_height) * sqrtf(sqrtf(vx * vx + vy * vy));
        const float eh = er.swing / R.max_foot_height - 1.f;
        const float at = fminf((er.air - R.air_thr_min) * first_contact, R.air_thr_max - R.air_thr_min);
        const bool ft = lane < 2;
        t[ODUCK_LIB_FEET_CLEARANCE] = nan_to_num(wsum(ft ? clr : 0.f));
        t[ODUCK_LIB_FEET_HEIGHT] = nan_to_num(wsum(ft ? eh * eh * first_contact : 0.f));
        t[ODUCK_LIB_FEET_AIR_TIME] = nan_to_num(wsum(ft ? at : 0.f) * (cmd_norm > 0.01f ? 1.f : 0.f));
        // gait-clocked terms (include/oduck.h): clock = the reference-motion phase counter after this step's increment
        const float tt = (float)er.imitation_i * dt;
        const float tgt = R.swing_amp * sinf(2.f * PI * R.swing_freq * tt);
        t[ODUCK_LIB_BASE_Y_SWING] = nan_to_num(expf(-((tgt - sd[4]) * (tgt - sd[4])) / c.tracking_sigma));
        float phi = 2.f * PI * (float)er.imitation_i / (float)c.nb_steps + (float)f * PI;
        phi -= 2.f * PI * floorf((phi + PI) / (2.f * PI));                                   // wrap to [-pi, pi)
        const float xg = (phi + PI) / (2.f * PI);                                            // gait.get_rz: cubic Bezier 0 -> h -> 0
        const float xb = xg <= 0.5f ? 2.f * xg : 2.f * xg - 1.f;
        const float bz = xb * xb * xb + 3.f * (xb * xb * (1.f - xb));
        const float rz = xg <= 0.5f ? R.max_foot_height * bz : R.max_foot_height + (0.f - R.max_foot_height) * bz;
        t[ODUCK_LIB_FEET_PHASE] = nan_to_num(expf(-wsum(ft ? (fz - rz) * (fz - rz) : 0.f) / 0.01f));
      }
      float libsum = 0.f;
#pragma unroll
      for (int k = 0; k < ODUCK_NLIBTERM; ++k) if (R.scale[k] != 0.f) libsum += t[k] * R.scale[k];
      total += libsum;
    }
    const float reward = fminf(fmaxf(total * dt, 0.f), 10000.f);
    // info updates (joystick.py:449-469)
    er.step += 1;
    er.push_step += 1;
    er.last_act[2] = er.last_act[1]; er.last_act[1] = er.last_act[0]; er.last_act[0] = act;
    RKey cmd_rng;
    { RKey o = rblock(er.rng, lane & 1); cmd_rng = kshfl(o, 1); er.rng = kshfl(o, 0); }
    if (er.step > 500) { const float nc = sample_command(c, cmd_rng, lane); er.cmd = lane < 7 ? nc : 0.f; }
    if (done || er.step > 500) er.step = 0;
    { const float nc = contact != 0.f ? 0.f : 1.f; er.air *= nc; er.lastc = contact; er.swing *= nc; }
    // metrics (joystick.py:470-477): scaled term, sign flipped for costs
    {
      float mv = 0.f;
#pragma unroll
      for (int k = 0; k < 7; ++k) if (lane == k) mv = sg[k] == 0.f ? 0.f : (sg[k] > 0.f ? scs[k] : -scs[k]);
      const float sw = 0.5f * (__shfl_sync(FULLMASK, er.swing, 0) + __shfl_sync(FULLMASK, er.swing, 1));
      if (lane == (c.task == ODUCK_TASK_STANDING ? 6 : 7)) mv = sw;       // swing_peak follows the last reward term
      if (lane < ODUCK_NMETRIC) p.metrics[(size_t)env * ODUCK_NMETRIC + lane] = mv;
    }
    // EpisodeWrapper (action_repeat = 1) + AutoReset
    er.steps += 1;
    const bool trunc = er.steps >= c.episode_length;
    const float d = done ? 1.f : 0.f;
    const float done_f = trunc ? 1.f : d;
    if (lane == 0) { p.reward[env] = reward; p.done[env] = done_f; p.trunc[env] = trunc ? 1.f - d : 0.f; }
    store_info(m, c, inf, er, pushv, lane);
    store_out(s, lane, p.out + (size_t)env * OUT_STRIDE);
    if (c.auto_reset && done_f != 0.f) {
      __syncwarp();
      const float* fp = p.first_phys + (size_t)env * PHYS_STRIDE;
      for (int i = lane; i < PHYS_CTRL; i += 32) ph[i] = fp[i];      // qpos, qvel, qacc_warmstart <- first_state
      const int a = m.d_act[lane];
      if (a >= 0) ph[PHYS_CTRL + a] = L.ctrl;
      const float* fs_ = p.first_obs_state + (size_t)env * ODUCK_OBS_STATE;
      for (int i = lane; i < ODUCK_OBS_STATE; i += 32) ost[i] = fs_[i];
      const float* fpv = p.first_obs_priv + (size_t)env * ODUCK_OBS_PRIV;
      for (int i = lane; i < ODUCK_OBS_PRIV; i += 32) opr[i] = fpv[i];
    } else {
      store_phys(m, s, L, lane, ph);
    }
    __syncwarp();
    if (p.sk_obs_p) {
      // A17: this step's Transition goes straight into the caller's rollout buffers (oduck_rollout_step): reward / done /
      // truncation of slot t, the observations the NEXT action is computed from (after auto-reset) into slot t + 1.  The rows
      // were written by this warp just above (L1 / L2 hits); the policy's raw action / log-prob were written by the actor's head.
      if (lane == 0) { p.sk_reward[env] = reward; p.sk_done[env] = done_f; p.sk_trunc[env] = trunc ? 1.f - d : 0.f; }
      float* so = p.sk_obs_p + (size_t)env * p.sk_dp;
      for (int i = lane; i < p.sk_dp; i += 32) so[i] = ost[i];
      float* sv = p.sk_obs_v + (size_t)env * p.sk_dv;
      for (int i = lane; i < p.sk_dv; i += 32) sv[i] = opr[i];
    }
  }
}

// slot 0 of the sink's observations <- the handle's current observations (first step of an unroll)
__global__ void k_sink_obs0(const float* __restrict__ ost, const float* __restrict__ opr, float* __restrict__ dp_, float* __restrict__ dv_, int N, int dp, int dv) {
  const int tot = N * (dp + dv);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += gridDim.x * blockDim.x) {
    if (idx < N * dp) { const int e = idx / dp, k = idx - e * dp; dp_[idx] = ost[(size_t)e * ODUCK_OBS_STATE + k]; }
    else { const int r = idx - N * dp, e = r / dv, k = r - e * dv; dv_[r] = opr[(size_t)e * ODUCK_OBS_PRIV + k]; }
  }
}

// A14: domain_randomize (common/randomize.py:39-106), one warp per env, one round per split.
__global__ void __launch_bounds__(WPB * 32, CTAS_PER_SM) k_randomize(Params p) {
  ODUCK_SMEM_RAW(raw);
  DevModel* mp; DevEnvCfg* cp; WarpSmem* ws;
  block_load_tables(p, raw, mp, cp, ws);
  const DevModel& m = *mp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int env = blockIdx.x * WPB + warp; env < p.N; env += gridDim.x * WPB) {
    float* dr = p.dr + (size_t)env * DR_STRIDE;
    RKey rng; rng.a = p.keys[2 * env]; rng.b = p.keys[2 * env + 1];
    float u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      RKey o = rblock(rng, lane & 1);
      RKey key = kshfl(o, 1);
      rng = kshfl(o, 0);
      RKey b = rblock(key, (uint32_t)lane);
      u[k] = bits_unit(b.a ^ b.b);
    }
    auto uni = [](float x, float lo, float hi) { return fmaxf(lo, x * (hi - lo) + lo); };
    if (lane == 0) dr[DR_FRIC0] = uni(u[0], 0.5f, 1.0f);
    // friction-loss dofs in dof order; row index = draw index
    for (int d = 0; d < m.nv; ++d) {
      const int r = m.d_frrow[d];
      if (r < 0 || r != lane) continue;
      dr[DR_FLOSS + d] = m.d_floss[d] * uni(u[1], 0.9f, 1.1f);
      dr[DR_ARM + d] = m.d_arm[d] * uni(u[2], 1.0f, 1.05f);
      dr[DR_QPOS0 + m.d_qadr[d]] = m.qpos0[m.d_qadr[d]] + uni(u[6], -0.03f, 0.03f);
    }
    if (lane < 3) dr[DR_IPOS1 + lane] = m.b_ipos[lane][1] + uni(u[3], -0.05f, 0.05f);
    const float dm1 = uni(__shfl_sync(FULLMASK, u[5], 0), -0.1f, 0.1f);
    if (lane < m.nbody) dr[lane] = m.b_mass[lane] * uni(u[4], 0.9f, 1.1f) + (lane == 1 ? dm1 : 0.f);
    if (lane < m.nu) dr[DR_KP + lane] = m.d_kp[m.act_dof[lane]] * uni(u[7], 0.9f, 1.1f);
  }
}

// ------------------------------------------------------------------------------------------------- host side
#include "oduck_build.h"   // build_dev_model, build_dev_ff: OduckModel -> device tables (pure host code, shared with tests/emu)

static void build_dev_cfg(const OduckEnvConfig& C, DevEnvCfg& D) {
  memset(&D, 0, sizeof(D));
  D.task = C.task; D.sc_orient = (float)C.scale_orientation; D.sc_head = (float)C.scale_head_pos; D.reset_qvel_noise = (float)C.reset_base_qvel_noise;
  D.n_substeps = C.n_substeps; D.episode_length = C.episode_length; D.use_imitation = C.use_imitation_reward; D.use_speed_limits = C.use_motor_speed_limits;
  D.push_enable = C.push_enable; D.act_min_delay = C.action_min_delay; D.act_max_delay = C.action_max_delay; D.imu_min_delay = C.imu_min_delay; D.imu_max_delay = C.imu_max_delay;
  D.auto_reset = C.auto_reset;
  D.ctrl_dt = (float)C.ctrl_dt; D.action_scale = (float)C.action_scale; D.dof_vel_scale = (float)C.dof_vel_scale; D.max_motor_velocity = (float)C.max_motor_velocity;
  D.noise_level = (float)C.noise_level; D.noise_gyro = (float)C.noise_gyro; D.noise_acc = (float)C.noise_accelerometer; D.noise_gravity = (float)C.noise_gravity; D.noise_joint_vel = (float)C.noise_joint_vel;
  for (int i = 0; i < 16; i++) D.qpos_noise_scale[i] = (float)C.qpos_noise_scale[i];
  D.sc_lin = (float)C.scale_tracking_lin_vel; D.sc_ang = (float)C.scale_tracking_ang_vel; D.sc_torques = (float)C.scale_torques; D.sc_rate = (float)C.scale_action_rate;
  D.sc_still = (float)C.scale_stand_still; D.sc_alive = (float)C.scale_alive; D.sc_imit = (float)C.scale_imitation; D.tracking_sigma = (float)C.tracking_sigma;
  for (int i = 0; i < 2; i++) { D.push_interval[i] = (float)C.push_interval_range[i]; D.push_magnitude[i] = (float)C.push_magnitude_range[i]; }
  for (int i = 0; i < 7; i++) { D.cmd_range[i][0] = (float)C.cmd_range[i][0]; D.cmd_range[i][1] = (float)C.cmd_range[i][1]; }
  D.ndx = C.ndx; D.ndy = C.ndy; D.ndth = C.ndth; D.nb_steps = C.nb_steps_in_period > 0 ? C.nb_steps_in_period : 1;
  for (int i = 0; i < 8; i++) { D.dxs[i] = (float)C.dxs[i]; D.dys[i] = (float)C.dys[i]; }
  for (int i = 0; i < 16; i++) D.dths[i] = (float)C.dthetas[i];
  for (int i = 0; i < 2; i++) { D.dx_range[i] = (float)C.dx_range[i]; D.dy_range[i] = (float)C.dy_range[i]; D.dth_range[i] = (float)C.dtheta_range[i]; }
}

static Params make_params(OduckHandle* h) {
  Params p;
  memset(&p, 0, sizeof(p));
  p.model = h->dmodel; p.cfg = h->dcfg; p.poly = h->poly;
  p.phys = h->phys; p.dr = h->dr; p.out = h->out; p.info = h->info; p.obs_state = h->obs_state; p.obs_priv = h->obs_priv;
  p.reward = h->reward; p.done = h->done; p.trunc = h->trunc; p.metrics = h->metrics;
  p.first_phys = h->first_phys; p.first_obs_state = h->first_obs_state; p.first_obs_priv = h->first_obs_priv; p.dbg = h->dbg;
  p.ffmodel = h->dff; p.ffscratch = h->ffscratch;
  p.hfmodel = h->dhf; p.hfscratch = h->hfscratch; p.rlib = h->drlib;
  p.N = h->n;
  return p;
}

template <typename K>
static int launch(OduckHandle* h, K kernel, const Params& p, void* stream) {
  CUDA_TRY(cudaSetDevice(h->device));
#ifdef ODUCK_WARP_EMU
  warp_emu::launch(kernel, h->grid, (size_t)h->smem_bytes, p);
  (void)stream;
#else
  kernel<<<h->grid, WPB * 32, h->smem_bytes, (cudaStream_t)stream>>>(p);
#endif
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return ODUCK_OK;
}

extern "C" {

int oduck_abi_version(void) { return ODUCK_ABI_VERSION; }
int oduck_sizeof_model(void) { return (int)sizeof(OduckModel); }
int oduck_sizeof_env_config(void) { return (int)sizeof(OduckEnvConfig); }
const char* oduck_last_error(void) { return g_err.c_str(); }
int oduck_num_envs(const OduckHandle* h) { return h ? h->n : 0; }
int64_t oduck_launch_count(const OduckHandle* h) { return h ? h->launches : 0; }

int oduck_destroy(OduckHandle* h) {
  if (!h) return ODUCK_OK;
  cudaSetDevice(h->device);
  void* ptrs[] = {h->dmodel, h->dcfg, h->poly, h->phys, h->dr, h->out, h->info, h->obs_state, h->obs_priv, h->reward, h->done, h->trunc,
                  h->metrics, h->first_phys, h->first_obs_state, h->first_obs_priv, h->dbg, h->policy_scratch, h->dff, h->ffscratch, h->dhf, h->hfdata, h->hfscratch, h->drlib, h->act_buf};
  for (void* q : ptrs) if (q) cudaFree(q);
  delete h;
  return ODUCK_OK;
}

int oduck_create(const OduckModel* model, const OduckEnvConfig* cfg, int num_envs, int device, OduckHandle** out) {
  if (!model || !cfg || !out || num_envs <= 0) return fail(ODUCK_ERR_ARG, "oduck_create: bad argument");
  if (model->abi_version != ODUCK_ABI_VERSION) return fail(ODUCK_ERR_MODEL, "oduck_create: model ABI version mismatch");
  if (cfg->task != ODUCK_TASK_JOYSTICK && cfg->task != ODUCK_TASK_STANDING) return fail(ODUCK_ERR_ARG, "oduck_create: unknown task");
  if (model->floor_is_hfield && (!model->hfield_data || model->hfield_nrow < 2 || model->hfield_ncol < 2))
    return fail(ODUCK_ERR_MODEL, "oduck_create: height-field floor without elevation data");
  if (model->floor_is_hfield && (model->hfield_ncol > 4096 || model->hfield_nrow > 8192))
    return fail(ODUCK_ERR_UNSUPPORTED, "oduck_create: height field larger than 8192 x 4096 samples (the collider packs row / column into one word)");
  if (cfg->action_max_delay > MAX_DELAY || cfg->imu_max_delay * 3 > 16 || cfg->action_max_delay < 1) return fail(ODUCK_ERR_ARG, "oduck_create: delay history out of range");
  bool use_lib = false;
  for (int k = 0; k < ODUCK_NLIBTERM; k++) use_lib |= cfg->lib.scale[k] != 0.0;
  if (use_lib && (cfg->lib.n_hip < 0 || cfg->lib.n_hip > 4 || cfg->lib.n_knee < 0 || cfg->lib.n_knee > 4)) return fail(ODUCK_ERR_ARG, "oduck_create: at most 4 hip / knee joints");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(ODUCK_ERR_CUDA, "oduck_create: no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));
  OduckHandle* h = new OduckHandle();
  memset(h, 0, sizeof(*h));
  h->n = num_envs; h->device = device; h->hm = *model; h->hcfg = *cfg;
  std::string err;
  if (build_dev_model(*model, h->hdm, err) != 0) { delete h; return fail(ODUCK_ERR_MODEL, "oduck_create: " + err); }
  build_dev_cfg(*cfg, h->hdc);
  h->nefc = h->hdm.nfr + h->hdm.nlim + 4 * NCON_ALL;
  if (h->nefc > 96) { delete h; return fail(ODUCK_ERR_MODEL, "oduck_create: too many constraint rows"); }
  const size_t N = (size_t)num_envs;
#define ALLOC(ptr, count) do { if (cudaMalloc((void**)&(ptr), (count) * sizeof(float)) != cudaSuccess) { oduck_destroy(h); return fail(ODUCK_ERR_ALLOC, "oduck_create: cudaMalloc failed"); } cudaMemset((ptr), 0, (count) * sizeof(float)); } while (0)
  if (cudaMalloc((void**)&h->dmodel, sizeof(DevModel)) != cudaSuccess || cudaMalloc((void**)&h->dcfg, sizeof(DevEnvCfg)) != cudaSuccess) { oduck_destroy(h); return fail(ODUCK_ERR_ALLOC, "oduck_create: cudaMalloc failed"); }
  CUDA_TRY(cudaMemcpy(h->dmodel, &h->hdm, sizeof(DevModel), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(h->dcfg, &h->hdc, sizeof(DevEnvCfg), cudaMemcpyHostToDevice));
  if (cfg->use_imitation_reward) {
    if (!cfg->poly_coef) { oduck_destroy(h); return fail(ODUCK_ERR_ARG, "oduck_create: imitation reward needs poly_coef"); }
    size_t np = (size_t)cfg->ndx * cfg->ndy * cfg->ndth * ODUCK_REF_DIM * ODUCK_POLY_DEG;
    std::vector<float> pf(np);
    for (size_t i = 0; i < np; i++) pf[i] = (float)cfg->poly_coef[i];   // jax float32 (x64 disabled)
    ALLOC(h->poly, np);
    CUDA_TRY(cudaMemcpy(h->poly, pf.data(), np * sizeof(float), cudaMemcpyHostToDevice));
  }
  h->hcfg.poly_coef = nullptr;
  {
    DevFF f;
    build_dev_ff(*model, f);
    if (cudaMalloc((void**)&h->dff, sizeof(DevFF)) != cudaSuccess) { oduck_destroy(h); return fail(ODUCK_ERR_ALLOC, "oduck_create: cudaMalloc failed"); }
    CUDA_TRY(cudaMemcpy(h->dff, &f, sizeof(DevFF), cudaMemcpyHostToDevice));
  }
  ALLOC(h->ffscratch, N * (FFJ_SIZE + FFV_SIZE));
  if (model->floor_is_hfield) {
    const size_t ns = (size_t)model->hfield_nrow * model->hfield_ncol;
    ALLOC(h->hfdata, ns);
    CUDA_TRY(cudaMemcpy(h->hfdata, model->hfield_data, ns * sizeof(float), cudaMemcpyHostToDevice));
    ALLOC(h->hfscratch, N * HF_SCRATCH);
    DevHF d;
    d.nrow = model->hfield_nrow; d.ncol = model->hfield_ncol;
    d.sx = (float)model->hfield_size[0]; d.sy = (float)model->hfield_size[1]; d.sz = (float)model->hfield_size[2];
    d.dx = (float)(2.0 * model->hfield_size[0] / (model->hfield_ncol - 1)); d.dy = (float)(2.0 * model->hfield_size[1] / (model->hfield_nrow - 1));
    d.data = h->hfdata;
    if (cudaMalloc((void**)&h->dhf, sizeof(DevHF)) != cudaSuccess) { oduck_destroy(h); return fail(ODUCK_ERR_ALLOC, "oduck_create: cudaMalloc failed"); }
    CUDA_TRY(cudaMemcpy(h->dhf, &d, sizeof(DevHF), cudaMemcpyHostToDevice));
  }
  h->hm.hfield_data = nullptr;      // the caller's array is not kept
  if (use_lib) {
    DevRewardLib r;
    memset(&r, 0, sizeof(r));
    const OduckRewardLibrary& L = cfg->lib;
    for (int k = 0; k < ODUCK_NLIBTERM; k++) r.scale[k] = (float)L.scale[k];
    r.base_height_target = (float)L.base_height_target; r.max_foot_height = (float)L.max_foot_height;
    r.air_thr_min = (float)L.air_time_threshold_min; r.air_thr_max = (float)L.air_time_threshold_max;
    r.swing_freq = (float)L.base_y_swing_freq; r.swing_amp = (float)L.base_y_swing_amplitude;
    for (int u = 0; u < model->nu; u++) { r.soft_lo[u] = (float)L.soft_lowers[u]; r.soft_hi[u] = (float)L.soft_uppers[u]; r.pose_w[u] = (float)L.pose_weights[u]; }
    for (int i = 0; i < L.n_hip; i++) r.hip_mask |= 1u << L.hip_indices[i];
    for (int i = 0; i < L.n_knee; i++) r.knee_mask |= 1u << L.knee_indices[i];
    if (cudaMalloc((void**)&h->drlib, sizeof(DevRewardLib)) != cudaSuccess) { oduck_destroy(h); return fail(ODUCK_ERR_ALLOC, "oduck_create: cudaMalloc failed"); }
    CUDA_TRY(cudaMemcpy(h->drlib, &r, sizeof(DevRewardLib), cudaMemcpyHostToDevice));
  }
  ALLOC(h->phys, N * PHYS_STRIDE); ALLOC(h->dr, N * DR_STRIDE); ALLOC(h->out, N * OUT_STRIDE); ALLOC(h->info, N * INFO_STRIDE);
  ALLOC(h->obs_state, N * ODUCK_OBS_STATE); ALLOC(h->obs_priv, N * ODUCK_OBS_PRIV);
  ALLOC(h->reward, N); ALLOC(h->done, N); ALLOC(h->trunc, N); ALLOC(h->metrics, N * ODUCK_NMETRIC);
  ALLOC(h->act_buf, N * ODUCK_MAX_NU);
  ALLOC(h->first_phys, N * PHYS_STRIDE); ALLOC(h->first_obs_state, N * ODUCK_OBS_STATE); ALLOC(h->first_obs_priv, N * ODUCK_OBS_PRIV);
#undef ALLOC
  // nominal state and nominal (un-randomised) per-env model
  {
    std::vector<float> ph(PHYS_STRIDE, 0.f), dr(DR_STRIDE, 0.f), inf(INFO_STRIDE, 0.f);
    for (int i = 0; i < model->nq; i++) { ph[i] = (float)model->key_qpos[i]; dr[DR_QPOS0 + i] = (float)model->qpos0[i]; }
    for (int u = 0; u < model->nu; u++) { ph[PHYS_CTRL + u] = (float)model->key_ctrl[u]; dr[DR_KP + u] = (float)model->act_kp[u]; }
    for (int b = 0; b < model->nbody; b++) dr[b] = (float)model->body_mass[b];
    for (int i = 0; i < 3; i++) dr[DR_IPOS1 + i] = (float)model->body_ipos[1][i];
    dr[DR_FRIC0] = 1.f;
    for (int d = 0; d < model->nv; d++) { dr[DR_FLOSS + d] = (float)model->dof_frictionloss[d]; dr[DR_ARM + d] = (float)model->dof_armature[d]; }
    int big = 1 << 30;
    memcpy(&inf[INFO_PUSH_INT], &big, 4);
    std::vector<float> all;
    auto fill = [&](float* dst, const std::vector<float>& rec) {
      all.resize(N * rec.size());
      for (size_t e = 0; e < N; e++) memcpy(&all[e * rec.size()], rec.data(), rec.size() * sizeof(float));
      return cudaMemcpy(dst, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice);
    };
    CUDA_TRY(fill(h->phys, ph)); CUDA_TRY(fill(h->first_phys, ph)); CUDA_TRY(fill(h->dr, dr)); CUDA_TRY(fill(h->info, inf));
  }
  h->smem_bytes = (int)(((sizeof(DevModel) + 15) & ~15u) + ((sizeof(DevEnvCfg) + 15) & ~15u) + WPB * sizeof(WarpSmem));
  if (model->floor_is_hfield) {
    CUDA_TRY(cudaFuncSetAttribute(k_physics<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(k_physics<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(k_reset<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(k_step<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(k_step<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
  } else {
    CUDA_TRY(cudaFuncSetAttribute(k_physics<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(k_physics<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(k_reset<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(k_step<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(k_step<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
  }
  CUDA_TRY(cudaFuncSetAttribute(k_randomize, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
  h->grid = (num_envs + WPB - 1) / WPB;
  CUDA_TRY(cudaDeviceSynchronize());
  *out = h;
  return ODUCK_OK;
}

int oduck_randomize(OduckHandle* h, const uint32_t* keys, void* stream) {
  if (!h || !keys) return fail(ODUCK_ERR_ARG, "oduck_randomize: bad argument");
  Params p = make_params(h);
  p.keys = keys;
  return launch(h, k_randomize, p, stream);
}
int oduck_reset(OduckHandle* h, const uint32_t* keys, const uint8_t* mask, void* stream) {
  if (!h || !keys) return fail(ODUCK_ERR_ARG, "oduck_reset: bad argument");
  Params p = make_params(h);
  p.keys = keys; p.mask = mask;
  return h->dhf ? launch(h, k_reset<true>, p, stream) : launch(h, k_reset<false>, p, stream);
}
int oduck_step(OduckHandle* h, const float* action, void* stream) {
  if (!h || !action) return fail(ODUCK_ERR_ARG, "oduck_step: bad argument");
  Params p = make_params(h);
  p.action = action;
  if (h->drlib) return h->dhf ? launch(h, k_step<true, true>, p, stream) : launch(h, k_step<false, true>, p, stream);
  return h->dhf ? launch(h, k_step<true, false>, p, stream) : launch(h, k_step<false, false>, p, stream);
}
int oduck_set_rollout_sink(OduckHandle* h, const OduckRolloutSink* sink) {
  if (!h) return fail(ODUCK_ERR_ARG, "oduck_set_rollout_sink: bad argument");
  if (!sink) { memset(&h->sink, 0, sizeof(h->sink)); return ODUCK_OK; }
  const int dp = h->hcfg.task == ODUCK_TASK_STANDING ? 85 : ODUCK_OBS_STATE, dv = h->hcfg.task == ODUCK_TASK_STANDING ? 153 : ODUCK_OBS_PRIV;
  if (sink->unroll < 1 || sink->env_offset < 0 || sink->env_offset + h->n > sink->num_envs) return fail(ODUCK_ERR_ARG, "oduck_set_rollout_sink: the handle's envs do not fit into the buffers");
  if (sink->policy_dim != dp || sink->value_dim != dv) return fail(ODUCK_ERR_ARG, "oduck_set_rollout_sink: obs row widths must be those of the task (Joystick 101 / 212, Standing 85 / 153)");
  if (!sink->obs_policy || !sink->obs_value || !sink->raw_action || !sink->log_prob || !sink->reward || !sink->done || !sink->truncation)
    return fail(ODUCK_ERR_ARG, "oduck_set_rollout_sink: null buffer");
  h->sink = *sink;
  return ODUCK_OK;
}
// oduck_step with the transition of unroll step t stored in the attached sink (called by oduck_rollout_step, oduck_policy.cu)
int oduck_step_into_sink(OduckHandle* h, const float* action, int t, void* stream) {
  const OduckRolloutSink& k = h->sink;
  if (!k.obs_policy || t < 0 || t >= k.unroll) return fail(ODUCK_ERR_ARG, "oduck_rollout_step: no sink attached or t outside the unroll");
  Params p = make_params(h);
  p.action = action;
  const size_t row = (size_t)t * k.num_envs + k.env_offset, next = row + k.num_envs;
  p.sk_obs_p = k.obs_policy + next * k.policy_dim; p.sk_obs_v = k.obs_value + next * k.value_dim;
  p.sk_reward = k.reward + row; p.sk_done = k.done + row; p.sk_trunc = k.truncation + row;
  p.sk_dp = k.policy_dim; p.sk_dv = k.value_dim;
  if (t == 0) {
    CUDA_TRY(cudaSetDevice(h->device));
#ifndef ODUCK_WARP_EMU
    k_sink_obs0<<<148, 256, 0, (cudaStream_t)stream>>>(h->obs_state, h->obs_priv, k.obs_policy + (size_t)k.env_offset * k.policy_dim,
                                                       k.obs_value + (size_t)k.env_offset * k.value_dim, h->n, k.policy_dim, k.value_dim);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
#else   // tests/emu: device memory is host memory
    for (int e = 0; e < h->n; e++) {
      for (int c = 0; c < k.policy_dim; c++) k.obs_policy[(size_t)(k.env_offset + e) * k.policy_dim + c] = h->obs_state[(size_t)e * ODUCK_OBS_STATE + c];
      for (int c = 0; c < k.value_dim; c++) k.obs_value[(size_t)(k.env_offset + e) * k.value_dim + c] = h->obs_priv[(size_t)e * ODUCK_OBS_PRIV + c];
    }
#endif
  }
  if (h->drlib) return h->dhf ? launch(h, k_step<true, true>, p, stream) : launch(h, k_step<false, true>, p, stream);
  return h->dhf ? launch(h, k_step<true, false>, p, stream) : launch(h, k_step<false, false>, p, stream);
}
int oduck_physics_substeps(OduckHandle* h, const float* ctrl, int n, void* stream) {
  if (!h || n < 0) return fail(ODUCK_ERR_ARG, "oduck_physics_substeps: bad argument");
  Params p = make_params(h);
  p.action = ctrl; p.nsub = n; p.integrate = 1;
  return h->dhf ? launch(h, k_physics<false, true>, p, stream) : launch(h, k_physics<false, false>, p, stream);
}
int oduck_forward(OduckHandle* h, void* stream) {
  if (!h) return fail(ODUCK_ERR_ARG, "oduck_forward: bad argument");
  Params p = make_params(h);
  p.nsub = 1; p.integrate = 0;
  return h->dhf ? launch(h, k_physics<false, true>, p, stream) : launch(h, k_physics<false, false>, p, stream);
}
// Diagnostic (not part of the reference surface): one forward with every intermediate dumped, DBG_STRIDE floats per env
// copied to host memory `out`.  Used by tests/test_parity_gpu.py to localise a parity failure.
int oduck_debug_forward(OduckHandle* h, float* out_host) {
  if (!h || !out_host) return fail(ODUCK_ERR_ARG, "oduck_debug_forward: bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->dbg) { CUDA_TRY(cudaMalloc((void**)&h->dbg, (size_t)h->n * DBG_STRIDE * sizeof(float))); }
  CUDA_TRY(cudaMemset(h->dbg, 0, (size_t)h->n * DBG_STRIDE * sizeof(float)));
  Params p = make_params(h);
  p.nsub = 1; p.integrate = 0;
  int rc = h->dhf ? launch(h, k_physics<true, true>, p, nullptr) : launch(h, k_physics<true, false>, p, nullptr);
  if (rc) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out_host, h->dbg, (size_t)h->n * DBG_STRIDE * sizeof(float), cudaMemcpyDeviceToHost));
  return ODUCK_OK;
}
int oduck_debug_stride(void) { return DBG_STRIDE; }

int oduck_set_state(OduckHandle* h, const float* qpos, const float* qvel, const float* qacc_warm, void* stream) {
  if (!h) return fail(ODUCK_ERR_ARG, "oduck_set_state: bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t N = (size_t)h->n, pitch = PHYS_STRIDE * sizeof(float);
  if (qpos) CUDA_TRY(cudaMemcpy2DAsync(h->phys, pitch, qpos, h->hm.nq * sizeof(float), h->hm.nq * sizeof(float), N, cudaMemcpyDeviceToDevice, st));
  if (qvel) CUDA_TRY(cudaMemcpy2DAsync(h->phys + PHYS_QVEL, pitch, qvel, h->hm.nv * sizeof(float), h->hm.nv * sizeof(float), N, cudaMemcpyDeviceToDevice, st));
  if (qacc_warm) CUDA_TRY(cudaMemcpy2DAsync(h->phys + PHYS_QACCW, pitch, qacc_warm, h->hm.nv * sizeof(float), h->hm.nv * sizeof(float), N, cudaMemcpyDeviceToDevice, st));
  return ODUCK_OK;
}

int oduck_get_buffer(OduckHandle* h, int id, void** ptr, int64_t* shape, int64_t* strides, int* dtype) {
  if (!h || !ptr || !shape || !strides || !dtype) return fail(ODUCK_ERR_ARG, "oduck_get_buffer: bad argument");
  const OduckModel& m = h->hm;
  float* p = nullptr;
  int64_t d1 = 0, d2 = 0, s0 = 0, s1 = 1;
  int dt = ODUCK_DTYPE_F32;
#define REC(base, stride, off, n1) p = (base) + (off); s0 = (stride); d1 = (n1);
  switch (id) {
    case ODUCK_BUF_QPOS: REC(h->phys, PHYS_STRIDE, 0, m.nq) break;
    case ODUCK_BUF_QVEL: REC(h->phys, PHYS_STRIDE, PHYS_QVEL, m.nv) break;
    case ODUCK_BUF_QACC_WARM: REC(h->phys, PHYS_STRIDE, PHYS_QACCW, m.nv) break;
    case ODUCK_BUF_CTRL: REC(h->phys, PHYS_STRIDE, PHYS_CTRL, m.nu) break;
    case ODUCK_BUF_QACC: REC(h->out, OUT_STRIDE, OUT_QACC, m.nv) break;
    case ODUCK_BUF_SENSORDATA: REC(h->out, OUT_STRIDE, OUT_SENS, 24) break;
    case ODUCK_BUF_EFC_FORCE: REC(h->out, OUT_STRIDE, OUT_EFC, h->nefc) break;
    case ODUCK_BUF_CONTACT_DIST: REC(h->out, OUT_STRIDE, OUT_CDIST, 12) break;
    case ODUCK_BUF_ACTUATOR_FORCE: REC(h->out, OUT_STRIDE, OUT_AFRC, m.nu) break;
    case ODUCK_BUF_SITE_XPOS_FEET: REC(h->out, OUT_STRIDE, OUT_FEET, 6) break;
    case ODUCK_BUF_OBS_STATE: REC(h->obs_state, ODUCK_OBS_STATE, 0, h->hcfg.task == ODUCK_TASK_STANDING ? 85 : ODUCK_OBS_STATE) break;
    case ODUCK_BUF_OBS_PRIV: REC(h->obs_priv, ODUCK_OBS_PRIV, 0, h->hcfg.task == ODUCK_TASK_STANDING ? 153 : ODUCK_OBS_PRIV) break;
    case ODUCK_BUF_REWARD: REC(h->reward, 1, 0, 0) break;
    case ODUCK_BUF_DONE: REC(h->done, 1, 0, 0) break;
    case ODUCK_BUF_TRUNCATION: REC(h->trunc, 1, 0, 0) break;
    case ODUCK_BUF_METRICS: REC(h->metrics, ODUCK_NMETRIC, 0, ODUCK_NMETRIC) break;
    case ODUCK_BUF_INFO_RNG: REC(h->info, INFO_STRIDE, INFO_RNG, 2) dt = ODUCK_DTYPE_U32; break;
    case ODUCK_BUF_INFO_COMMAND: REC(h->info, INFO_STRIDE, INFO_CMD, 7) break;
    case ODUCK_BUF_INFO_STEP: REC(h->info, INFO_STRIDE, INFO_STEP, 0) dt = ODUCK_DTYPE_I32; break;
    case ODUCK_BUF_INFO_STEPS: REC(h->info, INFO_STRIDE, INFO_STEPS, 0) dt = ODUCK_DTYPE_I32; break;
    case ODUCK_BUF_INFO_LAST_ACT: REC(h->info, INFO_STRIDE, INFO_LAST_ACT, 3) d2 = m.nu; s1 = 16; break;
    case ODUCK_BUF_INFO_MOTOR_TARGETS: REC(h->info, INFO_STRIDE, INFO_TARGETS, m.nu) break;
    case ODUCK_BUF_INFO_FEET_AIR_TIME: REC(h->info, INFO_STRIDE, INFO_AIR, 2) break;
    case ODUCK_BUF_INFO_LAST_CONTACT: REC(h->info, INFO_STRIDE, INFO_LASTC, 2) break;
    case ODUCK_BUF_INFO_SWING_PEAK: REC(h->info, INFO_STRIDE, INFO_SWING, 2) break;
    case ODUCK_BUF_INFO_PUSH: REC(h->info, INFO_STRIDE, INFO_PUSH, 2) break;
    case ODUCK_BUF_INFO_PUSH_STEP: REC(h->info, INFO_STRIDE, INFO_PUSH_STEP, 0) dt = ODUCK_DTYPE_I32; break;
    case ODUCK_BUF_INFO_PUSH_INTERVAL: REC(h->info, INFO_STRIDE, INFO_PUSH_INT, 0) dt = ODUCK_DTYPE_I32; break;
    case ODUCK_BUF_INFO_ACTION_HISTORY: REC(h->info, INFO_STRIDE, INFO_AHIST, h->hcfg.action_max_delay * m.nu) break;
    case ODUCK_BUF_INFO_IMU_HISTORY: REC(h->info, INFO_STRIDE, INFO_IMUHIST, h->hcfg.imu_max_delay * 3) break;
    case ODUCK_BUF_INFO_IMITATION_I: REC(h->info, INFO_STRIDE, INFO_IMIT_I, 0) dt = ODUCK_DTYPE_I32; break;
    case ODUCK_BUF_INFO_REF_MOTION: REC(h->info, INFO_STRIDE, INFO_REF, ODUCK_REF_DIM) break;
    case ODUCK_BUF_INFO_IMITATION_PHASE: REC(h->info, INFO_STRIDE, INFO_PHASE, 2) break;
    case ODUCK_BUF_DR_PARAMS: REC(h->dr, DR_STRIDE, 0, DR_STRIDE) break;
    case ODUCK_BUF_FIRST_QPOS: REC(h->first_phys, PHYS_STRIDE, 0, m.nq) break;
    case ODUCK_BUF_FIRST_QVEL: REC(h->first_phys, PHYS_STRIDE, PHYS_QVEL, m.nv) break;
    case ODUCK_BUF_FIRST_OBS_STATE: REC(h->first_obs_state, ODUCK_OBS_STATE, 0, h->hcfg.task == ODUCK_TASK_STANDING ? 85 : ODUCK_OBS_STATE) break;
    case ODUCK_BUF_FIRST_OBS_PRIV: REC(h->first_obs_priv, ODUCK_OBS_PRIV, 0, h->hcfg.task == ODUCK_TASK_STANDING ? 153 : ODUCK_OBS_PRIV) break;
    default: return fail(ODUCK_ERR_ARG, "oduck_get_buffer: unknown buffer id");
  }
#undef REC
  *ptr = p;
  shape[0] = h->n; shape[1] = d1; shape[2] = d2; shape[3] = 0;
  strides[0] = s0; strides[1] = d2 ? s1 : 1; strides[2] = 1; strides[3] = 0;
  *dtype = dt;
  return ODUCK_OK;
}

}  // extern "C"
