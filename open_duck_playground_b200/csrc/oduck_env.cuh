// oduck_env.cuh -- Joystick env logic around the physics, one warp per env.
// Restates reference open_duck_mini_v2/joystick.py:206-725 (reset / step / _get_obs / _get_reward / sample_command),
// common/rewards.py, open_duck_mini_v2/custom_rewards.py, common/poly_reference_motion.py:148-168 and the Brax
// Episode/AutoReset wrappers the runner applies (common/runner.py:117).  jax.random is Threefry-2x32 with
// jax_threefry_partitionable=True; independent draws of one call are spread over lanes ("rounds").
#pragma once
#include "oduck_physics.cuh"

#define INFO_STRIDE 256
#define INFO_RNG 0
#define INFO_STEP 2
#define INFO_STEPS 3
#define INFO_PUSH_STEP 4
#define INFO_PUSH_INT 5
#define INFO_IMIT_I 6
#define INFO_CMD 8
#define INFO_LAST_ACT 16      // [3][16]
#define INFO_TARGETS 64       // [16]
#define INFO_AIR 80
#define INFO_LASTC 82
#define INFO_SWING 84
#define INFO_PUSH 86
#define INFO_AHIST 88         // [action_max_delay * nu] (<= 64)
#define INFO_IMUHIST 152      // [imu_max_delay * 3] (<= 16)
#define INFO_REF 168          // [40]
#define INFO_PHASE 208
#define MAX_DELAY 4

struct alignas(16) DevEnvCfg {
  int task;                  // ODUCK_TASK_JOYSTICK / ODUCK_TASK_STANDING
  float sc_orient, sc_head, reset_qvel_noise;
  int n_substeps, episode_length, use_imitation, use_speed_limits, push_enable, act_min_delay, act_max_delay, imu_min_delay, imu_max_delay, auto_reset;
  float ctrl_dt, action_scale, dof_vel_scale, max_motor_velocity, noise_level, noise_gyro, noise_acc, noise_gravity, noise_joint_vel;
  float qpos_noise_scale[16];
  float sc_lin, sc_ang, sc_torques, sc_rate, sc_still, sc_alive, sc_imit, tracking_sigma;
  float push_interval[2], push_magnitude[2], cmd_range[7][2];
  int ndx, ndy, ndth, nb_steps;
  float dxs[8], dys[8], dths[16], dx_range[2], dy_range[2], dth_range[2];
};

// Reward-library parameters (include/oduck.h OduckRewardLibrary) for the RL instantiations of k_step; global memory
struct DevRewardLib {
  float scale[ODUCK_NLIBTERM];
  float base_height_target, max_foot_height, air_thr_min, air_thr_max, swing_freq, swing_amp;
  float soft_lo[16], soft_hi[16], pose_w[16];
  unsigned hip_mask, knee_mask;       // actuator lanes
};

// ---------------------------------------------------------------------------------- jax.random
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t& o0, uint32_t& o1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0; x1 += k1;
#define TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)  x0 += k1; x1 += k2 + 1u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24) x0 += k2; x1 += k0 + 2u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)  x0 += k0; x1 += k1 + 3u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24) x0 += k1; x1 += k2 + 4u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)  x0 += k2; x1 += k0 + 5u;
#undef TF_R
  o0 = x0; o1 = x1;
}
struct RKey { uint32_t a, b; };
// block (0, i) of key k: .a,.b = jax.random.split(k, n)[i];  a ^ b = jax.random.bits(k, (n,))[i]
__device__ __forceinline__ RKey rblock(RKey k, uint32_t i) { RKey o; threefry2x32(k.a, k.b, 0u, i, o.a, o.b); return o; }
__device__ __forceinline__ float bits_unit(uint32_t bits) { return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f; }
__device__ __forceinline__ float bits_uniform(uint32_t bits, float lo, float hi) { return fmaxf(lo, bits_unit(bits) * (hi - lo) + lo); }
__device__ __forceinline__ RKey kshfl(RKey k, int src) { RKey o; o.a = __shfl_sync(FULLMASK, k.a, src); o.b = __shfl_sync(FULLMASK, k.b, src); return o; }
// jax.random.randint(key, (1,), lo, hi)[0], evaluated by the whole warp (2 rounds)
__device__ __forceinline__ int warp_randint(RKey key, int lo, int hi, int lane) {
  RKey kk = rblock(key, lane & 1);            // lanes 0/1: split(key, 2)
  RKey bb = rblock(kk, 0u);
  uint32_t bits = bb.a ^ bb.b;
  uint32_t hb = __shfl_sync(FULLMASK, bits, 0), lb = __shfl_sync(FULLMASK, bits, 1);
  uint32_t span = hi <= lo ? 1u : (uint32_t)(hi - lo);
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  return lo + (int)(((hb % span) * mult + (lb % span)) % span);
}

// joystick.py:671-725; every lane returns command[min(lane, 6)]
__device__ __forceinline__ float sample_command(const DevEnvCfg& c, RKey rng, int lane) {
  RKey k = rblock(rng, lane & 7);              // split(rng, 8): lane i holds rng_{i+1}
  RKey b = rblock(k, 0u);
  const uint32_t bits = b.a ^ b.b;
  const float u4 = bits_unit(__shfl_sync(FULLMASK, bits, 3));   // rng4 -> bernoulli(p = 0.1)
  const int which = lane < 3 ? lane : lane + 1;                  // x,y,yaw <- rng1..3 ; neck..roll <- rng5..8
  const uint32_t mybits = __shfl_sync(FULLMASK, bits, which & 7);
  const int ci = lane < 7 ? lane : 6;
  float v = bits_uniform(mybits, c.cmd_range[ci][0], c.cmd_range[ci][1]);
  if (c.task == 1 && lane < 3) v = 0.f;                          // standing.py:648-655: no velocity command (same key usage)
  return (u4 < 0.1f) ? 0.f : v;
}

// poly_reference_motion.py:148-168: lane d gets ref[d] (d < 32) and ref[32 + d] (d < 8)
__device__ __forceinline__ void reference_motion(const DevEnvCfg& c, const float* __restrict__ poly, float dx, float dy, float dth, int i, int lane,
                                                 float& r_lo, float& r_hi) {
  auto nearest = [](float v, const float* grid, int n, const float* range) {
    v = fminf(fmaxf(v, range[0]), range[1]);
    int best = 0;
    float bd = __int_as_float(0x7f800000);
    for (int k = 0; k < n; ++k) { float d = fabsf(grid[k] - v); if (d < bd) { bd = d; best = k; } }
    return best;
  };
  const int ix = nearest(dx, c.dxs, c.ndx, c.dx_range), iy = nearest(dy, c.dys, c.ndy, c.dy_range), it = nearest(dth, c.dths, c.ndth, c.dth_range);
  float t = (float)(i % c.nb_steps) / (float)c.nb_steps;
  t = fminf(fmaxf(t, 0.f), 1.f);
  const float4* base = reinterpret_cast<const float4*>(poly + (size_t)((ix * c.ndy + iy) * c.ndth + it) * (40 * 16));
  auto horner = [&](int d) {
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 cf = __ldg(base + d * 4 + q);
      acc = acc * t + cf.x; acc = acc * t + cf.y; acc = acc * t + cf.z; acc = acc * t + cf.w;
    }
    return acc;
  };
  r_lo = horner(lane);
  r_hi = horner(32 + (lane & 7));
}

// Per-warp env registers shared by reset and step
struct EnvRegs {
  RKey rng;
  int step, steps, push_step, push_interval, imitation_i;
  float cmd;          // lane i < 7: command[i]
  float last_act[3];  // lane u < nu
  float targets;      // lane u < nu: motor_targets[u]
  float hist[MAX_DELAY];
  float air, lastc, swing;  // lane k < 2
  float ref_lo, ref_hi;     // lane d: ref[d], ref[32 + (d & 7)]
  float phase;              // lane k < 2: imitation_phase[k]
};

// _get_obs (joystick.py:487-620).  Reads the staged outputs of the last forward from s.outrec, advances er.rng by 5
// splits, writes obs["state"] (101) and obs["privileged_state"] (212).  contact: lane k < 2.
__device__ __forceinline__ void write_obs(const DevModel& m, const DevEnvCfg& c, WarpSmem& s, const float* orec, const Lane& L, EnvRegs& er,
                                          float contact, int lane, float* __restrict__ ost, float* __restrict__ opr, float* __restrict__ imu_hist) {
  // five chained splits: rng_{k+1}, noise_k = split(rng_k)
  RKey nk[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    RKey o = rblock(er.rng, lane & 1);
    nk[k] = kshfl(o, 1);
    er.rng = kshfl(o, 0);
  }
  // noise draws: lanes 0-2 gyro(nk0), 3-5 accel(nk1), 6-8 gravity(nk2); then 14 joint angles (nk3) and 14 joint vels (nk4)
  float n9;
  {
    const int g = lane / 3, e = lane % 3;
    RKey kk = g == 0 ? nk[0] : (g == 1 ? nk[1] : nk[2]);
    RKey b = rblock(kk, (uint32_t)e);
    n9 = 2.f * bits_unit(b.a ^ b.b) - 1.f;
  }
  float nja, njv;
  {
    RKey b = rblock(nk[3], (uint32_t)(lane & 15));
    nja = 2.f * bits_unit(b.a ^ b.b) - 1.f;
    RKey b2 = rblock(nk[4], (uint32_t)(lane & 15));
    njv = 2.f * bits_unit(b2.a ^ b2.b) - 1.f;
  }
  const float lvl = c.noise_level;
  const float* sd = orec + OUT_SENS;
  // lanes 0..8: [gyro, accel, gravity] + noise
  float base9 = 0.f, sc9 = 0.f;
  if (lane < 3) { base9 = sd[lane]; sc9 = c.noise_gyro; }
  else if (lane < 6) { base9 = sd[6 + lane - 3]; sc9 = c.noise_acc; }   // the +1.3 bias of joystick.py:502 is discarded by the reference
  else if (lane < 9) { base9 = -orec[OUT_IMUMAT + 6 + lane - 6]; sc9 = c.noise_gravity; }
  const float noisy9 = base9 + n9 * lvl * sc9;
  // imu history: roll by 3, head <- noisy gravity (never observed, kept for info parity)
  {
    const int nh = c.imu_max_delay * 3;
    float prev = (lane >= 3 && lane < nh) ? imu_hist[lane - 3] : 0.f;
    float ng = __shfl_sync(FULLMASK, noisy9, 6 + (lane % 3));
    __syncwarp();
    if (lane < nh) imu_hist[lane] = lane < 3 ? ng : prev;
  }
  // joint angles (+ backlash) and velocities, lane = actuator
  float ja = 0.f, jv = 0.f, afrc = 0.f;
  if (lane < m.nu) {
    ja = s.qpos[m.act_qadr[lane]];
    if (m.act_bl_qadr[lane] >= 0) ja += s.qpos[m.act_bl_qadr[lane]];
    afrc = orec[OUT_AFRC + lane];
  }
  {
    const int ad = lane < m.nu ? m.act_dof[lane] : 0;
    jv = __shfl_sync(FULLMASK, L.qvel, ad);
  }
  const float dflt = lane < m.nu ? m.key_ctrl[lane] : 0.f;
  const int nu = m.nu;
  const bool standing = c.task == 1;   // standing.py:526-566: no motor_targets / imitation_phase, empty reference motion
  if (lane < 6) { ost[lane] = noisy9; opr[lane] = noisy9; }
  if (lane < 7) { ost[6 + lane] = er.cmd; opr[6 + lane] = er.cmd; }
  if (lane < nu) {
    const float a = ja + nja * lvl * c.qpos_noise_scale[lane] - dflt;
    const float v = (jv + njv * lvl * c.noise_joint_vel) * c.dof_vel_scale;
    int p = 13;
    ost[p + lane] = a; opr[p + lane] = a; p += nu;
    ost[p + lane] = v; opr[p + lane] = v; p += nu;
#pragma unroll
    for (int k = 0; k < 3; ++k) { ost[p + lane] = er.last_act[k]; opr[p + lane] = er.last_act[k]; p += nu; }
    if (!standing) { ost[p + lane] = er.targets; opr[p + lane] = er.targets; }
  }
  const int p2 = 13 + (standing ? 5 : 6) * nu;
  if (lane < 2) {
    ost[p2 + lane] = contact; opr[p2 + lane] = contact;
    if (!standing) { ost[p2 + 2 + lane] = er.phase; opr[p2 + 2 + lane] = er.phase; }
  }
  // privileged tail
  int r = p2 + (standing ? 2 : 4);
  if (lane < 3) {
    opr[r + lane] = sd[lane];                         // gyro
    opr[r + 3 + lane] = sd[6 + lane];                 // accelerometer
    opr[r + 6 + lane] = -orec[OUT_IMUMAT + 6 + lane]; // gravity
    opr[r + 9 + lane] = sd[3 + lane];                 // local linvel
    opr[r + 12 + lane] = sd[12 + lane];               // global angvel
  }
  r += 15;
  if (lane < nu) { opr[r + lane] = ja - dflt; opr[r + nu + lane] = jv; }
  r += 2 * nu;
  if (lane == 0) opr[r] = s.qpos[2];
  r += 1;
  if (lane < nu) opr[r + lane] = afrc;
  r += nu;
  if (lane < 2) opr[r + lane] = contact;
  r += 2;
  if (lane < 6) opr[r + lane] = sd[15 + lane];
  r += 6;
  if (lane < 2) opr[r + lane] = er.air;
  r += 2;
  if (standing) return;
  opr[r + lane] = er.ref_lo;
  if (lane < 8) opr[r + 32 + lane] = er.ref_hi;
  r += 40;
  if (lane == 0) opr[r] = (float)er.imitation_i;
  if (lane < 2) opr[r + 1 + lane] = er.phase;
}

__device__ __forceinline__ float nan_to_num(float x) {
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0.f ? 3.4028234664e38f : -3.4028234664e38f;
  return x;
}
