// oduck_oracle.cpp -- CPU restatement (the ORACLE) of the Open Duck Mini V2 joystick hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under open_duck_playground_b200/ may import, link or call
// this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
// check / time the CUDA path against an independent implementation of the same algorithm.
//
// PARITY UNPINNED for the physics: the arithmetic of the reference's step lives in mujoco.mjx,
// which is not vendored in the reference, not installable here (no network), and the reference
// holds no golden vectors for it.  The physics below restates the published
// MuJoCo/MJX algorithms (MJX 3.x: forward.py, smooth.py, collision_convex.py plane_convex,
// constraint.py, solver.py, sensor.py) anchored on the reference's call sites:
//   mjx_env.step / mjx_env.init      open_duck_mini_v2/joystick.py:258,420
//   geoms_colliding                  open_duck_mini_v2/joystick.py:313-318,424-429
//   get_sensor_data                  open_duck_mini_v2/base.py:234-264
// PINNED parts (checked against the reference's own NumPy twins / data in tests/):
//   rewards            common/rewards.py:11-241           (twin: common/rewards_numpy.py)
//   imitation reward   open_duck_mini_v2/custom_rewards.py:4-149 (twin: custom_rewards_numpy.py)
//   reference motion   common/poly_reference_motion.py:148-168   (twin: poly_reference_motion_numpy.py)
// The env logic follows open_duck_mini_v2/joystick.py:206-725 line by line, jax.random is restated as
// Threefry-2x32 with jax_threefry_partitionable=True (JAX >= 0.5 default), and the Brax
// Episode/AutoReset wrappers (common/runner.py:117) are fused into oduck_step.
//
// Build: see oracle/Makefile.  ODUCK_REAL selects the arithmetic type (double = checker,
// float + OpenMP = the CPU baseline timed by bench.py).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../include/oduck.h"

#ifdef ODUCK_COUNT_FLOPS
// Op-counter build ("count from the oracle with an op-counter build"): `real` is a double that counts every
// arithmetic operation the restated algorithm executes (+ - * / and sqrt = 1 flop, transcendental = 1, comparisons / moves = 0).
// Single-threaded use only (ODUCK_THREADS=1); read with oduck_flop_count().
static uint64_t g_flops = 0;
struct CountReal {
  double v;
  CountReal() = default;
  CountReal(double x) : v(x) {}
  CountReal(float x) : v(x) {}
  CountReal(int x) : v(x) {}
  CountReal(unsigned x) : v(x) {}
  CountReal(long x) : v((double)x) {}
  CountReal(long long x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return (float)v; }
  explicit operator int() const { return (int)v; }
  explicit operator bool() const { return v != 0.0; }
  CountReal operator-() const { return CountReal(-v); }
  CountReal& operator+=(CountReal o) { ++g_flops; v += o.v; return *this; }
  CountReal& operator-=(CountReal o) { ++g_flops; v -= o.v; return *this; }
  CountReal& operator*=(CountReal o) { ++g_flops; v *= o.v; return *this; }
  CountReal& operator/=(CountReal o) { ++g_flops; v /= o.v; return *this; }
};
#define CR_BIN(op) \
  inline CountReal operator op(CountReal a, CountReal b) { ++g_flops; return CountReal(a.v op b.v); } \
  inline CountReal operator op(CountReal a, double b) { ++g_flops; return CountReal(a.v op b); } \
  inline CountReal operator op(double a, CountReal b) { ++g_flops; return CountReal(a op b.v); } \
  inline CountReal operator op(CountReal a, int b) { ++g_flops; return CountReal(a.v op b); } \
  inline CountReal operator op(int a, CountReal b) { ++g_flops; return CountReal(a op b.v); }
CR_BIN(+) CR_BIN(-) CR_BIN(*) CR_BIN(/)
#undef CR_BIN
#define CR_CMP(op) \
  inline bool operator op(CountReal a, CountReal b) { return a.v op b.v; } \
  inline bool operator op(CountReal a, double b) { return a.v op b; } \
  inline bool operator op(double a, CountReal b) { return a op b.v; } \
  inline bool operator op(CountReal a, int b) { return a.v op b; } \
  inline bool operator op(int a, CountReal b) { return a op b.v; }
CR_CMP(<) CR_CMP(>) CR_CMP(<=) CR_CMP(>=) CR_CMP(==) CR_CMP(!=)
#undef CR_CMP
namespace std {
#define CR_FN1(f) inline CountReal f(CountReal x) { ++g_flops; return CountReal(f(x.v)); }
CR_FN1(sqrt) CR_FN1(sin) CR_FN1(cos) CR_FN1(exp) CR_FN1(log) CR_FN1(log1p) CR_FN1(tanh)
#undef CR_FN1
inline CountReal fabs(CountReal x) { return CountReal(fabs(x.v)); }
inline CountReal nearbyint(CountReal x) { return CountReal(nearbyint(x.v)); }
inline CountReal pow(CountReal a, CountReal b) { ++g_flops; return CountReal(pow(a.v, b.v)); }
inline bool isnan(CountReal x) { return isnan(x.v); }
inline bool isinf(CountReal x) { return isinf(x.v); }
}  // namespace std
typedef CountReal real;
#else
#ifndef ODUCK_REAL
#define ODUCK_REAL double
#endif
typedef ODUCK_REAL real;
#endif

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }


#include <thread>
// Minimal parallel-for over envs (the image has no libgomp).  ODUCK_THREADS overrides the thread count.
static int g_threads_used = 1;
template <typename F> static void pfor(int n, F f) {
  int nt = (int)std::thread::hardware_concurrency();
  if (const char* s = getenv("ODUCK_THREADS")) nt = atoi(s);
  nt = std::max(1, std::min(nt, n));
  g_threads_used = nt;
  if (nt == 1) { for (int i = 0; i < n; i++) f(i); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++) th.emplace_back([=]() { for (int i = (int)((int64_t)n * t / nt); i < (int)((int64_t)n * (t + 1) / nt); i++) f(i); });
  for (auto& t : th) t.join();
}

static const real kMinVal = (real)1e-15;
static const real kMinImp = (real)0.0001, kMaxImp = (real)0.9999;

#define NB ODUCK_MAX_BODY
#define NJ ODUCK_MAX_JNT
#define NQ ODUCK_MAX_NQ
#define NV ODUCK_MAX_NV
#define NU ODUCK_MAX_NU
#define NCON ODUCK_MAX_CON
#define NEFC (ODUCK_MAX_NU + ODUCK_MAX_JNT + 4 * ODUCK_MAX_CON)

// ------------------------------------------------------------------------------------ small math
static inline void cross3(const real* a, const real* b, real* o) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void quat_mul(const real* a, const real* b, real* o) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
static inline void quat_norm(real* q) {
  real n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < kMinVal) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int i = 0; i < 4; i++) q[i] /= n;
}
static inline void quat2mat(const real* q, real* m) {  // row-major 3x3
  real w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}
static inline void mat_vec(const real* m, const real* v, real* o) {
  real x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
       z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void matT_vec(const real* m, const real* v, real* o) {
  real x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
       z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
// spatial (MuJoCo convention: [angular(3), linear(3)])
static inline void cross_motion(const real* v, const real* m, real* o) {  // mju_crossMotion
  real a[3], b[3], c[3];
  cross3(v, m, a);
  cross3(v, m + 3, b);
  cross3(v + 3, m, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
static inline void cross_force(const real* v, const real* f, real* o) {  // mju_crossForce
  real a[3], b[3], c[3];
  cross3(v, f, a);
  cross3(v + 3, f + 3, b);
  cross3(v, f + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
// 10-number spatial inertia about a reference point: [Ixx Iyy Izz Ixy Ixz Iyz, m*dx m*dy m*dz, m]
static inline void inert_mul(const real* I, const real* v, real* o) {  // mju_mulInertVec
  real r0 = I[0] * v[0] + I[3] * v[1] + I[4] * v[2] - I[8] * v[4] + I[7] * v[5];
  real r1 = I[3] * v[0] + I[1] * v[1] + I[5] * v[2] + I[8] * v[3] - I[6] * v[5];
  real r2 = I[4] * v[0] + I[5] * v[1] + I[2] * v[2] - I[7] * v[3] + I[6] * v[4];
  real r3 = I[8] * v[1] - I[7] * v[2] + I[9] * v[3];
  real r4 = I[6] * v[2] - I[8] * v[0] + I[9] * v[4];
  real r5 = I[7] * v[0] - I[6] * v[1] + I[9] * v[5];
  o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4; o[5] = r5;
}

// ------------------------------------------------------------------------------------ jax.random
static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* o0, uint32_t* o1) {
  const int R[8] = {13, 15, 26, 6, 17, 29, 16, 24};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0]; x1 += ks[1];
  for (int blk = 0; blk < 5; blk++) {
    const int* r = R + 4 * (blk & 1);
    for (int i = 0; i < 4; i++) { x0 += x1; x1 = rotl(x1, r[i]); x1 ^= x0; }
    x0 += ks[(blk + 1) % 3];
    x1 += ks[(blk + 2) % 3] + (uint32_t)(blk + 1);
  }
  *o0 = x0; *o1 = x1;
}
struct Key { uint32_t a, b; };
// jax.random.split(key, n)[i] under threefry_partitionable: both output words of block (0, i)
static inline Key key_split(Key k, uint32_t i) { Key o; threefry2x32(k.a, k.b, 0u, i, &o.a, &o.b); return o; }
// jax.random.bits(key, shape)[i], 32 bit: xor of the two output words of block (0, i)
static inline uint32_t key_bits(Key k, uint32_t i) { uint32_t a, b; threefry2x32(k.a, k.b, 0u, i, &a, &b); return a ^ b; }
static inline real bits_to_unit(uint32_t bits) {
  uint32_t u = (bits >> 9) | 0x3F800000u;
  float f; std::memcpy(&f, &u, 4);
  return (real)(f - 1.0f);
}
// jax.random.uniform(key, shape, minval, maxval)[i]
static inline real key_uniform(Key k, uint32_t i, real lo, real hi) {
  real v = bits_to_unit(key_bits(k, i)) * (hi - lo) + lo;
  return std::max(lo, v);
}
// jax.random.randint(key, (1,), lo, hi)[0]
static inline int key_randint(Key k, int lo, int hi) {
  Key k1 = key_split(k, 0), k2 = key_split(k, 1);
  uint32_t hb = key_bits(k1, 0), lb = key_bits(k2, 0);
  uint32_t span = hi <= lo ? 1u : (uint32_t)(hi - lo);
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  uint32_t off = ((hb % span) * mult + (lb % span)) % span;
  return lo + (int)off;
}

// ------------------------------------------------------------------------------------ state
struct EnvState {
  // mjx.Data subset
  real qpos[NQ], qvel[NV], qacc_warm[NV], qacc[NV], ctrl[NU];
  // per-env randomised model (common/randomize.py)
  real dr_geom_friction0;  // written, never read by physics: geom 0 is a visual mesh
  real body_mass[NB], body_ipos[NB][3], dof_frictionloss[NV], dof_armature[NV], qpos0[NQ], act_kp[NU];
  // outputs of the last forward()
  real sensordata[24], efc_force[NEFC], contact_dist[NCON], actuator_force[NU], site_xpos_feet[6], imu_xmat[9];
  // State
  real obs_state[ODUCK_OBS_STATE], obs_priv[ODUCK_OBS_PRIV], reward, done, truncation, metrics[ODUCK_NMETRIC];
  // info (joystick.py:278-302)
  Key rng;
  int32_t step, steps, push_step, push_interval_steps, imitation_i;
  real command[ODUCK_NCMD], last_act[3][NU], motor_targets[NU], feet_air_time[2], last_contact[2], swing_peak[2], push[2];
  real action_history[8 * NU], imu_history[8 * 3], ref_motion[ODUCK_REF_DIM], imitation_phase[2];
  // auto-reset target
  real first_qpos[NQ], first_qvel[NV], first_qacc_warm[NV], first_obs_state[ODUCK_OBS_STATE], first_obs_priv[ODUCK_OBS_PRIV];
  // forward() outputs that have to survive auto-reset bookkeeping are recomputed, not stored
};

struct OduckHandle {
  OduckModel m;
  OduckEnvConfig cfg;
  std::vector<double> poly;
  std::vector<float> hfield;   // [nrow][ncol] elevation in [0, 1] (floor_is_hfield)
  int n;
  std::vector<EnvState> env;
  int nefc_fr, nefc_lim;  // static row counts
  int fr_dof[NV], lim_jnt[NJ];
  int64_t launches;
  OduckRolloutSink sink = {};   // attached rollout buffers (obs_policy == null: none)
  std::vector<float> act_buf;
};

// ------------------------------------------------------------------------------------ physics scratch
struct Scratch {
  real xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9];
  real xanchor[NJ][3], xaxis[NJ][3];
  real site_xpos[ODUCK_MAX_SITE][3], site_xmat[ODUCK_MAX_SITE][9];
  real com[3];  // subtree_com of the robot's kinematic-tree root
  real cinert[NB][10], crb[NB][10], cdof[NV][6], cdof_dot[NV][6], cvel[NB][6];
  real M[NV][NV], L[NV][NV];
  real qfrc_bias[NV], qfrc_passive[NV], qfrc_actuator[NV], qfrc_smooth[NV], qacc_smooth[NV];
  // contacts
  real con_dist[NCON], con_pos[NCON][3], con_frame[NCON][9], con_mu[NCON];
  int con_b1[NCON], con_b2[NCON];
  // constraints
  int nefc;
  real J[NEFC][NV], D[NEFC], aref[NEFC], floss[NEFC];
  int rtype[NEFC];  // 0 friction, 1 limit/contact (one-sided)
};

static bool cholesky(int n, real A[NV][NV], real L[NV][NV]) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      real s = A[i][j];
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0)) return false;
        L[i][i] = std::sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  return true;
}
static void chol_solve(int n, real L[NV][NV], const real* b, real* x) {
  real y[NV];
  for (int i = 0; i < n; i++) {
    real s = b[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = n - 1; i >= 0; i--) {
    real s = y[i];
    for (int k = i + 1; k < n; k++) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
  }
}

// mjx/_src/smooth.py kinematics + com_pos
static void kinematics(const OduckModel& m, const EnvState& e, Scratch& s) {
  s.xpos[0][0] = s.xpos[0][1] = s.xpos[0][2] = 0;
  s.xquat[0][0] = 1; s.xquat[0][1] = s.xquat[0][2] = s.xquat[0][3] = 0;
  quat2mat(s.xquat[0], s.xmat[0]);
  for (int b = 1; b < m.nbody; b++) {
    int p = m.body_parentid[b];
    real pos[3], quat[4], bp[3] = {(real)m.body_pos[b][0], (real)m.body_pos[b][1], (real)m.body_pos[b][2]};
    real bq[4] = {(real)m.body_quat[b][0], (real)m.body_quat[b][1], (real)m.body_quat[b][2], (real)m.body_quat[b][3]};
    mat_vec(s.xmat[p], bp, pos);
    for (int i = 0; i < 3; i++) pos[i] += s.xpos[p][i];
    quat_mul(s.xquat[p], bq, quat);
    for (int k = 0; k < m.body_jntnum[b]; k++) {
      int j = m.body_jntadr[b] + k, qa = m.jnt_qposadr[j];
      if (m.jnt_type[j] == ODUCK_JNT_FREE) {
        for (int i = 0; i < 3; i++) pos[i] = e.qpos[qa + i];
        for (int i = 0; i < 4; i++) quat[i] = e.qpos[qa + 3 + i];
        quat_norm(quat);
        for (int i = 0; i < 3; i++) { s.xanchor[j][i] = pos[i]; s.xaxis[j][i] = (i == 2); }
      } else {
        real R[9], jp_[3] = {(real)m.jnt_pos[j][0], (real)m.jnt_pos[j][1], (real)m.jnt_pos[j][2]};
        real ja[3] = {(real)m.jnt_axis[j][0], (real)m.jnt_axis[j][1], (real)m.jnt_axis[j][2]}, t[3];
        quat2mat(quat, R);
        mat_vec(R, jp_, t);
        for (int i = 0; i < 3; i++) s.xanchor[j][i] = pos[i] + t[i];
        mat_vec(R, ja, s.xaxis[j]);
        real ang = e.qpos[qa] - e.qpos0[qa];
        real ql[4] = {std::cos(ang / 2), std::sin(ang / 2) * ja[0], std::sin(ang / 2) * ja[1], std::sin(ang / 2) * ja[2]}, qn[4];
        quat_mul(quat, ql, qn);
        for (int i = 0; i < 4; i++) quat[i] = qn[i];
        quat2mat(quat, R);
        mat_vec(R, jp_, t);
        for (int i = 0; i < 3; i++) pos[i] = s.xanchor[j][i] - t[i];
      }
    }
    quat_norm(quat);
    for (int i = 0; i < 3; i++) s.xpos[b][i] = pos[i];
    for (int i = 0; i < 4; i++) s.xquat[b][i] = quat[i];
    quat2mat(quat, s.xmat[b]);
    mat_vec(s.xmat[b], e.body_ipos[b], s.xipos[b]);
    for (int i = 0; i < 3; i++) s.xipos[b][i] += s.xpos[b][i];
    real iq[4] = {(real)m.body_iquat[b][0], (real)m.body_iquat[b][1], (real)m.body_iquat[b][2], (real)m.body_iquat[b][3]}, q2[4];
    quat_mul(quat, iq, q2);
    quat2mat(q2, s.ximat[b]);
  }
  for (int k = 0; k < m.nsite; k++) {
    int b = m.site_bodyid[k];
    real sp[3] = {(real)m.site_pos[k][0], (real)m.site_pos[k][1], (real)m.site_pos[k][2]};
    real sq[4] = {(real)m.site_quat[k][0], (real)m.site_quat[k][1], (real)m.site_quat[k][2], (real)m.site_quat[k][3]}, q2[4];
    mat_vec(s.xmat[b], sp, s.site_xpos[k]);
    for (int i = 0; i < 3; i++) s.site_xpos[k][i] += s.xpos[b][i];
    quat_mul(s.xquat[b], sq, q2);
    quat2mat(q2, s.site_xmat[k]);
  }
}

static bool body_moves(const OduckModel& m, int b) {
  while (b > 0) { if (m.body_dofnum[b] > 0) return true; b = m.body_parentid[b]; }
  return false;
}

static void com_pos(const OduckModel& m, const EnvState& e, Scratch& s) {
  // subtree COM of the (single) moving kinematic tree; static bodies (floor) refer to the world origin
  real mt = 0, c[3] = {0, 0, 0};
  for (int b = 1; b < m.nbody; b++)
    if (body_moves(m, b)) {
      mt += e.body_mass[b];
      for (int i = 0; i < 3; i++) c[i] += e.body_mass[b] * s.xipos[b][i];
    }
  for (int i = 0; i < 3; i++) s.com[i] = c[i] / std::max(mt, kMinVal);
  for (int b = 0; b < m.nbody; b++) {
    real* I = s.cinert[b];
    for (int i = 0; i < 10; i++) I[i] = 0;
    if (b == 0 || !body_moves(m, b)) continue;
    // mju_inertCom: rotate the principal inertia into the world, shift to the COM reference point
    const real* R = s.ximat[b];
    real din[3] = {(real)m.body_inertia[b][0], (real)m.body_inertia[b][1], (real)m.body_inertia[b][2]};
    real Iw[9];
    for (int r = 0; r < 3; r++)
      for (int cc = 0; cc < 3; cc++) Iw[3 * r + cc] = R[3 * r] * din[0] * R[3 * cc] + R[3 * r + 1] * din[1] * R[3 * cc + 1] + R[3 * r + 2] * din[2] * R[3 * cc + 2];
    real d[3] = {s.xipos[b][0] - s.com[0], s.xipos[b][1] - s.com[1], s.xipos[b][2] - s.com[2]};
    real ms = e.body_mass[b];
    I[0] = Iw[0] + ms * (d[1] * d[1] + d[2] * d[2]);
    I[1] = Iw[4] + ms * (d[0] * d[0] + d[2] * d[2]);
    I[2] = Iw[8] + ms * (d[0] * d[0] + d[1] * d[1]);
    I[3] = Iw[1] - ms * d[0] * d[1];
    I[4] = Iw[2] - ms * d[0] * d[2];
    I[5] = Iw[5] - ms * d[1] * d[2];
    I[6] = ms * d[0]; I[7] = ms * d[1]; I[8] = ms * d[2];
    I[9] = ms;
  }
  for (int j = 0; j < m.njnt; j++) {
    int da = m.jnt_dofadr[j], b = m.jnt_bodyid[j];
    if (m.jnt_type[j] == ODUCK_JNT_FREE) {
      for (int k = 0; k < 3; k++) {
        for (int i = 0; i < 6; i++) s.cdof[da + k][i] = 0;
        s.cdof[da + k][3 + k] = 1;
      }
      real off[3] = {s.com[0] - s.xpos[b][0], s.com[1] - s.xpos[b][1], s.com[2] - s.xpos[b][2]};
      for (int k = 0; k < 3; k++) {
        real ax[3] = {s.xmat[b][k], s.xmat[b][3 + k], s.xmat[b][6 + k]};
        for (int i = 0; i < 3; i++) s.cdof[da + 3 + k][i] = ax[i];
        cross3(ax, off, s.cdof[da + 3 + k] + 3);
      }
    } else {
      real off[3] = {s.com[0] - s.xanchor[j][0], s.com[1] - s.xanchor[j][1], s.com[2] - s.xanchor[j][2]};
      for (int i = 0; i < 3; i++) s.cdof[da][i] = s.xaxis[j][i];
      cross3(s.xaxis[j], off, s.cdof[da] + 3);
    }
  }
}

// composite rigid body + dense mass matrix (smooth.py crb), dense Cholesky (factor_m, dense branch)
static void crb(const OduckModel& m, const EnvState& e, Scratch& s) {
  for (int b = 0; b < m.nbody; b++)
    for (int i = 0; i < 10; i++) s.crb[b][i] = s.cinert[b][i];
  for (int b = m.nbody - 1; b > 0; b--) {
    int p = m.body_parentid[b];
    if (p > 0)
      for (int i = 0; i < 10; i++) s.crb[p][i] += s.crb[b][i];
  }
  for (int i = 0; i < m.nv; i++)
    for (int j = 0; j < m.nv; j++) s.M[i][j] = 0;
  for (int i = 0; i < m.nv; i++) {
    real buf[6];
    inert_mul(s.crb[m.dof_bodyid[i]], s.cdof[i], buf);
    s.M[i][i] = e.dof_armature[i];
    for (int j = i; j >= 0; j = m.dof_parentid[j]) {
      real v = 0;
      for (int k = 0; k < 6; k++) v += s.cdof[j][k] * buf[k];
      s.M[i][j] += v;
      if (j != i) s.M[j][i] = s.M[i][j];
    }
  }
}

static void com_vel(const OduckModel& m, const EnvState& e, Scratch& s) {
  for (int i = 0; i < 6; i++) s.cvel[0][i] = 0;
  for (int b = 1; b < m.nbody; b++) {
    real cv[6];
    for (int i = 0; i < 6; i++) cv[i] = s.cvel[m.body_parentid[b]][i];
    for (int k = 0; k < m.body_jntnum[b]; k++) {
      int j = m.body_jntadr[b] + k, da = m.jnt_dofadr[j];
      if (m.jnt_type[j] == ODUCK_JNT_FREE) {
        for (int d = 0; d < 3; d++) {
          for (int i = 0; i < 6; i++) { s.cdof_dot[da + d][i] = 0; cv[i] += s.cdof[da + d][i] * e.qvel[da + d]; }
        }
        for (int d = 3; d < 6; d++) cross_motion(cv, s.cdof[da + d], s.cdof_dot[da + d]);
        for (int d = 3; d < 6; d++)
          for (int i = 0; i < 6; i++) cv[i] += s.cdof[da + d][i] * e.qvel[da + d];
      } else {
        cross_motion(cv, s.cdof[da], s.cdof_dot[da]);
        for (int i = 0; i < 6; i++) cv[i] += s.cdof[da][i] * e.qvel[da];
      }
    }
    for (int i = 0; i < 6; i++) s.cvel[b][i] = cv[i];
  }
}

// recursive Newton-Euler bias force (smooth.py rne); with_acc adds cdof*qacc (rne_postconstraint cacc)
static void rne(const OduckModel& m, const EnvState& e, Scratch& s, const real* qacc, real cacc[NB][6], real* qfrc_bias) {
  real cfrc[NB][6];
  for (int i = 0; i < 3; i++) { cacc[0][i] = 0; cacc[0][3 + i] = -(real)m.gravity[i]; }
  for (int i = 0; i < 6; i++) cfrc[0][i] = 0;
  for (int b = 1; b < m.nbody; b++) {
    for (int i = 0; i < 6; i++) cacc[b][i] = cacc[m.body_parentid[b]][i];
    for (int d = m.body_dofadr[b]; d >= 0 && d < m.body_dofadr[b] + m.body_dofnum[b]; d++)
      for (int i = 0; i < 6; i++) cacc[b][i] += s.cdof_dot[d][i] * e.qvel[d] + (qacc ? s.cdof[d][i] * qacc[d] : 0);
    real t1[6], t2[6], t3[6];
    inert_mul(s.cinert[b], cacc[b], t1);
    inert_mul(s.cinert[b], s.cvel[b], t2);
    cross_force(s.cvel[b], t2, t3);
    for (int i = 0; i < 6; i++) cfrc[b][i] = t1[i] + t3[i];
  }
  if (!qfrc_bias) return;
  for (int b = m.nbody - 1; b > 0; b--)
    for (int i = 0; i < 6; i++) cfrc[m.body_parentid[b]][i] += cfrc[b][i];
  for (int d = 0; d < m.nv; d++) {
    real v = 0;
    for (int i = 0; i < 6; i++) v += s.cdof[d][i] * cfrc[m.dof_bodyid[d]][i];
    qfrc_bias[d] = v;
  }
}

// math.make_frame (MJX): orthonormal frame whose first row is the normal
static void make_frame(const real* n, real* f) {
  real a[3] = {n[0], n[1], n[2]};
  real nn = std::sqrt(dot3(a, a));
  for (int i = 0; i < 3; i++) a[i] /= nn;
  real y[3] = {0, 1, 0}, z[3] = {0, 0, 1};
  real* b0 = (-0.5 < a[1] && a[1] < 0.5) ? y : z;
  real b[3];
  real d = dot3(a, b0);
  for (int i = 0; i < 3; i++) b[i] = b0[i] - a[i] * d;
  real bn = std::sqrt(dot3(b, b));
  for (int i = 0; i < 3; i++) b[i] /= bn;
  real c[3];
  cross3(a, b, c);
  for (int i = 0; i < 3; i++) { f[i] = a[i]; f[3 + i] = b[i]; f[6 + i] = c[i]; }
}

// collision_convex.py _manifold_points: four polygon points of roughly maximal area.  The scores are rounded to float32
// before the argmax because MJX runs in float32, where `x + (-1e6)` ties for every masked-out vertex.
static void manifold_points(int nv, const real (*poly)[3], const bool* mask, const real* n, int* idx) {
  auto dmask = [&](int i) { return mask[i] ? (real)0 : (real)-1e6; };
  int a = 0;
  { real best = -std::numeric_limits<real>::infinity(); for (int i = 0; i < nv; i++) if (dmask(i) > best) { best = dmask(i); a = i; } }
  int b = 0;
  { real best = -std::numeric_limits<real>::infinity();
    for (int i = 0; i < nv; i++) { real d[3] = {poly[a][0] - poly[i][0], poly[a][1] - poly[i][1], poly[a][2] - poly[i][2]}; real v = (real)((float)dot3(d, d) + (float)dmask(i)); if (v > best) { best = v; b = i; } } }
  real ab[3], amb[3] = {poly[a][0] - poly[b][0], poly[a][1] - poly[b][1], poly[a][2] - poly[b][2]};
  cross3(n, amb, ab);
  int c = 0;
  { real best = -std::numeric_limits<real>::infinity();
    for (int i = 0; i < nv; i++) { real ap[3] = {poly[a][0] - poly[i][0], poly[a][1] - poly[i][1], poly[a][2] - poly[i][2]}; real v = (real)((float)std::fabs(dot3(ap, ab)) + (float)dmask(i)); if (v > best) { best = v; c = i; } } }
  real ac[3], bc[3], amc[3] = {poly[a][0] - poly[c][0], poly[a][1] - poly[c][1], poly[a][2] - poly[c][2]};
  real bmc[3] = {poly[b][0] - poly[c][0], poly[b][1] - poly[c][1], poly[b][2] - poly[c][2]};
  cross3(n, amc, ac);
  cross3(n, bmc, bc);
  int d = 0;
  { real best = -std::numeric_limits<real>::infinity();
    // argmax over concatenate([dist_bp, dist_ap]) % nv: first maximum, bp block first
    for (int i = 0; i < nv; i++) { real bp[3] = {poly[b][0] - poly[i][0], poly[b][1] - poly[i][1], poly[b][2] - poly[i][2]}; real v = (real)((float)std::fabs(dot3(bp, bc)) + (float)dmask(i)); if (v > best) { best = v; d = i; } }
    for (int i = 0; i < nv; i++) { real ap[3] = {poly[a][0] - poly[i][0], poly[a][1] - poly[i][1], poly[a][2] - poly[i][2]}; real v = (real)((float)std::fabs(dot3(ap, ac)) + (float)dmask(i)); if (v > best) { best = v; d = i; } } }
  idx[0] = a; idx[1] = b; idx[2] = c; idx[3] = d;
}

// collision_convex.py plane_convex for the flat floor (plane z = 0, normal +z) vs foot hull k
static void plane_convex(const OduckModel& m, const Scratch& s, int k, int slot0, Scratch& out) {
  int b = m.foot_body[k], nvt = m.foot_nvert;
  const real* R = s.xmat[b];
  // plane in the hull (body) frame
  real ppos[3], mp[3] = {-s.xpos[b][0], -s.xpos[b][1], -s.xpos[b][2]}, wn[3] = {0, 0, 1}, n[3];
  matT_vec(R, mp, ppos);
  matT_vec(R, wn, n);
  real vert[ODUCK_MAX_VERT][3], support[ODUCK_MAX_VERT], smax = -std::numeric_limits<real>::infinity();
  bool mask[ODUCK_MAX_VERT];
  for (int i = 0; i < nvt; i++) {
    for (int c = 0; c < 3; c++) vert[i][c] = (real)m.foot_vert[k][i][c];
    real d[3] = {ppos[0] - vert[i][0], ppos[1] - vert[i][1], ppos[2] - vert[i][2]};
    support[i] = dot3(d, n);
    smax = std::max(smax, support[i]);
  }
  real thr = std::max((real)0, smax - (real)1e-3);
  for (int i = 0; i < nvt; i++) mask[i] = support[i] > thr;
  int idx[4];
  manifold_points(nvt, vert, mask, n, idx);
  real frame[9];
  make_frame(wn, frame);
  for (int c = 0; c < 4; c++) {
    int sl = slot0 + c;
    // unique = first occurrence of this vertex index in idx
    bool unique = true;
    for (int p = 0; p < c; p++) unique &= idx[p] != idx[c];
    real dist = unique ? -support[idx[c]] : (real)1;
    real pw[3];
    mat_vec(R, vert[idx[c]], pw);
    for (int i = 0; i < 3; i++) out.con_pos[sl][i] = s.xpos[b][i] + pw[i] - (real)0.5 * dist * wn[i];
    out.con_dist[sl] = dist;
    for (int i = 0; i < 9; i++) out.con_frame[sl][i] = frame[i];
    out.con_b1[sl] = 0;  // floor body is static: zero Jacobian, zero invweight
    out.con_b2[sl] = b;
    out.con_mu[sl] = (real)m.floor_friction;
  }
}


// Convex-convex narrow phase (MJX collision_convex.py _convex_convex scheme) between two convex polytopes given by their
// world-frame vertices, polygon faces (outward normals, CCW vertex loops) and edges (vertex pair + the two adjacent faces).
// Used for the two feet (the mesh-mesh pair MJX instantiates: both collision meshes have contype = conaffinity = 1,
// SURVEY.md 2.1).  PARITY UNPINNED like the rest of the
// physics: MJX's source is not available, so this is a restatement of the documented scheme -- separating-axis test over
// the face normals of both hulls and the cross products of edge pairs, then a clipped-polygon manifold of at most 4 points:
//   * face axes: signed distance of the other hull's deepest vertex to every polygon face,
//   * edge axes: only edge pairs that span a face of the Minkowski difference (Gauss-map arc test), distance between the
//     two edge lines along their common normal,
//   * face contact: the incident face (most anti-parallel) is Sutherland-Hodgman clipped against the side planes of the
//     reference face; dist = signed distance to the reference plane, pos = midway; >4 points -> _manifold_points,
//   * edge contact (chosen only if it separates 0.1 mm better than the best face): closest points of the two edges.
// Normal points from hull A (geom1) to hull B (geom2).
#define HULL_MAXV ODUCK_MAX_VERT
#define HULL_MAXP ODUCK_MAX_PLANE
#define HULL_MAXE ODUCK_MAX_EDGE
struct Hull {
  int nv, np, ne;
  real V[HULL_MAXV][3], N[HULL_MAXP][3], c[3];
  int pnv[HULL_MAXP], pv[HULL_MAXP][ODUCK_MAX_PVERT], ev[HULL_MAXE][2], ep[HULL_MAXE][2];
};
struct HullContacts { real dist[4], pos[4][3], frame[9]; };   // dist = 1: inactive slot

static bool hull_hull(const Hull& A, const Hull& B, HullContacts& out) {
  for (int c = 0; c < 4; c++) { out.dist[c] = 1; out.pos[c][0] = out.pos[c][1] = out.pos[c][2] = 0; }
  for (int i = 0; i < 9; i++) out.frame[i] = (i == 0 || i == 4 || i == 8) ? 1 : 0;
  const real ninf = -std::numeric_limits<real>::infinity();
  real face_sep = ninf; int face_hull = 0, face_idx = 0;
  for (int h = 0; h < 2; h++) {
    const Hull& R = h ? B : A; const Hull& O = h ? A : B;
    for (int q = 0; q < R.np; q++) {
      const real* n = R.N[q];
      real off = dot3(n, R.V[R.pv[q][0]]), mn = std::numeric_limits<real>::infinity();
      for (int j = 0; j < O.nv; j++) mn = std::min(mn, dot3(n, O.V[j]));
      if (mn - off > face_sep) { face_sep = mn - off; face_hull = h; face_idx = q; }
    }
  }
  if (face_sep > 0) return false;
  real edge_sep = ninf, edge_n[3] = {0, 0, 1}; int eA_best = -1, eB_best = -1;
  for (int ea = 0; ea < A.ne; ea++) {
    const real *a = A.N[A.ep[ea][0]], *b = A.N[A.ep[ea][1]];
    const real *pa0 = A.V[A.ev[ea][0]], *pa1 = A.V[A.ev[ea][1]];
    real dA[3] = {pa1[0] - pa0[0], pa1[1] - pa0[1], pa1[2] - pa0[2]}, bxa[3];
    cross3(b, a, bxa);
    for (int eb = 0; eb < B.ne; eb++) {
      real c[3], d[3], dxc[3];
      for (int i = 0; i < 3; i++) { c[i] = -B.N[B.ep[eb][0]][i]; d[i] = -B.N[B.ep[eb][1]][i]; }
      cross3(d, c, dxc);
      real cba = dot3(c, bxa), dba = dot3(d, bxa), adc = dot3(a, dxc), bdc = dot3(b, dxc);
      if (!(cba * dba < 0 && adc * bdc < 0 && cba * bdc > 0)) continue;            // the two arcs do not cross on the Gauss map
      const real *pb0 = B.V[B.ev[eb][0]], *pb1 = B.V[B.ev[eb][1]];
      real dB[3] = {pb1[0] - pb0[0], pb1[1] - pb0[1], pb1[2] - pb0[2]}, n[3];
      cross3(dA, dB, n);
      real len2 = dot3(n, n);
      if (len2 < (real)1e-10 * dot3(dA, dA) * dot3(dB, dB)) continue;                  // parallel edges
      real inv = 1 / std::sqrt(len2);
      for (int i = 0; i < 3; i++) n[i] *= inv;
      real rel[3] = {pa0[0] - A.c[0], pa0[1] - A.c[1], pa0[2] - A.c[2]};
      if (dot3(n, rel) < 0) for (int i = 0; i < 3; i++) n[i] = -n[i];                 // away from hull A
      real w[3] = {pb0[0] - pa0[0], pb0[1] - pa0[1], pb0[2] - pa0[2]};
      real sep = dot3(n, w);
      if (sep > edge_sep) { edge_sep = sep; eA_best = ea; eB_best = eb; for (int i = 0; i < 3; i++) edge_n[i] = n[i]; }
    }
  }
  if (edge_sep > 0) return false;
  if (eA_best >= 0 && edge_sep > face_sep + (real)1e-4) {
    // closest points of the two supporting edges
    const real *p1 = A.V[A.ev[eA_best][0]], *q1 = A.V[A.ev[eA_best][1]], *p2 = B.V[B.ev[eB_best][0]], *q2 = B.V[B.ev[eB_best][1]];
    real d1[3], d2[3], r[3];
    for (int i = 0; i < 3; i++) { d1[i] = q1[i] - p1[i]; d2[i] = q2[i] - p2[i]; r[i] = p1[i] - p2[i]; }
    real a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), c = dot3(d1, r), b = dot3(d1, d2), den = a * e - b * b;
    real sA = den > (real)1e-18 ? std::min(std::max((b * f - c * e) / den, (real)0), (real)1) : 0;
    real tB = std::min(std::max((b * sA + f) / e, (real)0), (real)1);
    sA = std::min(std::max((b * tB - c) / a, (real)0), (real)1);
    out.dist[0] = edge_sep;
    for (int i = 0; i < 3; i++) out.pos[0][i] = (real)0.5 * (p1[i] + sA * d1[i] + p2[i] + tB * d2[i]);
    make_frame(edge_n, out.frame);
    return true;
  }
  // face contact: reference face on hull `face_hull`, incident face = most anti-parallel face of the other hull
  const Hull& Rh = face_hull ? B : A; const Hull& Ih = face_hull ? A : B;
  const real* nref = Rh.N[face_idx];
  int inc = 0; real best = std::numeric_limits<real>::infinity();
  for (int q = 0; q < Ih.np; q++) { real v = dot3(Ih.N[q], nref); if (v < best) { best = v; inc = q; } }
  real poly[2][2 * ODUCK_MAX_PVERT + 2][3];
  int cur = 0, cnt = Ih.pnv[inc];
  for (int k = 0; k < cnt; k++) for (int i = 0; i < 3; i++) poly[0][k][i] = Ih.V[Ih.pv[inc][k]][i];
  const int nr = Rh.pnv[face_idx];
  for (int k = 0; k < nr && cnt > 0; k++) {
    const real *r0 = Rh.V[Rh.pv[face_idx][k]], *r1 = Rh.V[Rh.pv[face_idx][(k + 1) % nr]];
    real e[3] = {r1[0] - r0[0], r1[1] - r0[1], r1[2] - r0[2]}, sd[3];
    cross3(e, nref, sd);                                                              // outward side-plane normal (loop is CCW from outside)
    int no = 0;
    for (int v = 0; v < cnt; v++) {
      const real *x0 = poly[cur][v], *x1 = poly[cur][(v + 1) % cnt];
      real w0[3] = {x0[0] - r0[0], x0[1] - r0[1], x0[2] - r0[2]}, w1[3] = {x1[0] - r0[0], x1[1] - r0[1], x1[2] - r0[2]};
      real d0 = dot3(sd, w0), d1 = dot3(sd, w1);
      if (d0 <= 0) { for (int i = 0; i < 3; i++) poly[1 - cur][no][i] = x0[i]; no++; }
      if ((d0 <= 0) != (d1 <= 0)) { real t = d0 / (d0 - d1); for (int i = 0; i < 3; i++) poly[1 - cur][no][i] = x0[i] + t * (x1[i] - x0[i]); no++; }
    }
    cur = 1 - cur; cnt = no;
  }
  if (cnt == 0) return false;
  real dist[2 * ODUCK_MAX_PVERT + 2];
  bool mask[2 * ODUCK_MAX_PVERT + 2];
  const real* r0 = Rh.V[Rh.pv[face_idx][0]];
  for (int v = 0; v < cnt; v++) { real w[3] = {poly[cur][v][0] - r0[0], poly[cur][v][1] - r0[1], poly[cur][v][2] - r0[2]}; dist[v] = dot3(nref, w); mask[v] = dist[v] < 0; }
  int idx[4] = {0, 1, 2, 3};
  if (cnt > 4) manifold_points(cnt, poly[cur], mask, nref, idx);
  real n12[3] = {face_hull ? -nref[0] : nref[0], face_hull ? -nref[1] : nref[1], face_hull ? -nref[2] : nref[2]};   // geom1 -> geom2
  make_frame(n12, out.frame);
  for (int c = 0; c < 4; c++) {
    int v = idx[c];
    bool ok = v < cnt;
    for (int p2 = 0; p2 < c && ok; p2++) ok = idx[p2] != v;                           // duplicates (cnt > 4 selection) are inactive
    if (!ok) continue;
    out.dist[c] = dist[v];
    for (int i = 0; i < 3; i++) out.pos[c][i] = poly[cur][v][i] - (real)0.5 * dist[v] * nref[i];
  }
  return true;
}

// world-frame polytope of foot k
static void foot_hull(const OduckModel& m, const Scratch& s, int k, Hull& H) {
  const int b = m.foot_body[k];
  auto xform = [&](const double* v, real* o) { real t[3] = {(real)v[0], (real)v[1], (real)v[2]}; mat_vec(s.xmat[b], t, o); for (int i = 0; i < 3; i++) o[i] += s.xpos[b][i]; };
  H.nv = m.foot_nvert; H.np = m.foot_nplane; H.ne = m.foot_nedge;
  xform(m.foot_center[k], H.c);
  for (int i = 0; i < H.nv; i++) xform(m.foot_vert[k][i], H.V[i]);
  for (int q = 0; q < H.np; q++) {
    real a[3] = {(real)m.foot_plane_normal[k][q][0], (real)m.foot_plane_normal[k][q][1], (real)m.foot_plane_normal[k][q][2]};
    mat_vec(s.xmat[b], a, H.N[q]);
    H.pnv[q] = m.foot_plane_nvert[q];
    for (int j = 0; j < H.pnv[q]; j++) H.pv[q][j] = m.foot_plane_vert[q][j];
  }
  for (int e = 0; e < H.ne; e++) for (int j = 0; j < 2; j++) { H.ev[e][j] = m.foot_edge_vert[e][j]; H.ep[e][j] = m.foot_edge_plane[e][j]; }
}

// the two feet against each other: contact slots 8..11
static void convex_convex(const OduckModel& m, const Scratch& s, Scratch& out) {
  const int bA = m.foot_body[0], bB = m.foot_body[1];
  for (int c = 8; c < 12; c++) { out.con_dist[c] = 1; out.con_b1[c] = bA; out.con_b2[c] = bB; out.con_mu[c] = (real)m.foot_friction; }
  static thread_local Hull A, B;
  foot_hull(m, s, 0, A);
  foot_hull(m, s, 1, B);
  real dc[3] = {B.c[0] - A.c[0], B.c[1] - A.c[1], B.c[2] - A.c[2]};
  if (std::sqrt(dot3(dc, dc)) > 2 * (real)m.foot_radius) return;                      // bounding spheres apart: no contact
  HullContacts hc;
  if (!hull_hull(A, B, hc)) return;
  for (int c = 0; c < 4; c++) {
    out.con_dist[8 + c] = hc.dist[c];
    for (int i = 0; i < 3; i++) out.con_pos[8 + c][i] = hc.pos[c][i];
    for (int i = 0; i < 9; i++) out.con_frame[8 + c][i] = hc.frame[i];
  }
}

// Height-field floor vs foot k.  MuJoCo / MJX treat a height field as a union of triangular prisms: every grid cell is split
// along the diagonal (c, r) - (c + 1, r + 1) into two triangles, each extruded down to z = -size[3]
// (engine_collision_convex.c mjc_ConvexHField; collision_convex.py hfield_convex **[upstream-memory]**).  PARITY UNPINNED, and
// here a deliberate simplification as well: a general convex-convex test of a prism against the foot also offers
// the prism's vertical side walls (and the foot's side faces) as contact axes, which yields horizontal contact normals
// whenever a foot overlaps a neighbouring prism by less than its penetration depth -- on perfectly flat ground too.  The side
// walls are interior to the terrain, so this restatement uses the terrain surface only:
//   * for every triangle under the foot's bounding sphere, every foot face that looks down onto it (face normal . triangle
//     normal < 0) is Sutherland-Hodgman clipped against the three vertical planes through the triangle's edges (all of
//     them rather than one "incident" face: the sole is several nearly coplanar polygons, and a single pick would flip);
//   * each clipped point below the triangle plane is a candidate: dist = signed distance to the plane, pos = midway, normal
//     = the triangle normal (terrain = geom1, so it points from the terrain into the foot);
//   * 4 of the candidates within 1 mm of the deepest one (plane_convex's support threshold) are kept by _manifold_points
//     with the mean candidate normal, exactly like the plane and mesh colliders; repeated picks are inactive.
// On an all-zero field this reduces to the plane collider's candidate set (sole vertices below z = 0) plus points on cell
// borders.  The hfield sits at the world origin (checked by the MJCF compiler).
static int g_hf_max_candidates = 0;   // largest candidate count seen by hfield_convex (tests: how far below MAXC the scenes stay)
static void hfield_convex(const OduckHandle& h, const Scratch& s, int k, int slot0, Scratch& out) {
  const OduckModel& m = h.m;
  const int b = m.foot_body[k], nrow = m.hfield_nrow, ncol = m.hfield_ncol;
  const real sx = (real)m.hfield_size[0], sy = (real)m.hfield_size[1], sz = (real)m.hfield_size[2];
  static thread_local Hull F;
  foot_hull(m, s, k, F);
  real frame0[9];
  real up[3] = {0, 0, 1};
  make_frame(up, frame0);
  for (int c = 0; c < 4; c++) {
    const int sl = slot0 + c;
    out.con_dist[sl] = 1; out.con_b1[sl] = 0; out.con_b2[sl] = b; out.con_mu[sl] = (real)m.floor_friction;
    for (int i = 0; i < 3; i++) out.con_pos[sl][i] = 0;
    for (int i = 0; i < 9; i++) out.con_frame[sl][i] = frame0[i];
  }
  const real dx = 2 * sx / (ncol - 1), dy = 2 * sy / (nrow - 1), rb = (real)m.foot_radius;
  int cmin = (int)std::floor((double)((F.c[0] - rb + sx) / dx)), cmax = (int)std::floor((double)((F.c[0] + rb + sx) / dx));
  int rmin = (int)std::floor((double)((F.c[1] - rb + sy) / dy)), rmax = (int)std::floor((double)((F.c[1] + rb + sy) / dy));
  cmin = std::max(cmin, 0); rmin = std::max(rmin, 0); cmax = std::min(cmax, ncol - 2); rmax = std::min(rmax, nrow - 2);
  // MAXC: the candidate list is bounded like the CUDA library's (HF_CAP in csrc/oduck_hfcollide.cuh): later candidates are
  // dropped but still count in the mean normal.  A resting foot has a few dozen; the cap is only reached by a foot sunk deep
  // below the terrain (a fallen robot just before its episode ends).
  constexpr int MAXC = 512, MAXP = ODUCK_MAX_PVERT + 4;
  static thread_local real cd[MAXC], cp[MAXC][3], cn[MAXC][3];
  static thread_local bool cm[MAXC];
  int nc = 0, nc_all = 0;
  real nmean[3] = {0, 0, 0};
  auto H = [&](int r, int c) { return (real)h.hfield[(size_t)r * ncol + c] * sz; };
  for (int r = rmin; r <= rmax; r++)
    for (int c = cmin; c <= cmax; c++) {
      const real x0 = c * dx - sx, x1 = (c + 1) * dx - sx, y0 = r * dy - sy, y1 = (r + 1) * dy - sy;
      const real t[2][3][3] = {{{x0, y1, H(r + 1, c)}, {x0, y0, H(r, c)}, {x1, y1, H(r + 1, c + 1)}},       // counter-clockwise seen from above
                               {{x0, y0, H(r, c)}, {x1, y0, H(r, c + 1)}, {x1, y1, H(r + 1, c + 1)}}};
      for (int i = 0; i < 2; i++) {
        const real (*T)[3] = t[i];
        real top = std::max(T[0][2], std::max(T[1][2], T[2][2]));
        if (F.c[2] - rb > top) continue;                             // the foot's bounding sphere is above this triangle
        real e1[3], e2[3], n[3];
        for (int a = 0; a < 3; a++) { e1[a] = T[1][a] - T[0][a]; e2[a] = T[2][a] - T[0][a]; }
        cross3(e1, e2, n);
        real ln = std::sqrt(dot3(n, n));
        for (int a = 0; a < 3; a++) n[a] /= ln;
        for (int q = 0; q < F.np; q++) {
          if (!(dot3(F.N[q], n) < 0)) continue;                      // only the faces that look down onto the triangle
          real poly[2][MAXP][3];
          int cur = 0, cnt = F.pnv[q];
          for (int v = 0; v < cnt; v++) for (int a = 0; a < 3; a++) poly[0][v][a] = F.V[F.pv[q][v]][a];
          for (int e = 0; e < 3 && cnt > 0; e++) {
            const real *r0 = T[e], *r1 = T[(e + 1) % 3];
            const real sd[2] = {r1[1] - r0[1], -(r1[0] - r0[0])};     // outward normal of the vertical plane through the edge
            int no = 0;
            for (int v = 0; v < cnt; v++) {
              const real *p0 = poly[cur][v], *p1 = poly[cur][(v + 1) % cnt];
              real d0 = sd[0] * (p0[0] - r0[0]) + sd[1] * (p0[1] - r0[1]), d1 = sd[0] * (p1[0] - r0[0]) + sd[1] * (p1[1] - r0[1]);
              if (d0 <= 0) { for (int a = 0; a < 3; a++) poly[1 - cur][no][a] = p0[a]; no++; }
              if ((d0 <= 0) != (d1 <= 0)) { real tt = d0 / (d0 - d1); for (int a = 0; a < 3; a++) poly[1 - cur][no][a] = p0[a] + tt * (p1[a] - p0[a]); no++; }
            }
            cur = 1 - cur; cnt = no;
          }
          for (int v = 0; v < cnt; v++) {
            real w[3] = {poly[cur][v][0] - T[0][0], poly[cur][v][1] - T[0][1], poly[cur][v][2] - T[0][2]};
            real dist = dot3(n, w);
            if (!(dist < 0)) continue;
            for (int a = 0; a < 3; a++) nmean[a] += n[a];
            nc_all++;
            if (nc >= MAXC) continue;
            cd[nc] = dist; cm[nc] = true;
            for (int a = 0; a < 3; a++) { cp[nc][a] = poly[cur][v][a] - (real)0.5 * dist * n[a]; cn[nc][a] = n[a]; }
            nc++;
          }
        }
      }
    }
  if (nc == 0) return;
  if (nc_all > g_hf_max_candidates) g_hf_max_candidates = nc_all;   // diagnostic (racy max across worker threads: good enough)
  real nn = std::sqrt(dot3(nmean, nmean));
  for (int a = 0; a < 3; a++) nmean[a] /= nn;                        // every triangle normal has n_z > 0
  real deepest = 0;
  for (int v = 0; v < nc; v++) deepest = std::min(deepest, cd[v]);
  for (int v = 0; v < nc; v++) cm[v] = cd[v] < std::min((real)0, deepest + (real)1e-3);   // plane_convex's rule: within 1 mm of the deepest
  // Twins: a clipped point on an edge shared by two terrain triangles (or on a cell border) is emitted once per triangle, the
  // copies ~1e-7 m apart but with different triangle normals.  Which twin a later arg-max picks would be decided by rounding,
  // and the pick matters through the normal.  An in-threshold candidate whose clipped point lies within HF_TWIN = 1e-5 m
  // (max-norm; hull vertices are millimetres apart) of an EARLIER in-threshold candidate is therefore masked out before the
  // manifold selection -- the first copy in list order (triangle, face, polygon vertex) stands for all of them.
  {
#ifdef ODUCK_HF_NO_TWIN
    const real HF_TWIN = (real)-1;
#else
    const real HF_TWIN = (real)1e-5;
#endif
    static thread_local bool twin[MAXC];
    for (int j = 0; j < nc; j++) {
      twin[j] = false;
      if (!cm[j]) continue;
      for (int i = 0; i < j && !twin[j]; i++) {
        if (!cm[i]) continue;
        real dmax = 0;
        for (int a = 0; a < 3; a++) {
          const real pj = cp[j][a] + (real)0.5 * cd[j] * cn[j][a], pi = cp[i][a] + (real)0.5 * cd[i] * cn[i][a];   // the clipped points themselves
          dmax = std::max(dmax, std::fabs(pj - pi));
        }
        if (dmax < HF_TWIN) twin[j] = true;
      }
    }
    for (int j = 0; j < nc; j++) if (twin[j]) cm[j] = false;
  }
  int idx[4];
  manifold_points(nc, cp, cm, nmean, idx);
  for (int c = 0; c < 4; c++) {
    const int sl = slot0 + c, v = idx[c];
    bool unique = true;
    for (int p2 = 0; p2 < c; p2++) unique &= idx[p2] != v;
    if (!unique) continue;
    out.con_dist[sl] = cd[v];
    for (int i = 0; i < 3; i++) out.con_pos[sl][i] = cp[v][i];
    make_frame(cn[v], out.con_frame[sl]);
  }
}

// constraint.py _kbi / _efc_row
static void kbi(const OduckModel& m, real pos, real* k, real* b, real* imp) {
  real timeconst = std::max((real)m.solref[0], 2 * (real)m.timestep), dampratio = (real)m.solref[1];
  real dmin = std::min(std::max((real)m.solimp[0], kMinImp), kMaxImp), dmax = std::min(std::max((real)m.solimp[1], kMinImp), kMaxImp);
  real width = std::max(kMinVal, (real)m.solimp[2]);
  real mid = std::min(std::max((real)m.solimp[3], kMinImp), kMaxImp), power = std::max((real)1, (real)m.solimp[4]);
  *k = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  *b = 2 / (dmax * timeconst);
  if (m.solref[0] <= 0) *k = -(real)m.solref[0] / (dmax * dmax);
  if (m.solref[1] <= 0) *b = -(real)m.solref[1] / dmax;
  real x = std::fabs(pos) / width;
  real ia = (1 / std::pow(mid, power - 1)) * std::pow(x, power);
  real ib = 1 - (1 / std::pow(1 - mid, power - 1)) * std::pow(1 - x, power);
  real y = x < mid ? ia : ib;
  real im = dmin + y * (dmax - dmin);
  im = std::min(std::max(im, dmin), dmax);
  if (x > 1) im = dmax;
  *imp = im;
}

static void make_constraint(const OduckHandle& h, const EnvState& e, Scratch& s) {
  const OduckModel& m = h.m;
  int r = 0;
  auto zero_row = [&](int row) { for (int j = 0; j < m.nv; j++) s.J[row][j] = 0; s.D[row] = 0; s.aref[row] = 0; s.floss[row] = 0; };
  // dof friction loss (constraint.py _instantiate_friction)
  for (int i = 0; i < h.nefc_fr; i++, r++) {
    int d = h.fr_dof[i];
    zero_row(r);
    s.rtype[r] = 0;
    real k, b, imp;
    kbi(m, 0, &k, &b, &imp);
    s.J[r][d] = 1;
    real rr = std::max((real)m.dof_invweight0[d] * (1 - imp) / imp, kMinVal);
    s.D[r] = 1 / rr;
    s.aref[r] = -b * e.qvel[d];
    s.floss[r] = e.dof_frictionloss[d];
  }
  // hinge limits (_instantiate_limit_slide_hinge)
  for (int i = 0; i < h.nefc_lim; i++, r++) {
    int j = h.lim_jnt[i], d = m.jnt_dofadr[j];
    zero_row(r);
    s.rtype[r] = 1;
    real q = e.qpos[m.jnt_qposadr[j]];
    real dmin = q - (real)m.jnt_range[j][0], dmax = (real)m.jnt_range[j][1] - q;
    real pos = std::min(dmin, dmax);
    if (!(pos < 0)) continue;
    real sign = dmin < dmax ? (real)1 : (real)-1;
    real k, b, imp;
    kbi(m, pos, &k, &b, &imp);
    s.J[r][d] = sign;
    real rr = std::max((real)m.dof_invweight0[d] * (1 - imp) / imp, kMinVal);
    s.D[r] = 1 / rr;
    s.aref[r] = -b * (sign * e.qvel[d]) - k * imp * pos;
  }
  // pyramidal frictional contacts (_instantiate_contact, condim 3)
  int ncon = ODUCK_CON_PER_PAIR * (2 + (m.enable_foot_foot ? 1 : 0));
  for (int c = 0; c < ncon; c++) {
    real dist = s.con_dist[c];
    bool active = dist < 0;
    real diff[3][NV];  // (jacp(body2) - jacp(body1)) rotated into the contact frame
    for (int j = 0; j < m.nv; j++) diff[0][j] = diff[1][j] = diff[2][j] = 0;
    if (active) {
      real off[3] = {s.con_pos[c][0] - s.com[0], s.con_pos[c][1] - s.com[1], s.con_pos[c][2] - s.com[2]};
      for (int side = 0; side < 2; side++) {
        int b = side ? s.con_b2[c] : s.con_b1[c];
        real sg = side ? (real)1 : (real)-1;
        if (b <= 0) continue;
        for (int d = m.body_dofadr[b] + m.body_dofnum[b] - 1; d >= 0; d = m.dof_parentid[d]) {
          real jp_[3], t[3];
          cross3(s.cdof[d], off, t);
          for (int i = 0; i < 3; i++) jp_[i] = s.cdof[d][3 + i] + t[i];
          for (int a = 0; a < 3; a++) diff[a][d] += sg * dot3(s.con_frame[c] + 3 * a, jp_);
        }
      }
    }
    real mu = s.con_mu[c];
    real t = (real)m.body_invweight0[s.con_b1[c]][0] + (real)m.body_invweight0[s.con_b2[c]][0];
    real invw = (t + mu * mu * t) * 2 * mu * mu / (real)m.impratio;
    real k, b, imp;
    kbi(m, dist, &k, &b, &imp);
    for (int e4 = 0; e4 < 4; e4++, r++) {
      zero_row(r);
      s.rtype[r] = 1;
      if (!active) continue;
      int dim = 1 + e4 / 2;
      real sg = (e4 & 1) ? (real)-1 : (real)1;
      real vel = 0;
      for (int j = 0; j < m.nv; j++) { s.J[r][j] = diff[0][j] + sg * mu * diff[dim][j]; vel += s.J[r][j] * e.qvel[j]; }
      real rr = std::max(invw * (1 - imp) / imp, kMinVal);
      s.D[r] = 1 / rr;
      s.aref[r] = -b * vel - k * imp * dist;
    }
  }
  s.nefc = r;
}

struct DebugCapture { bool on; real search[NV], grad[NV], H[NV][NV], costw, costs, alpha; int ls_it; };
static thread_local DebugCapture g_dbg = {false};

// ------------------------------------------------------------------------------------ solver.py (Newton)
struct Ctx {
  real qacc[NV], Ma[NV], Jaref[NEFC], efc_force[NEFC], qfrc_constraint[NV], grad[NV], Mgrad[NV], search[NV];
  bool active[NEFC];
  real gauss, cost, prev_cost;
};

static void update_constraint(const OduckModel& m, const Scratch& s, Ctx& c) {
  real cost = 0;
  for (int r = 0; r < s.nefc; r++) {
    real x = c.Jaref[r];
    if (s.rtype[r] == 0) {
      real f = s.floss[r], rf = (1 / (s.D[r] + (s.D[r] == 0 ? kMinVal : 0))) * f;
      if (x <= -rf) { c.efc_force[r] = f; c.active[r] = false; cost += f * (-(real)0.5 * rf - x); }
      else if (x >= rf) { c.efc_force[r] = -f; c.active[r] = false; cost += f * (-(real)0.5 * rf + x); }
      else { c.efc_force[r] = -s.D[r] * x; c.active[r] = true; cost += (real)0.5 * s.D[r] * x * x; }
    } else {
      c.active[r] = x < 0;
      c.efc_force[r] = c.active[r] ? -s.D[r] * x : 0;
      if (c.active[r]) cost += (real)0.5 * s.D[r] * x * x;
    }
  }
  for (int j = 0; j < m.nv; j++) c.qfrc_constraint[j] = 0;
  for (int r = 0; r < s.nefc; r++) {
    if (c.efc_force[r] == 0) continue;
    for (int j = 0; j < m.nv; j++) c.qfrc_constraint[j] += s.J[r][j] * c.efc_force[r];
  }
  real g = 0;
  for (int j = 0; j < m.nv; j++) g += (c.Ma[j] - s.qfrc_smooth[j]) * (c.qacc[j] - s.qacc_smooth[j]);
  c.gauss = (real)0.5 * g;
  c.prev_cost = c.cost;
  c.cost = cost + c.gauss;
}

static void update_gradient(const OduckModel& m, Scratch& s, Ctx& c) {
  static thread_local real H[NV][NV], LH[NV][NV];
  for (int j = 0; j < m.nv; j++) c.grad[j] = c.Ma[j] - s.qfrc_smooth[j] - c.qfrc_constraint[j];
  for (int i = 0; i < m.nv; i++)
    for (int j = 0; j < m.nv; j++) H[i][j] = s.M[i][j];
  for (int r = 0; r < s.nefc; r++) {
    if (!c.active[r] || s.D[r] == 0) continue;
    for (int i = 0; i < m.nv; i++) {
      if (s.J[r][i] == 0) continue;
      real w = s.J[r][i] * s.D[r];
      for (int j = 0; j < m.nv; j++) H[i][j] += w * s.J[r][j];
    }
  }
  if (g_dbg.on) for (int i = 0; i < m.nv; i++) for (int j = 0; j < m.nv; j++) g_dbg.H[i][j] = H[i][j];
  if (!cholesky(m.nv, H, LH)) {
    for (int j = 0; j < m.nv; j++) c.Mgrad[j] = std::numeric_limits<real>::quiet_NaN();
    return;
  }
  chol_solve(m.nv, LH, c.grad, c.Mgrad);
}

static void ctx_create(const OduckModel& m, Scratch& s, const real* qacc, Ctx& c, bool grad) {
  for (int j = 0; j < m.nv; j++) {
    c.qacc[j] = qacc[j];
    real v = 0;
    for (int k = 0; k < m.nv; k++) v += s.M[j][k] * qacc[k];
    c.Ma[j] = v;
    c.search[j] = 0;
  }
  for (int r = 0; r < s.nefc; r++) {
    if (s.D[r] == 0) { c.Jaref[r] = 0; continue; }   // empty (inactive) row: J = 0, aref = 0
    real v = 0;
    for (int j = 0; j < m.nv; j++) v += s.J[r][j] * qacc[j];
    c.Jaref[r] = v - s.aref[r];
  }
  c.cost = std::numeric_limits<real>::infinity();
  c.prev_cost = 0;
  update_constraint(m, s, c);
  if (grad) {
    update_gradient(m, s, c);
    for (int j = 0; j < m.nv; j++) c.search[j] = -c.Mgrad[j];
  }
}

struct LSPoint { real alpha, cost, d0, d1; };

static void linesearch(const OduckModel& m, const Scratch& s, Ctx& c) {
  real mv[NV], jv[NEFC], snorm = 0;
  for (int j = 0; j < m.nv; j++) snorm += c.search[j] * c.search[j];
  snorm = std::sqrt(snorm);
  real gtol = (real)m.tolerance * (real)m.ls_tolerance * snorm * (real)m.meaninertia * std::max(1, m.nv);
  for (int j = 0; j < m.nv; j++) {
    real v = 0;
    for (int k = 0; k < m.nv; k++) v += s.M[j][k] * c.search[k];
    mv[j] = v;
  }
  for (int r = 0; r < s.nefc; r++) {
    real v = 0;
    if (s.D[r] != 0) for (int j = 0; j < m.nv; j++) v += s.J[r][j] * c.search[j];
    jv[r] = v;
  }
  real qg[3] = {c.gauss, 0, 0};
  for (int j = 0; j < m.nv; j++) { qg[1] += c.search[j] * c.Ma[j] - c.search[j] * s.qfrc_smooth[j]; qg[2] += (real)0.5 * c.search[j] * mv[j]; }
  auto point = [&](real alpha) {
    real q0 = qg[0], q1 = qg[1], q2 = qg[2];
    for (int r = 0; r < s.nefc; r++) {
      real x = c.Jaref[r] + alpha * jv[r], D = s.D[r];
      if (s.rtype[r] == 0) {
        real f = s.floss[r], rf = (1 / (D + (D == 0 ? kMinVal : 0))) * f;
        if (x <= -rf) { q0 += f * (-(real)0.5 * rf - c.Jaref[r]); q1 += -f * jv[r]; }
        else if (x >= rf) { q0 += f * (-(real)0.5 * rf + c.Jaref[r]); q1 += f * jv[r]; }
        else { q0 += (real)0.5 * D * c.Jaref[r] * c.Jaref[r]; q1 += D * jv[r] * c.Jaref[r]; q2 += (real)0.5 * D * jv[r] * jv[r]; }
      } else if (x < 0) {
        q0 += (real)0.5 * D * c.Jaref[r] * c.Jaref[r]; q1 += D * jv[r] * c.Jaref[r]; q2 += (real)0.5 * D * jv[r] * jv[r];
      }
    }
    LSPoint p;
    p.alpha = alpha;
    p.cost = alpha * alpha * q2 + alpha * q1 + q0;
    p.d0 = 2 * alpha * q2 + q1;
    p.d1 = 2 * q2 + (q2 == 0 ? kMinVal : 0);
    return p;
  };
  LSPoint p0 = point(0);
  LSPoint lo = point(p0.alpha - p0.d0 / p0.d1), hi;
  if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
  bool swap = true;
  int it = 0;
  const bool trace = getenv("ODUCK_LS_TRACE") != nullptr;
  if (trace) printf("LS p0: a=%g cost=%.9g d0=%g d1=%g | gtol=%g\n", (double)p0.alpha, (double)p0.cost, (double)p0.d0, (double)p0.d1, (double)gtol);
  while (true) {
    if (trace) printf("  it%d lo(a=%.6g c=%.9g d0=%g) hi(a=%.6g c=%.9g d0=%g) swap=%d\n", it, (double)lo.alpha, (double)lo.cost, (double)lo.d0, (double)hi.alpha, (double)hi.cost, (double)hi.d0, (int)swap);
    bool done = it >= m.ls_iterations;
    done |= !swap;
    done |= (lo.d0 < 0) && (lo.d0 > -gtol);
    done |= (hi.d0 > 0) && (hi.d0 < gtol);
    if (done) break;
    LSPoint lo_next = point(lo.alpha - lo.d0 / lo.d1);
    LSPoint hi_next = point(hi.alpha - hi.d0 / hi.d1);
    LSPoint mid = point((real)0.5 * (lo.alpha + hi.alpha));
    // Bracket update.  solver.py swaps an end "if 1) it is not correctly at a bracket boundary (e.g. lo.deriv_0 > 0), OR
    // 2) moving to next or mid narrows the bracket".  Restated so that the outcome does not depend on the rounding sign of a
    // candidate that lands exactly on the root: each candidate c (Newton step from lo, midpoint, Newton step from hi) becomes
    // the new lo if its slope is negative and (lo is on the wrong side or c is closer to the root), symmetrically for hi.
    swap = false;
    const LSPoint cand_lo[3] = {lo_next, mid, hi_next}, cand_hi[3] = {hi_next, mid, lo_next};
    for (int k = 0; k < 3; k++) {
      const LSPoint& c = cand_lo[k];
      if (c.d0 < 0 && (lo.d0 > 0 || c.d0 > lo.d0)) { lo = c; swap = true; }
    }
    for (int k = 0; k < 3; k++) {
      const LSPoint& c = cand_hi[k];
      if (c.d0 >= 0 && (hi.d0 < 0 || c.d0 < hi.d0)) { hi = c; swap = true; }
    }
    it++;
  }
  bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  real alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (g_dbg.on) { g_dbg.alpha = improved ? alpha : 0; g_dbg.ls_it = it; }
  if (improved) {
    for (int j = 0; j < m.nv; j++) { c.qacc[j] += c.search[j] * alpha; c.Ma[j] += mv[j] * alpha; }
    for (int r = 0; r < s.nefc; r++) c.Jaref[r] += jv[r] * alpha;
  }
}

static void solve(const OduckModel& m, EnvState& e, Scratch& s) {
  static thread_local Ctx warm, smth, c;
  ctx_create(m, s, e.qacc_warm, warm, false);
  ctx_create(m, s, s.qacc_smooth, smth, false);
  const real* start = warm.cost < smth.cost ? e.qacc_warm : s.qacc_smooth;
  ctx_create(m, s, start, c, true);
  if (g_dbg.on) { g_dbg.costw = warm.cost; g_dbg.costs = smth.cost; for (int j = 0; j < m.nv; j++) { g_dbg.search[j] = c.search[j]; g_dbg.grad[j] = c.grad[j]; } }
  real scale = 1 / ((real)m.meaninertia * std::max(1, m.nv));
  for (int it = 0; it < m.iterations; it++) {
    if (m.iterations > 1) {
      real gn = 0;
      for (int j = 0; j < m.nv; j++) gn += c.grad[j] * c.grad[j];
      if ((c.prev_cost - c.cost) * scale < (real)m.tolerance || std::sqrt(gn) * scale < (real)m.tolerance) break;
    }
    linesearch(m, s, c);
    update_constraint(m, s, c);
    update_gradient(m, s, c);
    for (int j = 0; j < m.nv; j++) c.search[j] = -c.Mgrad[j];
  }
  for (int j = 0; j < m.nv; j++) { e.qacc[j] = c.qacc[j]; e.qacc_warm[j] = c.qacc[j]; }
  for (int r = 0; r < NEFC; r++) e.efc_force[r] = r < s.nefc ? c.efc_force[r] : 0;
}

// ------------------------------------------------------------------------------------ forward / step
static void sensors(const OduckModel& m, EnvState& e, Scratch& s) {
  real cacc[NB][6];
  rne(m, e, s, e.qacc, cacc, nullptr);
  auto site_vel = [&](int site, real* ang, real* lin) {  // world-frame velocity of the site point
    int b = m.site_bodyid[site];
    real diff[3] = {s.site_xpos[site][0] - s.com[0], s.site_xpos[site][1] - s.com[1], s.site_xpos[site][2] - s.com[2]}, t[3];
    cross3(diff, s.cvel[b], t);
    for (int i = 0; i < 3; i++) { ang[i] = s.cvel[b][i]; lin[i] = s.cvel[b][3 + i] - t[i]; }
  };
  int imu = m.imu_site, ib = m.site_bodyid[imu];
  const real* R = s.site_xmat[imu];
  real ang[3], lin[3], angl[3], linl[3];
  site_vel(imu, ang, lin);
  matT_vec(R, ang, angl);
  matT_vec(R, lin, linl);
  real diff[3] = {s.site_xpos[imu][0] - s.com[0], s.site_xpos[imu][1] - s.com[1], s.site_xpos[imu][2] - s.com[2]}, t[3], acc[3], accl[3], corr[3];
  cross3(diff, cacc[ib], t);
  for (int i = 0; i < 3; i++) acc[i] = cacc[ib][3 + i] - t[i];
  matT_vec(R, acc, accl);
  cross3(angl, linl, corr);
  real* sd = e.sensordata;
  for (int i = 0; i < 3; i++) {
    sd[0 + i] = angl[i];               // gyro
    sd[3 + i] = linl[i];               // local_linvel (velocimeter)
    sd[6 + i] = accl[i] + corr[i];     // accelerometer
    sd[9 + i] = R[3 * i + 2];          // upvector (framezaxis)
    sd[12 + i] = ang[i];               // global_angvel
  }
  for (int k = 0; k < 2; k++) {
    real a2[3], l2[3];
    site_vel(m.foot_site[k], a2, l2);
    for (int i = 0; i < 3; i++) { sd[15 + 3 * k + i] = l2[i]; e.site_xpos_feet[3 * k + i] = s.site_xpos[m.foot_site[k]][i]; }
  }
  for (int i = 21; i < 24; i++) sd[i] = 0;
  for (int i = 0; i < 9; i++) e.imu_xmat[i] = R[i];
}

static void forward(const OduckHandle& h, EnvState& e, Scratch& s) {
  const OduckModel& m = h.m;
  kinematics(m, e, s);
  com_pos(m, e, s);
  crb(m, e, s);
  bool pd = cholesky(m.nv, s.M, s.L);
  // collision
  int ncon = ODUCK_CON_PER_PAIR * (2 + (m.enable_foot_foot ? 1 : 0));
  for (int c = 0; c < NCON; c++) { s.con_dist[c] = 1; s.con_b1[c] = s.con_b2[c] = 0; s.con_mu[c] = 0; for (int i = 0; i < 3; i++) s.con_pos[c][i] = 0; for (int i = 0; i < 9; i++) s.con_frame[c][i] = (i % 4 == 0); }
  if (m.floor_is_hfield) { hfield_convex(h, s, 0, 0, s); hfield_convex(h, s, 1, 4, s); }
  else { plane_convex(m, s, 0, 0, s); plane_convex(m, s, 1, 4, s); }
  if (m.enable_foot_foot) convex_convex(m, s, s);
  (void)ncon;
  make_constraint(h, e, s);
  com_vel(m, e, s);
  real cacc[NB][6];
  rne(m, e, s, nullptr, cacc, s.qfrc_bias);
  for (int d = 0; d < m.nv; d++) { s.qfrc_passive[d] = -(real)m.dof_damping[d] * e.qvel[d]; s.qfrc_actuator[d] = 0; }
  for (int u = 0; u < m.nu; u++) {
    int j = m.act_jntid[u];
    real c = std::min(std::max(e.ctrl[u], (real)m.act_ctrlrange[u][0]), (real)m.act_ctrlrange[u][1]);
    real f = e.act_kp[u] * c - e.act_kp[u] * e.qpos[m.jnt_qposadr[j]] - (real)m.act_kv[u] * e.qvel[m.jnt_dofadr[j]];
    f = std::min(std::max(f, (real)m.act_forcerange[u][0]), (real)m.act_forcerange[u][1]);
    e.actuator_force[u] = f;
    s.qfrc_actuator[m.jnt_dofadr[j]] += f;
  }
  for (int d = 0; d < m.nv; d++) s.qfrc_smooth[d] = s.qfrc_passive[d] - s.qfrc_bias[d] + s.qfrc_actuator[d];
  if (pd) chol_solve(m.nv, s.L, s.qfrc_smooth, s.qacc_smooth);
  else for (int d = 0; d < m.nv; d++) s.qacc_smooth[d] = std::numeric_limits<real>::quiet_NaN();
  solve(m, e, s);
  sensors(m, e, s);
  for (int c = 0; c < NCON; c++) e.contact_dist[c] = s.con_dist[c];
}

static void euler(const OduckModel& m, EnvState& e) {
  real dt = (real)m.timestep;
  for (int d = 0; d < m.nv; d++) e.qvel[d] += dt * e.qacc[d];
  for (int j = 0; j < m.njnt; j++) {
    int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    if (m.jnt_type[j] == ODUCK_JNT_FREE) {
      for (int i = 0; i < 3; i++) e.qpos[qa + i] += dt * e.qvel[da + i];
      real w[3] = {e.qvel[da + 3], e.qvel[da + 4], e.qvel[da + 5]};
      real nrm = std::sqrt(dot3(w, w));
      real ax[3] = {0, 0, 0};
      if (nrm > 0) for (int i = 0; i < 3; i++) ax[i] = w[i] / nrm;
      real ang = dt * nrm;
      real qr[4] = {std::cos(ang / 2), std::sin(ang / 2) * ax[0], std::sin(ang / 2) * ax[1], std::sin(ang / 2) * ax[2]}, qn[4];
      quat_mul(e.qpos + qa + 3, qr, qn);
      quat_norm(qn);
      for (int i = 0; i < 4; i++) e.qpos[qa + 3 + i] = qn[i];
    } else {
      e.qpos[qa] += dt * e.qvel[da];
    }
  }
}

// ------------------------------------------------------------------------------------ env logic
static void poly_reference_motion(const OduckHandle& h, real dx, real dy, real dth, int i, real* out) {
  // poly_reference_motion.py:148-168
  const OduckEnvConfig& c = h.cfg;
  auto nearest = [](real v, const double* grid, int n, const double* range) {
    v = std::min(std::max(v, (real)range[0]), (real)range[1]);
    int best = 0;
    real bd = std::numeric_limits<real>::infinity();
    for (int k = 0; k < n; k++) { real d = std::fabs((real)grid[k] - v); if (d < bd) { bd = d; best = k; } }
    return best;
  };
  int ix = nearest(dx, c.dxs, c.ndx, c.dx_range), iy = nearest(dy, c.dys, c.ndy, c.dy_range), it = nearest(dth, c.dthetas, c.ndth, c.dtheta_range);
  real t = (real)(i % c.nb_steps_in_period) / (real)c.nb_steps_in_period;
  t = std::min(std::max(t, (real)0), (real)1);
  const double* base = h.poly.data() + (((size_t)ix * c.ndy + iy) * c.ndth + it) * ODUCK_REF_DIM * ODUCK_POLY_DEG;
  for (int d = 0; d < ODUCK_REF_DIM; d++) {
    real acc = 0;
    for (int k = 0; k < ODUCK_POLY_DEG; k++) acc = acc * t + (real)base[d * ODUCK_POLY_DEG + k];
    out[d] = acc;
  }
}

static void sample_command(const OduckHandle& h, Key rng, real* cmd) {  // joystick.py:671-725
  const OduckEnvConfig& c = h.cfg;
  Key k[8];
  for (int i = 0; i < 8; i++) k[i] = key_split(rng, i);
  const int which[7] = {0, 1, 2, 4, 5, 6, 7};
  for (int i = 0; i < 7; i++) cmd[i] = key_uniform(k[which[i]], 0, (real)c.cmd_range[i][0], (real)c.cmd_range[i][1]);
  if (c.task == ODUCK_TASK_STANDING) cmd[0] = cmd[1] = cmd[2] = 0;   // standing.py:648-655: no velocity command, same key usage
  bool zero = key_uniform(k[3], 0, 0, 1) < (real)0.1;
  if (zero) for (int i = 0; i < 7; i++) cmd[i] = 0;
}




static real nan_to_num(real x) {
  if (std::isnan(x)) return 0;
  if (std::isinf(x)) return x > 0 ? std::numeric_limits<float>::max() : -std::numeric_limits<float>::max();
  return x;
}

static void actuated(const OduckModel& m, const real* qpos, const real* qvel, real* q, real* qd) {
  for (int u = 0; u < m.nu; u++) { int j = m.act_jntid[u]; if (q) q[u] = qpos[m.jnt_qposadr[j]]; if (qd) qd[u] = qvel[m.jnt_dofadr[j]]; }
}

// The seven reward terms of Joystick._get_reward (joystick.py:622-669), unscaled:
//   [tracking_lin_vel, tracking_ang_vel, torques, action_rate, stand_still, alive, imitation]
// rewards.py:11-31,68-79,93-125 and custom_rewards.py:4-149.  Checked against the reference's NumPy twins
// (tests/golden/rewards.npz, made by tools/make_golden.py).
static void compute_rewards(const OduckHandle& h, const real* command, const real* local_linvel, const real* gyro, const real* actuator_force,
                            const real* action, const real* last_act, const real* base_qvel, const real* q, const real* qd, const real* contact,
                            const real* ref, real* out) {
  const OduckModel& m = h.m;
  const OduckEnvConfig& c = h.cfg;
  real ex = (command[0] - local_linvel[0]) * (command[0] - local_linvel[0]);
  real ey = std::max(std::fabs(local_linvel[1] - command[1]) - (real)0.1, (real)0);
  out[0] = nan_to_num(std::exp(-(ex + ey * ey) / (real)c.tracking_sigma));
  real ea = (command[2] - gyro[2]) * (command[2] - gyro[2]);
  out[1] = nan_to_num(std::exp(-ea / (real)c.tracking_sigma));
  real c_torque = 0, c_rate = 0;
  for (int u = 0; u < m.nu; u++) { c_torque += actuator_force[u] * actuator_force[u]; real d = action[u] - last_act[u]; c_rate += d * d; }
  out[2] = nan_to_num(c_torque);
  out[3] = nan_to_num(c_rate);
  real cmd_norm = std::sqrt(command[0] * command[0] + command[1] * command[1] + command[2] * command[2]);
  real pose = 0, vel = 0;
  for (int u = 0; u < m.nu; u++) { pose += std::fabs(q[u] - (real)m.key_ctrl[u]); vel += std::fabs(qd[u]); }
  out[4] = nan_to_num(pose + vel) * (cmd_norm < (real)0.01 ? 1 : 0);
  out[5] = 1;
  out[6] = 0;
  if (c.use_imitation_reward) {
    real lxy = (base_qvel[0] - ref[34]) * (base_qvel[0] - ref[34]) + (base_qvel[1] - ref[35]) * (base_qvel[1] - ref[35]);
    real lz = (base_qvel[2] - ref[36]) * (base_qvel[2] - ref[36]);
    real axy = (base_qvel[3] - ref[37]) * (base_qvel[3] - ref[37]) + (base_qvel[4] - ref[38]) * (base_qvel[4] - ref[38]);
    real az = (base_qvel[5] - ref[39]) * (base_qvel[5] - ref[39]);
    real jp_ = 0, jv_ = 0;
    for (int k = 0; k < 10; k++) {
      int u = k < 5 ? k : k + 4;       // joints_qpos[:5] ++ joints_qpos[9:]
      int rr = k < 5 ? k : k + 6;      // ref[:5] ++ ref[11:16]
      jp_ += (q[u] - ref[rr]) * (q[u] - ref[rr]);
      jv_ += (qd[u] - ref[16 + rr]) * (qd[u] - ref[16 + rr]);
    }
    real crew = 0;
    for (int i = 0; i < 2; i++) crew += (contact[i] == (ref[32 + i] > (real)0.5 ? (real)1 : (real)0)) ? 1 : 0;
    real rew = std::exp(-8 * lxy) + std::exp(-8 * lz) + (real)0.5 * std::exp(-2 * axy) + (real)0.5 * std::exp(-2 * az) - 15 * jp_ - (real)1e-3 * jv_ + crew;
    rew *= cmd_norm > (real)0.01 ? 1 : 0;
    out[6] = nan_to_num(rew);
  }
}

// The six reward terms of Standing._get_reward (standing.py:573-606), unscaled, in the order of its dict:
//   [orientation, torques, action_rate, alive, stand_still (ignore_head=True), head_pos]
// rewards.py:45-46 (cost_orientation), :68-79, :93-117, :124-125, :131-147 (cost_head_pos).  Checked against the NumPy twins
// (tests/golden/rewards_standing.npz).
static void compute_rewards_standing(const OduckHandle& h, const real* command, const real* upvector, const real* actuator_force, const real* action,
                                     const real* last_act, const real* q, const real* qd, real* out) {
  const OduckModel& m = h.m;
  out[0] = nan_to_num(upvector[0] * upvector[0] + upvector[1] * upvector[1]);
  real c_torque = 0, c_rate = 0;
  for (int u = 0; u < m.nu; u++) { c_torque += actuator_force[u] * actuator_force[u]; real d = action[u] - last_act[u]; c_rate += d * d; }
  out[1] = nan_to_num(c_torque);
  out[2] = nan_to_num(c_rate);
  out[3] = 1;
  real cmd_norm = std::sqrt(command[0] * command[0] + command[1] * command[1] + command[2] * command[2]);
  real pose = 0, vel = 0;
  for (int u = 0; u < m.nu; u++) {
    if (u >= 5 && u < 9) continue;                     // qpos[:5] and qpos[9:]: the head joints are ignored
    pose += std::fabs(q[u] - (real)m.key_ctrl[u]); vel += std::fabs(qd[u]);
  }
  out[4] = nan_to_num(pose + vel) * (cmd_norm < (real)0.01 ? 1 : 0);
  real herr = 0;
  for (int k = 0; k < 4; k++) { real d = q[5 + k] - command[3 + k]; herr += d * d; }
  out[5] = nan_to_num(herr) * (cmd_norm > (real)0.01 ? 1 : 0);
}

// The rest of the reward library (common/rewards.py:37-90,120,152-241; twin common/rewards_numpy.py): terms that neither
// Joystick nor Standing wires in.  Pure functions of one env's quantities, argument for argument like the
// reference; checked against the NumPy twins (tests/golden/rewards_library.npz).  Not yet reachable from OduckEnvConfig.
struct RewardLibrary {
  static real sq(real x) { return x * x; }
  static real cost_lin_vel_z(const real* global_linvel) { return nan_to_num(sq(global_linvel[2])); }                       // :37-38
  static real cost_ang_vel_xy(const real* global_angvel) { return nan_to_num(sq(global_angvel[0]) + sq(global_angvel[1])); }   // :41-42
  static real cost_base_height(real base_height, real target) { return nan_to_num(sq(base_height - target)); }              // :49-50
  static real reward_base_y_swing(real base_y_speed, real freq, real amplitude, real t, real tracking_sigma) {              // :53-65
    const real target = amplitude * std::sin(2 * (real)M_PI * freq * t);
    return nan_to_num(std::exp(-sq(target - base_y_speed) / tracking_sigma));
  }
  static real cost_energy(int n, const real* qvel, const real* qfrc_actuator) {                                             // :73-74
    real a = 0;
    for (int i = 0; i < n; i++) a += std::fabs(qvel[i]) * std::fabs(qfrc_actuator[i]);
    return nan_to_num(a);
  }
  static real cost_joint_pos_limits(int n, const real* qpos, const real* soft_lowers, const real* soft_uppers) {            // :85-90
    real a = 0;
    for (int i = 0; i < n; i++) a += -std::min(qpos[i] - soft_lowers[i], (real)0) + std::max(qpos[i] - soft_uppers[i], (real)0);
    return nan_to_num(a);
  }
  static real cost_termination(real done) { return done; }                                                                   // :120-121
  static real cost_joint_deviation_hip(const real* qpos, const real* cmd, int n_hip, const int* hip, const real* default_pose) {   // :152-158
    real a = 0;
    for (int i = 0; i < n_hip; i++) a += std::fabs(qpos[hip[i]] - default_pose[hip[i]]);
    return nan_to_num(a * (std::fabs(cmd[1]) > (real)0.1 ? 1 : 0));
  }
  static real cost_joint_deviation_knee(const real* qpos, int n_knee, const int* knee, const real* default_pose) {          // :161-167
    real a = 0;
    for (int i = 0; i < n_knee; i++) a += std::fabs(qpos[knee[i]] - default_pose[knee[i]]);
    return nan_to_num(a);
  }
  static real cost_pose(int n, const real* qpos, const real* default_pose, const real* weights) {                           // :170-175
    real a = 0;
    for (int i = 0; i < n; i++) a += sq(qpos[i] - default_pose[i]) * weights[i];
    return nan_to_num(a);
  }
  // :180-183 -- as written upstream: the norm of the BASE's horizontal velocity, counted once per foot in contact
  static real cost_feet_slip(const real* contact, const real* global_linvel) {
    const real v = std::sqrt(sq(global_linvel[0]) + sq(global_linvel[1]));
    return nan_to_num(v * contact[0] + v * contact[1]);
  }
  static real cost_feet_clearance(const real (*feet_vel)[3], const real (*foot_pos)[3], real max_foot_height) {             // :187-198
    real a = 0;
    for (int f = 0; f < 2; f++) a += std::fabs(foot_pos[f][2] - max_foot_height) * std::sqrt(std::sqrt(sq(feet_vel[f][0]) + sq(feet_vel[f][1])));
    return nan_to_num(a);
  }
  static real cost_feet_height(const real* swing_peak, const real* first_contact, real max_foot_height) {                   // :202-208
    real a = 0;
    for (int f = 0; f < 2; f++) a += sq(swing_peak[f] / max_foot_height - 1) * first_contact[f];
    return nan_to_num(a);
  }
  static real reward_feet_air_time(const real* air_time, const real* first_contact, const real* commands, real threshold_min, real threshold_max) {   // :212-224
    const real cmd_norm = std::sqrt(sq(commands[0]) + sq(commands[1]) + sq(commands[2]));
    real a = 0;
    for (int f = 0; f < 2; f++) a += std::min((air_time[f] - threshold_min) * first_contact[f], threshold_max - threshold_min);
    return nan_to_num(a * (cmd_norm > (real)0.01 ? 1 : 0));
  }
  static real reward_feet_phase(const real (*foot_pos)[3], const real* rz) {                                                // :228-241
    return nan_to_num(std::exp(-(sq(foot_pos[0][2] - rz[0]) + sq(foot_pos[1][2] - rz[1])) / (real)0.01));
  }
};

// mujoco_playground gait.get_rz(phi, swing_height) restated from upstream (not vendored in the reference): desired foot height over
// the gait phase phi in [-pi, pi) -- cubic Bezier 0 -> h over the first half of the period, h -> 0 over the second.
static real gait_rz(real phi, real swing_height) {
  auto bezier = [](real y0, real y1, real x) { return y0 + (y1 - y0) * (x * x * x + 3 * (x * x * (1 - x))); };
  const real x = (phi + (real)M_PI) / (2 * (real)M_PI);
  return x <= (real)0.5 ? bezier(0, swing_height, 2 * x) : bezier(swing_height, 0, 2 * x - 1);
}
extern "C" double oduck_test_gait_rz(double phi, double swing_height) { return (double)gait_rz((real)phi, (real)swing_height); }

// The library terms a task switched on through OduckEnvConfig.lib.scale (include/oduck.h), scaled and summed in enum order.
// Inputs follow the reference's accessors: joints = the nu actuated joints (base.py:193-215), imu-site sensors
// (base.py:234-264; global_linvel = site_xmat * local_linvel), feet sites (base.py:266-271), step-local contact /
// first_contact / feet_air_time (after `+= dt`) / swing_peak (joystick.py:424-435).
static real reward_library_sum(const OduckHandle& h, const EnvState& e, const real* q, const real* qd, const real* contact,
                               const real* first_contact, bool done) {
  const OduckModel& m = h.m;
  const OduckRewardLibrary& L = h.cfg.lib;
  bool any = false;
  for (int k = 0; k < ODUCK_NLIBTERM; k++) any |= L.scale[k] != 0;
  if (!any) return 0;
  typedef RewardLibrary R;
  const real* sd = e.sensordata;
  real glv[3];
  for (int i = 0; i < 3; i++) glv[i] = e.imu_xmat[3 * i] * sd[3] + e.imu_xmat[3 * i + 1] * sd[4] + e.imu_xmat[3 * i + 2] * sd[5];
  real lo[NU], hi[NU], w[NU], dflt[NU];
  for (int u = 0; u < m.nu; u++) { lo[u] = (real)L.soft_lowers[u]; hi[u] = (real)L.soft_uppers[u]; w[u] = (real)L.pose_weights[u]; dflt[u] = (real)m.key_ctrl[u]; }
  real feet_vel[2][3], foot_pos[2][3];
  for (int f = 0; f < 2; f++) for (int i = 0; i < 3; i++) { feet_vel[f][i] = sd[15 + 3 * f + i]; foot_pos[f][i] = e.site_xpos_feet[3 * f + i]; }
  const real up2[3] = {sd[9], sd[10], sd[11]};
  real t[ODUCK_NLIBTERM];
  t[ODUCK_LIB_ORIENTATION] = nan_to_num(up2[0] * up2[0] + up2[1] * up2[1]);                 // rewards.py:45-46
  t[ODUCK_LIB_LIN_VEL_Z] = R::cost_lin_vel_z(glv);
  t[ODUCK_LIB_ANG_VEL_XY] = R::cost_ang_vel_xy(sd + 12);
  t[ODUCK_LIB_BASE_HEIGHT] = R::cost_base_height(e.qpos[2], (real)L.base_height_target);
  t[ODUCK_LIB_ENERGY] = R::cost_energy(m.nu, qd, e.actuator_force);
  t[ODUCK_LIB_JOINT_POS_LIMITS] = R::cost_joint_pos_limits(m.nu, q, lo, hi);
  t[ODUCK_LIB_TERMINATION] = R::cost_termination(done ? (real)1 : (real)0);
  t[ODUCK_LIB_JOINT_DEVIATION_HIP] = R::cost_joint_deviation_hip(q, e.command, L.n_hip, L.hip_indices, dflt);
  t[ODUCK_LIB_JOINT_DEVIATION_KNEE] = R::cost_joint_deviation_knee(q, L.n_knee, L.knee_indices, dflt);
  t[ODUCK_LIB_POSE] = R::cost_pose(m.nu, q, dflt, w);
  t[ODUCK_LIB_FEET_SLIP] = R::cost_feet_slip(contact, glv);
  t[ODUCK_LIB_FEET_CLEARANCE] = R::cost_feet_clearance(feet_vel, foot_pos, (real)L.max_foot_height);
  t[ODUCK_LIB_FEET_HEIGHT] = R::cost_feet_height(e.swing_peak, first_contact, (real)L.max_foot_height);
  t[ODUCK_LIB_FEET_AIR_TIME] = R::reward_feet_air_time(e.feet_air_time, first_contact, e.command, (real)L.air_time_threshold_min, (real)L.air_time_threshold_max);
  {
    // gait clock (include/oduck.h): the reference-motion phase counter after this step's increment (joystick.py:352-356)
    const int period = h.cfg.nb_steps_in_period > 0 ? h.cfg.nb_steps_in_period : 1;
    const real tt = (real)e.imitation_i * (real)h.cfg.ctrl_dt;
    t[ODUCK_LIB_BASE_Y_SWING] = R::reward_base_y_swing(sd[4], (real)L.base_y_swing_freq, (real)L.base_y_swing_amplitude, tt, (real)h.cfg.tracking_sigma);
    real rz[2];
    for (int k = 0; k < 2; k++) {
      real phi = 2 * (real)M_PI * (real)e.imitation_i / (real)period + (real)k * (real)M_PI;
      phi -= 2 * (real)M_PI * std::floor((phi + (real)M_PI) / (2 * (real)M_PI));           // wrap to [-pi, pi)
      rz[k] = gait_rz(phi, (real)L.max_foot_height);
    }
    t[ODUCK_LIB_FEET_PHASE] = R::reward_feet_phase(foot_pos, rz);
  }
  real sum = 0;
  for (int k = 0; k < ODUCK_NLIBTERM; k++) if (L.scale[k] != 0) sum += t[k] * (real)L.scale[k];
  return sum;
}

// joystick.py:487-620.  Advances e.rng exactly as the reference (5 splits).
static void get_obs(const OduckHandle& h, EnvState& e, const real* contact) {
  const OduckModel& m = h.m;
  const OduckEnvConfig& c = h.cfg;
  const real* sd = e.sensordata;
  real lvl = (real)c.noise_level;
  auto split_noise = [&]() { Key nk = key_split(e.rng, 1); e.rng = key_split(e.rng, 0); return nk; };
  Key nk = split_noise();
  real noisy_gyro[3], noisy_acc[3], gravity[3], noisy_grav[3];
  for (int i = 0; i < 3; i++) noisy_gyro[i] = sd[i] + (2 * key_uniform(nk, i, 0, 1) - 1) * lvl * (real)c.noise_gyro;
  nk = split_noise();  // accelerometer: the +1.3 bias of joystick.py:502 is discarded by the reference (quirk #2)
  for (int i = 0; i < 3; i++) noisy_acc[i] = sd[6 + i] + (2 * key_uniform(nk, i, 0, 1) - 1) * lvl * (real)c.noise_accelerometer;
  for (int i = 0; i < 3; i++) gravity[i] = -e.imu_xmat[6 + i];  // site_xmat.T @ [0,0,-1]
  nk = split_noise();
  for (int i = 0; i < 3; i++) noisy_grav[i] = gravity[i] + (2 * key_uniform(nk, i, 0, 1) - 1) * lvl * (real)c.noise_gravity;
  int nh = c.imu_max_delay * 3;
  for (int i = nh - 1; i >= 3; i--) e.imu_history[i] = e.imu_history[i - 3];  // jp.roll(hist, 3).at[:3].set(...)
  for (int i = 0; i < 3 && i < nh; i++) e.imu_history[i] = noisy_grav[i];
  (void)key_randint(nk, c.imu_min_delay, c.imu_max_delay);  // imu delay index: computed, never observed (quirk #4)
  real q[NU], qd[NU], ja[NU];
  actuated(m, e.qpos, e.qvel, q, qd);
  for (int u = 0; u < m.nu; u++) {
    ja[u] = q[u];
    // backlash joint = the joint right after the actuated one on the same body (base.py:121-125, joystick.py:535-541)
    int j = m.act_jntid[u];
    if (j + 1 < m.njnt && m.jnt_bodyid[j + 1] == m.jnt_bodyid[j] && m.jnt_type[j + 1] == ODUCK_JNT_HINGE) ja[u] += e.qpos[m.jnt_qposadr[j + 1]];
  }
  real nja[NU], njv[NU];
  nk = split_noise();
  for (int u = 0; u < m.nu; u++) nja[u] = ja[u] + (2 * key_uniform(nk, u, 0, 1) - 1) * lvl * (real)c.qpos_noise_scale[u];
  nk = split_noise();
  for (int u = 0; u < m.nu; u++) njv[u] = qd[u] + (2 * key_uniform(nk, u, 0, 1) - 1) * lvl * (real)c.noise_joint_vel;
  real* o = e.obs_state;
  int p = 0;
  for (int i = 0; i < 3; i++) o[p++] = noisy_gyro[i];
  for (int i = 0; i < 3; i++) o[p++] = noisy_acc[i];
  for (int i = 0; i < 7; i++) o[p++] = e.command[i];
  for (int u = 0; u < m.nu; u++) o[p++] = nja[u] - (real)m.key_ctrl[u];
  for (int u = 0; u < m.nu; u++) o[p++] = njv[u] * (real)c.dof_vel_scale;
  const bool standing = c.task == ODUCK_TASK_STANDING;     // standing.py:526-542: no motor_targets / phase; the reference motion is empty
  for (int k = 0; k < 3; k++) for (int u = 0; u < m.nu; u++) o[p++] = e.last_act[k][u];
  if (!standing) for (int u = 0; u < m.nu; u++) o[p++] = e.motor_targets[u];
  for (int i = 0; i < 2; i++) o[p++] = contact[i];
  if (!standing) for (int i = 0; i < 2; i++) o[p++] = e.imitation_phase[i];
  real* pr = e.obs_priv;
  int r = 0;
  for (int i = 0; i < p; i++) pr[r++] = o[i];
  for (int i = 0; i < 3; i++) pr[r++] = sd[i];
  for (int i = 0; i < 3; i++) pr[r++] = sd[6 + i];
  for (int i = 0; i < 3; i++) pr[r++] = gravity[i];
  for (int i = 0; i < 3; i++) pr[r++] = sd[3 + i];
  for (int i = 0; i < 3; i++) pr[r++] = sd[12 + i];
  for (int u = 0; u < m.nu; u++) pr[r++] = ja[u] - (real)m.key_ctrl[u];
  for (int u = 0; u < m.nu; u++) pr[r++] = qd[u];
  pr[r++] = e.qpos[2];
  for (int u = 0; u < m.nu; u++) pr[r++] = e.actuator_force[u];
  for (int i = 0; i < 2; i++) pr[r++] = contact[i];
  for (int i = 0; i < 6; i++) pr[r++] = sd[15 + i];  // [left, right] foot linvel (joystick.py:173-181)
  for (int i = 0; i < 2; i++) pr[r++] = e.feet_air_time[i];
  if (!standing) {
    for (int i = 0; i < ODUCK_REF_DIM; i++) pr[r++] = e.ref_motion[i];
    pr[r++] = (real)e.imitation_i;
    for (int i = 0; i < 2; i++) pr[r++] = e.imitation_phase[i];
  }
  for (int i = p; i < ODUCK_OBS_STATE; i++) o[i] = 0;
  for (int i = r; i < ODUCK_OBS_PRIV; i++) pr[i] = 0;
}

static void geoms_colliding(const EnvState& e, real* contact) {
  for (int k = 0; k < 2; k++) {
    real dmin = (real)1e4;
    for (int c = 0; c < 4; c++) dmin = std::min(dmin, e.contact_dist[4 * k + c]);
    contact[k] = dmin < 0 ? 1 : 0;
  }
}

static void env_reset(OduckHandle& h, EnvState& e, Key rng) {  // joystick.py:206-321
  const OduckModel& m = h.m;
  const OduckEnvConfig& c = h.cfg;
  static thread_local Scratch s;
  for (int i = 0; i < m.nq; i++) e.qpos[i] = (real)m.key_qpos[i];
  for (int i = 0; i < m.nv; i++) { e.qvel[i] = 0; e.qacc_warm[i] = 0; e.qacc[i] = 0; }
  Key key;
  key = key_split(rng, 1); rng = key_split(rng, 0);
  for (int i = 0; i < 2; i++) e.qpos[i] = (real)m.key_qpos[i] + key_uniform(key, i, (real)-0.05, (real)0.05);
  key = key_split(rng, 1); rng = key_split(rng, 0);
  real yaw = key_uniform(key, 0, (real)-3.14, (real)3.14);
  real qy[4] = {std::cos(yaw / 2), 0, 0, std::sin(yaw / 2)}, q0[4] = {(real)m.key_qpos[3], (real)m.key_qpos[4], (real)m.key_qpos[5], (real)m.key_qpos[6]}, qn[4];
  quat_mul(q0, qy, qn);
  for (int i = 0; i < 4; i++) e.qpos[3 + i] = qn[i];
  key = key_split(rng, 1); rng = key_split(rng, 0);
  for (int u = 0; u < m.nu; u++) { int qa = m.jnt_qposadr[m.act_jntid[u]]; e.qpos[qa] = e.qpos[qa] * key_uniform(key, u, (real)0.5, (real)1.5); }
  key = key_split(rng, 1); rng = key_split(rng, 0);
  for (int i = 0; i < 6; i++) e.qvel[i] = key_uniform(key, i, -(real)c.reset_base_qvel_noise, (real)c.reset_base_qvel_noise);
  for (int u = 0; u < m.nu; u++) e.ctrl[u] = e.qpos[m.jnt_qposadr[m.act_jntid[u]]];
  forward(h, e, s);  // mjx_env.init
  Key cmd_rng = key_split(rng, 1); rng = key_split(rng, 0);
  sample_command(h, cmd_rng, e.command);
  Key push_rng = key_split(rng, 1); rng = key_split(rng, 0);
  real pi_ = key_uniform(push_rng, 0, (real)c.push_interval_range[0], (real)c.push_interval_range[1]);
  e.push_interval_steps = (int32_t)std::nearbyint(pi_ / (real)c.ctrl_dt);
  if (c.use_imitation_reward) poly_reference_motion(h, e.command[0], e.command[1], e.command[2], 0, e.ref_motion);
  else for (int i = 0; i < ODUCK_REF_DIM; i++) e.ref_motion[i] = 0;
  e.rng = rng;
  e.step = 0; e.steps = 0; e.push_step = 0; e.imitation_i = 0;
  for (int k = 0; k < 3; k++) for (int u = 0; u < NU; u++) e.last_act[k][u] = 0;
  for (int u = 0; u < m.nu; u++) e.motor_targets[u] = c.task == ODUCK_TASK_STANDING ? (real)0 : (real)m.key_ctrl[u];   // standing.py:279
  for (int i = 0; i < 2; i++) { e.feet_air_time[i] = 0; e.last_contact[i] = 0; e.swing_peak[i] = 0; e.push[i] = 0; e.imitation_phase[i] = 0; }
  for (int i = 0; i < 8 * NU; i++) e.action_history[i] = 0;
  for (int i = 0; i < 24; i++) e.imu_history[i] = 0;
  for (int i = 0; i < ODUCK_NMETRIC; i++) e.metrics[i] = 0;
  real contact[2];
  geoms_colliding(e, contact);
  get_obs(h, e, contact);
  e.reward = 0; e.done = 0; e.truncation = 0;
  // BraxAutoResetWrapper.reset: first_state / first_obs
  for (int i = 0; i < NQ; i++) e.first_qpos[i] = e.qpos[i];
  for (int i = 0; i < NV; i++) { e.first_qvel[i] = e.qvel[i]; e.first_qacc_warm[i] = e.qacc_warm[i]; }
  for (int i = 0; i < ODUCK_OBS_STATE; i++) e.first_obs_state[i] = e.obs_state[i];
  for (int i = 0; i < ODUCK_OBS_PRIV; i++) e.first_obs_priv[i] = e.obs_priv[i];
}

static void env_step(OduckHandle& h, EnvState& e, const float* action_f) {  // joystick.py:323-481 + wrappers
  const OduckModel& m = h.m;
  const OduckEnvConfig& c = h.cfg;
  static thread_local Scratch s;
  const real pi = (real)3.14159265358979323846;
  real action[NU];
  for (int u = 0; u < m.nu; u++) action[u] = (real)action_f[u];
  // AutoResetWrapper.step prologue: steps = where(done, 0, steps); done = 0
  if (e.done != 0) e.steps = 0;
  real dt = (real)c.ctrl_dt;
  if (c.use_imitation_reward) {
    e.imitation_i = (e.imitation_i + 1) % c.nb_steps_in_period;
    real ph = ((real)e.imitation_i / (real)c.nb_steps_in_period) * 2 * pi;
    e.imitation_phase[0] = std::cos(ph);
    e.imitation_phase[1] = std::sin(ph);
    poly_reference_motion(h, e.command[0], e.command[1], e.command[2], e.imitation_i, e.ref_motion);
  } else {
    e.imitation_i = 0;
  }
  Key push1 = key_split(e.rng, 1), push2 = key_split(e.rng, 2), delay = key_split(e.rng, 3);
  e.rng = key_split(e.rng, 0);
  int nh = c.action_max_delay * m.nu;
  for (int i = nh - 1; i >= m.nu; i--) e.action_history[i] = e.action_history[i - m.nu];
  for (int u = 0; u < m.nu && u < nh; u++) e.action_history[u] = action[u];
  int aidx = key_randint(delay, c.action_min_delay, c.action_max_delay);
  const real* act_delayed = e.action_history + aidx * m.nu;
  real theta = key_uniform(push1, 0, 0, 2 * pi);
  real mag = key_uniform(push2, 0, (real)c.push_magnitude_range[0], (real)c.push_magnitude_range[1]);
  real push[2] = {std::cos(theta), std::sin(theta)};
  bool fire = ((e.push_step + 1) % e.push_interval_steps) == 0;  // jp.mod of positive ints
  for (int i = 0; i < 2; i++) { push[i] *= fire ? 1 : 0; push[i] *= c.push_enable ? 1 : 0; e.qvel[i] += push[i] * mag; }
  real targets[NU];
  for (int u = 0; u < m.nu; u++) {
    targets[u] = (real)m.key_ctrl[u] + act_delayed[u] * (real)c.action_scale;
    if (c.use_motor_speed_limits) {
      real lim = (real)c.max_motor_velocity * dt;
      targets[u] = std::min(std::max(targets[u], e.motor_targets[u] - lim), e.motor_targets[u] + lim);
    }
  }
  for (int u = 0; u < m.nu; u++) e.ctrl[u] = targets[u];
  for (int k = 0; k < c.n_substeps; k++) { forward(h, e, s); euler(m, e); }
  for (int u = 0; u < m.nu; u++) e.motor_targets[u] = targets[u];
  real contact[2], first_contact[2];
  geoms_colliding(e, contact);
  for (int i = 0; i < 2; i++) {
    real filt = (contact[i] != 0 || e.last_contact[i] != 0) ? 1 : 0;
    first_contact[i] = (e.feet_air_time[i] > 0 ? 1 : 0) * filt;
    e.feet_air_time[i] += dt;
    e.swing_peak[i] = std::max(e.swing_peak[i], e.site_xpos_feet[3 * i + 2]);
  }
  get_obs(h, e, contact);
  bool nan_state = false;
  for (int i = 0; i < m.nq; i++) nan_state |= std::isnan(e.qpos[i]);
  for (int i = 0; i < m.nv; i++) nan_state |= std::isnan(e.qvel[i]);
  bool done = (e.sensordata[9 + 2] < 0) || nan_state;
  // rewards (common/rewards.py, custom_rewards.py)
  real q[NU], qd[NU];
  actuated(m, e.qpos, e.qvel, q, qd);
  real sc[7] = {0, 0, 0, 0, 0, 0, 0}, total;
  double scales[7] = {0, 0, 0, 0, 0, 0, 0};   // metric slots: config order of reward_config.scales
  if (c.task == ODUCK_TASK_STANDING) {
    real t6[6];
    compute_rewards_standing(h, e.command, e.sensordata + 9, e.actuator_force, action, e.last_act[0], q, qd, t6);
    const real s_or = t6[0] * (real)c.scale_orientation, s_tq = t6[1] * (real)c.scale_torques, s_ar = t6[2] * (real)c.scale_action_rate;
    const real s_al = t6[3] * (real)c.scale_alive, s_ss = t6[4] * (real)c.scale_stand_still, s_hp = t6[5] * (real)c.scale_head_pos;
    total = s_or + s_tq + s_ar + s_al + s_ss + s_hp;          // sum order of the rewards dict (standing.py:585-604)
    sc[0] = s_or; sc[1] = s_tq; sc[2] = s_ar; sc[3] = s_ss; sc[4] = s_al; sc[5] = s_hp;
    scales[0] = c.scale_orientation; scales[1] = c.scale_torques; scales[2] = c.scale_action_rate; scales[3] = c.scale_stand_still; scales[4] = c.scale_alive; scales[5] = c.scale_head_pos;
  } else {
    real terms[7];
    compute_rewards(h, e.command, e.sensordata + 3, e.sensordata, e.actuator_force, action, e.last_act[0], e.qvel, q, qd, contact, e.ref_motion, terms);
    const real r_lin = terms[0], r_ang = terms[1], c_torque = terms[2], c_rate = terms[3], c_still = terms[4], r_alive = terms[5], r_imit = terms[6];
    sc[0] = r_lin * (real)c.scale_tracking_lin_vel; sc[1] = r_ang * (real)c.scale_tracking_ang_vel; sc[2] = c_torque * (real)c.scale_torques;
    sc[3] = c_rate * (real)c.scale_action_rate; sc[4] = c_still * (real)c.scale_stand_still; sc[5] = r_alive * (real)c.scale_alive; sc[6] = r_imit * (real)c.scale_imitation;
    // sum order of the rewards dict (joystick.py:634-667): lin, ang, torques, action_rate, alive, imitation, stand_still
    total = sc[0] + sc[1] + sc[2] + sc[3] + sc[5] + sc[6] + sc[4];
    scales[0] = c.scale_tracking_lin_vel; scales[1] = c.scale_tracking_ang_vel; scales[2] = c.scale_torques; scales[3] = c.scale_action_rate;
    scales[4] = c.scale_stand_still; scales[5] = c.scale_alive; scales[6] = c.scale_imitation;
  }
  total += reward_library_sum(h, e, q, qd, contact, first_contact, done);
  real reward = std::min(std::max(total * dt, (real)0), (real)10000);
  for (int i = 0; i < 2; i++) e.push[i] = push[i];
  e.step += 1;
  e.push_step += 1;
  for (int u = 0; u < m.nu; u++) { e.last_act[2][u] = e.last_act[1][u]; e.last_act[1][u] = e.last_act[0][u]; e.last_act[0][u] = action[u]; }
  Key cmd_rng = key_split(e.rng, 1);
  e.rng = key_split(e.rng, 0);
  if (e.step > 500) sample_command(h, cmd_rng, e.command);
  if (done || e.step > 500) e.step = 0;
  for (int i = 0; i < 2; i++) {
    real nc = contact[i] != 0 ? 0 : 1;
    e.feet_air_time[i] *= nc;
    e.last_contact[i] = contact[i];
    e.swing_peak[i] *= nc;
  }
  // metrics: reward/<k> = v, cost/<k> = -v for negative scales (joystick.py:470-477); swing_peak follows the last term
  for (int k = 0; k < 7; k++) e.metrics[k] = scales[k] == 0 ? 0 : (scales[k] > 0 ? sc[k] : -sc[k]);
  if (c.task == ODUCK_TASK_STANDING) { e.metrics[6] = (e.swing_peak[0] + e.swing_peak[1]) / 2; e.metrics[7] = 0; }
  else e.metrics[7] = (e.swing_peak[0] + e.swing_peak[1]) / 2;
  e.reward = reward;
  // EpisodeWrapper.step (action_repeat = 1)
  e.steps += 1;
  real d = done ? 1 : 0;
  bool trunc = e.steps >= c.episode_length;
  e.truncation = trunc ? 1 - d : 0;
  e.done = trunc ? 1 : d;
  // BraxAutoResetWrapper.step: data/obs <- first_state where done (info is NOT reset)
  if (c.auto_reset && e.done != 0) {
    for (int i = 0; i < NQ; i++) e.qpos[i] = e.first_qpos[i];
    for (int i = 0; i < NV; i++) { e.qvel[i] = e.first_qvel[i]; e.qacc_warm[i] = e.first_qacc_warm[i]; }
    for (int i = 0; i < ODUCK_OBS_STATE; i++) e.obs_state[i] = e.first_obs_state[i];
    for (int i = 0; i < ODUCK_OBS_PRIV; i++) e.obs_priv[i] = e.first_obs_priv[i];
  }
}

static void env_randomize(OduckHandle& h, EnvState& e, Key rng) {  // common/randomize.py:39-106
  const OduckModel& m = h.m;
  Key key;
  key = key_split(rng, 1); rng = key_split(rng, 0);
  e.dr_geom_friction0 = key_uniform(key, 0, (real)0.5, (real)1.0);
  key = key_split(rng, 1); rng = key_split(rng, 0);
  for (int i = 0; i < h.nefc_fr; i++) e.dof_frictionloss[h.fr_dof[i]] = (real)m.dof_frictionloss[h.fr_dof[i]] * key_uniform(key, i, (real)0.9, (real)1.1);
  key = key_split(rng, 1); rng = key_split(rng, 0);
  for (int i = 0; i < h.nefc_fr; i++) e.dof_armature[h.fr_dof[i]] = (real)m.dof_armature[h.fr_dof[i]] * key_uniform(key, i, (real)1.0, (real)1.05);
  key = key_split(rng, 1); rng = key_split(rng, 0);
  for (int i = 0; i < 3; i++) e.body_ipos[1][i] = (real)m.body_ipos[1][i] + key_uniform(key, i, (real)-0.05, (real)0.05);
  key = key_split(rng, 1); rng = key_split(rng, 0);
  for (int b = 0; b < m.nbody; b++) e.body_mass[b] = (real)m.body_mass[b] * key_uniform(key, b, (real)0.9, (real)1.1);
  key = key_split(rng, 1); rng = key_split(rng, 0);
  e.body_mass[1] += key_uniform(key, 0, (real)-0.1, (real)0.1);
  key = key_split(rng, 1); rng = key_split(rng, 0);
  for (int i = 0; i < h.nefc_fr; i++) { int qa = m.jnt_qposadr[m.dof_jntid[h.fr_dof[i]]]; e.qpos0[qa] = (real)m.qpos0[qa] + key_uniform(key, i, (real)-0.03, (real)0.03); }
  key = key_split(rng, 1); rng = key_split(rng, 0);
  for (int u = 0; u < m.nu; u++) e.act_kp[u] = (real)m.act_kp[u] * key_uniform(key, u, (real)0.9, (real)1.1);
}

static void env_nominal(const OduckHandle& h, EnvState& e) {
  const OduckModel& m = h.m;
  std::memset(&e, 0, sizeof(e));
  e.dr_geom_friction0 = 1;
  for (int b = 0; b < NB; b++) { e.body_mass[b] = (real)m.body_mass[b]; for (int i = 0; i < 3; i++) e.body_ipos[b][i] = (real)m.body_ipos[b][i]; }
  for (int d = 0; d < NV; d++) { e.dof_frictionloss[d] = (real)m.dof_frictionloss[d]; e.dof_armature[d] = (real)m.dof_armature[d]; }
  for (int i = 0; i < NQ; i++) { e.qpos0[i] = (real)m.qpos0[i]; e.qpos[i] = (real)m.key_qpos[i]; e.first_qpos[i] = e.qpos[i]; }
  for (int u = 0; u < NU; u++) { e.act_kp[u] = (real)m.act_kp[u]; e.ctrl[u] = (real)m.key_ctrl[u]; }
  e.push_interval_steps = 1 << 30;
  for (int c = 0; c < NCON; c++) e.contact_dist[c] = 1;
}

// ------------------------------------------------------------------------------------ C ABI
extern "C" {

int oduck_abi_version(void) { return ODUCK_ABI_VERSION; }
#ifdef ODUCK_COUNT_FLOPS
uint64_t oduck_flop_count(void) { return g_flops; }
void oduck_flop_reset(void) { g_flops = 0; }
#endif
int oduck_sizeof_model(void) { return (int)sizeof(OduckModel); }
int oduck_sizeof_env_config(void) { return (int)sizeof(OduckEnvConfig); }
const char* oduck_last_error(void) { return g_err.c_str(); }

int oduck_create(const OduckModel* model, const OduckEnvConfig* cfg, int num_envs, int device, OduckHandle** out) {
  (void)device;
  if (!model || !cfg || !out || num_envs <= 0) return fail(ODUCK_ERR_ARG, "oduck_create: bad argument");
  if (model->abi_version != ODUCK_ABI_VERSION) return fail(ODUCK_ERR_MODEL, "oduck_create: model ABI version mismatch");
  if (model->nv > NV || model->nq > NQ || model->nbody > NB || model->nu > NU) return fail(ODUCK_ERR_MODEL, "oduck_create: model too large");
  if (model->floor_is_hfield && (!model->hfield_data || model->hfield_nrow < 2 || model->hfield_ncol < 2))
    return fail(ODUCK_ERR_MODEL, "oduck_create: height-field floor without elevation data");
  if (cfg->action_max_delay > 8 || cfg->imu_max_delay > 8) return fail(ODUCK_ERR_ARG, "oduck_create: delay history too long");
  OduckHandle* h = new OduckHandle();
  h->m = *model;
  h->cfg = *cfg;
  size_t npoly = (size_t)cfg->ndx * cfg->ndy * cfg->ndth * ODUCK_REF_DIM * ODUCK_POLY_DEG;
  if (cfg->use_imitation_reward) {
    if (!cfg->poly_coef || npoly == 0) { delete h; return fail(ODUCK_ERR_ARG, "oduck_create: imitation reward needs poly_coef"); }
    h->poly.assign(cfg->poly_coef, cfg->poly_coef + npoly);
  }
  h->cfg.poly_coef = nullptr;
  if (model->floor_is_hfield) h->hfield.assign(model->hfield_data, model->hfield_data + (size_t)model->hfield_nrow * model->hfield_ncol);
  h->m.hfield_data = nullptr;
  h->n = num_envs;
  h->launches = 0;
  h->nefc_fr = h->nefc_lim = 0;
  for (int d = 0; d < model->nv; d++) if (model->dof_frictionloss[d] > 0) h->fr_dof[h->nefc_fr++] = d;
  for (int j = 0; j < model->njnt; j++) if (model->jnt_limited[j] && model->jnt_type[j] == ODUCK_JNT_HINGE) h->lim_jnt[h->nefc_lim++] = j;
  h->env.resize(num_envs);
  for (auto& e : h->env) env_nominal(*h, e);
  *out = h;
  return ODUCK_OK;
}
int oduck_destroy(OduckHandle* h) { delete h; return ODUCK_OK; }
int oduck_num_envs(const OduckHandle* h) { return h ? h->n : 0; }
int64_t oduck_launch_count(const OduckHandle* h) { return h ? h->launches : 0; }
int oduck_policy_invalidate(OduckHandle*) { return ODUCK_OK; }   // no packed copy on the CPU

int oduck_randomize(OduckHandle* h, const uint32_t* keys, void*) {
  if (!h || !keys) return fail(ODUCK_ERR_ARG, "oduck_randomize: bad argument");
  pfor(h->n, [&](int i) { env_randomize(*h, h->env[i], Key{keys[2 * i], keys[2 * i + 1]}); });
  return ODUCK_OK;
}
int oduck_reset(OduckHandle* h, const uint32_t* keys, const uint8_t* mask, void*) {
  if (!h || !keys) return fail(ODUCK_ERR_ARG, "oduck_reset: bad argument");
  pfor(h->n, [&](int i) { if (!mask || mask[i]) env_reset(*h, h->env[i], Key{keys[2 * i], keys[2 * i + 1]}); });
  return ODUCK_OK;
}
int oduck_step(OduckHandle* h, const float* action, void*) {
  if (!h || !action) return fail(ODUCK_ERR_ARG, "oduck_step: bad argument");
  pfor(h->n, [&](int i) { env_step(*h, h->env[i], action + (size_t)i * h->m.nu); });
  return ODUCK_OK;
}
int oduck_set_rollout_sink(OduckHandle* h, const OduckRolloutSink* sink) {
  if (!h) return fail(ODUCK_ERR_ARG, "oduck_set_rollout_sink: bad argument");
  if (!sink) { h->sink = OduckRolloutSink{}; return ODUCK_OK; }
  const int dp = h->cfg.task == ODUCK_TASK_STANDING ? 85 : ODUCK_OBS_STATE, dv = h->cfg.task == ODUCK_TASK_STANDING ? 153 : ODUCK_OBS_PRIV;
  if (sink->unroll < 1 || sink->env_offset < 0 || sink->env_offset + h->n > sink->num_envs) return fail(ODUCK_ERR_ARG, "oduck_set_rollout_sink: the handle's envs do not fit into the buffers");
  if (sink->policy_dim != dp || sink->value_dim != dv) return fail(ODUCK_ERR_ARG, "oduck_set_rollout_sink: obs row widths must be those of the task (Joystick 101 / 212, Standing 85 / 153)");
  if (!sink->obs_policy || !sink->obs_value || !sink->raw_action || !sink->log_prob || !sink->reward || !sink->done || !sink->truncation)
    return fail(ODUCK_ERR_ARG, "oduck_set_rollout_sink: null buffer");
  h->sink = *sink;
  return ODUCK_OK;
}
int oduck_policy_forward(OduckHandle* h, const OduckPolicyWeights* w, const float* obs, const uint32_t* keys, int deterministic,
                         float* action, float* raw_action, float* log_prob, void*);
// A17: one step of Brax generate_unroll (common/runner.py:104-118): Transition(observation = obs before the step, action,
// reward, discount / truncation, policy extras = raw action + log-prob); the observation after the step is the next slot's.
int oduck_rollout_step(OduckHandle* h, const OduckPolicyWeights* w, const uint32_t* keys, int t, void* stream) {
  if (!h || !w || !keys) return fail(ODUCK_ERR_ARG, "oduck_rollout_step: bad argument");
  const OduckRolloutSink& k = h->sink;
  if (!k.obs_policy || t < 0 || t >= k.unroll) return fail(ODUCK_ERR_ARG, "oduck_rollout_step: no sink attached or t outside the unroll");
  if (w->obs_dim != k.policy_dim) return fail(ODUCK_ERR_ARG, "oduck_rollout_step: the policy's input width is not the sink's policy_dim");
  const int nu = h->m.nu;
  const size_t row = (size_t)t * k.num_envs + k.env_offset;
  auto put_obs = [&](size_t r0) {
    for (int i = 0; i < h->n; i++) {
      for (int c = 0; c < k.policy_dim; c++) k.obs_policy[(r0 + i) * k.policy_dim + c] = (float)h->env[i].obs_state[c];
      for (int c = 0; c < k.value_dim; c++) k.obs_value[(r0 + i) * k.value_dim + c] = (float)h->env[i].obs_priv[c];
    }
  };
  if (t == 0) put_obs((size_t)k.env_offset);
  h->act_buf.resize((size_t)h->n * nu);
  int rc = oduck_policy_forward(h, w, nullptr, keys, 0, h->act_buf.data(), k.raw_action + row * nu, k.log_prob + row, stream);
  if (rc) return rc;
  rc = oduck_step(h, h->act_buf.data(), stream);
  if (rc) return rc;
  for (int i = 0; i < h->n; i++) { k.reward[row + i] = (float)h->env[i].reward; k.done[row + i] = (float)h->env[i].done; k.truncation[row + i] = (float)h->env[i].truncation; }
  put_obs(row + k.num_envs);
  return ODUCK_OK;
}
int oduck_physics_substeps(OduckHandle* h, const float* ctrl, int n, void*) {
  if (!h || n < 0) return fail(ODUCK_ERR_ARG, "oduck_physics_substeps: bad argument");
  pfor(h->n, [&](int i) {
    static thread_local Scratch s;
    EnvState& e = h->env[i];
    if (ctrl) for (int u = 0; u < h->m.nu; u++) e.ctrl[u] = (real)ctrl[(size_t)i * h->m.nu + u];
    for (int k = 0; k < n; k++) { forward(*h, e, s); euler(h->m, e); }
  });
  return ODUCK_OK;
}
int oduck_forward(OduckHandle* h, void*) {
  if (!h) return fail(ODUCK_ERR_ARG, "oduck_forward: bad argument");
  pfor(h->n, [&](int i) { { static thread_local Scratch s; forward(*h, h->env[i], s); } });
  return ODUCK_OK;
}
int oduck_set_state(OduckHandle* h, const float* qpos, const float* qvel, const float* qacc_warm, void*) {
  if (!h) return fail(ODUCK_ERR_ARG, "oduck_set_state: bad argument");
  for (int i = 0; i < h->n; i++) {
    EnvState& e = h->env[i];
    if (qpos) for (int k = 0; k < h->m.nq; k++) e.qpos[k] = (real)qpos[(size_t)i * h->m.nq + k];
    if (qvel) for (int k = 0; k < h->m.nv; k++) e.qvel[k] = (real)qvel[(size_t)i * h->m.nv + k];
    if (qacc_warm) for (int k = 0; k < h->m.nv; k++) e.qacc_warm[k] = (real)qacc_warm[(size_t)i * h->m.nv + k];
  }
  return ODUCK_OK;
}

// Brax policy MLP: normalise -> 3 x (Dense + swish) -> Dense -> NormalTanhDistribution (common/export_onnx.py:64-72)
int oduck_policy_forward(OduckHandle* h, const OduckPolicyWeights* w, const float* obs, const uint32_t* keys, int deterministic,
                         float* action, float* raw_action, float* log_prob, void*) {
  if (!h || !w) return fail(ODUCK_ERR_ARG, "oduck_policy_forward: bad argument");
  if (!deterministic && !keys) return fail(ODUCK_ERR_ARG, "oduck_policy_forward: stochastic policy needs keys");
  int dims[5] = {w->obs_dim, w->hidden[0], w->hidden[1], w->hidden[2], w->out_dim};
  int na = w->out_dim / 2;
  pfor(h->n, [&](int i) {
    std::vector<real> x(dims[0]), y;
    for (int k = 0; k < dims[0]; k++) {
      real o = obs ? (real)obs[(size_t)i * dims[0] + k] : h->env[i].obs_state[k];
      x[k] = (o - (real)w->obs_mean[k]) / (real)w->obs_std[k];
    }
    for (int l = 0; l < 4; l++) {
      // y[o] = b[o] + sum_k x[k] W[k][o], accumulated over k in order for every o (row-wise axpy: contiguous, vectorisable --
      // the same additions in the same order as the textbook o-outer loop)
      const int no = dims[l + 1];
      y.resize(no);
      for (int o = 0; o < no; o++) y[o] = (real)w->b[l][o];
      for (int k = 0; k < dims[l]; k++) {
        const real xk = x[k];
        const float* wr = w->w[l] + (size_t)k * no;
        for (int o = 0; o < no; o++) y[o] += xk * (real)wr[o];
      }
      if (l < 3) for (int o = 0; o < no; o++) y[o] = y[o] / (1 + std::exp(-y[o]));
      x = y;
    }
    real lp = 0;
    for (int a = 0; a < na; a++) {
      real loc = x[a], scale = std::log1p(std::exp(x[na + a])) + (real)0.001;  // softplus + min_std
      real raw = loc;
      if (!deterministic) {
        // jax.random.normal: sqrt(2) * erf_inv(uniform(-1 + ulp, 1)); erf_inv = the f32 polynomial XLA uses (M. Giles)
        Key k = Key{keys[2 * i], keys[2 * i + 1]};
        const real lo = (real)-0.99999994f;
        real u = std::max(lo, bits_to_unit(key_bits(k, a)) * (1 - lo) + lo);
        real w2 = -std::log((1 - u) * (1 + u)), pp;
        if (w2 < 5) {
          w2 -= (real)2.5;
          pp = (real)2.81022636e-08; pp = (real)3.43273939e-07 + pp * w2; pp = (real)-3.5233877e-06 + pp * w2; pp = (real)-4.39150654e-06 + pp * w2;
          pp = (real)0.00021858087 + pp * w2; pp = (real)-0.00125372503 + pp * w2; pp = (real)-0.00417768164 + pp * w2; pp = (real)0.246640727 + pp * w2;
          pp = (real)1.50140941 + pp * w2;
        } else {
          w2 = std::sqrt(w2) - 3;
          pp = (real)-0.000200214257; pp = (real)0.000100950558 + pp * w2; pp = (real)0.00134934322 + pp * w2; pp = (real)-0.00367342844 + pp * w2;
          pp = (real)0.00573950773 + pp * w2; pp = (real)-0.0076224613 + pp * w2; pp = (real)0.00943887047 + pp * w2; pp = (real)1.00167406 + pp * w2;
          pp = (real)2.83297682 + pp * w2;
        }
        real z = (real)1.41421356237 * pp * u;
        raw = loc + scale * z;
        real lpn = -(real)0.5 * z * z - std::log(scale) - (real)0.5 * std::log(2 * (real)3.14159265358979323846);
        real ldj = 2 * (std::log((real)2) - raw - std::log1p(std::exp(-2 * raw)));
        lp += lpn - ldj;
      }
      if (action) action[(size_t)i * na + a] = (float)std::tanh(raw);
      if (raw_action) raw_action[(size_t)i * na + a] = (float)raw;
    }
    if (log_prob) log_prob[i] = (float)lp;
  });
  return ODUCK_OK;
}

// Test-only exports (oracle only): direct access to the pieces pinned by tests/golden/*.npz.
int oduck_test_reference_motion(OduckHandle* h, double dx, double dy, double dth, int i, double* out40) {
  if (!h || !out40) return fail(ODUCK_ERR_ARG, "oduck_test_reference_motion: bad argument");
  real o[ODUCK_REF_DIM];
  poly_reference_motion(*h, (real)dx, (real)dy, (real)dth, i, o);
  for (int k = 0; k < ODUCK_REF_DIM; k++) out40[k] = (double)(o[k]);
  return ODUCK_OK;
}
// in: command[7] local_linvel[3] gyro[3] actuator_force[nu] action[nu] last_act[nu] base_qvel[6] q[nu] qd[nu] contact[2] ref[40]
int oduck_test_rewards(OduckHandle* h, const double* in, double* out7) {
  if (!h || !in || !out7) return fail(ODUCK_ERR_ARG, "oduck_test_rewards: bad argument");
  const int nu = h->m.nu;
  std::vector<real> v(in, in + 7 + 3 + 3 + 3 * nu + 6 + 2 * nu + 2 + ODUCK_REF_DIM);
  const real* p = v.data();
  const real *cmd = p, *lv = p + 7, *gy = p + 10, *af = p + 13, *ac = af + nu, *la = ac + nu, *bq = la + nu, *q = bq + 6, *qd = q + nu, *ct = qd + nu, *rf = ct + 2;
  real o[7];
  compute_rewards(*h, cmd, lv, gy, af, ac, la, bq, q, qd, ct, rf, o);
  for (int k = 0; k < 7; k++) out7[k] = (double)(o[k]);
  return ODUCK_OK;
}

// in: command[7] upvector[3] actuator_force[nu] action[nu] last_act[nu] q[nu] qd[nu]; out: the six Standing terms (dict order)
int oduck_test_rewards_standing(OduckHandle* h, const double* in, double* out6) {
  if (!h || !in || !out6) return fail(ODUCK_ERR_ARG, "oduck_test_rewards_standing: bad argument");
  const int nu = h->m.nu;
  std::vector<real> v(in, in + 7 + 3 + 5 * nu);
  const real* p = v.data();
  const real *cmd = p, *up = p + 7, *af = p + 10, *ac = af + nu, *la = ac + nu, *q = la + nu, *qd = q + nu;
  real o[6];
  compute_rewards_standing(*h, cmd, up, af, ac, la, q, qd, o);
  for (int k = 0; k < 6; k++) out6[k] = (double)(o[k]);
  return ODUCK_OK;
}

// Reward-library terms (RewardLibrary above), one env per call.  `in` (doubles), nu = ODUCK joints:
//   global_linvel[3] global_angvel[3] base_height target base_y_speed freq amplitude t tracking_sigma qvel[nu] qfrc_actuator[nu]
//   qpos[nu] soft_lowers[nu] soft_uppers[nu] done command[7] default_pose[nu] n_hip hip[4] n_knee knee[4] weights[nu] contact[2]
//   feet_vel[2][3] foot_pos[2][3] max_foot_height swing_peak[2] first_contact[2] air_time[2] threshold_min threshold_max rz[2]
// out15: lin_vel_z ang_vel_xy base_height base_y_swing energy joint_pos_limits termination joint_deviation_hip
//        joint_deviation_knee pose feet_slip feet_clearance feet_height feet_air_time feet_phase
int oduck_test_reward_library(int nu, const double* in, double* out15) {
  if (!in || !out15 || nu < 1 || nu > ODUCK_MAX_NU) return fail(ODUCK_ERR_ARG, "oduck_test_reward_library: bad argument");
  const int total = 13 + 5 * nu + 1 + 7 + nu + 10 + nu + 2 + 12 + 1 + 6 + 2 + 2;
  std::vector<real> v(in, in + total);
  const real* p = v.data();
  auto take = [&](int n) { const real* q = p; p += n; return q; };
  const real *glv = take(3), *gav = take(3), *sc = take(7), *qvel = take(nu), *qfrc = take(nu), *qpos = take(nu), *lo = take(nu), *hi = take(nu);
  const real *done = take(1), *cmd = take(7), *dflt = take(nu), *hipr = take(5), *kneer = take(5), *w = take(nu), *contact = take(2);
  const real (*feet_vel)[3] = reinterpret_cast<const real (*)[3]>(take(6));
  const real (*foot_pos)[3] = reinterpret_cast<const real (*)[3]>(take(6));
  const real *mfh = take(1), *peak = take(2), *first = take(2), *air = take(2), *thr = take(2), *rz = take(2);
  int hip[4], knee[4];
  const int n_hip = (int)hipr[0], n_knee = (int)kneer[0];
  if (n_hip < 0 || n_hip > 4 || n_knee < 0 || n_knee > 4) return fail(ODUCK_ERR_ARG, "oduck_test_reward_library: at most 4 hip / knee indices");
  for (int i = 0; i < 4; i++) { hip[i] = (int)hipr[1 + i]; knee[i] = (int)kneer[1 + i]; }
  typedef RewardLibrary R;
  const real o[15] = {R::cost_lin_vel_z(glv), R::cost_ang_vel_xy(gav), R::cost_base_height(sc[0], sc[1]), R::reward_base_y_swing(sc[2], sc[3], sc[4], sc[5], sc[6]),
                      R::cost_energy(nu, qvel, qfrc), R::cost_joint_pos_limits(nu, qpos, lo, hi), R::cost_termination(done[0]),
                      R::cost_joint_deviation_hip(qpos, cmd, n_hip, hip, dflt), R::cost_joint_deviation_knee(qpos, n_knee, knee, dflt), R::cost_pose(nu, qpos, dflt, w),
                      R::cost_feet_slip(contact, glv), R::cost_feet_clearance(feet_vel, foot_pos, mfh[0]), R::cost_feet_height(peak, first, mfh[0]),
                      R::reward_feet_air_time(air, first, cmd, thr[0], thr[1]), R::reward_feet_phase(foot_pos, rz)};
  for (int k = 0; k < 15; k++) out15[k] = (double)(o[k]);
  return ODUCK_OK;
}

int oduck_test_hf_max_candidates(int reset) { int v = g_hf_max_candidates; if (reset) g_hf_max_candidates = 0; return v; }

// Diagnostic twin of liboduck_cuda's oduck_debug_forward: same layout (DBG_STRIDE reals per env), values in double.
int oduck_debug_stride(void) { return 5120; }
int oduck_debug_forward(OduckHandle* h, double* out) {
  if (!h || !out) return fail(ODUCK_ERR_ARG, "oduck_debug_forward: bad argument");
  const OduckModel& m = h->m;
  for (int i = 0; i < h->n; i++) {
    static thread_local Scratch s;
    EnvState e = h->env[i];   // copy: diagnostic forward does not advance the warm start
    double* d = out + (size_t)i * 5120;
    for (int k = 0; k < 5120; k++) d[k] = (double)(0);
    g_dbg.on = true;
    forward(*h, e, s);
    g_dbg.on = false;
    h->env[i] = e;
    for (int a = 0; a < m.nv; a++) for (int b = 0; b < m.nv; b++) { d[a * 32 + b] = (double)(s.M[a][b]); d[4096 + a * 32 + b] = (double)(g_dbg.H[a][b]); }
    for (int a = 0; a < m.nv; a++) {
      d[1024 + a] = (double)(s.qfrc_bias[a]); d[1056 + a] = (double)(s.qfrc_smooth[a]); d[1088 + a] = (double)(s.qacc_smooth[a]);
      d[1376 + a] = (double)(g_dbg.search[a]); d[1408 + a] = (double)(g_dbg.grad[a]); d[1736 + a] = (double)(e.qacc[a]);
      for (int k = 0; k < 6; k++) d[1540 + a * 6 + k] = (double)(s.cdof[a][k]);
    }
    for (int c = 0; c < NCON; c++) { d[1120 + c] = (double)(s.con_dist[c]); for (int k = 0; k < 3; k++) { d[1136 + 3 * c + k] = (double)(s.con_pos[c][k]); d[2560 + 3 * c + k] = (double)(s.con_frame[c][k]); } }
    int r = 0;
    for (int k = 0; k < h->nefc_fr; k++, r++) { d[1184 + h->fr_dof[k]] = (double)(s.D[r]); d[1264 + h->fr_dof[k]] = (double)(s.aref[r]); }
    for (int k = 0; k < h->nefc_lim; k++, r++) { int dd = m.jnt_dofadr[h->lim_jnt[k]]; d[1216 + dd] = (double)(s.D[r]); d[1296 + dd] = (double)(s.aref[r]); }
    for (int c = 0; c < NCON && r + 3 < s.nefc + 4; c++) { d[1248 + c] = (double)(s.D[r]); for (int k = 0; k < 4; k++, r++) d[1328 + 4 * c + k] = r < s.nefc ? (double)(s.aref[r]) : 0.0; }
    for (int b = 0; b < m.nbody; b++) for (int k = 0; k < 3; k++) d[1440 + 3 * b + k] = (double)(s.xpos[b][k]);
    for (int k = 0; k < 3; k++) d[1536 + k] = (double)(s.com[k]);
    d[2536] = (double)(g_dbg.costw); d[2537] = (double)(g_dbg.costs); d[2538] = (double)(g_dbg.alpha); d[2539] = (double)(g_dbg.ls_it);
  }
  return ODUCK_OK;
}

int oduck_get_buffer(OduckHandle* h, int id, void** ptr, int64_t* shape, int64_t* strides, int* dtype) {
  if (!h || !ptr || !shape || !strides || !dtype) return fail(ODUCK_ERR_ARG, "oduck_get_buffer: bad argument");
  EnvState* e0 = h->env.data();
  const OduckModel& m = h->m;
  const int64_t es = sizeof(EnvState) / sizeof(real);
  static_assert(sizeof(EnvState) % sizeof(real) == 0, "EnvState must be a whole number of reals");
  int64_t d1 = 0, d2 = 0, s1 = 1, s2 = 1;
  void* p = nullptr;
  int dt = sizeof(real) == 8 ? ODUCK_DTYPE_F64 : ODUCK_DTYPE_F32;
  const int64_t es32 = sizeof(EnvState) / 4;
  bool is32 = false;
#define FIELD(f, n1) p = (void*)(e0->f); d1 = (n1);
  switch (id) {
    case ODUCK_BUF_QPOS: FIELD(qpos, m.nq) break;
    case ODUCK_BUF_QVEL: FIELD(qvel, m.nv) break;
    case ODUCK_BUF_QACC_WARM: FIELD(qacc_warm, m.nv) break;
    case ODUCK_BUF_QACC: FIELD(qacc, m.nv) break;
    case ODUCK_BUF_CTRL: FIELD(ctrl, m.nu) break;
    case ODUCK_BUF_OBS_STATE: FIELD(obs_state, h->cfg.task == ODUCK_TASK_STANDING ? 85 : ODUCK_OBS_STATE) break;
    case ODUCK_BUF_OBS_PRIV: FIELD(obs_priv, h->cfg.task == ODUCK_TASK_STANDING ? 153 : ODUCK_OBS_PRIV) break;
    case ODUCK_BUF_REWARD: p = &e0->reward; break;
    case ODUCK_BUF_DONE: p = &e0->done; break;
    case ODUCK_BUF_TRUNCATION: p = &e0->truncation; break;
    case ODUCK_BUF_METRICS: FIELD(metrics, ODUCK_NMETRIC) break;
    case ODUCK_BUF_EFC_FORCE: FIELD(efc_force, h->nefc_fr + h->nefc_lim + 4 * ODUCK_CON_PER_PAIR * (2 + (m.enable_foot_foot ? 1 : 0))) break;
    case ODUCK_BUF_CONTACT_DIST: FIELD(contact_dist, NCON) break;
    case ODUCK_BUF_SENSORDATA: FIELD(sensordata, 24) break;
    case ODUCK_BUF_ACTUATOR_FORCE: FIELD(actuator_force, m.nu) break;
    case ODUCK_BUF_SITE_XPOS_FEET: FIELD(site_xpos_feet, 6) break;
    case ODUCK_BUF_INFO_RNG: p = &e0->rng; d1 = 2; dt = ODUCK_DTYPE_U32; is32 = true; break;
    case ODUCK_BUF_INFO_COMMAND: FIELD(command, ODUCK_NCMD) break;
    case ODUCK_BUF_INFO_STEP: p = &e0->step; dt = ODUCK_DTYPE_I32; is32 = true; break;
    case ODUCK_BUF_INFO_STEPS: p = &e0->steps; dt = ODUCK_DTYPE_I32; is32 = true; break;
    case ODUCK_BUF_INFO_LAST_ACT: p = (void*)e0->last_act; d1 = 3; d2 = m.nu; s1 = NU; break;
    case ODUCK_BUF_INFO_MOTOR_TARGETS: FIELD(motor_targets, m.nu) break;
    case ODUCK_BUF_INFO_FEET_AIR_TIME: FIELD(feet_air_time, 2) break;
    case ODUCK_BUF_INFO_LAST_CONTACT: FIELD(last_contact, 2) break;
    case ODUCK_BUF_INFO_SWING_PEAK: FIELD(swing_peak, 2) break;
    case ODUCK_BUF_INFO_PUSH: FIELD(push, 2) break;
    case ODUCK_BUF_INFO_PUSH_STEP: p = &e0->push_step; dt = ODUCK_DTYPE_I32; is32 = true; break;
    case ODUCK_BUF_INFO_PUSH_INTERVAL: p = &e0->push_interval_steps; dt = ODUCK_DTYPE_I32; is32 = true; break;
    case ODUCK_BUF_INFO_ACTION_HISTORY: FIELD(action_history, h->cfg.action_max_delay * m.nu) break;
    case ODUCK_BUF_INFO_IMU_HISTORY: FIELD(imu_history, h->cfg.imu_max_delay * 3) break;
    case ODUCK_BUF_INFO_IMITATION_I: p = &e0->imitation_i; dt = ODUCK_DTYPE_I32; is32 = true; break;
    case ODUCK_BUF_INFO_REF_MOTION: FIELD(ref_motion, ODUCK_REF_DIM) break;
    case ODUCK_BUF_INFO_IMITATION_PHASE: FIELD(imitation_phase, 2) break;
    case ODUCK_BUF_DR_PARAMS: p = &e0->dr_geom_friction0; d1 = (int64_t)(offsetof(EnvState, sensordata) - offsetof(EnvState, dr_geom_friction0)) / (int64_t)sizeof(real); break;
    case ODUCK_BUF_FIRST_QPOS: FIELD(first_qpos, m.nq) break;
    case ODUCK_BUF_FIRST_QVEL: FIELD(first_qvel, m.nv) break;
    case ODUCK_BUF_FIRST_OBS_STATE: FIELD(first_obs_state, h->cfg.task == ODUCK_TASK_STANDING ? 85 : ODUCK_OBS_STATE) break;
    case ODUCK_BUF_FIRST_OBS_PRIV: FIELD(first_obs_priv, h->cfg.task == ODUCK_TASK_STANDING ? 153 : ODUCK_OBS_PRIV) break;
    default: return fail(ODUCK_ERR_ARG, "oduck_get_buffer: unknown buffer id");
  }
#undef FIELD
  *ptr = p;
  shape[0] = h->n; shape[1] = d1; shape[2] = d2; shape[3] = 0;
  strides[0] = is32 ? es32 : es; strides[1] = d2 ? s1 : 1; strides[2] = 1; strides[3] = 0;
  (void)s2;
  *dtype = dt;
  return ODUCK_OK;
}

}  // extern "C"
