// oduck_device.cuh -- device-side model tables and per-warp physics for the Open Duck step (sm_100a).
//
// Mapping: ONE WARP PER ENV.  Depending on the phase a lane is a dof (nv <= 32),
// a body (nbody <= 32), a hull vertex (<= 32) or a contact-Jacobian row.  Tree recurrences run as
// parallel prefix sums over the dof tree with warp shuffles; the joint-space inertia M and the Newton
// Hessian H live in shared memory as packed lower triangles (bank-conflict free by row and by column,
// because triangular numbers mod 32 are a permutation) and are factorised leaf-to-root (M = L^T L),
// which has no fill-in on a kinematic tree.  Restates what mjx.step computes for the reference at
// open_duck_mini_v2/joystick.py:420 (algorithms: SURVEY.md section 3.3 / 9; oracle/oduck_oracle.cpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FULLMASK 0xffffffffu
#define NLANE 32
#define TRI(i) (((i) * ((i) + 1)) >> 1)
#define CON_STRIDE 16     // floats per contact record in shared memory
#define NCON_FLOOR 8
#define NCON_ALL 12

// record strides (floats) of the per-env state in HBM -- one contiguous record per env per group
#define PHYS_STRIDE 128   // qpos[36] qvel[32] qacc_warm[32] ctrl[16] pad
#define PHYS_QVEL 36
#define PHYS_QACCW 68
#define PHYS_CTRL 100
#define DR_STRIDE 144     // mass[20] ipos1[3] fric0 frictionloss[32] armature[32] qpos0[36] kp[16]
#define DR_IPOS1 20
#define DR_FRIC0 23
#define DR_FLOSS 24
#define DR_ARM 56
#define DR_QPOS0 88
#define DR_KP 124
#define OUT_STRIDE 224    // qacc[32] sensordata[24] efc_force[96] contact_dist[12] actuator_force[16] site_xpos_feet[6] pad imu_xmat[9] pad
#define OUT_QACC 0
#define OUT_SENS 32
#define OUT_EFC 56
#define OUT_CDIST 152
#define OUT_AFRC 164
#define OUT_FEET 180
#define OUT_IMUMAT 188
#define DBG_STRIDE 5120

// flags of a dof lane
#define DF_TRANS 1     // free-joint translation
#define DF_ROT 2       // free-joint rotation (body-local axis)
#define DF_HINGE 4
#define DF_LIMITED 8
#define DF_FLOSS 16

struct alignas(16) DevModel {   // 16-byte multiple: staged into shared memory with one TMA bulk copy per CTA
  int nbody, njnt, nq, nv, nu, maxdepth, prefix_rounds, nfr, nlim, nvert, enable_ff;
  int imu_body, foot_body[2], foot_site_body[2], foot_chain[2];  // foot_chain: dof bitmask of the foot's ancestor chain
  int imu_chain;
  int iterations, ls_iterations;
  float imu_pos[3], imu_rot[9], foot_site_pos[2][3];
  float timestep, gravity[3], tolerance, ls_tolerance, meaninertia, impratio;
  float sol_k, sol_b, dmin, dmax, width, mid, power;
  float floor_mu, foot_mu;
  // body tables (lane = body)
  int b_parent[NLANE], b_depth[NLANE], b_jtype[NLANE], b_qadr0[NLANE], b_qadr1[NLANE], b_dofadr[NLANE], b_lastdof[NLANE];
  int b_submask[NLANE];
  float b_pos[3][NLANE], b_quat[4][NLANE], b_ipos[3][NLANE], b_Ib[6][NLANE], b_mass[NLANE], b_ax0[3][NLANE], b_ax1[3][NLANE], b_invw0[NLANE];
  // dof tables (lane = dof)
  int d_body[NLANE], d_parent[NLANE], d_vparent[NLANE], d_flags[NLANE], d_ancmask[NLANE], d_bsubmask[NLANE], d_qadr[NLANE], d_act[NLANE];
  int d_frrow[NLANE], d_limrow[NLANE];
  float d_damping[NLANE], d_invw0[NLANE], d_lo[NLANE], d_hi[NLANE], d_Dfric[NLANE], d_floss[NLANE], d_arm[NLANE];
  float d_kv[NLANE], d_kp[NLANE], d_clo[NLANE], d_chi[NLANE], d_flo[NLANE], d_fhi[NLANE];
  float vert[2][3][NLANE];
  float key_qpos[36], key_ctrl[16], qpos0[36];
  int act_dof[16], act_qadr[16], act_bl_qadr[16];
  // factorisation tables: anc[k] = dof ancestors of k, root first (anc[k][l] is the ancestor at depth l); d_depth[k] = their count
  unsigned char anc[NLANE][NLANE];
  int d_depth[NLANE];
  int max_dof_depth;
  unsigned short pair_ab[512];   // p -> (a << 8 | b), b <= a, p = a (a + 1) / 2 + b
  // chain plan of the factorisation: root chain of plan_nbase dofs + pure chains attached to its last dof
  int plan_ok, plan_nbase, plan_pair_len, plan_pair_start[2], plan_single_len, plan_single_start;
  unsigned int chol_tab[1536];    // per pivot k, per ancestor pair: target | src_a << 10 | src_b << 20 (offsets into the packed triangle)
  int chol_ofs[NLANE + 1];
  unsigned short mpair[512];     // structural non-zeros of M: (i << 8 | j), j an ancestor of i or i itself
  int n_mpairs, body_rounds;
  int b_sameaxis[NLANE];
  // subtree sums by chain scan: b_next = the only child of a body (-1: leaf or branching), scan_rounds = log2 of the longest
  // single-child chain; branching bodies deepest first with their children and the bodies of the chain that ends in them
  int scan_ok, scan_rounds, n_branch;
  int b_next[NLANE];
  int br_nchild[4], br_child[4][4], br_chain[4];
};

// per-warp shared memory
struct WarpSmem {
  float A[528];               // M, packed lower
  float H[528];               // chol(M), then H = M + J^T D J and its factor.  Its first 256 floats double as the staging records
                              // of the M pair pass (crb[body_i] * cdof_i (6) | armature | - per dof): H is not live before M exists
  float xpos[3][NLANE];
  float xmat[9][NLANE];
  float cdof[NLANE][8];       // motion axis per dof, one 32-byte record: ang (3), lin (3), - , -  (read back as two float4)
  float qpos[36];
  float qpos0[36];
  // contact records, read back as float4 broadcasts: [0] dist, [1..3] pos | [4..6] force (n, t1, t2) | [8..12] Hessian
  // weights (nn, n1, n2, 11, 22) of the contact in its own frame
  float con[NCON_ALL][CON_STRIDE];
  float misc[64];
  float rhs[NLANE];            // right-hand side swept leaf-to-root inside chol_rev
  float rowbuf[NLANE];         // pivot rows of the register-blocked factorisation (one 16-float row per half-warp), float4-aligned
  float outrec[OUT_STRIDE];     // staged outputs of the last forward (copied to HBM once per launch)
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
// Sums of K lane-local values over the warp with a folding butterfly: at every level a lane keeps the values whose index
// bit matches its lane bit and sends the others, so K values cost about K + 3 shuffles instead of 5 K.  Every lane ends up
// with the total of value number (lane & (P - 1)), P = K rounded up to a power of two; read value k with wfold_get(t, k).
template <int N, int BIT>
struct WFold {
  static __device__ __forceinline__ float run(float* v, const int lane) {
    if constexpr (BIT >= 32) {
      return v[0];
    } else if constexpr (N == 1) {
      v[0] += __shfl_xor_sync(FULLMASK, v[0], BIT);
      return WFold<1, BIT * 2>::run(v, lane);
    } else {
      const bool up = (lane & BIT) != 0;
#pragma unroll
      for (int k = 0; k < N / 2; ++k) {
        const float keep = up ? v[2 * k + 1] : v[2 * k];
        const float send = up ? v[2 * k] : v[2 * k + 1];
        v[k] = keep + __shfl_xor_sync(FULLMASK, send, BIT);
      }
      if constexpr (N & 1) {   // unpaired value: its partner is an implicit zero
        const float x = v[N - 1];
        v[N / 2] = (up ? 0.f : x) + __shfl_xor_sync(FULLMASK, up ? x : 0.f, BIT);
      }
      return WFold<(N + 1) / 2, BIT * 2>::run(v, lane);
    }
  }
};
template <int K>
__device__ __forceinline__ float wfold(float (&v)[K], const int lane) { return WFold<K, 1>::run(v, lane); }
__device__ __forceinline__ float wfold_get(const float t, const int k) { return __shfl_sync(FULLMASK, t, k); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Warp max on the redux unit (sm_100a: redux.sync.max.f32 -> CREDUX.MAX.F32, one instruction instead of a 5-step shuffle
// butterfly; NaN lanes are ignored).
__device__ __forceinline__ float wmaxf(float v) {
#ifdef ODUCK_WARP_EMU   // tests/emu: CPU emulation of one warp (test infrastructure)
  return emu_wmaxf(v);
#else
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
#endif
}
// argmax with first-index tie break (jp.argmax semantics); every lane gets the winner: max on the redux unit, then the
// lowest lane that holds it (ballot + ffs) -- 2 instructions on the shuffle/vote path instead of 10 shuffles
__device__ __forceinline__ int wargmax(float v, int lane) {
  const float m = wmaxf(v);
  const unsigned b = __ballot_sync(FULLMASK, v == m);
  return b ? __ffs(b) - 1 : 0;
}
__device__ __forceinline__ float wargmax_val(float v, int lane, int* out_idx) {
  const float m = wmaxf(v);
  const unsigned b = __ballot_sync(FULLMASK, v == m);
  *out_idx = b ? __ffs(b) - 1 : 0;
  return m;
}

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
struct Q4 { float w, x, y, z; };
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
  V3 u = v3(q.x, q.y, q.z);
  V3 t = 2.f * cross(u, v);
  return v + q.w * t + cross(u, t);
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n < 1e-15f) { q.w = 1.f; q.x = q.y = q.z = 0.f; return q; }
  float inv = 1.f / n;
  q.w *= inv; q.x *= inv; q.y *= inv; q.z *= inv;
  return q;
}
__device__ __forceinline__ Q4 axis_angle(V3 ax, float ang) {
  // half joint angles stay within (-pi, pi): the fast-path intrinsic (abs error 2^-21.4 there) without sincosf's out-of-line
  // argument reduction (~850 instructions of code per inlined call)
  float s, c;
  __sincosf(0.5f * ang, &s, &c);
  Q4 q; q.w = c; q.x = s * ax.x; q.y = s * ax.y; q.z = s * ax.z;
  return q;
}

struct S6 { float a0, a1, a2, l0, l1, l2; };   // spatial vector [angular, linear]
__device__ __forceinline__ S6 s6zero() { S6 r; r.a0 = r.a1 = r.a2 = r.l0 = r.l1 = r.l2 = 0.f; return r; }
__device__ __forceinline__ S6 s6add(S6 a, S6 b) { S6 r; r.a0 = a.a0 + b.a0; r.a1 = a.a1 + b.a1; r.a2 = a.a2 + b.a2; r.l0 = a.l0 + b.l0; r.l1 = a.l1 + b.l1; r.l2 = a.l2 + b.l2; return r; }
__device__ __forceinline__ S6 s6scale(S6 a, float s) { S6 r; r.a0 = a.a0 * s; r.a1 = a.a1 * s; r.a2 = a.a2 * s; r.l0 = a.l0 * s; r.l1 = a.l1 * s; r.l2 = a.l2 * s; return r; }
__device__ __forceinline__ float s6dot(S6 a, S6 b) { return a.a0 * b.a0 + a.a1 * b.a1 + a.a2 * b.a2 + a.l0 * b.l0 + a.l1 * b.l1 + a.l2 * b.l2; }
__device__ __forceinline__ S6 s6shfl(S6 a, int src) {
  S6 r;
  r.a0 = __shfl_sync(FULLMASK, a.a0, src); r.a1 = __shfl_sync(FULLMASK, a.a1, src); r.a2 = __shfl_sync(FULLMASK, a.a2, src);
  r.l0 = __shfl_sync(FULLMASK, a.l0, src); r.l1 = __shfl_sync(FULLMASK, a.l1, src); r.l2 = __shfl_sync(FULLMASK, a.l2, src);
  return r;
}
__device__ __forceinline__ S6 cross_motion(S6 v, S6 m) {   // mju_crossMotion
  V3 va = v3(v.a0, v.a1, v.a2), vl = v3(v.l0, v.l1, v.l2), ma = v3(m.a0, m.a1, m.a2), ml = v3(m.l0, m.l1, m.l2);
  V3 a = cross(va, ma), l = cross(va, ml) + cross(vl, ma);
  S6 r; r.a0 = a.x; r.a1 = a.y; r.a2 = a.z; r.l0 = l.x; r.l1 = l.y; r.l2 = l.z;
  return r;
}
__device__ __forceinline__ S6 cross_force(S6 v, S6 f) {    // mju_crossForce
  V3 va = v3(v.a0, v.a1, v.a2), vl = v3(v.l0, v.l1, v.l2), fa = v3(f.a0, f.a1, f.a2), fl = v3(f.l0, f.l1, f.l2);
  V3 a = cross(va, fa) + cross(vl, fl), l = cross(va, fl);
  S6 r; r.a0 = a.x; r.a1 = a.y; r.a2 = a.z; r.l0 = l.x; r.l1 = l.y; r.l2 = l.z;
  return r;
}
// 10-number spatial inertia about the common reference point (world axes)
struct I10 { float xx, yy, zz, xy, xz, yz, hx, hy, hz, m; };
__device__ __forceinline__ S6 inert_mul(const I10& I, S6 v) {  // mju_mulInertVec
  S6 r;
  r.a0 = I.xx * v.a0 + I.xy * v.a1 + I.xz * v.a2 - I.hz * v.l1 + I.hy * v.l2;
  r.a1 = I.xy * v.a0 + I.yy * v.a1 + I.yz * v.a2 + I.hz * v.l0 - I.hx * v.l2;
  r.a2 = I.xz * v.a0 + I.yz * v.a1 + I.zz * v.a2 - I.hy * v.l0 + I.hx * v.l1;
  r.l0 = I.hz * v.a1 - I.hy * v.a2 + I.m * v.l0;
  r.l1 = I.hx * v.a2 - I.hz * v.a0 + I.m * v.l1;
  r.l2 = I.hy * v.a0 - I.hx * v.a1 + I.m * v.l2;
  return r;
}

// ---------------------------------------------------------------------------------- dense kernels on packed triangles
// In-place leaf-to-root factorisation A = L^T L (L lower, stored where A was), fused with the leaf-to-root sweep
// rhs <- L^-T rhs.  For pivot k only rows/columns in S(k) are touched, S(k) = dof ancestors of k (tree == true: exact for
// the sparsity of a kinematic tree, no fill-in) or all j < k (dense fallback).  The |S|(|S|+1)/2 rank-1 updates of one
// pivot are independent, so they are spread over the warp 32 at a time (pair table), instead of one ancestor row per step.
static __device__ __noinline__ void chol_rev(const DevModel& m, float* A, float* rhs, int n, int lane, bool tree) {
  for (int k = n - 1; k >= 0; --k) {
    const int rk = TRI(k);
    const int d = tree ? m.d_depth[k] : k;
    const unsigned char* ak = m.anc[k];
    __syncwarp();                                  // rank-1 updates of the previous pivot are visible
    const float akk = A[rk + k];
    const float inv = rsqrtf(akk);
    const float yk = rhs[k] * inv;
    __syncwarp();                                  // everyone holds akk / rhs[k] before they are overwritten
    if (lane < d) {
      const int i = tree ? ak[lane] : lane;
      const float l = A[rk + i] * inv;
      A[rk + i] = l;
      rhs[i] -= l * yk;
    } else if (lane == d) {
      A[rk + k] = akk * inv;
      rhs[k] = yk;
    }
    __syncwarp();
    if (tree) {
      const int e0 = m.chol_ofs[k], e1 = m.chol_ofs[k + 1];
      for (int p = e0 + lane; p < e1; p += 32) {
        const unsigned e = m.chol_tab[p];
        A[e & 1023u] -= A[(e >> 10) & 1023u] * A[e >> 20];
      }
    } else {
      const int npairs = (d * (d + 1)) >> 1;
      for (int p = lane; p < npairs; p += 32) {
        const unsigned ab = m.pair_ab[p];
        const int a = ab >> 8, b = ab & 255;
        A[TRI(a) + b] -= A[rk + a] * A[rk + b];
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------- register-blocked chain factorisation
// The dof tree of the duck is a root chain (the 6 free-joint dofs) with pure chains (left leg, head, right leg) attached to
// its last dof.  Restricted to {root chain + one branch} the matrix is a dense lower triangle, so a branch is eliminated
// leaf-to-root entirely in registers: lane = column (half-warp local), a[i] = entry (row i, own column); the rank-1 update of
// pivot k is k shuffles + k FMAs.  The two equal-length legs run simultaneously in the two half-warps.  Each branch leaves
// its Schur-complement contribution to the root block (accumulated from zero) in a[0..5]; the root block is finished last.
// Same arithmetic as chol_rev(tree = true) up to summation order; rhs is swept leaf-to-root alongside.
#define CH_NB 6
// The row of the pivot (L[k][0..k-1], one value per column lane) is handed to all columns through a 16-float shared-memory
// row per half-warp and read back as float4 broadcasts: 1 store + ceil(k/4) loads on the shuffle/shared-memory pipe instead
// of k shuffles (that pipe, not the FP32 pipes, bounds the step kernel).
template <int NLOC>
__device__ __forceinline__ void branch_elim(float* H, float* rhs, const int lane, const int gstart, const bool store, float (&a)[NLOC], float& r, float* rowbuf) {
  const int j = lane & 15, hb = lane & 16;
  const int gj = j < CH_NB ? j : gstart + j - CH_NB;
  const bool col = j < NLOC;
#pragma unroll
  for (int i = 0; i < NLOC; ++i) {
    const int gi = i < CH_NB ? i : gstart + i - CH_NB;
    a[i] = (i >= CH_NB && j <= i && col) ? H[TRI(gi) + gj] : 0.f;
  }
  r = (j >= CH_NB && col) ? rhs[gj] : 0.f;
  // Square-root-free form (measured on B200: +0.4 % env-steps/s over the Cholesky form, profiles/r02b_bench_n1_ldl.json).  The pivot row stays unnormalised (u_kj, u_kk = d_k), so
  // its hand-over through shared memory no longer waits for the pivot's broadcast and reciprocal root: per pivot the chain
  // SHFL -> RSQ -> MUL -> STS -> LDS -> FMA becomes max(SHFL -> RCP -> MUL, STS -> LDS) -> FMA.  H then holds U with
  // M = U^T D^-1 U and rhs the unscaled sweep r = sqrt(D) y; chol_rev_back reads both forms with the same code
  // (x_k = (r_k - sum_j u_kj x_j) / u_kk), and nothing else consumes the factor.
#pragma unroll
  for (int k = NLOC - 1; k >= CH_NB; --k) {
    const float uk = j < k ? a[k] : 0.f;                       // u_kj, final
    __syncwarp();                                               // the previous pivot's row has been read by every lane
    rowbuf[hb | j] = uk;
    const float invd = 1.f / __shfl_sync(FULLMASK, a[k], hb | k);
    const float w = uk * invd;                                  // u_kj / d_k, zero for j >= k
    r = fmaf(-w, __shfl_sync(FULLMASK, r, hb | k), r);
    __syncwarp();
#pragma unroll
    for (int i4 = 0; i4 < (k + 3) / 4; ++i4) {                  // A[i][j] -= u_ki u_kj / d_k
      const float4 v = lds4(rowbuf + hb + 4 * i4);
      a[4 * i4] = fmaf(-v.x, w, a[4 * i4]);
      if (4 * i4 + 1 < NLOC) a[4 * i4 + 1] = fmaf(-v.y, w, a[4 * i4 + 1]);
      if (4 * i4 + 2 < NLOC) a[4 * i4 + 2] = fmaf(-v.z, w, a[4 * i4 + 2]);
      if (4 * i4 + 3 < NLOC) a[4 * i4 + 3] = fmaf(-v.w, w, a[4 * i4 + 3]);
    }
  }
  if (store && col) {
#pragma unroll
    for (int i = CH_NB; i < NLOC; ++i)
      if (j <= i) H[TRI(gstart + i - CH_NB) + gj] = a[i];
    if (j >= CH_NB) rhs[gj] = r;
  }
}

template <int NPAIR, int NSINGLE>   // branch lengths
static __device__ __noinline__ void chol_rev_chain(const DevModel& m, float* H, float* rhs, const int lane, float* rowbuf) {
  const int j = lane & 15;
  float db[CH_NB], dr;
  {
    float a[CH_NB + NPAIR], r;
    branch_elim<CH_NB + NPAIR>(H, rhs, lane, m.plan_pair_start[lane >> 4], true, a, r, rowbuf);
#pragma unroll
    for (int i = 0; i < CH_NB; ++i) db[i] = a[i] + __shfl_xor_sync(FULLMASK, a[i], 16);
    dr = r + __shfl_xor_sync(FULLMASK, r, 16);
  }
  if (NSINGLE > 0) {
    float a[CH_NB + (NSINGLE > 0 ? NSINGLE : 1)], r;
    branch_elim<CH_NB + (NSINGLE > 0 ? NSINGLE : 1)>(H, rhs, lane, m.plan_single_start, lane < 16, a, r, rowbuf);   // both halves compute, half 0 stores
#pragma unroll
    for (int i = 0; i < CH_NB; ++i) db[i] += a[i];
    dr += r;
  }
  __syncwarp();
  // root block (pivots 5 .. 0): true entries + the branches' Schur complements
  float b[CH_NB], rb = 0.f;
#pragma unroll
  for (int i = 0; i < CH_NB; ++i) b[i] = (j <= i && j < CH_NB) ? H[TRI(i) + j] + db[i] : 0.f;
  if (j < CH_NB) rb = rhs[j] + dr;
  const int hb = lane & 16;
#pragma unroll
  for (int k = CH_NB - 1; k >= 0; --k) {
    const float uk = j < k ? b[k] : 0.f;
    const float w = uk * (1.f / __shfl_sync(FULLMASK, b[k], hb | k));
    rb = fmaf(-w, __shfl_sync(FULLMASK, rb, hb | k), rb);
#pragma unroll
    for (int i = 0; i < k; ++i) b[i] = fmaf(-__shfl_sync(FULLMASK, uk, hb | i), w, b[i]);   // the row's hand-over does not wait for the reciprocal
  }
  __syncwarp();                          // both half-warps have read the root block (lanes 16..21 mirror lanes 0..5) before it is overwritten:
                                         // the shuffles above converge the warp but are no memory barrier (compute-sanitizer racecheck, r02b)
  if (lane < CH_NB) {
#pragma unroll
    for (int i = 0; i < CH_NB; ++i)
      if (j <= i) H[TRI(i) + j] = b[i];
    rhs[j] = rb;
  }
  __syncwarp();
}
// factorise + sweep: register-blocked chains when the model matches a compiled plan, else the generic pair-table version
__device__ __forceinline__ void chol_rev_tree(const DevModel& m, float* H, float* rhs, int n, int lane, float* rowbuf) {
  if (m.plan_ok == 1) chol_rev_chain<10, 4>(m, H, rhs, lane, rowbuf);
  else if (m.plan_ok == 2) chol_rev_chain<5, 4>(m, H, rhs, lane, rowbuf);
  else chol_rev(m, H, rhs, n, lane, true);
}

// Root-to-leaf sweep x <- L^-1 y with y one element per lane (y = rhs after chol_rev).  tree: all dofs of one depth level
// are finished together (their ancestors are done), so the dependency chain is max_dof_depth long instead of n.
static __device__ __noinline__ float chol_rev_back(const DevModel& m, const float* L, int n, int lane, float y, bool tree) {
  if (tree) {
    const int dep = lane < n ? m.d_depth[lane] : -1;
    const float invd = lane < n ? 1.f / L[TRI(lane) + lane] : 0.f;
    const unsigned char* al = m.anc[lane];
    const int ri = TRI(lane);
    for (int lev = 0; lev <= m.max_dof_depth; ++lev) {
      if (dep == lev) y *= invd;
      const int j = dep > lev ? al[lev] : lane;
      const float xj = __shfl_sync(FULLMASK, y, j);
      if (dep > lev) y -= L[ri + j] * xj;
    }
    return y;
  }
  for (int j = 0; j < n; ++j) {
    float xj = __shfl_sync(FULLMASK, y, j) / L[TRI(j) + j];
    if (lane == j) y = xj;
    else if (lane > j && lane < n) y -= L[TRI(lane) + j] * xj;
  }
  return y;
}
// chol_rev_back for the chain plan (root chain of CH_NB dofs + pure chains attached to its last dof): the ancestor of a dof at
// level lev is lev itself on the root chain and (first dof of the lane's chain) + lev - CH_NB below it, so the sweep needs no
// ancestor table -- the byte reads of m.anc were an 8-way shared-memory bank conflict per level (lanes 32 B apart) -- and the
// fully unrolled loop lets the factor entries be fetched ahead of the dependent shuffle / FMA chain.  Same arithmetic, same order.
template <int NCHAIN>
static __device__ __noinline__ float chol_rev_back_chain(const DevModel& m, const float* L, int n, int lane, float y) {
  const int dep = lane < n ? m.d_depth[lane] : -1;
  const float* Lr = L + TRI(lane);
  const float invd = lane < n ? 1.f / Lr[lane] : 0.f;
  const int cs = lane - (dep - CH_NB);                          // first dof of this lane's chain (meaningful for dep >= CH_NB)
#pragma unroll
  for (int lev = 0; lev < CH_NB + NCHAIN; ++lev) {
    if (dep == lev) y *= invd;
    const int j = lev < CH_NB ? lev : cs + (lev - CH_NB);
    const bool below = dep > lev;
    const float l = below ? Lr[j] : 0.f;
    const float xj = __shfl_sync(FULLMASK, y, below ? j : lane);
    y = fmaf(-l, xj, y);
  }
  return y;
}
// y = A x for the packed symmetric matrix; one element per lane.  x is handed to all lanes through a 32-float shared-memory
// row read back as float4 broadcasts (8 loads instead of 30 shuffles).
static __device__ __noinline__ float symv(const float* A, int n, int lane, float x, float* xbuf) {
  const int ri = TRI(lane);
  __syncwarp();
  xbuf[lane] = lane < n ? x : 0.f;
  __syncwarp();
  // The column loop is fully unrolled and branch-free on four independent FMA chains: j is a compile-time constant, so TRI(j)
  // and both address forms fold into LDS immediates (a rolled loop spends ~14 issue slots per element on index arithmetic and
  // divergence bookkeeping for the j < n test: measured on B200 at 4096 envs, k_step 0.666 -> 0.621 ms, profiles/r02a_*; a
  // rolled four-chain loop gained nothing, 0.667 ms).  No per-element bound test: columns
  // n .. 4 ceil(n / 4) - 1 multiply x = 0 (xbuf is zero-padded) with the zero padding rows of A (load_env clears all 528
  // floats and only ancestor pairs are rewritten), both addresses stay inside the 528-float triangle for every lane.
  float acc4[4] = {0.f, 0.f, 0.f, 0.f};
  const float* Arow = A + ri;
  const float* Acol = A + lane;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    if (4 * j4 >= n) break;                                    // warp-uniform
    const float4 xv = lds4(xbuf + 4 * j4);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * j4 + e;
      const float* pj = (j <= lane) ? Arow + j : Acol + TRI(j);
      acc4[e] = fmaf(*pj, xs[e], acc4[e]);
    }
  }
  const float acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
  __syncwarp();
  return lane < n ? acc : 0.f;
}
// constraint.py _kbi impedance for a signed distance.  The general-power branch (two powf expansions, ~1.5 k instructions each
// time it is inlined) is never taken by the reference scenes (solimp power = 2): it lives out of line so that the substep's hot
// instruction stream, which already exceeds the instruction cache, does not carry it twice.
static __device__ __noinline__ float impedance_general_power(float x, float mid, float power) {
  const float ia = (1.f / powf(mid, power - 1.f)) * powf(x, power);
  const float ib = 1.f - (1.f / powf(1.f - mid, power - 1.f)) * powf(1.f - x, power);
  return x < mid ? ia : ib;
}
__device__ __forceinline__ float impedance(const DevModel& m, float pos) {
  float x = fabsf(pos) / m.width;
  float y;
  if (m.power == 2.f) y = x < m.mid ? x * x / m.mid : 1.f - (1.f - x) * (1.f - x) / (1.f - m.mid);
  else y = impedance_general_power(x, m.mid, m.power);
  float imp = m.dmin + y * (m.dmax - m.dmin);
  imp = fminf(fmaxf(imp, m.dmin), m.dmax);
  if (x > 1.f) imp = m.dmax;
  return imp;
}
