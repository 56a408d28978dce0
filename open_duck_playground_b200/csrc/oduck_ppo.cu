// oduck_ppo.cu -- the PPO learner step on the device (include/oduck_ppo.h).
// Restates what Brax's `ppo.train` does per minibatch for the reference's runner (playground/common/runner.py:86-118):
// compute_ppo_loss (losses.py: value/policy forward, compute_gae, clipped surrogate, sampled NormalTanh entropy),
// its gradient, optax.clip_by_global_norm + optax.adam.  Checker: the PyTorch fp32 twin (ppo.py) in tests/test_ppo_device.py.
//
// Launch sequence of one minibatch (B trajectories x T steps; Mp = T B policy rows, Mv = (T + 1) B value rows):
//   k_ppo_pack x2        gather rows by env index, normalise, write R(X0) and R(X0^T) (hi/lo tf32 operand blocks)
//   k_gemm_tc  x8        forward: 3 x (Dense + swish) + head, per net; epilogues emit the next operands in both orientations
//   k_ppo_gae            per-trajectory GAE scan, advantage statistics (last-CTA reduction in CTA order: deterministic)
//   k_ppo_loss           per-row NormalTanh log-prob / ratio / clip / entropy / value error -> head gradients as operands
//   k_gemm_tc  x14       backward: dW_l = X_l^T dZ_l (split-K over the batch) and dZ_{l-1} = (dZ_l W_l^T) * swish'(Z_{l-1})
//   k_ppo_grad_reduce    sum split-K / per-warp partials into the flat gradient, partial sums of squares
//   k_ppo_adam           global-norm clip, Adam, master weights + both packed operand forms of every kernel matrix
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/oduck.h"
#include "../../include/oduck_ppo.h"
#include "oduck_env.cuh"
#include "oduck_gemm_tc.cuh"
#include <cooperative_groups.h>

extern int oduck_fail(int code, const std::string& msg);
#define PPO_TRY(x)                                                                                             \
  do {                                                                                                         \
    cudaError_t e_ = (x);                                                                                      \
    if (e_ != cudaSuccess) return oduck_fail(ODUCK_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

#define PPO_NL 4            // dense layers per net
#define PPO_MAXT 64
#define PPO_HEADW 32        // padded head width (one NT = 32 tile)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// one tensor of the flat parameter vector
struct Seg {
  long long off;            // offset in the flat vector
  int net, layer, bias, K, N;   // kernel: K x N row-major ([in][out]); bias: N
  int nt;                   // forward B-operand tile height (128 hidden, 32 head)
  long long wf, wb;         // offsets of the packed operands inside the pack buffer (wb < 0: not needed)
  long long dwpart, dbpart; // offsets inside the partial-gradient buffer
  int ldo, nsplit, ldb, nwarprows;
  long long split_stride;
};
#define PPO_NSEG (2 * PPO_NL * 2)
struct SegTable { Seg s[PPO_NSEG]; int n; long long total; };

struct NetBuf {
  int dims[PPO_NL + 1];
  int M, Mpad, mtiles;                  // rows of this net's batch
  float* Xr[PPO_NL];                    // R(X_l)   A operand of the forward GEMM of layer l
  float* Xt[PPO_NL];                    // R(X_l^T) A operand of the dW GEMM of layer l
  float* Z[PPO_NL - 1];                 // pre-activations, blocked [mtiles][d_{l+1} / 32][128 x 32]
  float* out;                           // head output [Mpad][32]
  float* dZr[PPO_NL];                   // R(dZ_l)   A operand of the dX GEMM
  float* dZt[PPO_NL];                   // R(dZ_l^T) B operand of the dW GEMM
};

struct OduckPpo {
  OduckPpoConfig cfg;
  int device;
  int B, T, na;
  SegTable seg;
  SegTable* dseg;
  NetBuf net[2];
  float *params, *grads, *adam_m, *adam_v, *packed, *partial;
  long long P, packed_floats, partial_floats;
  float *adv, *vs, *adv_norm;
  double* losses;           // [8]
  double* stats;            // [4 + 2 * gae CTAs] adv mean, std | per-CTA partial sums
  float* sumsq_part;        // [grid of grad_reduce]
  int* step;                // [3]: adam step count, finished-block tickets of k_ppo_adam and k_ppo_gae
  int reduce_blocks;
  int64_t launches;
  bool pdl;                 // programmatic dependent launch along the kernel chain (oduck_gemm_tc.cuh): OFF by default, ODUCK_PPO_PDL=1 turns it on.
                            // Measured on B200 (profiles/r02g_bench_ppo_*_pdl{0,1}.json): the update takes 38.3 ms with it against 37.0 ms without
                            // (3xTF32) and 30.8 against 30.2 ms (one-pass TF32) -- inside a captured graph the kernel-to-kernel hand-over is
                            // already short, and early-launched CTAs hold SMs / TMEM that the other streams' kernels could use.
  cudaStream_t side;        // the value net's chain runs here, concurrently with the policy net's chain on the caller's stream
  cudaStream_t side_w[2];   // weight-gradient GEMMs of the policy / value net (off the dZ critical path)
  // input prefetch (oduck_ppo_prefetch): two sets of layer-0 operand buffers per net; x0_sel = the set the current minibatch reads
  float* X0r[2][2];         // [net][set]
  float* X0t[2][2];
  int x0_sel;
  const int32_t* pf_idx;    // env_idx pointer the other set was packed for (nullptr: nothing prefetched)
  cudaStream_t pf_stream;
  cudaEvent_t ev_mb_start, ev_pf_done;
  cudaEvent_t ev_fork, ev_join, ev_join_w[2], ev_dz[2][PPO_NL];
  int coop_blocks;          // co-resident CTAs of the fused reduce + Adam kernel (0: cooperative launch unavailable)
  int num_sms;
  std::vector<void*> allocs;
};

// ------------------------------------------------------------------------------------------------- kernels
// float offset of (time t, env e) in a rollout field whose rows are `width` floats wide (blocked layout: include/oduck_ppo.h)
__device__ __forceinline__ size_t ro_off(const int be, const long long bstride, const int t, const int e, const int width) {
  const int blk = e / be, el = e - blk * be;
  return (size_t)blk * (size_t)bstride + ((size_t)t * be + el) * width;
}
// Gather the minibatch rows (row = t * B + b <- env idx[b] at time t), normalise, emit R(X) and R(X^T).
// One thread per 4 x 4 micro-tile (4 rows x 4 features): both layouts keep 4 consecutive k (R(X)) / 4 consecutive rows
// (R(X^T)) contiguous, so every store is a float4.
__global__ void k_ppo_pack(const float* __restrict__ obs, int be, long long bstride, int ld, int K, const int* __restrict__ idx, int B, int M, int Mpad, int kch,
                           const float* __restrict__ mean, const float* __restrict__ stdv, float* __restrict__ Xr, float* __restrict__ Xt) {
  pdl_launch_dependents();
  pdl_wait();
  const int K4 = kch * (TC_KC / 4), ytn = Mpad / TC_KC;
  const int total = (Mpad / 4) * K4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kg = i % K4, rg = i / K4;                  // consecutive threads: consecutive feature groups of one row group (coalesced gathers)
    const int row0 = 4 * rg, k0 = 4 * kg;
    float v[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = row0 + a;
      const float* src = nullptr;
      if (row < M) { const int t = row / B, b = row - t * B; src = obs + ro_off(be, bstride, t, idx[b], ld); }
#pragma unroll
      for (int e = 0; e < 4; ++e) { const int k = k0 + e; v[a][e] = (src && k < K) ? (src[k] - mean[k]) / stdv[k] : 0.f; }
    }
    float* blk = Xr + ((size_t)(row0 >> 7) * kch + (k0 >> 5)) * GBLK_A;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float4 h4, l4;
      float* ph = reinterpret_cast<float*>(&h4);
      float* pl = reinterpret_cast<float*>(&l4);
#pragma unroll
      for (int e = 0; e < 4; ++e) gsplit_tf32(v[a][e], ph[e], pl[e]);
      const int off = gblk_off((row0 + a) & 127, k0 & 31);
      *reinterpret_cast<float4*>(blk + off) = h4;
      *reinterpret_cast<float4*>(blk + TC_M * TC_KC + off) = l4;
    }
    blk = Xt + ((size_t)(k0 >> 7) * ytn + (row0 >> 5)) * GBLK_A;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float4 h4, l4;
      float* ph = reinterpret_cast<float*>(&h4);
      float* pl = reinterpret_cast<float*>(&l4);
#pragma unroll
      for (int a = 0; a < 4; ++a) gsplit_tf32(v[a][e], ph[a], pl[a]);
      const int off = gblk_off((k0 + e) & 127, row0 & 31);
      *reinterpret_cast<float4*>(blk + off) = h4;
      *reinterpret_cast<float4*>(blk + TC_M * TC_KC + off) = l4;
    }
  }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += sh[w];
  return t;
}

// brax losses.compute_gae, one trajectory per thread, one warp per CTA (B / 32 CTAs).  The gathers of a trajectory are issued
// up front into a thread-local window (independent loads, the scan itself is a chain of T dependent steps).  Advantage
// statistics: every CTA leaves its partial sums, the last CTA to finish adds them in CTA order (deterministic) and publishes
// mean / std; k_ppo_loss normalises on the fly.
#define GAE_THREADS 32
__global__ void __launch_bounds__(GAE_THREADS) k_ppo_gae(const float* __restrict__ values /*[Mv_pad][32]*/, OduckRollout ro, const int* __restrict__ idx, int B, int T,
                                                         float discount, float lambda, float reward_scaling, float* __restrict__ adv, float* __restrict__ vs,
                                                         double* __restrict__ stats /*[4 + 2 * gridDim.x]*/, int* __restrict__ ticket, double* __restrict__ losses) {
  pdl_launch_dependents();
  pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x < 8) losses[threadIdx.x] = 0.0;
  const int be = ro.block_envs > 0 ? ro.block_envs : ro.num_envs;
  const int b = blockIdx.x * GAE_THREADS + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (b < B) {
    const int env = idx[b];
    float lv[PPO_MAXT + 1], lr[PPO_MAXT], ld[PPO_MAXT], lt[PPO_MAXT];
    for (int t = 0; t < T; ++t) {
      const size_t g = ro_off(be, ro.block_stride, t, env, 1);
      lv[t] = values[(size_t)(t * B + b) * PPO_HEADW];
      lr[t] = ro.reward[g]; ld[t] = ro.done[g]; lt[t] = ro.truncation[g];
    }
    lv[T] = values[(size_t)(T * B + b) * PPO_HEADW];
    float acc = 0.f, v_next = lv[T], vs_next = lv[T];
    for (int t = T - 1; t >= 0; --t) {
      const float trunc = lt[t], done = ld[t], rew = lr[t] * reward_scaling;
      const float term = done * (1.f - trunc), mask = 1.f - trunc;
      const float v = lv[t];
      const float delta = (rew + discount * (1.f - term) * v_next - v) * mask;
      acc = delta + discount * (1.f - term) * mask * lambda * acc;
      const float vst = acc + v;
      const float a = (rew + discount * (1.f - term) * vs_next - v) * mask;
      vs[t * B + b] = vst;
      adv[t * B + b] = a;
      s1 += a; s2 += (double)a * a;
      v_next = v; vs_next = vst;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if (threadIdx.x == 0) {
    stats[4 + 2 * blockIdx.x] = s1; stats[5 + 2 * blockIdx.x] = s2;
    __threadfence();
    if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
      __threadfence();
      double S1 = 0.0, S2 = 0.0;
      for (int k = 0; k < (int)gridDim.x; ++k) { S1 += ((volatile double*)stats)[4 + 2 * k]; S2 += ((volatile double*)stats)[5 + 2 * k]; }
      const double n = (double)B * T, mean = S1 / n, var = fmax(S2 / n - mean * mean, 0.0);
      stats[0] = mean; stats[1] = sqrt(var);
      *ticket = 0;
    }
  }
}

__device__ __forceinline__ float ppo_erfinv(float x) {   // XLA's f32 erf_inv polynomial (as in oduck_policy.cu)
  float w = -__logf((1.0f - x) * (1.0f + x)), p;
  if (w < 5.0f) {
    w -= 2.5f;
    p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w; p = 0.00021858087f + p * w;
    p = -0.00125372503f + p * w; p = -0.00417768164f + p * w; p = 0.246640727f + p * w; p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w; p = 0.00573950773f + p * w;
    p = -0.0076224613f + p * w; p = 0.00943887047f + p * w; p = 1.00167406f + p * w; p = 2.83297682f + p * w;
  }
  return p * x;
}
__device__ __forceinline__ float softplusf(float x) { return x > 20.f ? x : log1pf(expf(x)); }

struct LossParams {
  const float* logits;      // [Mp_pad][32]
  const float* values;      // [Mv_pad][32]
  const float* adv; const float* vs;
  const double* stats;      // [0] mean, [1] std of the raw advantages (k_ppo_gae)
  float* adv_norm;          // [Mp] normalised advantages (published for the parity tests)
  int normalize;
  OduckRollout ro;
  const int* idx;
  const float* noise;       // [Mp][na] or null
  const uint32_t* key;      // device u32[2] (used when noise == null)
  int B, T, na, Mp, Mp_pad, Mv, Mv_pad;
  float clip_eps, entropy_cost;
  float *dZr_p, *dZt_p, *db_p;   // policy head gradient: R(dZ) [Mp_pad/128][1], Rb32(dZ^T) [1][Mp_pad/32], column sums [Mp_pad/32][32]
  float *dZr_v, *dZt_v, *db_v;
  double* losses;
};

// Sixteen lanes per row (lane a < na = one action), 16 rows per CTA.  Policy rows r < Mp (the value baseline rows are the same
// indices); bootstrap rows Mp <= r < Mv get zero gradient.  Outputs are the head gradients already in operand form:
// R(dZ) (A operand of the dX GEMM), R_32(dZ^T) (B operand of the dW GEMM) and one bias-gradient partial per CTA.
#define LOSS_ROWS 16
__global__ void __launch_bounds__(LOSS_ROWS * 16) k_ppo_loss(LossParams p) {
  __shared__ double sh[32];
  __shared__ float gp[LOSS_ROWS][PPO_HEADW], gv[LOSS_ROWS];
  pdl_launch_dependents();
  pdl_wait();
  const int sub = threadIdx.x >> 4, a = threadIdx.x & 15;        // row slot in the CTA, action lane
  const int r = blockIdx.x * LOSS_ROWS + sub;
  double l_pol = 0.0, l_val = 0.0, l_ent = 0.0, l_clip = 0.0, l_adv = 0.0;
  const float invM = 1.f / (float)p.Mp;
  const bool prow = r < p.Mp_pad;                                // CTA-uniform (Mp_pad is a multiple of 128)
  // ---------------- policy head
  float g_loc = 0.f, g_sc = 0.f;
  if (r < p.Mp) {
    const int t = r / p.B, b = r - t * p.B;
    const int be = p.ro.block_envs > 0 ? p.ro.block_envs : p.ro.num_envs;
    const size_t gi = ro_off(be, p.ro.block_stride, t, p.idx[b], 1);
    const size_t ga = ro_off(be, p.ro.block_stride, t, p.idx[b], p.na);
    const float* lg = p.logits + (size_t)r * PPO_HEADW;
    float A = p.adv[r];
    if (p.normalize) A = (A - (float)p.stats[0]) / ((float)p.stats[1] + 1e-8f);
    if (a == 0) p.adv_norm[r] = A;
    float logp = 0.f, ent = 0.f, dlp_loc = 0.f, dlp_sc = 0.f, dent_loc = 0.f, dent_sc = 0.f, sig = 0.f;
    if (a < p.na) {
      const float loc = lg[a], sp = lg[p.na + a];
      const float scale = softplusf(sp) + 0.001f;
      const float raw = p.ro.raw_action[ga + a];
      const float zz = (raw - loc) / scale;
      logp = -0.5f * zz * zz - logf(scale) - 0.91893853320467f - 2.f * (0.69314718056f - raw - softplusf(-2.f * raw));
      float eps;
      if (p.noise) eps = p.noise[(size_t)r * p.na + a];
      else {
        RKey k0; k0.a = p.key[0]; k0.b = p.key[1];
        const RKey bk = rblock(rblock(k0, (uint32_t)r), (uint32_t)a);
        const float lo = -0.99999994f;
        eps = 1.41421356237f * ppo_erfinv(fmaxf(lo, bits_unit(bk.a ^ bk.b) * (1.0f - lo) + lo));
      }
      const float x = loc + scale * eps;
      ent = 0.5f + 0.91893853320467f + logf(scale) + 2.f * (0.69314718056f - x - softplusf(-2.f * x));
      const float th = tanhf(x);
      dlp_loc = zz / scale; dlp_sc = (zz * zz - 1.f) / scale;
      dent_loc = -2.f * th; dent_sc = 1.f / scale - 2.f * th * eps;
      sig = 1.f / (1.f + expf(-sp));
    }
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);   // this row's half-warp (the other half may be a row >= Mp)
#pragma unroll
    for (int o = 8; o; o >>= 1) { logp += __shfl_xor_sync(hmask, logp, o); ent += __shfl_xor_sync(hmask, ent, o); }
    const float rho = expf(logp - p.ro.log_prob[gi]);
    const float lo = 1.f - p.clip_eps, hi = 1.f + p.clip_eps;
    const float s1 = rho * A, s2 = fminf(fmaxf(rho, lo), hi) * A;
    const bool inr = rho >= lo && rho <= hi;
    const float w = s1 < s2 ? 1.f : (s1 > s2 ? (inr ? 1.f : 0.f) : 0.5f + (inr ? 0.5f : 0.f));   // ties split like jnp / torch minimum
    const float glp = -A * rho * w * invM;              // d loss / d logp
    const float ge = -p.entropy_cost * invM;            // d loss / d entropy_row
    g_loc = glp * dlp_loc + ge * dent_loc;
    g_sc = (glp * dlp_sc + ge * dent_sc) * sig;
    if (a == 0) { l_pol = -(double)fminf(s1, s2); l_ent = ent; l_clip = inr ? 0.0 : 1.0; l_adv = fabsf(A); }
  }
  if (prow) {
    // lane c writes columns c and c + 16 of the padded head: [loc gradients | scale gradients | zeros]
    const int base = threadIdx.x & 16;                    // first lane of this row's half-warp
    float* blk = p.dZr_p + (size_t)(r >> 7) * GBLK_A;
    float* bt = p.dZt_p + (size_t)(r >> 5) * gblk_b(PPO_HEADW);
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int col = a + 16 * h2;
      const int src = col < p.na ? col : col - p.na;                       // lane holding this column's gradient
      const float vl = __shfl_sync(0xffffffffu, g_loc, base | (src & 15)), vs2 = __shfl_sync(0xffffffffu, g_sc, base | (src & 15));
      const float g = col < p.na ? vl : (col < 2 * p.na ? vs2 : 0.f);
      float hi, lo;
      gsplit_tf32(g, hi, lo);
      int off = gblk_off(r & 127, col);
      blk[off] = hi; blk[TC_M * TC_KC + off] = lo;
      off = gblk_off(col, r & 31);
      bt[off] = hi; bt[PPO_HEADW * TC_KC + off] = lo;
      gp[sub][col] = g;
    }
  }
  // ---------------- value head: v_loss = 0.5 * 0.5 * mean((vs - v)^2)
  {
    float dv = 0.f;
    if (r < p.Mp && a == 0) {
      const float err = p.vs[r] - p.values[(size_t)r * PPO_HEADW];
      dv = -0.5f * err * invM;
      l_val = 0.25 * (double)err * err;
    }
    float hi, lo;
    gsplit_tf32(dv, hi, lo);
    float* blk = p.dZr_v + (size_t)(r >> 7) * GBLK_A;
    float* bt = p.dZt_v + (size_t)(r >> 5) * gblk_b(PPO_HEADW);
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int col = a + 16 * h2;
      const float xh = col == 0 ? hi : 0.f, xl = col == 0 ? lo : 0.f;
      int off = gblk_off(r & 127, col);
      blk[off] = xh; blk[TC_M * TC_KC + off] = xl;
      off = gblk_off(col, r & 31);
      bt[off] = xh; bt[PPO_HEADW * TC_KC + off] = xl;
    }
    if (a == 0) gv[sub] = dv;
  }
  __syncthreads();
  if (threadIdx.x < PPO_HEADW) {                          // bias-gradient partials of this CTA's 16 rows
    const int c = threadIdx.x;
    if (prow) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < LOSS_ROWS; ++q) t += gp[q][c];
      p.db_p[(size_t)blockIdx.x * PPO_HEADW + c] = t;
    }
    float t = 0.f;
    if (c == 0) {
#pragma unroll
      for (int q = 0; q < LOSS_ROWS; ++q) t += gv[q];
    }
    p.db_v[(size_t)blockIdx.x * PPO_HEADW + c] = t;
  }
  const double inv = 1.0 / (double)p.Mp;
  const double a0 = block_sum(l_pol, sh), a1 = block_sum(l_val, sh), a2 = block_sum(l_ent, sh), a3 = block_sum(l_clip, sh), a4 = block_sum(l_adv, sh);
  if (threadIdx.x == 0) {
    atomicAdd(p.losses + 1, a0 * inv); atomicAdd(p.losses + 2, a1 * inv); atomicAdd(p.losses + 3, a2 * inv);
    atomicAdd(p.losses + 0, (a0 + a1 - (double)p.entropy_cost * a2) * inv);
    atomicAdd(p.losses + 4, a4 * inv); atomicAdd(p.losses + 5, a3 * inv);
  }
}

__device__ __forceinline__ const Seg& find_seg(const SegTable& tb, long long i) {
  int k = 0;
#pragma unroll 1
  for (int s = 1; s < tb.n; ++s) if (i >= tb.s[s].off) k = s;
  return tb.s[k];
}

// gradient of flat element i: sum of its split-K partials (kernels) or per-CTA column sums (biases); four independent
// accumulators keep four loads in flight per thread
__device__ __forceinline__ float reduce_partials(const Seg& s, const float* __restrict__ partial, long long jj) {
  const float* q;
  long long stride;
  int cnt;
  const unsigned j = (unsigned)jj;                             // offsets inside a tensor fit 32 bits: 32-bit division (a 64-bit one costs ~100 instructions)
  if (!s.bias) { const unsigned k = j / (unsigned)s.N, n = j - k * (unsigned)s.N; q = partial + s.dwpart + (size_t)k * s.ldo + n; stride = s.split_stride; cnt = s.nsplit; }
  else { q = partial + s.dbpart + j; stride = s.ldb; cnt = s.nwarprows; }
  // even splits into g0, odd splits into g1, eight loads in flight: the same summation order as reduce_quad
  float g0 = 0.f, g1 = 0.f;
  int z = 0;
  for (; z + 8 <= cnt; z += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(q + (size_t)(z + u) * stride);
#pragma unroll
    for (int u = 0; u < 8; u += 2) { g0 += v[u]; g1 += v[u + 1]; }
  }
  for (; z < cnt; ++z) g0 += __ldcg(q + (size_t)z * stride);
  return g0 + g1;
}

__device__ __forceinline__ void adam_element(const Seg& s, long long i, float g, float scale, float c1, float c2, float lr, float b1, float b2, float eps,
                                             float* __restrict__ params, float* __restrict__ m1, float* __restrict__ m2, float* __restrict__ packed, bool update) {
  float w = params[i];
  if (update) {
    g *= scale;
    const float mm = b1 * m1[i] + (1.f - b1) * g, vv = b2 * m2[i] + (1.f - b2) * g * g;
    m1[i] = mm; m2[i] = vv;
    w -= lr * (mm * c1) / (sqrtf(vv * c2) + eps);
    params[i] = w;
  }
  if (s.bias) return;
  const unsigned j = (unsigned)(i - s.off);
  const int k = (int)(j / (unsigned)s.N), n = (int)(j - (unsigned)k * (unsigned)s.N);
  float hi, lo;
  gsplit_tf32(w, hi, lo);
  {
    // forward B operand: R_nt(W^T), rows = out feature n, contraction over the in feature k (nt is 128 or 32)
    const int kch = (s.K + TC_KC - 1) / TC_KC;
    const int ntile = s.nt == 128 ? n >> 7 : n >> 5, nrow = n & (s.nt - 1);
    float* blk = packed + s.wf + ((size_t)ntile * kch + (k >> 5)) * gblk_b(s.nt);
    const int off = gblk_off(nrow, k & 31);
    blk[off] = hi; blk[s.nt * TC_KC + off] = lo;
  }
  if (s.wb >= 0) {
    // dX B operand: R(W), rows = in feature k, contraction over the out feature n
    const int nch = (s.N + TC_KC - 1) / TC_KC;
    float* blk = packed + s.wb + ((size_t)(k >> 7) * nch + (n >> 5)) * GBLK_A;
    const int off = gblk_off(k & 127, n & 31);
    blk[off] = hi; blk[TC_M * TC_KC + off] = lo;
  }
}

// sum of the per-block partial sums of squares, by one warp in a fixed order (deterministic); every lane gets the total
__device__ __forceinline__ double warp_total(const float* __restrict__ part, int n, int lane) {
  double t = 0.0;
  for (int k = lane; k < n; k += 32) t += (double)__ldcg(part + k);
#pragma unroll
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// Bias gradients: element j of a bias tensor is the sum of one partial per producer row block (168 - 336 of them); one warp
// per element, partial rows spread over the lanes, shuffle tree in a fixed order (deterministic).  A thread-per-element loop
// would chain up to 42 dependent load batches on the few threads that own a bias element.
__device__ __forceinline__ void reduce_biases(const SegTable& tb, const float* __restrict__ partial, float* __restrict__ grads, double& ss) {
  const int lane = threadIdx.x & 31;
  const int wid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nw = (int)((gridDim.x * blockDim.x) >> 5);
  int total = 0;
  for (int k = 0; k < tb.n; ++k) if (tb.s[k].bias) total += tb.s[k].N;
  for (int b = wid; b < total; b += nw) {
    int k = 0, j = b;
    for (;; ++k) { if (!tb.s[k].bias) continue; if (j < tb.s[k].N) break; j -= tb.s[k].N; }
    const Seg& sg = tb.s[k];
    const float* q = partial + sg.dbpart + j;
    float g = 0.f;
    for (int w = lane; w < sg.nwarprows; w += 32) g += __ldcg(q + (size_t)w * sg.ldb);
#pragma unroll
    for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
    if (lane == 0) { grads[sg.off + j] = g; ss += (double)g * g; }
  }
}

// Single-GPU product path: gradient reduce -> grid barrier -> global-norm clip + Adam + repack, one cooperative launch.
// Work unit = a quad of 4 consecutive flat elements (every tensor starts 4-aligned and, except the 1-wide value head, has a
// multiple-of-4 row length): the split-K partials come in as float4 loads, all splits of a quad in flight at once (the
// phase is latency-bound otherwise), and the gradients stay in registers across the barrier.
#define RA_QUADS 2
// kernel-matrix elements only: bias elements return 0 with their bit set in `skip` (reduce_biases owns them)
__device__ __forceinline__ float4 reduce_quad(const SegTable& tb, const float* __restrict__ partial, long long i0, unsigned& skip) {
  const Seg& s = find_seg(tb, i0);
  const unsigned j = (unsigned)(i0 - s.off);
  const unsigned len = s.bias ? (unsigned)s.N : (unsigned)s.K * (unsigned)s.N;
  const bool vec = (s.N & 3) == 0 && (j & 3) == 0 && j + 4 <= len;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  skip = 0u;
  if (vec && s.bias) { skip = 15u; return g; }
  if (vec) {
    const float* q; long long stride; int cnt;
    { const unsigned k = j / (unsigned)s.N, n = j - k * (unsigned)s.N; q = partial + s.dwpart + (size_t)k * s.ldo + n; stride = s.split_stride; cnt = s.nsplit; }
    float4 a0 = g, a1 = g;
    int z = 0;
    for (; z + 8 <= cnt; z += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(q + (size_t)(z + u) * stride));
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        a0.x += v[u].x; a0.y += v[u].y; a0.z += v[u].z; a0.w += v[u].w;
        a1.x += v[u + 1].x; a1.y += v[u + 1].y; a1.z += v[u + 1].z; a1.w += v[u + 1].w;
      }
    }
    for (; z < cnt; ++z) { const float4 v = __ldcg(reinterpret_cast<const float4*>(q + (size_t)z * stride)); a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w; }
    g = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
  } else {
    float* pg = reinterpret_cast<float*>(&g);
    for (int e = 0; e < 4; ++e) {
      const long long i = i0 + e;
      if (i >= tb.total) { skip |= 1u << e; continue; }
      const Seg& se = find_seg(tb, i);
      if (se.bias) skip |= 1u << e;
      else pg[e] = reduce_partials(se, partial, i - se.off);
    }
  }
  return g;
}

__global__ void __launch_bounds__(256) k_ppo_reduce_adam(const SegTable* __restrict__ tbp, const float* __restrict__ partial, float* __restrict__ grads,
                                                         float* __restrict__ sumsq_part, float* __restrict__ params, float* __restrict__ m1, float* __restrict__ m2,
                                                         float* __restrict__ packed, int* __restrict__ step, float lr, float b1, float b2, float eps, float max_norm) {
  __shared__ double sh[32];
  __shared__ SegTable tb;
  __shared__ float s_scale, s_c1, s_c2;
  for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) reinterpret_cast<int*>(&tb)[i] = reinterpret_cast<const int*>(tbp)[i];
  __syncthreads();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const long long nquad = (tb.total + 3) >> 2;
  float4 g[RA_QUADS];
  double ss = 0.0;
  auto quad_grad = [&](long long qd) {
    unsigned skip;
    const float4 x = reduce_quad(tb, partial, qd * 4, skip);
    const float* px = reinterpret_cast<const float*>(&x);
    if (skip == 0u) *reinterpret_cast<float4*>(grads + qd * 4) = x;
    else for (int e = 0; e < 4; ++e) if (!((skip >> e) & 1u)) grads[qd * 4 + e] = px[e];
    for (int e = 0; e < 4; ++e) if (!((skip >> e) & 1u)) ss += (double)px[e] * px[e];
    return x;
  };
#pragma unroll
  for (int u = 0; u < RA_QUADS; ++u) {
    const long long qd = tid + (long long)u * nth;
    g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qd < nquad) g[u] = quad_grad(qd);
  }
  for (long long qd = tid + (long long)RA_QUADS * nth; qd < nquad; qd += nth) quad_grad(qd);      // (only if P > 4 RA_QUADS threads)
  reduce_biases(tb, partial, grads, ss);
  const double t = block_sum(ss, sh);
  if (threadIdx.x == 0) sumsq_part[blockIdx.x] = (float)t;
  __threadfence();
  cooperative_groups::this_grid().sync();
  if (threadIdx.x < 32) {
    const double tot = warp_total(sumsq_part, (int)gridDim.x, threadIdx.x);
    if (threadIdx.x == 0) {
      const float gn = (float)sqrt(tot);
      s_scale = (max_norm > 0.f && gn >= max_norm) ? max_norm / gn : 1.f;
      const int tt = step[0] + 1;
      s_c1 = 1.f / (1.f - powf(b1, (float)tt));
      s_c2 = 1.f / (1.f - powf(b2, (float)tt));
    }
  }
  __syncthreads();
  auto quad_adam = [&](long long qd, const float4& x) {
    const float* px = reinterpret_cast<const float*>(&x);
    for (int e = 0; e < 4; ++e) {
      const long long i = qd * 4 + e;
      if (i >= tb.total) continue;
      const Seg& se = find_seg(tb, i);
      adam_element(se, i, se.bias ? __ldcg(grads + i) : px[e], s_scale, s_c1, s_c2, lr, b1, b2, eps, params, m1, m2, packed, true);   // bias gradients were reduced by other warps
    }
  };
#pragma unroll
  for (int u = 0; u < RA_QUADS; ++u) {
    const long long qd = tid + (long long)u * nth;
    if (qd < nquad) quad_adam(qd, g[u]);
  }
  for (long long qd = tid + (long long)RA_QUADS * nth; qd < nquad; qd += nth) {
    float4 x;
    float* px = reinterpret_cast<float*>(&x);
    for (int e = 0; e < 4; ++e) px[e] = qd * 4 + e < tb.total ? grads[qd * 4 + e] : 0.f;
    quad_adam(qd, x);
  }
  cooperative_groups::this_grid().sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) step[0] += 1;
}

// flat gradient <- split-K partials (kernels) / per-warp column sums (biases); per-block sum of squares for the global norm.
// Quads of 4 consecutive elements like the fused kernel above (float4 partial loads, 8 in flight; same summation order as the
// scalar reduce_partials): the scalar version of this kernel took 29 us per minibatch (profiles/r02d_launches_ppo_tf32.csv).
__global__ void __launch_bounds__(256) k_ppo_grad_reduce(const SegTable* __restrict__ tbp, const float* __restrict__ partial, float* __restrict__ grads, float* __restrict__ sumsq_part) {
  __shared__ double sh[32];
  __shared__ SegTable tb;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) reinterpret_cast<int*>(&tb)[i] = reinterpret_cast<const int*>(tbp)[i];
  __syncthreads();
  double ss = 0.0;
  const long long nquad = (tb.total + 3) >> 2;
  for (long long qd = (long long)blockIdx.x * blockDim.x + threadIdx.x; qd < nquad; qd += (long long)gridDim.x * blockDim.x) {
    unsigned skip;
    const float4 x = reduce_quad(tb, partial, qd * 4, skip);               // bias elements are skipped: reduce_biases owns them
    const float* px = reinterpret_cast<const float*>(&x);
    if (skip == 0u) *reinterpret_cast<float4*>(grads + qd * 4) = x;
    else for (int e = 0; e < 4; ++e) if (!((skip >> e) & 1u)) grads[qd * 4 + e] = px[e];
    for (int e = 0; e < 4; ++e) if (!((skip >> e) & 1u)) ss += (double)px[e] * px[e];
  }
  reduce_biases(tb, partial, grads, ss);
  const double t = block_sum(ss, sh);
  if (threadIdx.x == 0) sumsq_part[blockIdx.x] = (float)t;
}

__global__ void __launch_bounds__(256) k_ppo_gradnorm(const float* __restrict__ grads, long long total, float* __restrict__ sumsq_part) {
  __shared__ double sh[32];
  pdl_launch_dependents();
  pdl_wait();
  double ss = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) ss += (double)grads[i] * grads[i];
  const double t = block_sum(ss, sh);
  if (threadIdx.x == 0) sumsq_part[blockIdx.x] = (float)t;
}

// optax.chain(clip_by_global_norm(max_norm), adam(lr)); update = 0: only (re)write the packed operand forms from `params`.
__global__ void __launch_bounds__(256) k_ppo_adam(const SegTable* __restrict__ tbp, float* __restrict__ params, const float* __restrict__ grads, float* __restrict__ m1,
                                                  float* __restrict__ m2, float* __restrict__ packed, const float* __restrict__ sumsq_part, int nparts, int* __restrict__ step,
                                                  float lr, float b1, float b2, float eps, float max_norm, int update) {
  __shared__ SegTable tb;
  __shared__ float s_scale, s_c1, s_c2;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) reinterpret_cast<int*>(&tb)[i] = reinterpret_cast<const int*>(tbp)[i];
  if (threadIdx.x < 32 && update) {
    const double ss = warp_total(sumsq_part, nparts, threadIdx.x);
    if (threadIdx.x == 0) {
      const float gn = (float)sqrt(ss);
      s_scale = (max_norm > 0.f && gn >= max_norm) ? max_norm / gn : 1.f;
      const int t = step[0] + 1;
      s_c1 = 1.f / (1.f - powf(b1, (float)t));
      s_c2 = 1.f / (1.f - powf(b2, (float)t));
    }
  }
  __syncthreads();
  // quads of 4 consecutive elements: one segment lookup, float4 loads / stores of the gradient, the master weights and both
  // moments, and the dX operand form R(W) of a quad is one float4 per half (4 consecutive out features of one in feature)
  const long long nquad = (tb.total + 3) >> 2;
  for (long long qd = (long long)blockIdx.x * blockDim.x + threadIdx.x; qd < nquad; qd += (long long)gridDim.x * blockDim.x) {
    const long long i0 = qd * 4;
    const Seg& sg = find_seg(tb, i0);
    const unsigned j = (unsigned)(i0 - sg.off);
    const unsigned len = sg.bias ? (unsigned)sg.N : (unsigned)sg.K * (unsigned)sg.N;
    if (!((sg.N & 3) == 0 && j + 4 <= len && !sg.bias)) {             // ragged quad (a bias tensor, a 1-wide head, a segment border): element by element
      for (int e = 0; e < 4; ++e) {
        const long long i = i0 + e;
        if (i < tb.total) adam_element(find_seg(tb, i), i, update ? grads[i] : 0.f, s_scale, s_c1, s_c2, lr, b1, b2, eps, params, m1, m2, packed, update != 0);
      }
      continue;
    }
    float4 w4 = *reinterpret_cast<const float4*>(params + i0);
    float* w = reinterpret_cast<float*>(&w4);
    if (update) {
      const float4 g4 = *reinterpret_cast<const float4*>(grads + i0);
      float4 a4 = *reinterpret_cast<const float4*>(m1 + i0), v4 = *reinterpret_cast<const float4*>(m2 + i0);
      const float* g = reinterpret_cast<const float*>(&g4);
      float* mm = reinterpret_cast<float*>(&a4);
      float* vv = reinterpret_cast<float*>(&v4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float ge = g[e] * s_scale;
        mm[e] = b1 * mm[e] + (1.f - b1) * ge; vv[e] = b2 * vv[e] + (1.f - b2) * ge * ge;
        w[e] -= lr * (mm[e] * s_c1) / (sqrtf(vv[e] * s_c2) + eps);
      }
      *reinterpret_cast<float4*>(m1 + i0) = a4; *reinterpret_cast<float4*>(m2 + i0) = v4; *reinterpret_cast<float4*>(params + i0) = w4;
    }
    const int k = (int)(j / (unsigned)sg.N), n = (int)(j - (unsigned)k * (unsigned)sg.N);
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) gsplit_tf32(w[e], hi[e], lo[e]);
    {
      // forward B operand R_nt(W^T): rows = out feature n (4 different rows), contraction over the in feature k
      const int kch = (sg.K + TC_KC - 1) / TC_KC;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ne = n + e;
        const int ntile = sg.nt == 128 ? ne >> 7 : ne >> 5, nrow = ne & (sg.nt - 1);
        float* blk = packed + sg.wf + ((size_t)ntile * kch + (k >> 5)) * gblk_b(sg.nt);
        const int off = gblk_off(nrow, k & 31);
        blk[off] = hi[e]; blk[sg.nt * TC_KC + off] = lo[e];
      }
    }
    if (sg.wb >= 0) {
      // dX B operand R(W): row = in feature k, the 4 out features n .. n + 3 are one 16-byte core-matrix row segment
      const int nch = (sg.N + TC_KC - 1) / TC_KC;
      float* blk = packed + sg.wb + ((size_t)(k >> 7) * nch + (n >> 5)) * GBLK_A;
      const int off = gblk_off(k & 127, n & 31);
      *reinterpret_cast<float4*>(blk + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(blk + TC_M * TC_KC + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  if (update) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int ticket = atomicAdd(step + 1, 1);
      if (ticket == (int)gridDim.x - 1) { step[0] += 1; step[1] = 0; }     // every block has read step[0] by now
    }
  }
}

// ------------------------------------------------------------------------------------------------- host side
template <typename T>
static int dev_alloc(OduckPpo* h, T*& ptr, size_t count) {
  if (cudaMalloc((void**)&ptr, count * sizeof(T)) != cudaSuccess) return -1;
  cudaMemset(ptr, 0, count * sizeof(T));
  h->allocs.push_back((void*)ptr);
  return 0;
}

extern "C" {

int oduck_ppo_destroy(OduckPpo* h) {
  if (!h) return ODUCK_OK;
  cudaSetDevice(h->device);
  for (void* q : h->allocs) cudaFree(q);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->pf_stream) cudaStreamDestroy(h->pf_stream);
  if (h->ev_mb_start) cudaEventDestroy(h->ev_mb_start);
  if (h->ev_pf_done) cudaEventDestroy(h->ev_pf_done);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  for (int k = 0; k < 2; ++k) {
    if (h->side_w[k]) cudaStreamDestroy(h->side_w[k]);
    if (h->ev_join_w[k]) cudaEventDestroy(h->ev_join_w[k]);
    for (int l = 0; l < PPO_NL; ++l) if (h->ev_dz[k][l]) cudaEventDestroy(h->ev_dz[k][l]);
  }
  delete h;
  return ODUCK_OK;
}

int oduck_ppo_create(const OduckPpoConfig* cfg, int device, OduckPpo** out) {
  if (!cfg || !out) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_create: bad argument");
  const int B = cfg->batch_envs, T = cfg->unroll, na = cfg->num_actions;
  if (B <= 0 || T <= 0 || T > PPO_MAXT || na <= 0 || na > 16) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_create: batch_envs / unroll / num_actions out of range");
  if (cfg->policy_dims[4] != 2 * na || cfg->value_dims[4] != 1) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_create: heads must be 2 * num_actions (policy) and 1 (value)");
  for (int net = 0; net < 2; ++net) {
    const int32_t* d = net ? cfg->value_dims : cfg->policy_dims;
    if (d[0] <= 0) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_create: bad observation size");
    for (int l = 1; l <= 3; ++l)
      if (d[l] <= 0 || d[l] % 128) return oduck_fail(ODUCK_ERR_UNSUPPORTED, "oduck_ppo_create: hidden sizes must be multiples of 128 (reference: 512, 256, 128)");
  }
  int ndev = 0;
  PPO_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return oduck_fail(ODUCK_ERR_CUDA, "oduck_ppo_create: no such CUDA device");
  PPO_TRY(cudaSetDevice(device));
  OduckPpo* h = new OduckPpo();
  h->cfg = *cfg; h->device = device; h->B = B; h->T = T; h->na = na; h->launches = 0;
  { const char* e = getenv("ODUCK_PPO_PDL"); h->pdl = e && e[0] == '1'; }
  h->side = nullptr; h->ev_fork = h->ev_join = nullptr; h->coop_blocks = 0;
  h->pf_stream = nullptr; h->ev_mb_start = h->ev_pf_done = nullptr; h->x0_sel = 0; h->pf_idx = nullptr; memset(h->X0r, 0, sizeof(h->X0r)); memset(h->X0t, 0, sizeof(h->X0t));
  memset(h->side_w, 0, sizeof(h->side_w)); memset(h->ev_join_w, 0, sizeof(h->ev_join_w)); memset(h->ev_dz, 0, sizeof(h->ev_dz));
  {
    bool sok = cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < 2 && sok; ++k) {
      sok = cudaStreamCreateWithFlags(&h->side_w[k], cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&h->ev_join_w[k], cudaEventDisableTiming) == cudaSuccess;
      for (int l = 0; l < PPO_NL && sok; ++l) sok = cudaEventCreateWithFlags(&h->ev_dz[k][l], cudaEventDisableTiming) == cudaSuccess;
    }
    sok = sok && cudaStreamCreateWithFlags(&h->pf_stream, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&h->ev_mb_start, cudaEventDisableTiming) == cudaSuccess &&
          cudaEventCreateWithFlags(&h->ev_pf_done, cudaEventDisableTiming) == cudaSuccess;
    if (!sok) { oduck_ppo_destroy(h); return oduck_fail(ODUCK_ERR_CUDA, "oduck_ppo_create: stream / event creation failed"); }
  }
  // ---- parameter segments, packed-operand and partial-gradient layouts
  SegTable& tb = h->seg;
  memset(&tb, 0, sizeof(tb));
  long long off = 0, poff = 0, goff = 0;
  for (int net = 0; net < 2; ++net) {
    NetBuf& nb = h->net[net];
    const int32_t* d = net ? cfg->value_dims : cfg->policy_dims;
    for (int l = 0; l <= PPO_NL; ++l) nb.dims[l] = d[l];
    nb.M = (net ? T + 1 : T) * B;
    nb.Mpad = round_up(nb.M, 128);
    nb.mtiles = nb.Mpad / 128;
    for (int l = 0; l < PPO_NL; ++l) {
      const int K = d[l], N = d[l + 1];
      const int nt = l == PPO_NL - 1 ? PPO_HEADW : 128;
      const int Npad = round_up(N, nt), kch = ceil_div(K, TC_KC);
      Seg w;
      memset(&w, 0, sizeof(w));
      w.off = off; w.net = net; w.layer = l; w.bias = 0; w.K = K; w.N = N; w.nt = nt;
      w.wf = poff; poff += (long long)(Npad / nt) * kch * gblk_b(nt);
      w.wb = -1;
      if (l > 0) { w.wb = poff; poff += (long long)ceil_div(K, 128) * ceil_div(N, TC_KC) * GBLK_A; }
      // dW partials: [nsplit][round_up(K, 128)][Npad]; about 8 splits over the batch chunks: fewer splits mean less partial-sum traffic
      // for the reduce kernel on the critical path, more mean fuller grids for the dW GEMMs beside it.  Measured on B200 (update of
      // configs[2], two runs each, profiles/r02aa_bench_ppo_*.json): 8 vs 16 splits 26.41 / 26.42 vs 27.25 / 27.29 ms (TF32),
      // 33.37 / 33.29 vs 33.73 / 33.71 ms (3xTF32); 6 / 10 / 12 splits 27.9 / 28.0 / 27.1 ms (TF32, r02y / r02z).  ODUCK_PPO_SPLITS overrides.
      int want_splits = 8;
      { const char* e = getenv("ODUCK_PPO_SPLITS"); if (e && atoi(e) > 0) want_splits = atoi(e); }
      const int bch = nb.Mpad / TC_KC, cps = ceil_div(bch, want_splits);
      w.nsplit = ceil_div(bch, cps); w.ldo = Npad; w.split_stride = (long long)round_up(K, 128) * Npad;
      w.dwpart = goff; goff += w.split_stride * w.nsplit;
      off += (long long)K * N;
      Seg b = w;
      b.off = off; b.bias = 1; b.K = 1; b.wf = b.wb = -1;
      b.ldb = round_up(N, 32); b.nwarprows = l == PPO_NL - 1 ? nb.Mpad / LOSS_ROWS : nb.Mpad / 32;   // head partials come from k_ppo_loss
      b.dbpart = goff; goff += (long long)b.ldb * b.nwarprows;
      off += N;
      tb.s[tb.n++] = w; tb.s[tb.n++] = b;
    }
  }
  tb.total = off;
  h->P = off; h->packed_floats = poff; h->partial_floats = goff;
  bool ok = true;
  ok = ok && dev_alloc(h, h->dseg, 1) == 0;
  ok = ok && dev_alloc(h, h->params, (size_t)h->P) == 0 && dev_alloc(h, h->grads, (size_t)h->P) == 0 && dev_alloc(h, h->adam_m, (size_t)h->P) == 0 && dev_alloc(h, h->adam_v, (size_t)h->P) == 0;
  ok = ok && dev_alloc(h, h->packed, (size_t)poff) == 0 && dev_alloc(h, h->partial, (size_t)goff) == 0;
  ok = ok && dev_alloc(h, h->adv, (size_t)T * B) == 0 && dev_alloc(h, h->vs, (size_t)T * B) == 0 && dev_alloc(h, h->adv_norm, (size_t)T * B) == 0;
  ok = ok && dev_alloc(h, h->losses, 8) == 0 && dev_alloc(h, h->stats, 4 + 2 * (size_t)ceil_div(B, GAE_THREADS)) == 0 && dev_alloc(h, h->step, 4) == 0;
  h->reduce_blocks = 296;
  ok = ok && dev_alloc(h, h->sumsq_part, (size_t)h->reduce_blocks) == 0;
  for (int net = 0; net < 2 && ok; ++net) {
    NetBuf& nb = h->net[net];
    const size_t bch = nb.Mpad / TC_KC;
    for (int l = 0; l < PPO_NL && ok; ++l) {
      const int K = nb.dims[l];
      ok = ok && dev_alloc(h, nb.Xr[l], (size_t)nb.mtiles * ceil_div(K, TC_KC) * GBLK_A) == 0;
      ok = ok && dev_alloc(h, nb.Xt[l], (size_t)ceil_div(K, 128) * bch * GBLK_A) == 0;
      if (l == 0) {                                                    // second input set for oduck_ppo_prefetch
        h->X0r[net][0] = nb.Xr[0]; h->X0t[net][0] = nb.Xt[0];
        ok = ok && dev_alloc(h, h->X0r[net][1], (size_t)nb.mtiles * ceil_div(K, TC_KC) * GBLK_A) == 0;
        ok = ok && dev_alloc(h, h->X0t[net][1], (size_t)ceil_div(K, 128) * bch * GBLK_A) == 0;
      }
      const int N = nb.dims[l + 1];
      if (l < PPO_NL - 1) {
        ok = ok && dev_alloc(h, nb.Z[l], (size_t)nb.Mpad * N) == 0;
        ok = ok && dev_alloc(h, nb.dZr[l], (size_t)nb.mtiles * (N / TC_KC) * GBLK_A) == 0;
        ok = ok && dev_alloc(h, nb.dZt[l], (size_t)(N / 128) * bch * GBLK_A) == 0;
      } else {
        ok = ok && dev_alloc(h, nb.dZr[l], (size_t)nb.mtiles * GBLK_A) == 0;
        ok = ok && dev_alloc(h, nb.dZt[l], bch * gblk_b(PPO_HEADW)) == 0;
      }
    }
    ok = ok && dev_alloc(h, nb.out, (size_t)nb.Mpad * PPO_HEADW) == 0;
  }
  if (!ok) { oduck_ppo_destroy(h); return oduck_fail(ODUCK_ERR_ALLOC, "oduck_ppo_create: cudaMalloc failed"); }
  PPO_TRY(cudaMemcpy(h->dseg, &h->seg, sizeof(SegTable), cudaMemcpyHostToDevice));
  {
    int coop = 0, per_sm = 0, sms = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    h->num_sms = sms > 0 ? sms : 148;
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ppo_reduce_adam, 256, 0) == cudaSuccess && per_sm > 0)
      h->coop_blocks = std::min(h->reduce_blocks, per_sm * sms);
  }
  PPO_TRY(cudaDeviceSynchronize());
  *out = h;
  return ODUCK_OK;
}

int64_t oduck_ppo_num_params(const OduckPpo* h) { return h ? h->P : 0; }
int64_t oduck_ppo_launch_count(const OduckPpo* h) { return h ? h->launches : 0; }

int oduck_ppo_param_info(const OduckPpo* h, int net, int layer, int which, int64_t* offset, int64_t* rows, int64_t* cols) {
  if (!h || net < 0 || net > 1 || layer < 0 || layer >= PPO_NL || which < 0 || which > 1 || !offset || !rows || !cols) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_param_info: bad argument");
  const Seg& s = h->seg.s[(net * PPO_NL + layer) * 2 + which];
  *offset = s.off; *rows = which ? 1 : s.K; *cols = s.N;
  return ODUCK_OK;
}

int oduck_ppo_packed_weights(OduckPpo* h, int net, int layer, const float** ptr) {
  if (!h || net < 0 || net > 1 || layer < 0 || layer >= PPO_NL || !ptr) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_packed_weights: bad argument");
  *ptr = h->packed + h->seg.s[(net * PPO_NL + layer) * 2].wf;
  return ODUCK_OK;
}

static int launch_adam(OduckPpo* h, int update, cudaStream_t st) {
  const OduckPpoConfig& c = h->cfg;
  PPO_TRY(launch_kernel(k_ppo_adam, dim3(h->reduce_blocks), dim3(256), 0, st, h->pdl, h->dseg, h->params, h->grads, h->adam_m, h->adam_v, h->packed, h->sumsq_part,
                        h->reduce_blocks, h->step, c.learning_rate, c.adam_b1, c.adam_b2, c.adam_eps, c.max_grad_norm, update));
  h->launches++;
  return ODUCK_OK;
}

int oduck_ppo_set_params(OduckPpo* h, const float* flat, int reset_opt, void* stream) {
  if (!h || !flat) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_set_params: bad argument");
  PPO_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (flat != h->params) PPO_TRY(cudaMemcpyAsync(h->params, flat, (size_t)h->P * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (reset_opt) {
    PPO_TRY(cudaMemsetAsync(h->adam_m, 0, (size_t)h->P * sizeof(float), st));
    PPO_TRY(cudaMemsetAsync(h->adam_v, 0, (size_t)h->P * sizeof(float), st));
    PPO_TRY(cudaMemsetAsync(h->step, 0, 4 * sizeof(int), st));
  }
  return launch_adam(h, 0, st);
}

int oduck_ppo_get_buffer(OduckPpo* h, int id, void** ptr, int64_t* count, int* dtype) {
  if (!h || !ptr || !count || !dtype) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_get_buffer: bad argument");
  *dtype = ODUCK_DTYPE_F32;
  switch (id) {
    case ODUCK_PPO_BUF_PARAMS: *ptr = h->params; *count = h->P; break;
    case ODUCK_PPO_BUF_GRADS: *ptr = h->grads; *count = h->P; break;
    case ODUCK_PPO_BUF_ADAM_M: *ptr = h->adam_m; *count = h->P; break;
    case ODUCK_PPO_BUF_ADAM_V: *ptr = h->adam_v; *count = h->P; break;
    case ODUCK_PPO_BUF_LOGITS: *ptr = h->net[0].out; *count = (int64_t)h->net[0].Mpad * PPO_HEADW; break;
    case ODUCK_PPO_BUF_VALUES: *ptr = h->net[1].out; *count = (int64_t)h->net[1].Mpad * PPO_HEADW; break;
    case ODUCK_PPO_BUF_LOSSES: *ptr = h->losses; *count = 8; *dtype = ODUCK_DTYPE_F64; break;
    case ODUCK_PPO_BUF_ADV: *ptr = h->adv_norm; *count = (int64_t)h->T * h->B; break;
    case ODUCK_PPO_BUF_VS: *ptr = h->vs; *count = (int64_t)h->T * h->B; break;
    case ODUCK_PPO_BUF_STEP: *ptr = h->step; *count = 1; *dtype = ODUCK_DTYPE_I32; break;
    default: return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_get_buffer: unknown buffer id");
  }
  return ODUCK_OK;
}

#define GEMM_TRY(call)                                                                                                  \
  do {                                                                                                                  \
    cudaError_t e_ = (call);                                                                                            \
    if (e_ != cudaSuccess) return oduck_fail(ODUCK_ERR_CUDA, std::string("oduck_ppo_minibatch launch: ") + cudaGetErrorString(e_)); \
    h->launches++;                                                                                                      \
  } while (0)

static int net_forward(OduckPpo* h, int net, bool simt, cudaStream_t st) {
  NetBuf& nb = h->net[net];
  for (int l = 0; l < PPO_NL; ++l) {
    const Seg& w = h->seg.s[(net * PPO_NL + l) * 2];
    const Seg& b = h->seg.s[(net * PPO_NL + l) * 2 + 1];
    GemmParams g;
    memset(&g, 0, sizeof(g));
    g.single = h->cfg.matmul_tf32 ? 1 : 0;
    g.A = nb.Xr[l]; g.B = h->packed + w.wf;
    g.nchunks = ceil_div(w.K, TC_KC); g.cps = g.nchunks;
    g.bias = h->params + b.off; g.nvalid = w.N;
    if (l < PPO_NL - 1) {
      g.Z = nb.Z[l]; g.z_nch = w.N / TC_KC;
      g.Yr = nb.Xr[l + 1]; g.yr_nch = w.N / TC_KC;
      g.Yt = nb.Xt[l + 1]; g.yt_nch = nb.Mpad / TC_KC;
      GEMM_TRY((launch_gemm<128, 3, EPI_FWD>(g, nb.mtiles, w.N / 128, simt, st, h->pdl)));
    } else {
      g.out = nb.out; g.ldo = PPO_HEADW;
      GEMM_TRY((launch_gemm<PPO_HEADW, 4, EPI_OUT>(g, nb.mtiles, 1, simt, st, h->pdl)));
    }
  }
  return ODUCK_OK;
}

// Backward of one net on two streams: the dZ chain (dX GEMMs, the critical path) on `sx`, the weight-gradient GEMMs on `sw`;
// dW_l waits for the event that marks dZ_l complete (dZ of the head layer comes from k_ppo_loss, before the fork).
static int net_backward(OduckPpo* h, int net, bool simt, cudaStream_t sx, cudaStream_t sw) {
  NetBuf& nb = h->net[net];
  for (int l = PPO_NL - 1; l >= 0; --l) {
    const Seg& w = h->seg.s[(net * PPO_NL + l) * 2];
    {
      // dW_l = X_l^T dZ_l: rows = in features, columns = out features, contraction over the batch (split-K)
      GemmParams g;
      memset(&g, 0, sizeof(g));
    g.single = h->cfg.matmul_tf32 ? 1 : 0;
      g.A = nb.Xt[l]; g.B = nb.dZt[l];
      g.nchunks = nb.Mpad / TC_KC; g.cps = ceil_div(g.nchunks, w.nsplit);
      g.out = h->partial + w.dwpart; g.ldo = w.ldo; g.out_split = w.split_stride;
      const int mt = ceil_div(w.K, 128);
      if (l == PPO_NL - 1) GEMM_TRY((launch_gemm<PPO_HEADW, 4, EPI_DW>(g, mt, 1, simt, sw, h->pdl)));
      else GEMM_TRY((launch_gemm<128, 3, EPI_DW>(g, mt, w.N / 128, simt, sw, h->pdl)));
    }
    if (l > 0) {
      // dZ_{l-1} = (dZ_l W_l^T) * swish'(Z_{l-1}): columns = in features of layer l, contraction over its out features
      const Seg& bprev = h->seg.s[(net * PPO_NL + l - 1) * 2 + 1];
      GemmParams g;
      memset(&g, 0, sizeof(g));
    g.single = h->cfg.matmul_tf32 ? 1 : 0;
      g.A = nb.dZr[l]; g.B = h->packed + w.wb;
      g.nchunks = ceil_div(w.N, TC_KC); g.cps = g.nchunks;
      g.Z = nb.Z[l - 1]; g.z_nch = w.K / TC_KC;
      g.Yr = nb.dZr[l - 1]; g.yr_nch = w.K / TC_KC;
      g.Yt = nb.dZt[l - 1]; g.yt_nch = nb.Mpad / TC_KC;
      g.dbpart = h->partial + bprev.dbpart; g.ldb = bprev.ldb;
      g.nvalid = w.K;
      GEMM_TRY((launch_gemm<128, 3, EPI_DX>(g, nb.mtiles, w.K / 128, simt, sx, h->pdl)));
      if (sw != sx) {
        PPO_TRY(cudaEventRecord(h->ev_dz[net][l - 1], sx));
        PPO_TRY(cudaStreamWaitEvent(sw, h->ev_dz[net][l - 1], 0));
      }
    }
  }
  return ODUCK_OK;
}

int oduck_ppo_prefetch(OduckPpo* h, const OduckRollout* ro, const OduckNormalizer* nm, const int32_t* next_env_idx, void* stream) {
  if (!h || !ro || !nm || !next_env_idx) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_prefetch: bad argument");
  if (ro->unroll != h->T || ro->num_envs < 1) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_prefetch: rollout shape does not match the learner");
  if (ro->block_envs < 0 || (ro->block_envs > 0 && ro->num_envs % ro->block_envs)) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_prefetch: num_envs must be a multiple of block_envs");
  if (ro->obs_policy_ld != 0 && ro->obs_policy_ld < h->net[0].dims[0]) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_prefetch: obs_policy_ld smaller than the policy observation");
  PPO_TRY(cudaSetDevice(h->device));
  (void)stream;                                  // (kept in the signature: the fork point is an event of that stream, see below)
  const int be = ro->block_envs > 0 ? ro->block_envs : ro->num_envs;
  NetBuf& np = h->net[0];
  NetBuf& nv = h->net[1];
  const int alt = h->x0_sel ^ 1;                 // the set the current minibatch does not read (its last reader, the previous
                                                 // minibatch's layer-0 dW GEMM, is ordered before the current minibatch's start)
  // fork from the START of the current minibatch (recorded by oduck_ppo_minibatch at its FORWARD stage), not from the tail of
  // `stream`: the pack then runs beside the current minibatch's kernels
  PPO_TRY(cudaStreamWaitEvent(h->pf_stream, h->ev_mb_start, 0));
  PPO_TRY(launch_kernel(k_ppo_pack, dim3(296), dim3(256), 0, h->pf_stream, false, ro->obs_value, be, (long long)ro->block_stride, nv.dims[0], nv.dims[0], next_env_idx, h->B, nv.M, nv.Mpad,
                        ceil_div(nv.dims[0], TC_KC), nm->value_mean, nm->value_std, h->X0r[1][alt], h->X0t[1][alt]));
  PPO_TRY(launch_kernel(k_ppo_pack, dim3(296), dim3(256), 0, h->pf_stream, false, ro->obs_policy, be, (long long)ro->block_stride, ro->obs_policy_ld > 0 ? ro->obs_policy_ld : np.dims[0], np.dims[0],
                        next_env_idx, h->B, np.M, np.Mpad, ceil_div(np.dims[0], TC_KC), nm->policy_mean, nm->policy_std, h->X0r[0][alt], h->X0t[0][alt]));
  PPO_TRY(cudaEventRecord(h->ev_pf_done, h->pf_stream));
  h->launches += 2;
  h->pf_idx = next_env_idx;
  return ODUCK_OK;
}

int oduck_ppo_minibatch(OduckPpo* h, const OduckRollout* ro, const OduckNormalizer* nm, const int32_t* env_idx, const float* noise,
                        const uint32_t* key, int stages, void* stream) {
  if (!h || !ro || !nm || !env_idx) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_minibatch: bad argument");
  if (ro->unroll != h->T || ro->num_envs < 1) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_minibatch: rollout shape does not match the learner");
  if (ro->obs_policy_ld != 0 && ro->obs_policy_ld < h->net[0].dims[0]) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_minibatch: obs_policy_ld smaller than the policy observation");
  if (ro->block_envs < 0 || (ro->block_envs > 0 && ro->num_envs % ro->block_envs)) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_minibatch: num_envs must be a multiple of block_envs");
  const int be = ro->block_envs > 0 ? ro->block_envs : ro->num_envs;
  if ((stages & ODUCK_PPO_STAGE_LOSS) && !noise && !key) return oduck_fail(ODUCK_ERR_ARG, "oduck_ppo_minibatch: entropy term needs noise or a key");
  PPO_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const bool simt = (stages & ODUCK_PPO_DEBUG_SIMT) != 0;
  const OduckPpoConfig& c = h->cfg;
  NetBuf& np = h->net[0];
  NetBuf& nv = h->net[1];
  // The two nets are independent until the loss and again after it: the value chain runs on the side stream, forked from
  // and joined back into the caller's stream with events (capturable into a CUDA graph like any other launch).
#define FORK() { PPO_TRY(cudaEventRecord(h->ev_fork, st)); PPO_TRY(cudaStreamWaitEvent(h->side, h->ev_fork, 0)); }
#define JOIN() { PPO_TRY(cudaEventRecord(h->ev_join, h->side)); PPO_TRY(cudaStreamWaitEvent(st, h->ev_join, 0)); }
  if (stages & ODUCK_PPO_STAGE_FORWARD) {
    // input operands: the set oduck_ppo_prefetch packed for this very minibatch (wait for its stream), or pack them now
    const bool prefetched = h->pf_idx != nullptr && h->pf_idx == env_idx;
    if (prefetched) {
      h->x0_sel ^= 1;
      PPO_TRY(cudaStreamWaitEvent(st, h->ev_pf_done, 0));
    }
    h->pf_idx = nullptr;
    for (int net = 0; net < 2; ++net) { h->net[net].Xr[0] = h->X0r[net][h->x0_sel]; h->net[net].Xt[0] = h->X0t[net][h->x0_sel]; }
    PPO_TRY(cudaEventRecord(h->ev_mb_start, st));                     // fork point of a following oduck_ppo_prefetch
    FORK()
    if (!prefetched)
      GEMM_TRY(launch_kernel(k_ppo_pack, dim3(296), dim3(256), 0, h->side, h->pdl, ro->obs_value, be, (long long)ro->block_stride, nv.dims[0], nv.dims[0], env_idx, h->B, nv.M, nv.Mpad, ceil_div(nv.dims[0], TC_KC), nm->value_mean, nm->value_std, nv.Xr[0], nv.Xt[0]));
    int rc = net_forward(h, 1, simt, h->side);
    if (rc) return rc;
    if (!prefetched)
      GEMM_TRY(launch_kernel(k_ppo_pack, dim3(296), dim3(256), 0, st, h->pdl, ro->obs_policy, be, (long long)ro->block_stride, ro->obs_policy_ld > 0 ? ro->obs_policy_ld : np.dims[0], np.dims[0], env_idx, h->B, np.M, np.Mpad, ceil_div(np.dims[0], TC_KC), nm->policy_mean, nm->policy_std, np.Xr[0], np.Xt[0]));
    rc = net_forward(h, 0, simt, st);
    if (rc) return rc;
    JOIN()
  }
  if (stages & ODUCK_PPO_STAGE_LOSS) {
    GEMM_TRY(launch_kernel(k_ppo_gae, dim3(ceil_div(h->B, GAE_THREADS)), dim3(GAE_THREADS), 0, st, h->pdl, nv.out, *ro, env_idx, h->B, h->T, c.discounting, c.gae_lambda, c.reward_scaling, h->adv, h->vs, h->stats, h->step + 2, h->losses));
    LossParams lp;
    memset(&lp, 0, sizeof(lp));
    lp.logits = np.out; lp.values = nv.out; lp.adv = h->adv; lp.vs = h->vs; lp.stats = h->stats; lp.adv_norm = h->adv_norm; lp.normalize = c.normalize_advantage; lp.ro = *ro; lp.idx = env_idx; lp.noise = noise; lp.key = key;
    lp.B = h->B; lp.T = h->T; lp.na = h->na; lp.Mp = np.M; lp.Mp_pad = np.Mpad; lp.Mv = nv.M; lp.Mv_pad = nv.Mpad;
    lp.clip_eps = c.clipping_epsilon; lp.entropy_cost = c.entropy_cost;
    const Seg& bp = h->seg.s[(0 * PPO_NL + PPO_NL - 1) * 2 + 1];
    const Seg& bv = h->seg.s[(1 * PPO_NL + PPO_NL - 1) * 2 + 1];
    lp.dZr_p = np.dZr[PPO_NL - 1]; lp.dZt_p = np.dZt[PPO_NL - 1]; lp.db_p = h->partial + bp.dbpart;
    lp.dZr_v = nv.dZr[PPO_NL - 1]; lp.dZt_v = nv.dZt[PPO_NL - 1]; lp.db_v = h->partial + bv.dbpart;
    lp.losses = h->losses;
    GEMM_TRY(launch_kernel(k_ppo_loss, dim3(nv.Mpad / LOSS_ROWS), dim3(LOSS_ROWS * 16), 0, st, h->pdl, lp));
  }
  if (stages & ODUCK_PPO_STAGE_BACKWARD) {
    FORK()
    for (int k = 0; k < 2; ++k) PPO_TRY(cudaStreamWaitEvent(h->side_w[k], h->ev_fork, 0));
    int rc = net_backward(h, 1, simt, h->side, h->side_w[1]);
    if (rc) return rc;
    rc = net_backward(h, 0, simt, st, h->side_w[0]);
    if (rc) return rc;
    JOIN()
    for (int k = 0; k < 2; ++k) { PPO_TRY(cudaEventRecord(h->ev_join_w[k], h->side_w[k])); PPO_TRY(cudaStreamWaitEvent(st, h->ev_join_w[k], 0)); }
    if ((stages & ODUCK_PPO_STAGE_ADAM) && h->coop_blocks > 0 && !(stages & ODUCK_PPO_NO_COOP)) {
      // product path on one GPU: gradient reduce, global norm and Adam in one cooperative launch (grid barrier in between)
      const OduckPpoConfig& cc = h->cfg;
      const SegTable* a0 = h->dseg; const float* a1 = h->partial; float* a2 = h->grads; float* a3 = h->sumsq_part; float* a4 = h->params; float* a5 = h->adam_m;
      float* a6 = h->adam_v; float* a7 = h->packed; int* a8 = h->step;
      float lr = cc.learning_rate, b1 = cc.adam_b1, b2 = cc.adam_b2, eps = cc.adam_eps, mx = cc.max_grad_norm;
      void* args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &lr, &b1, &b2, &eps, &mx};
      PPO_TRY(cudaLaunchCooperativeKernel((void*)k_ppo_reduce_adam, dim3(h->coop_blocks), dim3(256), args, 0, st));
      h->launches++;
      return ODUCK_OK;
    }
    GEMM_TRY(launch_kernel(k_ppo_grad_reduce, dim3(h->reduce_blocks), dim3(256), 0, st, h->pdl, h->dseg, h->partial, h->grads, h->sumsq_part));
  }
  if (stages & ODUCK_PPO_STAGE_ADAM) {
    if (!(stages & ODUCK_PPO_STAGE_BACKWARD)) {           // gradients were all-reduced by the caller: recompute their norm
      GEMM_TRY(launch_kernel(k_ppo_gradnorm, dim3(h->reduce_blocks), dim3(256), 0, st, h->pdl, h->grads, (long long)h->P, h->sumsq_part));
    }
    return launch_adam(h, 1, st);
  }
  return ODUCK_OK;
}

}  // extern "C"
