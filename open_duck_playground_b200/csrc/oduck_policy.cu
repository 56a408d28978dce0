// oduck_policy.cu -- A15: actor-MLP forward of the Brax PPO policy (common/runner.py:94-100; architecture mirrored in
// common/export_onnx.py:14-72): x = (obs - mean) / std; 3 x (Dense + swish); Dense -> (loc, scale);
// action = tanh(loc + (softplus(scale) + 0.001) * N(0,1)), log-prob with the tanh Jacobian (Brax NormalTanhDistribution).
// Tensor-core version (tcgen05 + TMEM + TMA bulk copies), 3xTF32 for fp32-faithful logits: see oduck_policy_tc.cuh.
#include <cuda_runtime.h>

#include <string.h>
#include <string>

#include "oduck_handle.cuh"
#include "oduck_policy_tc.cuh"
#include "oduck_gemm_tc.cuh"
#include <stdlib.h>

extern int oduck_fail(int code, const std::string& msg);

__device__ __forceinline__ float erfinv_giles(float x) {   // the f32 polynomial XLA uses for lax.erf_inv (M. Giles)
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

// ------------------------------------------------------------------------------------------------- tcgen05 path
// Operands live in HBM/L2 already in UMMA form: per (row-tile, k-chunk) one contiguous block
//   [hi | lo][rows/8][8 k-groups][8 rows][4 floats]      (hi = value truncated to tf32, lo = value - hi)
// so a pipeline stage is two 1-D TMA bulk copies.  3xTF32: D += A_hi B_hi + A_lo B_hi + A_hi B_lo keeps fp32-level accuracy
// (the rollout log-prob must agree with the learner's fp32 recomputation; plain tf32 is off by ~1e-2 in the logits).
#define BLK_A (2 * TC_M * TC_KC)                 // floats per A block (hi + lo)
__host__ __device__ constexpr int blk_b(int nt) { return 2 * nt * TC_KC; }

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = v - hi;
}
__device__ __forceinline__ int blk_off(int r, int kk) { return (r >> 3) * 256 + (kk >> 2) * 32 + (r & 7) * 4 + (kk & 3); }

// weights [K][N] (flax) -> blocked hi/lo tiles [N/NT][Kpad/32]
__global__ void k_pack_weights(const float* __restrict__ W, float* __restrict__ out, int K, int N, int NT) {
  const int nchunks = (K + TC_KC - 1) / TC_KC, ntiles = (N + NT - 1) / NT;
  const int total = ntiles * nchunks * NT * TC_KC;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx % NT, kk = (idx / NT) % TC_KC, c = (idx / (NT * TC_KC)) % nchunks, t = idx / (NT * TC_KC * nchunks);
    const int n = t * NT + r, k = c * TC_KC + kk;
    const float v = (n < N && k < K) ? W[(size_t)k * N + n] : 0.f;
    float hi, lo;
    split_tf32(v, hi, lo);
    float* blk = out + (size_t)(t * nchunks + c) * blk_b(NT);
    blk[blk_off(r, kk)] = hi;
    blk[NT * TC_KC + blk_off(r, kk)] = lo;
  }
}
// observations -> normalised, blocked hi/lo A operand of the first layer.  One thread per (row, group of 4 k): the blocked
// layout keeps those 4 values contiguous, so hi and lo go out as one float4 each, 128 bytes per 8 consecutive rows.
__global__ void k_pack_obs(const float* __restrict__ obs, int ld, const float* __restrict__ mean, const float* __restrict__ stdv, float* __restrict__ out, int M, int K) {
  const int nchunks = (K + TC_KC - 1) / TC_KC, mtiles = (M + TC_M - 1) / TC_M;
  const int total = mtiles * nchunks * TC_M * (TC_KC / 4);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx % TC_M, kg = (idx / TC_M) % (TC_KC / 4), c = (idx / (TC_M * (TC_KC / 4))) % nchunks, t = idx / (TC_M * (TC_KC / 4) * nchunks);
    const int row = t * TC_M + r, k0 = c * TC_KC + 4 * kg;
    float4 h4, l4;
    float* ph = reinterpret_cast<float*>(&h4);
    float* pl = reinterpret_cast<float*>(&l4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + e;
      const float v = (row < M && k < K) ? (obs[(size_t)row * ld + k] - mean[k]) / stdv[k] : 0.f;
      split_tf32(v, ph[e], pl[e]);
    }
    float* blk = out + (size_t)(t * nchunks + c) * BLK_A;
    *reinterpret_cast<float4*>(blk + blk_off(r, 4 * kg)) = h4;
    *reinterpret_cast<float4*>(blk + TC_M * TC_KC + blk_off(r, 4 * kg)) = l4;
  }
}

struct DenseParams {
  const float *Xb, *Wb, *B;   // blocked A [mtiles][nchunks][BLK_A], blocked W [ntiles][nchunks][blk_b(NT)], bias [N]
  float* Yb;                  // blocked A operand of the next layer (K_next = N), or null for the head
  int M, K, N;
  const uint32_t* keys;
  float *action, *raw, *logp;
  int deterministic, na;
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

template <int NT, bool HEAD>
__global__ void __launch_bounds__(TC_THREADS) k_dense_tc(DenseParams p) {
  extern __shared__ __align__(128) unsigned char raw_smem[];
  __shared__ __align__(8) uint64_t full[2], empty[2], done;
  __shared__ uint32_t tmem_base_s;
  constexpr int STAGE = BLK_A + blk_b(NT);                       // floats per stage
  float* stage[2] = {reinterpret_cast<float*>(raw_smem), reinterpret_cast<float*>(raw_smem) + STAGE};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;
  constexpr uint32_t kCols = NT < 32 ? 32 : NT;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&empty[0], 1); mbar_init(&empty[1], 1); mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const int nchunks = (p.K + TC_KC - 1) / TC_KC;
  if (threadIdx.x == 0) {
    // producer + MMA issuer in one thread, software-pipelined: the copies of chunk c+1 fly while the MMAs of chunk c run
    const uint32_t idesc = umma_idesc_tf32(NT);
    const float* xa = p.Xb + (size_t)mt * nchunks * BLK_A;
    const float* wb = p.Wb + (size_t)nt * nchunks * blk_b(NT);
    auto issue = [&](int c) {
      const int s = c & 1;
      mbar_expect_tx(&full[s], (uint32_t)(STAGE * sizeof(float)));
      bulk_g2s(stage[s], xa + (size_t)c * BLK_A, BLK_A * sizeof(float), &full[s]);
      bulk_g2s(stage[s] + BLK_A, wb + (size_t)c * blk_b(NT), blk_b(NT) * sizeof(float), &full[s]);
    };
    issue(0);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 1;
      if (c + 1 < nchunks) {
        if (c + 1 >= 2) mbar_wait(&empty[(c + 1) & 1], (uint32_t)((((c + 1) >> 1) - 1) & 1));   // MMAs of chunk c-1 retired
        issue(c + 1);
      }
      mbar_wait(&full[s], (uint32_t)((c >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t a_hi = smem_u32(stage[s]), a_lo = a_hi + TC_M * TC_KC * 4, b_hi = a_hi + BLK_A * 4, b_lo = b_hi + NT * TC_KC * 4;
#pragma unroll
      for (int kk = 0; kk < TC_KC / 8; ++kk) {
        const uint32_t o = kk * 256;
        umma_tf32(tmem_base, umma_smem_desc(a_hi + o, 128, 1024), umma_smem_desc(b_hi + o, 128, 1024), idesc, (c > 0 || kk > 0) ? 1u : 0u);
        umma_tf32(tmem_base, umma_smem_desc(a_lo + o, 128, 1024), umma_smem_desc(b_hi + o, 128, 1024), idesc, 1u);
        umma_tf32(tmem_base, umma_smem_desc(a_hi + o, 128, 1024), umma_smem_desc(b_lo + o, 128, 1024), idesc, 1u);
      }
      umma_commit(&empty[s]);
    }
    umma_commit(&done);                                            // arrives when every MMA above has retired
  }
  if (warp == 0) __syncwarp();                                     // lanes 1..31 of the issuing warp park here instead of spinning
  // a single-use barrier: waiting for a *future* phase of the ping-pong barriers would return at once (parity aliasing)
  mbar_wait(&done, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const int rt = 32 * warp + lane, row = mt * TC_M + rt;          // output row = TMEM lane
  const uint32_t tlane = (uint32_t)(32 * warp) << 16;
  if (!HEAD) {
    const int nchunks_next = p.N / TC_KC;
#pragma unroll 1
    for (int j = 0; j < NT / 32; ++j) {
      float v[32];
      tmem_ld32(tmem_base + tlane + (uint32_t)(j * 32), v);
      float* blk = p.Yb + ((size_t)mt * nchunks_next + (nt * NT) / TC_KC + j) * BLK_A;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 h4, l4;
        float* ph = reinterpret_cast<float*>(&h4);
        float* pl = reinterpret_cast<float*>(&l4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = v[4 * q + e] + __ldg(p.B + nt * NT + j * 32 + 4 * q + e);
          x = x / (1.f + __expf(-x));                                // swish
          split_tf32(x, ph[e], pl[e]);
        }
        *reinterpret_cast<float4*>(blk + blk_off(rt, 4 * q)) = h4;
        *reinterpret_cast<float4*>(blk + TC_M * TC_KC + blk_off(rt, 4 * q)) = l4;
      }
    }
  } else {
    float v[32];
    tmem_ld32(tmem_base + tlane, v);
    if (row < p.M) {
      const int na = p.na;
      float lp = 0.f;
      RKey key; key.a = 0; key.b = 0;
      if (!p.deterministic) { key.a = p.keys[2 * row]; key.b = p.keys[2 * row + 1]; }
#pragma unroll
      for (int a = 0; a < 16; ++a) {
        if (a < na) {
          const float loc = v[a] + __ldg(p.B + a);
          float rawv = loc;
          if (!p.deterministic) {
            const float sp = v[na + a] + __ldg(p.B + na + a);
            const float scale = (sp > 20.f ? sp : log1pf(__expf(sp))) + 0.001f;
            RKey blk = rblock(key, (uint32_t)a);
            const float lo = -0.99999994f;
            const float u = fmaxf(lo, bits_unit(blk.a ^ blk.b) * (1.0f - lo) + lo);
            const float z = 1.41421356237f * erfinv_giles(u);
            rawv = loc + scale * z;
            const float m2 = -2.f * rawv;
            lp += -0.5f * z * z - __logf(scale) - 0.91893853320467f - 2.f * (0.69314718056f - rawv - (m2 > 20.f ? m2 : log1pf(__expf(m2))));
          }
          if (p.action) p.action[(size_t)row * na + a] = tanhf(rawv);
          if (p.raw) p.raw[(size_t)row * na + a] = rawv;
        }
      }
      if (p.logp) p.logp[row] = lp;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kCols) : "memory");
}

template <int NT, bool HEAD>
static cudaError_t launch_dense(const DenseParams& p, cudaStream_t st) {
  const int smem = 2 * (BLK_A + blk_b(NT)) * (int)sizeof(float);
  static unsigned long long attr_done = 0ull;                      // one bit per device: the opt-in is per device (context)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done >> (dev & 63)) & 1ull)) { cudaFuncSetAttribute(k_dense_tc<NT, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_done |= 1ull << (dev & 63); }
  dim3 grid((p.M + TC_M - 1) / TC_M, (p.N + NT - 1) / NT);
  k_dense_tc<NT, HEAD><<<grid, TC_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

extern "C" int oduck_policy_invalidate(OduckHandle* h) {
  if (!h) return oduck_fail(ODUCK_ERR_ARG, "oduck_policy_invalidate: bad argument");
  h->policy_packed_for = nullptr;
  return ODUCK_OK;
}

extern "C" int oduck_policy_forward(OduckHandle* h, const OduckPolicyWeights* w, const float* obs, const uint32_t* keys, int deterministic,
                                    float* action, float* raw_action, float* log_prob, void* stream) {
  if (!h || !w) return oduck_fail(ODUCK_ERR_ARG, "oduck_policy_forward: bad argument");
  if (!deterministic && !keys) return oduck_fail(ODUCK_ERR_ARG, "oduck_policy_forward: stochastic policy needs keys");
  const bool tc_shape = w->hidden[0] % 128 == 0 && w->hidden[1] % 128 == 0 && w->hidden[2] % 128 == 0 && w->out_dim <= 32 && (w->out_dim & 1) == 0 && w->out_dim / 2 <= 16;
  if (!tc_shape) return oduck_fail(ODUCK_ERR_UNSUPPORTED, "oduck_policy_forward: hidden sizes must be multiples of 128 and out_dim <= 32 (reference: 512, 256, 128 -> 28)");
  if (cudaSetDevice(h->device) != cudaSuccess) return oduck_fail(ODUCK_ERR_CUDA, "cudaSetDevice failed");
  const float* obs_in = obs ? obs : h->obs_state;
  // repacked weights (cached per weight pointer), pack obs, 3 dense layers, dense + NormalTanh head
  cudaStream_t st = (cudaStream_t)stream;
  const int dims[5] = {w->obs_dim, w->hidden[0], w->hidden[1], w->hidden[2], w->out_dim};
  const int nts[4] = {128, 128, 128, 32};
  const int mtiles = (h->n + TC_M - 1) / TC_M;
  size_t woff[5] = {0}, aoff[5] = {0};
  for (int l = 0; l < 4; l++) {
    const int nch = (dims[l] + TC_KC - 1) / TC_KC, ntl = (dims[l + 1] + nts[l] - 1) / nts[l];
    woff[l + 1] = woff[l] + (size_t)ntl * nch * blk_b(nts[l]);
    aoff[l + 1] = aoff[l] + (size_t)mtiles * nch * BLK_A;
  }
  const size_t need = woff[4] + aoff[4];
  if (h->policy_scratch_floats < need) {
    if (h->policy_scratch) cudaFree(h->policy_scratch);
    if (cudaMalloc((void**)&h->policy_scratch, need * sizeof(float)) != cudaSuccess) { h->policy_scratch = nullptr; h->policy_scratch_floats = 0; return oduck_fail(ODUCK_ERR_ALLOC, "oduck_policy_forward: scratch cudaMalloc failed"); }
    h->policy_scratch_floats = need;
    h->policy_packed_for = nullptr;
  }
  float* wbuf = h->policy_scratch; float* abuf = wbuf + woff[4];
  cudaError_t e = cudaSuccess;
  const bool prepacked = w->packed[0] && w->packed[1] && w->packed[2] && w->packed[3];   // the learner's operand blocks, always current
  if (!prepacked && h->policy_packed_for != w->w[0]) {          // new weight buffers (PolicyWeights.refresh hands out fresh ones): repack once
    for (int l = 0; l < 4 && e == cudaSuccess; l++) {
      k_pack_weights<<<64, 256, 0, st>>>(w->w[l], wbuf + woff[l], dims[l], dims[l + 1], nts[l]);
      e = cudaGetLastError();
      h->launches++;
    }
    h->policy_packed_for = w->w[0];
  }
  if (e == cudaSuccess) { k_pack_obs<<<128, 256, 0, st>>>(obs_in, obs ? w->obs_dim : ODUCK_OBS_STATE /* the handle's records are ODUCK_OBS_STATE apart */, w->obs_mean, w->obs_std, abuf + aoff[0], h->n, dims[0]); e = cudaGetLastError(); }
  DenseParams d;
  memset(&d, 0, sizeof(d));
  d.M = h->n;
  for (int l = 0; l < 3 && e == cudaSuccess; l++) {
    // hidden layers: the learner's GEMM kernel (3-stage TMA ring, 8-warp epilogue) with the activation-only epilogue
    GemmParams g;
    memset(&g, 0, sizeof(g));
    g.A = abuf + aoff[l]; g.B = prepacked ? w->packed[l] : wbuf + woff[l];
    g.nchunks = (dims[l] + TC_KC - 1) / TC_KC; g.cps = g.nchunks;
    g.bias = w->b[l]; g.nvalid = dims[l + 1];
    g.Yr = abuf + aoff[l + 1]; g.yr_nch = dims[l + 1] / TC_KC;
    e = launch_gemm<128, 3, EPI_ACT>(g, mtiles, dims[l + 1] / 128, false, st);
  }
  d.Xb = abuf + aoff[3]; d.Wb = prepacked ? w->packed[3] : wbuf + woff[3]; d.B = w->b[3]; d.Yb = nullptr; d.K = dims[3]; d.N = dims[4];
  d.keys = keys; d.action = action; d.raw = raw_action; d.logp = log_prob; d.deterministic = deterministic; d.na = w->out_dim / 2;
  if (e == cudaSuccess) e = launch_dense<32, true>(d, st);
  if (e != cudaSuccess) return oduck_fail(ODUCK_ERR_CUDA, std::string("oduck_policy_forward launch: ") + cudaGetErrorString(e));
  h->launches += 5;
  return ODUCK_OK;
}

// A17: one step of the PPO unroll (common/runner.py:104-118 -> Brax generate_unroll): sample from the handle's own obs["state"],
// step, store the Transition in the attached sink.  The actor's head writes raw action / log-prob into slot t, k_step the rest.
extern "C" int oduck_step_into_sink(OduckHandle* h, const float* action, int t, void* stream);
extern "C" int oduck_rollout_step(OduckHandle* h, const OduckPolicyWeights* w, const uint32_t* keys, int t, void* stream) {
  if (!h || !w || !keys) return oduck_fail(ODUCK_ERR_ARG, "oduck_rollout_step: bad argument");
  const OduckRolloutSink& k = h->sink;
  if (!k.obs_policy || t < 0 || t >= k.unroll) return oduck_fail(ODUCK_ERR_ARG, "oduck_rollout_step: no sink attached or t outside the unroll");
  if (w->obs_dim != k.policy_dim) return oduck_fail(ODUCK_ERR_ARG, "oduck_rollout_step: the policy's input width is not the sink's policy_dim");
  const int nu = w->out_dim / 2;
  const size_t row = (size_t)t * k.num_envs + k.env_offset;
  int rc = oduck_policy_forward(h, w, nullptr, keys, 0, h->act_buf, k.raw_action + row * nu, k.log_prob + row, stream);
  if (rc) return rc;
  return oduck_step_into_sink(h, h->act_buf, t, stream);
}
