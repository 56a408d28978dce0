// oduck_policy.cu -- A15: actor-MLP forward of the Brax PPO policy (common/runner.py:94-100; architecture mirrored in
// common/export_onnx.py:14-72): x = (obs - mean) / std; 3 x (Dense + swish); Dense -> (loc, scale);
// action = tanh(loc + (softplus(scale) + 0.001) * N(0,1)), log-prob with the tanh Jacobian (Brax NormalTanhDistribution).
// fp32 SIMT version: one CTA = 16 envs, activations ping-pong in shared memory, every weight is read once per CTA
// (coalesced over the output column) and reused for the 16 rows from registers.
#include <cuda_runtime.h>

#include <string>

#include "oduck_handle.cuh"

#define PT 16          // envs per CTA
#define PTHREADS 256

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

struct PolicyParams {
  const float *obs, *mean, *std, *w[4], *b[4];
  const uint32_t* keys;
  float *action, *raw, *logp;
  int n, dims[5], deterministic, obs_stride;
};

template <int COLS>   // columns per thread
__device__ __forceinline__ void dense_layer(const float* __restrict__ W, const float* __restrict__ B, const float* xin, float* xout, int K, int Nc, bool act) {
  // thread t owns columns t, t + 256, ... (COLS of them) for all PT rows
  float acc[COLS][PT];
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    const int col = threadIdx.x + c * PTHREADS;
    const float bias = col < Nc ? B[col] : 0.f;
#pragma unroll
    for (int r = 0; r < PT; ++r) acc[c][r] = bias;
  }
  for (int k = 0; k < K; ++k) {
    float wv[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) { const int col = threadIdx.x + c * PTHREADS; wv[c] = col < Nc ? __ldg(W + (size_t)k * Nc + col) : 0.f; }
#pragma unroll
    for (int r = 0; r < PT; ++r) {
      const float xv = xin[r * 512 + k];
#pragma unroll
      for (int c = 0; c < COLS; ++c) acc[c][r] = fmaf(xv, wv[c], acc[c][r]);
    }
  }
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    const int col = threadIdx.x + c * PTHREADS;
    if (col < Nc) {
#pragma unroll
      for (int r = 0; r < PT; ++r) {
        float v = acc[c][r];
        if (act) v = v / (1.f + __expf(-v));   // swish
        xout[r * 512 + col] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(PTHREADS, 2) k_policy(PolicyParams p) {
  extern __shared__ float sm[];
  float* x0 = sm;              // [PT][512]
  float* x1 = sm + PT * 512;
  const int env0 = blockIdx.x * PT;
  for (int i = threadIdx.x; i < PT * p.dims[0]; i += PTHREADS) {
    const int r = i / p.dims[0], k = i % p.dims[0], e = env0 + r;
    float v = 0.f;
    if (e < p.n) v = (p.obs[(size_t)e * p.obs_stride + k] - p.mean[k]) / p.std[k];
    x0[r * 512 + k] = v;
  }
  __syncthreads();
  dense_layer<2>(p.w[0], p.b[0], x0, x1, p.dims[0], p.dims[1], true);
  __syncthreads();
  dense_layer<1>(p.w[1], p.b[1], x1, x0, p.dims[1], p.dims[2], true);
  __syncthreads();
  dense_layer<1>(p.w[2], p.b[2], x0, x1, p.dims[2], p.dims[3], true);
  __syncthreads();
  dense_layer<1>(p.w[3], p.b[3], x1, x0, p.dims[3], p.dims[4], false);
  __syncthreads();
  // NormalTanh head: thread = (row, action)
  const int na = p.dims[4] / 2;
  for (int i = threadIdx.x; i < PT * 32; i += PTHREADS) {
    const int r = i >> 5, a = i & 31, e = env0 + r;
    float lp = 0.f;
    if (a < na && e < p.n) {
      const float loc = x0[r * 512 + a];
      float raw = loc;
      if (!p.deterministic) {
        const float sp = x0[r * 512 + na + a];
        const float scale = (sp > 20.f ? sp : log1pf(__expf(sp))) + 0.001f;
        RKey k; k.a = p.keys[2 * e]; k.b = p.keys[2 * e + 1];
        RKey blk = rblock(k, (uint32_t)a);
        // jax.random.normal: sqrt(2) * erf_inv(uniform(-1 + ulp, 1))
        const float lo = -0.99999994f;
        float u = fmaxf(lo, bits_unit(blk.a ^ blk.b) * (1.0f - lo) + lo);
        const float z = 1.41421356237f * erfinv_giles(u);
        raw = loc + scale * z;
        const float lpn = -0.5f * z * z - __logf(scale) - 0.91893853320467f;
        const float ldj = 2.f * (0.69314718056f - raw - (-2.f * raw > 20.f ? -2.f * raw : log1pf(__expf(-2.f * raw))));
        lp = lpn - ldj;
      }
      if (p.action) p.action[(size_t)e * na + a] = tanhf(raw);
      if (p.raw) p.raw[(size_t)e * na + a] = raw;
    }
    // log-prob = sum over the env's actions (one warp = one row)
#pragma unroll
    for (int o = 16; o; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
    if (a == 0 && e < p.n && p.logp) p.logp[e] = lp;
  }
}

extern "C" int oduck_policy_forward(OduckHandle* h, const OduckPolicyWeights* w, const float* obs, const uint32_t* keys, int deterministic,
                                    float* action, float* raw_action, float* log_prob, void* stream) {
  if (!h || !w) return oduck_fail(ODUCK_ERR_ARG, "oduck_policy_forward: bad argument");
  if (!deterministic && !keys) return oduck_fail(ODUCK_ERR_ARG, "oduck_policy_forward: stochastic policy needs keys");
  if (w->obs_dim > 512 || w->hidden[0] > 512 || w->hidden[1] > 256 || w->hidden[2] > 256 || w->out_dim > 64 || (w->out_dim & 1))
    return oduck_fail(ODUCK_ERR_UNSUPPORTED, "oduck_policy_forward: layer sizes exceed the kernel's tile (101 -> 512 -> 256 -> 128 -> 28 class)");
  PolicyParams p;
  p.obs = obs ? obs : h->obs_state;
  p.obs_stride = w->obs_dim;
  p.mean = w->obs_mean; p.std = w->obs_std;
  for (int l = 0; l < 4; l++) { p.w[l] = w->w[l]; p.b[l] = w->b[l]; }
  p.keys = keys; p.action = action; p.raw = raw_action; p.logp = log_prob;
  p.n = h->n; p.deterministic = deterministic;
  p.dims[0] = w->obs_dim; p.dims[1] = w->hidden[0]; p.dims[2] = w->hidden[1]; p.dims[3] = w->hidden[2]; p.dims[4] = w->out_dim;
  if (cudaSetDevice(h->device) != cudaSuccess) return oduck_fail(ODUCK_ERR_CUDA, "cudaSetDevice failed");
  const int smem = 2 * PT * 512 * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_policy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_set = true; }
  k_policy<<<(h->n + PT - 1) / PT, PTHREADS, smem, (cudaStream_t)stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return oduck_fail(ODUCK_ERR_CUDA, std::string("k_policy launch: ") + cudaGetErrorString(e));
  h->launches++;
  return ODUCK_OK;
}
