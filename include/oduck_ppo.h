/* oduck_ppo.h -- C ABI of the on-device PPO learner step: the step right after the rollout path.
 *
 * What it replaces in the reference: the body of Brax `ppo.train`'s `minibatch_step` as driven from
 * playground/common/runner.py:104-118 -- `compute_ppo_loss` (value + policy forward, GAE, clipped surrogate,
 * entropy bonus), `jax.grad`, `optax.clip_by_global_norm(max_grad_norm)` and `optax.adam(learning_rate)` -- for the
 * network factory of runner.py:94-100 (policy 101-512-256-128-28, value 212-512-256-128-1, swish) with the
 * hyper-parameters of `locomotion_params.brax_ppo_config` (runner.py:87-89).  The arithmetic lives in the un-vendored
 * third-party package `brax`; the checker for this path is the PyTorch fp32 twin in
 * open_duck_playground_b200/ppo.py (`PPOTrainer._minibatch_loss` + autograd), see tests/test_ppo_device.py.
 *
 * One `oduck_ppo_minibatch` call = one SGD step on one minibatch of B env trajectories of T transitions:
 *   gather + normalise + pack -> 2 x 4 dense layers (tcgen05, 3xTF32) -> GAE + loss + head gradients ->
 *   2 x 7 backward GEMMs (tcgen05, split-K) -> gradient reduce (+ optional all-reduce by the caller) -> clip + Adam +
 *   repack of the weights for the next forward.  Everything is asynchronous on `stream`; no allocation after create.
 *
 * Conventions as in oduck.h: return 0 on success, negative code otherwise, message via oduck_last_error().
 * All pointers are CUDA device pointers unless stated otherwise.
 */
#ifndef ODUCK_PPO_H_
#define ODUCK_PPO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODUCK_PPO_POLICY 0
#define ODUCK_PPO_VALUE 1

typedef struct OduckPpoConfig {
  int32_t batch_envs;           /* B: env trajectories per minibatch (num_envs / num_minibatches = 256) */
  int32_t unroll;               /* T: unroll_length = 20 */
  int32_t num_actions;          /* nu = 14 (policy head = 2 * nu) */
  int32_t policy_dims[5];       /* 101, 512, 256, 128, 28 */
  int32_t value_dims[5];        /* 212, 512, 256, 128, 1 */
  int32_t normalize_advantage;  /* Brax default: 1 */
  float discounting;            /* 0.97 */
  float gae_lambda;             /* 0.95 */
  float clipping_epsilon;       /* 0.2 */
  float entropy_cost;           /* 0.005 */
  float reward_scaling;         /* 1.0 */
  float learning_rate;          /* 3e-4 */
  float max_grad_norm;          /* 1.0; <= 0 disables clipping */
  float adam_b1, adam_b2, adam_eps;   /* optax.adam defaults 0.9, 0.999, 1e-8 */
  int32_t matmul_tf32;          /* 0 (default): fp32-faithful GEMMs, 3 tensor-core passes on hi / lo tf32 halves (the parity-tested mode).
                                 * 1: ONE tf32 pass (operands truncated to tf32, fp32 accumulate) -- what XLA runs for f32 dots on NVIDIA
                                 * GPUs at jax's default matmul precision, i.e. the reference's own arithmetic for ppo.train
                                 * (common/runner.py:104-118); a third of the MMAs and half the operand traffic.  The rollout actor
                                 * always runs the fp32-faithful form (its log-prob must match the learner's recomputation). */
} OduckPpoConfig;

/* The rollout of one training step, time-major like Brax's Transition pytree: [T(+1)][N][...]. */
typedef struct OduckRollout {
  int32_t num_envs;             /* N */
  int32_t unroll;               /* T (must equal cfg.unroll) */
  const float* obs_policy;      /* [T+1][N][policy_dims[0]]  obs["state"] before each step, + the final one */
  const float* obs_value;       /* [T+1][N][value_dims[0]]   obs["privileged_state"] */
  const float* raw_action;      /* [T][N][nu]   pre-tanh action sampled by the behaviour policy */
  const float* log_prob;        /* [T][N]       behaviour log-prob */
  const float* reward;          /* [T][N] */
  const float* done;            /* [T][N]       1 - discount */
  const float* truncation;      /* [T][N] */
  /* Blocked layout (the buffer one NCCL all-gather leaves behind): the N envs come as N / block_envs blocks of
   * block_envs envs, block b = one rank's rollout buffers, every field [T(+1)][block_envs][...] inside it, the blocks
   * block_stride floats apart; the pointers above address block 0.  Env e of field f at time t:
   * f + (e / block_envs) * block_stride + (t * block_envs + e % block_envs) * width.  block_envs = 0: one block of N envs. */
  int32_t block_envs;
  int64_t block_stride;
  /* Floats between consecutive env rows of obs_policy; 0 = the policy observation size (a dense [T + 1][N][policy_dim] tensor).
   * Both reference envs build the value observation as hstack([state, privileged tail]) (joystick.py:596-615, standing.py), so the
   * policy observation is its first policy_dim columns: a caller may pass obs_policy = obs_value with obs_policy_ld = value_dim and
   * keep / all-gather ONE observation tensor instead of two (30 % fewer bytes in the exchange of SURVEY 8e). */
  int32_t obs_policy_ld;
} OduckRollout;

/* Observation normaliser (brax running_statistics): obs_n = (obs - mean) / std per feature. */
typedef struct OduckNormalizer {
  const float* policy_mean; const float* policy_std;   /* [policy_dims[0]] */
  const float* value_mean;  const float* value_std;    /* [value_dims[0]]  */
} OduckNormalizer;

typedef struct OduckPpo OduckPpo;

/* stages of oduck_ppo_minibatch (bit mask); ODUCK_PPO_ALL is the product path, the others exist for the parity tests */
#define ODUCK_PPO_STAGE_FORWARD 1   /* gather/pack + both forward passes            -> buffers LOGITS, VALUES */
#define ODUCK_PPO_STAGE_LOSS 2      /* GAE + loss + head gradients                   -> buffer LOSSES */
#define ODUCK_PPO_STAGE_BACKWARD 4  /* backward GEMMs + gradient reduce              -> buffer GRADS */
#define ODUCK_PPO_STAGE_ADAM 8      /* global-norm clip + Adam + weight repack       -> buffer PARAMS */
#define ODUCK_PPO_ALL 15
#define ODUCK_PPO_DEBUG_SIMT 256    /* run the GEMMs on CUDA cores (same operands, same epilogues): bisects tcgen05 problems */
#define ODUCK_PPO_NO_COOP 512       /* use the two-kernel reduce + Adam tail instead of the fused cooperative launch (for callers
                                     * that capture the call into a CUDA graph and prefer plain kernel nodes) */

typedef enum {
  ODUCK_PPO_BUF_PARAMS = 0,   /* f32 [P]  master weights, flat: for net in (policy, value): for layer: W[in][out] (flax), b[out] */
  ODUCK_PPO_BUF_GRADS,        /* f32 [P]  gradient of the last minibatch (before clipping), same layout */
  ODUCK_PPO_BUF_ADAM_M,       /* f32 [P] */
  ODUCK_PPO_BUF_ADAM_V,       /* f32 [P] */
  ODUCK_PPO_BUF_LOGITS,       /* f32 [Mp_pad][32]  policy head output of the last minibatch, row = t * B + b */
  ODUCK_PPO_BUF_VALUES,       /* f32 [Mv_pad][32]  column 0 = value, row = t * B + b, t = 0..T */
  ODUCK_PPO_BUF_LOSSES,       /* f64 [8]  total, policy, value, entropy, mean |adv| (raw), clip fraction, -, - of the last minibatch */
  ODUCK_PPO_BUF_ADV,          /* f32 [T][B] normalised advantages of the last minibatch */
  ODUCK_PPO_BUF_VS,           /* f32 [T][B] value targets */
  ODUCK_PPO_BUF_STEP,         /* i32 [1]  Adam step count */
  ODUCK_PPO_BUF_COUNT
} OduckPpoBufferId;

int oduck_ppo_create(const OduckPpoConfig* cfg, int device, OduckPpo** out);
int oduck_ppo_destroy(OduckPpo* p);
/* Number of parameters P and offset/shape of one tensor inside the flat layout (which: 0 = kernel [in][out], 1 = bias). */
int64_t oduck_ppo_num_params(const OduckPpo* p);
int oduck_ppo_param_info(const OduckPpo* p, int net, int layer, int which, int64_t* offset, int64_t* rows, int64_t* cols);
/* Copy flat parameters in (device pointer, layout above), reset Adam state if reset_opt != 0, and repack the GEMM operands. */
int oduck_ppo_set_params(OduckPpo* p, const float* flat_params, int reset_opt, void* stream);
/* The forward-pass operand form of one kernel matrix (hi/lo tf32 blocks), rewritten by every Adam step: hand it to
 * OduckPolicyWeights.packed[] and the rollout actor reads the learner's weights without a repack or a copy. */
int oduck_ppo_packed_weights(OduckPpo* p, int net, int layer, const float** ptr);
/* Zero-copy view of a learner buffer: ptr, element count, dtype (ODUCK_DTYPE_*). */
int oduck_ppo_get_buffer(OduckPpo* p, int id, void** ptr, int64_t* count, int* dtype);
/* One SGD step on the minibatch made of env trajectories env_idx[0..B) (i32, device).  entropy_noise: optional f32
 * [T * B][nu] standard normals for Brax's sampled entropy term (row = t * B + b); NULL = drawn in-kernel from
 * entropy_key (u32[2] in DEVICE memory, so that a captured CUDA graph sees a fresh key on every replay).  stages: ODUCK_PPO_ALL, or a prefix of the pipeline (parity tests).
 * With world_size > 1 the caller runs stages FORWARD|LOSS|BACKWARD, all-reduces buffer GRADS (mean), then ADAM. */
int oduck_ppo_minibatch(OduckPpo* p, const OduckRollout* rollout, const OduckNormalizer* norm, const int32_t* env_idx,
                        const float* entropy_noise, const uint32_t* entropy_key, int stages, void* stream);
/* Optional, call it right AFTER the oduck_ppo_minibatch of the current minibatch (same stream, stages including FORWARD): packs the
 * observation operands of the NEXT minibatch (env trajectories next_env_idx[0..B)) into the learner's alternate input buffers on
 * a side stream forked from the START of the current minibatch, i.e. beside its kernels.  The next oduck_ppo_minibatch call
 * whose env_idx pointer equals next_env_idx waits for that side stream and skips its own gather / normalise / pack pass (13 us
 * on the critical path of a 0.2 ms step).  The rollout, the normaliser and next_env_idx must not change until then.  Capturable
 * like oduck_ppo_minibatch (inside one capture: never before the first oduck_ppo_minibatch, never after the last). */
int oduck_ppo_prefetch(OduckPpo* p, const OduckRollout* rollout, const OduckNormalizer* norm, const int32_t* next_env_idx, void* stream);
/* Kernels launched since create (bench `gpu_launches`). */
int64_t oduck_ppo_launch_count(const OduckPpo* p);

#ifdef __cplusplus
}
#endif
#endif /* ODUCK_PPO_H_ */
