// oduck_handle.cuh -- host-side handle shared by the translation units of liboduck_cuda.so
#pragma once
#include <stdint.h>

#include "../../include/oduck.h"
#include "oduck_env.cuh"

struct Params {
  const DevModel* model;
  const DevEnvCfg* cfg;
  const float* poly;
  float *phys, *dr, *out, *info, *obs_state, *obs_priv, *reward, *done, *trunc, *metrics;
  float *first_phys, *first_obs_state, *first_obs_priv, *dbg;
  const float* action;     // step: [N, nu];  physics: ctrl [N, nu] or null
  const uint32_t* keys;
  const uint8_t* mask;
  const DevFF* ffmodel;
  float* ffscratch;
  int N, nsub, integrate;
  const DevHF* hfmodel;    // height-field floor (HF kernel instantiations only); last, so that the flat kernels' argument layout is unchanged
  float* hfscratch;        // per env: HF_SCRATCH floats of contact candidates
  const DevRewardLib* rlib;   // reward-library parameters (RL kernel instantiations only)
  // rollout sink (oduck_rollout_step): rows of THIS step, already offset to the handle's first env; null = no sink
  float *sk_obs_p, *sk_obs_v;      // slot t + 1, row widths sk_dp / sk_dv
  float *sk_reward, *sk_done, *sk_trunc;   // slot t
  int sk_dp, sk_dv;
};

struct OduckHandle {
  int n, device;
  OduckModel hm;
  OduckEnvConfig hcfg;
  DevModel hdm;
  DevEnvCfg hdc;
  DevModel* dmodel;
  DevEnvCfg* dcfg;
  float* poly;
  float *phys, *dr, *out, *info, *obs_state, *obs_priv, *reward, *done, *trunc, *metrics, *first_phys, *first_obs_state, *first_obs_priv, *dbg;
  int nefc, smem_bytes, grid;
  int64_t launches;
  DevFF* dff;
  float* ffscratch;               // per env: 12 x 32 foot-foot Jacobian rows + SAT operands (rare path)
  DevHF* dhf;                     // height-field floor: descriptor, elevation samples, per-env candidate lists
  float* hfdata;
  float* hfscratch;
  DevRewardLib* drlib;            // reward-library parameters; null when no library term is switched on
  float* policy_scratch;          // hidden activations of the actor MLP (tensor-core path)
  size_t policy_scratch_floats;
  const float* policy_packed_for;  // w[0] pointer the packed weight copy was made from
  OduckRolloutSink sink;          // attached rollout buffers (sink.obs_policy == null: none)
  float* act_buf;                 // [N, nu] squashed actions between the actor's head and k_step inside oduck_rollout_step
};

