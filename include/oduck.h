/*
 * oduck.h -- C-ABI of the B200-native batched Open Duck Mini V2 joystick hot path.
 *
 * Drop-in boundary for the one data-parallel path of apirrone/Open_Duck_Playground
 * (SURVEY.md section 8b).  The reference has no FFI of its own (it is pure Python
 * on top of MJX/Brax), so every entry point cites the *Python* interface it
 * replaces (paths relative to the reference repo, playground/...):
 *
 *   oduck_create            OpenDuckMiniV2Env.__init__ + mjx.put_model     open_duck_mini_v2/base.py:44-61
 *                           Joystick._post_init                            open_duck_mini_v2/joystick.py:121-204
 *   oduck_randomize         randomize.domain_randomize                     common/randomize.py:26-146
 *   oduck_reset             Joystick.reset (+ wrapper first_state store)   open_duck_mini_v2/joystick.py:206-321   (Standing: standing.py:200-318)
 *   oduck_step              Joystick.step wrapped by wrap_for_brax_training open_duck_mini_v2/joystick.py:323-481, common/runner.py:117   (Standing: standing.py:320-443)
 *   oduck_physics_substeps  mjx_env.step(model, data, ctrl, n_substeps)    open_duck_mini_v2/joystick.py:420
 *   oduck_forward           mjx_env.init's mjx.forward                     open_duck_mini_v2/joystick.py:258
 *   oduck_policy_forward    Brax make_ppo_networks policy apply            common/runner.py:94-100, common/export_onnx.py:64-72
 *   oduck_rollout_step      one step of Brax generate_unroll (policy +     common/runner.py:104-118
 *   oduck_set_rollout_sink    env.step + Transition store) of ppo.train
 *   oduck_get_buffer        attribute access on mjx.Data / State / info    open_duck_mini_v2/joystick.py:278-321
 *   oduck_set_state         state.data.replace(qpos=..., qvel=...)         open_duck_mini_v2/joystick.py:399
 *
 * Conventions: every call returns 0 on success or a negative OduckStatus; the
 * message is available from oduck_last_error() (thread local).  The library
 * owns all per-env state (device memory, one record per env);
 * the caller owns action / key / weight / output pointers (device pointers for
 * the CUDA library, host pointers for the CPU oracle).  `stream` is a
 * cudaStream_t (ignored by the oracle).  Calls are asynchronous on that stream:
 * no hidden synchronisation, no allocation after oduck_create.  A handle is not
 * thread-safe; distinct handles are independent.  No torch types appear here.
 *
 * The same header is implemented twice:
 *   open_duck_playground_b200/csrc/  -> liboduck_cuda.so  (the product, sm_100a)
 *   oracle/                          -> liboduck_oracle.so (CPU fp64 checker; test infrastructure only)
 */
#ifndef ODUCK_H_
#define ODUCK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODUCK_ABI_VERSION 9

#define ODUCK_MAX_BODY 20
#define ODUCK_MAX_JNT 28
#define ODUCK_MAX_NQ 36
#define ODUCK_MAX_NV 32
#define ODUCK_MAX_NU 16
#define ODUCK_MAX_SITE 8
#define ODUCK_MAX_VERT 32      /* convex-hull vertices per foot */
#define ODUCK_MAX_FACE 64      /* convex-hull faces per foot (triangles) */
#define ODUCK_MAX_PLANE 32     /* merged (coplanar) polygon faces per foot */
#define ODUCK_MAX_PVERT 8      /* vertices per polygon face */
#define ODUCK_MAX_EDGE 48      /* hull edges between two different polygon faces */
#define ODUCK_NFEET 2
#define ODUCK_CON_PER_PAIR 4   /* MJX emits 4 manifold points per geom pair */
#define ODUCK_MAX_CON 12       /* 2 x plane/hfield-foot + 1 x foot-foot */
#define ODUCK_REF_DIM 40       /* reference-motion frame width */
#define ODUCK_POLY_DEG 16      /* coefficients per polynomial */
#define ODUCK_OBS_STATE 101    /* Joystick obs["state"]; also the capacity of the per-env obs records (Standing: 85) */
#define ODUCK_OBS_PRIV 212     /* Joystick obs["privileged_state"] (Standing: 153) */
#define ODUCK_NMETRIC 8
#define ODUCK_TASK_JOYSTICK 0  /* open_duck_mini_v2/joystick.py */
#define ODUCK_TASK_STANDING 1  /* open_duck_mini_v2/standing.py: same physics, other rewards / observation */
#define ODUCK_NCMD 7

typedef enum {
  ODUCK_OK = 0,
  ODUCK_ERR_ARG = -1,
  ODUCK_ERR_CUDA = -2,
  ODUCK_ERR_MODEL = -3,
  ODUCK_ERR_ALLOC = -4,
  ODUCK_ERR_UNSUPPORTED = -5
} OduckStatus;

/* Joint types use MuJoCo's enum values (base.py:99 relies on free == 0, hinge == 3). */
#define ODUCK_JNT_FREE 0
#define ODUCK_JNT_HINGE 3

/* Compiled model: the subset of mjModel the path reads.  Filled
 * by open_duck_playground_b200/mjcf.py; all reals are double, the CUDA library
 * narrows to float at create time. */
typedef struct OduckModel {
  int32_t abi_version;
  int32_t nbody, njnt, nq, nv, nu, nsite;
  /* bodies (depth-first ids, world = 0) */
  int32_t body_parentid[ODUCK_MAX_BODY];
  int32_t body_jntadr[ODUCK_MAX_BODY];
  int32_t body_jntnum[ODUCK_MAX_BODY];
  int32_t body_dofadr[ODUCK_MAX_BODY];
  int32_t body_dofnum[ODUCK_MAX_BODY];
  double body_pos[ODUCK_MAX_BODY][3];
  double body_quat[ODUCK_MAX_BODY][4];
  double body_ipos[ODUCK_MAX_BODY][3];
  double body_iquat[ODUCK_MAX_BODY][4];
  double body_mass[ODUCK_MAX_BODY];
  double body_inertia[ODUCK_MAX_BODY][3];
  double body_invweight0[ODUCK_MAX_BODY][2];
  /* joints */
  int32_t jnt_type[ODUCK_MAX_JNT];
  int32_t jnt_qposadr[ODUCK_MAX_JNT];
  int32_t jnt_dofadr[ODUCK_MAX_JNT];
  int32_t jnt_bodyid[ODUCK_MAX_JNT];
  int32_t jnt_limited[ODUCK_MAX_JNT];
  double jnt_pos[ODUCK_MAX_JNT][3];
  double jnt_axis[ODUCK_MAX_JNT][3];
  double jnt_range[ODUCK_MAX_JNT][2];
  double qpos0[ODUCK_MAX_NQ];
  /* dofs */
  int32_t dof_bodyid[ODUCK_MAX_NV];
  int32_t dof_jntid[ODUCK_MAX_NV];
  int32_t dof_parentid[ODUCK_MAX_NV];
  double dof_armature[ODUCK_MAX_NV];
  double dof_damping[ODUCK_MAX_NV];
  double dof_frictionloss[ODUCK_MAX_NV];
  double dof_invweight0[ODUCK_MAX_NV];
  /* position actuators: force = kp*ctrl - kp*q - kv*qd, clipped */
  int32_t act_jntid[ODUCK_MAX_NU];
  double act_kp[ODUCK_MAX_NU];
  double act_kv[ODUCK_MAX_NU];
  double act_ctrlrange[ODUCK_MAX_NU][2];
  double act_forcerange[ODUCK_MAX_NU][2];
  /* sites */
  int32_t site_bodyid[ODUCK_MAX_SITE];
  double site_pos[ODUCK_MAX_SITE][3];
  double site_quat[ODUCK_MAX_SITE][4];
  int32_t imu_site;
  int32_t foot_site[ODUCK_NFEET];
  /* collision: floor (plane z=0 of the world, or height field) vs two convex feet */
  int32_t floor_is_hfield;
  double floor_friction;          /* wins by priority=1 for foot-floor pairs */
  int32_t foot_body[ODUCK_NFEET];
  int32_t foot_nvert;
  double foot_vert[ODUCK_NFEET][ODUCK_MAX_VERT][3]; /* hull vertices, BODY frame */
  int32_t foot_nface;
  int32_t foot_face[ODUCK_MAX_FACE][3];             /* hull triangles (shared topology) */
  /* polygon faces and edges of the hull (topology shared by both feet; normals per foot, BODY frame) for convex-convex */
  int32_t foot_nplane;
  int32_t foot_plane_nvert[ODUCK_MAX_PLANE];
  int32_t foot_plane_vert[ODUCK_MAX_PLANE][ODUCK_MAX_PVERT];   /* counter-clockwise seen from outside */
  double foot_plane_normal[ODUCK_NFEET][ODUCK_MAX_PLANE][3];
  int32_t foot_nedge;
  int32_t foot_edge_vert[ODUCK_MAX_EDGE][2];
  int32_t foot_edge_plane[ODUCK_MAX_EDGE][2];
  double foot_center[ODUCK_NFEET][3];                          /* bounding sphere (BODY frame) */
  double foot_radius;
  double foot_friction;           /* foot-foot pair: max of the two geoms */
  int32_t enable_foot_foot;       /* 1 = instantiate the convex-convex pair */
  /* options */
  double timestep;
  double gravity[3];
  double tolerance, ls_tolerance, impratio, meaninertia;
  int32_t iterations, ls_iterations;
  double solref[2];
  double solimp[5];
  /* keyframe "home" */
  double key_qpos[ODUCK_MAX_NQ];
  double key_ctrl[ODUCK_MAX_NU];
  /* height-field floor (floor_is_hfield; rough_terrain scenes, xmls/scene_rough_terrain_backlash.xml:22-27): MuJoCo hfield asset,
   * grid x in [-size[0], size[0]] over ncol samples, y in [-size[1], size[1]] over nrow samples, height = data * size[2],
   * solid down to -size[3].  Both libraries collide the foot faces with the terrain triangles under the foot;
   * the CUDA library runs the HF instantiations of its kernels for such models. */
  int32_t hfield_nrow, hfield_ncol;
  double hfield_size[4];         /* radius_x, radius_y, elevation_z, base_z */
  const float* hfield_data;      /* HOST pointer, elevation normalised to [0, 1], row-major [nrow][ncol]; copied by oduck_create */
} OduckModel;

/* Environment constants: Joystick.default_config() (joystick.py:49-102) plus the
 * tables _post_init derives (joystick.py:121-204). */
/* Reward library: the terms of common/rewards.py that neither Joystick nor Standing wires in (rewards.py:37-90,120,152-224),
 * selectable per task like the reference selects terms -- by a non-zero entry of reward_config.scales.  Inputs follow the
 * reference's accessors (open_duck_mini_v2/base.py:193-271): "joints" are the nu actuated joints, sensors are the imu-site
 * sensors, feet are the two foot sites.  Scaled terms are added to the task's own sum in enum order, before `* dt` and the
 * clip (joystick.py:444-447).  The terms are not reported in the metrics buffer.
 * reward_base_y_swing and reward_feet_phase take a time / a per-foot target height that no reference task computes (neither is
 * called anywhere in the reference; the second carries a FIXME).  They are offered with an EXPLICIT gait clock: the env's own
 * reference-motion phase counter i = info["imitation_i"] after this step's increment (joystick.py:352-356, period
 * nb_steps_in_period control steps = 0.54 s; 0 when the imitation reward is off).  t = i * ctrl_dt;  foot k's phase
 * phi_k = wrap(2 pi i / period + k pi) in [-pi, pi), rz_k = mujoco_playground gait.get_rz(phi_k, swing_height = max_foot_height)
 * (cubic Bezier 0 -> h over the first half period, h -> 0 over the second; restated from upstream). */
enum OduckLibTerm {
  ODUCK_LIB_ORIENTATION = 0,       /* cost_orientation(gravity sensor)          -- Joystick (joystick.py:645, commented out upstream) */
  ODUCK_LIB_LIN_VEL_Z,             /* cost_lin_vel_z(global_linvel)             rewards.py:37 */
  ODUCK_LIB_ANG_VEL_XY,            /* cost_ang_vel_xy(global_angvel)            rewards.py:41 */
  ODUCK_LIB_BASE_HEIGHT,           /* cost_base_height(qpos[2], target)         rewards.py:49 */
  ODUCK_LIB_ENERGY,                /* cost_energy(joint qvel, actuator_force)   rewards.py:73 */
  ODUCK_LIB_JOINT_POS_LIMITS,      /* cost_joint_pos_limits(joint qpos, soft)   rewards.py:85 */
  ODUCK_LIB_TERMINATION,           /* cost_termination(done)                    rewards.py:120 */
  ODUCK_LIB_JOINT_DEVIATION_HIP,   /* rewards.py:152 */
  ODUCK_LIB_JOINT_DEVIATION_KNEE,  /* rewards.py:161 */
  ODUCK_LIB_POSE,                  /* cost_pose(joint qpos, default, weights)   rewards.py:170 */
  ODUCK_LIB_FEET_SLIP,             /* cost_feet_slip(contact, global_linvel)    rewards.py:180 */
  ODUCK_LIB_FEET_CLEARANCE,        /* cost_feet_clearance(feet linvel sensors, feet site pos, max_foot_height)  rewards.py:187 */
  ODUCK_LIB_FEET_HEIGHT,           /* cost_feet_height(swing_peak, first_contact, max_foot_height)              rewards.py:202 */
  ODUCK_LIB_FEET_AIR_TIME,         /* reward_feet_air_time(feet_air_time, first_contact, command, thresholds)   rewards.py:212 */
  ODUCK_LIB_BASE_Y_SWING,          /* reward_base_y_swing(local_linvel y, freq, amplitude, t, tracking_sigma)    rewards.py:53; t = gait clock (below) */
  ODUCK_LIB_FEET_PHASE,            /* reward_feet_phase(feet site pos, rz)                                        rewards.py:228; rz from the gait clock */
  ODUCK_NLIBTERM
};
typedef struct OduckRewardLibrary {
  double scale[ODUCK_NLIBTERM];    /* 0 = term off (the default: the shipped tasks use none) */
  double base_height_target, max_foot_height, air_time_threshold_min, air_time_threshold_max;
  double base_y_swing_freq, base_y_swing_amplitude;   /* Hz (default 1 / gait period), m/s */
  double soft_lowers[ODUCK_MAX_NU], soft_uppers[ODUCK_MAX_NU], pose_weights[ODUCK_MAX_NU];
  int32_t n_hip, hip_indices[4], n_knee, knee_indices[4];   /* actuator indices */
} OduckRewardLibrary;

typedef struct OduckEnvConfig {
  int32_t task;                  /* ODUCK_TASK_*: which env class the step implements */
  int32_t n_substeps;            /* ctrl_dt / sim_dt = 10 */
  int32_t episode_length;        /* 1000 (EpisodeWrapper) */
  int32_t use_imitation_reward;  /* joystick.py:45 */
  int32_t use_motor_speed_limits;/* joystick.py:46 */
  int32_t push_enable;
  int32_t action_min_delay, action_max_delay;
  int32_t imu_min_delay, imu_max_delay;
  int32_t auto_reset;            /* 1 = BraxAutoResetWrapper semantics fused into oduck_step */
  double ctrl_dt;
  double action_scale;
  double dof_vel_scale;
  double max_motor_velocity;
  double noise_level;
  double noise_gyro, noise_accelerometer, noise_gravity, noise_joint_vel;
  double qpos_noise_scale[ODUCK_MAX_NU];   /* joystick.py:184-200, quirk #3 of SURVEY 2.1 */
  double scale_tracking_lin_vel, scale_tracking_ang_vel, scale_torques, scale_action_rate;
  double scale_stand_still, scale_alive, scale_imitation;
  double scale_orientation, scale_head_pos;   /* Standing only (standing.py:75-84; rewards.py:45-46,131-147) */
  double reset_base_qvel_noise;               /* joystick.py:253: 0.05, standing.py:247: 0.5 */
  double tracking_sigma;
  double push_interval_range[2];
  double push_magnitude_range[2];
  double cmd_range[ODUCK_NCMD][2];         /* lin_vel_x, lin_vel_y, ang_vel_yaw, neck_pitch, head_pitch, head_yaw, head_roll */
  /* reference motion table (poly_reference_motion.py:74-146): float64 host array
   * [ndx][ndy][ndth][40][16], highest power first (jp.polyval order) */
  int32_t ndx, ndy, ndth, nb_steps_in_period;
  double dxs[8], dys[8], dthetas[16];
  double dx_range[2], dy_range[2], dtheta_range[2];  /* ranges include 0 (poly_reference_motion.py:59-61,105-110) */
  const double* poly_coef;
  OduckRewardLibrary lib;
} OduckEnvConfig;

typedef struct OduckHandle OduckHandle;

/* Policy weights for oduck_policy_forward (Brax MLP, swish, NormalTanh head).
 * Row-major W_l[in][out] like flax Dense kernels; pointers follow the library's
 * memory space (device for CUDA). */
typedef struct OduckPolicyWeights {
  int32_t obs_dim;               /* 101 */
  int32_t hidden[3];             /* 512, 256, 128 */
  int32_t out_dim;               /* 2 * nu = 28 */
  const float* obs_mean;         /* [obs_dim] running-statistics normaliser */
  const float* obs_std;          /* [obs_dim] */
  const float* w[4];
  const float* b[4];
  const float* packed[4];        /* optional (CUDA library): the kernels already in tensor-core operand form, as kept up to date by
                                  * the device learner (oduck_ppo_packed_weights); NULL = the library packs w[] itself and caches it */
} OduckPolicyWeights;

/* Rollout sink (A17: the PPO unroll of common/runner.py:104-118 -> Brax acting.generate_unroll stores a Transition per step).
 * Caller-owned buffers in the library's memory space, time-major like Brax's stacked Transition; when a sink is attached,
 * oduck_rollout_step makes the KERNELS write each step's transition straight into slot t (no pack / copy pass): the actor's
 * head writes raw_action / log_prob, the step kernel writes reward / done / truncation of slot t and the NEW observations into
 * slot t + 1 (after auto-reset: the first observation of the next episode, as Brax's AutoResetWrapper returns it).  Slot 0 of
 * the observations is filled from the handle's current observations when t == 0.  `num_envs` / `env_offset` place the handle's
 * envs inside wider buffers (several handles of one rank, or a rank's slice of the all-gather buffer of SURVEY.md 8e). */
typedef struct OduckRolloutSink {
  int32_t unroll;                /* T */
  int32_t num_envs;              /* env count of the buffers (>= env_offset + the handle's envs) */
  int32_t env_offset;            /* first env of this handle inside the buffers */
  int32_t policy_dim, value_dim; /* row widths: Joystick 101 / 212, Standing 85 / 153 */
  float* obs_policy;             /* [T + 1][num_envs][policy_dim]  <- obs["state"] */
  float* obs_value;              /* [T + 1][num_envs][value_dim]   <- obs["privileged_state"] */
  float* raw_action;             /* [T][num_envs][nu]  pre-tanh sample */
  float* log_prob;               /* [T][num_envs] */
  float* reward;                 /* [T][num_envs] */
  float* done;                   /* [T][num_envs] */
  float* truncation;             /* [T][num_envs] */
} OduckRolloutSink;

typedef enum {
  ODUCK_BUF_QPOS = 0,        /* f32 [N, nq]  */
  ODUCK_BUF_QVEL,            /* f32 [N, nv]  */
  ODUCK_BUF_QACC_WARM,       /* f32 [N, nv]  */
  ODUCK_BUF_QACC,            /* f32 [N, nv]  last solver output */
  ODUCK_BUF_CTRL,            /* f32 [N, nu]  motor targets applied */
  ODUCK_BUF_OBS_STATE,       /* f32 [N, 101]  (Standing: [N, 85], rows 101 apart) */
  ODUCK_BUF_OBS_PRIV,        /* f32 [N, 212]  (Standing: [N, 153], rows 212 apart) */
  ODUCK_BUF_REWARD,          /* f32 [N]      */
  ODUCK_BUF_DONE,            /* f32 [N]      */
  ODUCK_BUF_TRUNCATION,      /* f32 [N]      */
  ODUCK_BUF_METRICS,         /* f32 [N, 8]: reward/tracking_lin_vel, reward/tracking_ang_vel, cost/torques, cost/action_rate, cost/stand_still, reward/alive, reward/imitation, swing_peak
                              *   Standing: cost/orientation, cost/torques, cost/action_rate, cost/stand_still, reward/alive, cost/head_pos, swing_peak, (unused) */
  ODUCK_BUF_EFC_FORCE,       /* f32 [N, nefc] rows: friction dofs, joint limits, contacts x4 */
  ODUCK_BUF_CONTACT_DIST,    /* f32 [N, 12]  */
  ODUCK_BUF_SENSORDATA,      /* f32 [N, 24]: gyro3 local_linvel3 accelerometer3 upvector3 global_angvel3 left_foot_linvel3 right_foot_linvel3 pad3 */
  ODUCK_BUF_ACTUATOR_FORCE,  /* f32 [N, nu]  */
  ODUCK_BUF_SITE_XPOS_FEET,  /* f32 [N, 6]   */
  ODUCK_BUF_INFO_RNG,        /* u32 [N, 2]   */
  ODUCK_BUF_INFO_COMMAND,    /* f32 [N, 7]   */
  ODUCK_BUF_INFO_STEP,       /* i32 [N]      */
  ODUCK_BUF_INFO_STEPS,      /* i32 [N]  EpisodeWrapper counter */
  ODUCK_BUF_INFO_LAST_ACT,   /* f32 [N, 3, nu] last, last_last, last_last_last */
  ODUCK_BUF_INFO_MOTOR_TARGETS, /* f32 [N, nu] */
  ODUCK_BUF_INFO_FEET_AIR_TIME, /* f32 [N, 2] */
  ODUCK_BUF_INFO_LAST_CONTACT,  /* f32 [N, 2] (0/1) */
  ODUCK_BUF_INFO_SWING_PEAK,    /* f32 [N, 2] */
  ODUCK_BUF_INFO_PUSH,          /* f32 [N, 2] */
  ODUCK_BUF_INFO_PUSH_STEP,     /* i32 [N] */
  ODUCK_BUF_INFO_PUSH_INTERVAL, /* i32 [N] */
  ODUCK_BUF_INFO_ACTION_HISTORY,/* f32 [N, action_max_delay*nu] */
  ODUCK_BUF_INFO_IMU_HISTORY,   /* f32 [N, imu_max_delay*3] */
  ODUCK_BUF_INFO_IMITATION_I,   /* i32 [N] */
  ODUCK_BUF_INFO_REF_MOTION,    /* f32 [N, 40] */
  ODUCK_BUF_INFO_IMITATION_PHASE, /* f32 [N, 2] */
  ODUCK_BUF_DR_PARAMS,          /* f32 [N, dr_stride] per-env randomised model */
  ODUCK_BUF_FIRST_QPOS,         /* f32 [N, nq] auto-reset target */
  ODUCK_BUF_FIRST_QVEL,         /* f32 [N, nv] */
  ODUCK_BUF_FIRST_OBS_STATE,    /* f32 [N, 101] */
  ODUCK_BUF_FIRST_OBS_PRIV,     /* f32 [N, 212] */
  ODUCK_BUF_COUNT
} OduckBufferId;

#define ODUCK_DTYPE_F32 0
#define ODUCK_DTYPE_I32 1
#define ODUCK_DTYPE_U32 2
#define ODUCK_DTYPE_F64 3   /* the oracle's buffers */

int oduck_abi_version(void);
int oduck_sizeof_model(void);
int oduck_sizeof_env_config(void);

/* device = CUDA ordinal (ignored by the oracle). */
int oduck_create(const OduckModel* model, const OduckEnvConfig* cfg, int num_envs, int device, OduckHandle** out);
int oduck_destroy(OduckHandle* h);
int oduck_num_envs(const OduckHandle* h);

/* keys: u32 [N,2] threefry keys (jax.random key data). */
int oduck_randomize(OduckHandle* h, const uint32_t* keys, void* stream);
/* mask: optional u8 [N]; only envs with mask != 0 are reset (NULL = all).  Stores first_state/first_obs. */
int oduck_reset(OduckHandle* h, const uint32_t* keys, const uint8_t* mask, void* stream);
/* action: f32 [N, nu].  Fuses Joystick.step + EpisodeWrapper + AutoReset (cfg.auto_reset). */
int oduck_step(OduckHandle* h, const float* action, void* stream);
/* ctrl: f32 [N, nu] (NULL = keep the stored ctrl); n forward+Euler substeps. */
int oduck_physics_substeps(OduckHandle* h, const float* ctrl, int n, void* stream);
/* One mjx.forward on the stored qpos/qvel/ctrl (no integration); refreshes sensordata/efc_force/contact. */
int oduck_forward(OduckHandle* h, void* stream);
/* qpos [N,nq], qvel [N,nv], qacc_warm [N,nv]; f32, any may be NULL. */
int oduck_set_state(OduckHandle* h, const float* qpos, const float* qvel, const float* qacc_warm, void* stream);
/* obs NULL = the handle's own OBS_STATE buffer.  keys u32 [N,2] (ignored if deterministic).
 * Outputs f32: action [N,nu] (tanh-squashed), raw_action [N,nu] (pre-tanh), log_prob [N]; any may be NULL. */
int oduck_policy_forward(OduckHandle* h, const OduckPolicyWeights* w, const float* obs, const uint32_t* keys,
                         int deterministic, float* action, float* raw_action, float* log_prob, void* stream);
/* The library caches a tensor-core repack of the weights keyed by w->w[0]; call this after the weights behind the same
 * pointers changed in place (the device learner of oduck_ppo.h updates them every SGD step). */
int oduck_policy_invalidate(OduckHandle* h);
/* Attach (or with NULL detach) the rollout buffers.  The struct is copied; the buffers must outlive the attachment. */
int oduck_set_rollout_sink(OduckHandle* h, const OduckRolloutSink* sink);
/* Step t (0 <= t < sink.unroll) of the unroll: sample an action from the handle's own obs["state"] with `w` / `keys`
 * (u32 [N,2]), step the envs with it, store the transition in the sink.  Same kernels as oduck_policy_forward + oduck_step. */
int oduck_rollout_step(OduckHandle* h, const OduckPolicyWeights* w, const uint32_t* keys, int t, void* stream);
/* Zero-copy view.  shape[4] (unused dims = 0), strides in ELEMENTS. */
int oduck_get_buffer(OduckHandle* h, int id, void** ptr, int64_t* shape, int64_t* strides, int* dtype);
/* Number of kernels this library has launched on the handle since create (bench `gpu_launches`). */
int64_t oduck_launch_count(const OduckHandle* h);

const char* oduck_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* ODUCK_H_ */
