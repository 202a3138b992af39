/*
 * b200gym.h -- C ABI of libb200gym.so: the B200-native replacement for the part of the
 * reference that lives in the closed Isaac Gym binary (gymapi + PhysX) on the hot path
 * `VecTask.step` (reference: isaacgymenvs/tasks/base/vec_task.py:360-408).
 *
 * Every entry point cites the Isaac Gym call (and the reference call site) it replaces.
 * Conventions:
 *   - plain C types only: pointers, sizes, PODs. No torch / C++ types cross this boundary.
 *   - every call returns int: 0 = OK, <0 = error code; b2g_last_error() gives the message.
 *   - the sim owns all device memory; pointers returned by b2g_sim_tensor() stay valid until
 *     b2g_sim_destroy(). Python wraps them as non-owning torch tensors through DLPack.
 *   - all device work is enqueued on the caller-supplied cudaStream_t (passed as void*); the
 *     library never synchronises except in the *_host convenience calls, which say so.
 *   - one thread drives one sim (same as the reference's process-global singleton,
 *     vec_task.py:55-64). Not thread-safe.
 *   - there is NO CPU fallback: if no CUDA device is usable b2g_sim_create fails with
 *     B2G_ERR_CUDA.
 */
#ifndef B200GYM_H
#define B200GYM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2G_ABI_VERSION 2

#define B2G_MAX_DOF 24
#define B2G_MAX_LINKS (B2G_MAX_DOF + 1)
#define B2G_MAX_BODIES 32
#define B2G_MAX_CHAINS 8
#define B2G_MAX_CHAIN_LEN 6        /* chains of a floating-base robot */
#define B2G_MAX_FIXED_CHAIN_LEN 7  /* the single chain of a fixed-base arm (Manipulator: 7-DOF Franka) */
#define B2G_MAX_CPTS 128
#define B2G_LINK_SCALE_COLS 6
#define B2G_MAX_CONTACTS_PER_CHAIN 8   /* upper bound of b2g_sim_params::max_contacts_per_chain */
#define B2G_DEFAULT_CONTACTS_PER_CHAIN 4

enum b2g_status {
    B2G_OK = 0,
    B2G_ERR_ARG = -1,      /* bad argument / bad handle              */
    B2G_ERR_CUDA = -2,     /* CUDA runtime error (message has detail) */
    B2G_ERR_STATE = -3,    /* call out of order (e.g. step before prepare) */
    B2G_ERR_UNSUPPORTED = -4
};

/* Isaac Gym gymapi.DOF_MODE_* (reference: tasks/anymal.py:201, tasks/cartpole.py:110-111) */
enum b2g_drive_mode { B2G_DOF_MODE_NONE = 0, B2G_DOF_MODE_POS = 1, B2G_DOF_MODE_VEL = 2, B2G_DOF_MODE_EFFORT = 4 };

enum b2g_joint_type { B2G_JOINT_REVOLUTE = 0, B2G_JOINT_PRISMATIC = 1 };

/*
 * Compiled articulation ("asset"): one root link + n_chains serial chains (star topology).
 * Replaces the result of gym.load_asset (reference: tasks/anymal.py:183). Produced by the Python
 * model compiler (isaacgymenv_b200/model/urdf.py). Link 0 is the root; link 1+d is the child link of
 * DOF d. DOFs are numbered chain after chain, which is Isaac Gym's depth-first order.
 * Inertia is about the link's centre of mass in link axes: xx, yy, zz, xy, xz, yz.
 */
typedef struct b2g_model {
    int32_t fixed_base;
    int32_t n_dof;
    int32_t n_bodies;
    int32_t n_chains;
    int32_t n_cpts;
    int32_t chain_start[B2G_MAX_CHAINS];
    int32_t chain_len[B2G_MAX_CHAINS];
    float link_mass[B2G_MAX_LINKS];
    float link_com[B2G_MAX_LINKS][3];
    float link_inertia[B2G_MAX_LINKS][6];
    int32_t joint_type[B2G_MAX_DOF];
    float joint_pos[B2G_MAX_DOF][3];   /* joint frame origin in the parent link frame          */
    float joint_quat[B2G_MAX_DOF][4];  /* joint frame rotation in the parent frame, xyzw, q = 0 */
    float joint_axis[B2G_MAX_DOF][3];  /* unit axis in the child (= joint) frame               */
    float lower[B2G_MAX_DOF];          /* -inf / +inf when the joint has no limits              */
    float upper[B2G_MAX_DOF];
    float effort[B2G_MAX_DOF];
    float vel_limit[B2G_MAX_DOF];
    float armature[B2G_MAX_DOF];
    /* API bodies = rows of the rigid-body-state and net-contact-force tensors */
    int32_t body_link[B2G_MAX_BODIES];
    float body_pos[B2G_MAX_BODIES][3];
    float body_quat[B2G_MAX_BODIES][4];
    /* contact spheres; cp_chain = the chain (lane) that owns the candidate in the solver */
    int32_t cp_link[B2G_MAX_CPTS];
    int32_t cp_body[B2G_MAX_CPTS];
    int32_t cp_chain[B2G_MAX_CPTS];
    float cp_pos[B2G_MAX_CPTS][3];
    float cp_radius[B2G_MAX_CPTS];
} b2g_model;

/* gymapi.SimParams + PlaneParams subset (reference: vec_task.py:514-562, cfg/task/Anymal.yaml:81-100,
 * tasks/anymal.py:159-164). */
typedef struct b2g_sim_params {
    float dt;
    int32_t substeps;
    float gravity[3];
    int32_t num_position_iterations;
    int32_t num_velocity_iterations;
    float contact_offset;
    float rest_offset;
    float bounce_threshold_velocity;
    float max_depenetration_velocity;
    float plane_static_friction;
    float plane_dynamic_friction;
    float plane_restitution;
    int32_t has_ground;               /* add_ground was called */
    float joint_limit_stiffness;      /* joint limits are one-sided implicit spring-dampers [N m/rad, N m s/rad]  */
    float joint_limit_damping;
    int32_t max_contacts_per_chain;   /* contact slots per solver lane (chain), 1..B2G_MAX_CONTACTS_PER_CHAIN; 0 = 4.  PhysX keeps every
                                         contact a body has; here candidates beyond the slots are dropped and counted
                                         (b2g_sim_contact_stats) -- pick the smallest value whose drop count stays negligible: the
                                         slots live in shared memory ([slot][field][thread]) and take L1 capacity from the kernels */
    float max_linear_velocity;        /* gymapi.AssetOptions.max_linear_velocity / max_angular_velocity (defaults 1000 m/s, 64 rad/s; the     */
    float max_angular_velocity;       /* hot-path tasks leave them alone): the root body's velocity is clamped to them after the contact solve,  */
                                      /* as PhysX clamps a body's velocity.  0 = no limit                                                        */
    int32_t self_collision;           /* 1: the candidate spheres of every link that does not hang off the root directly also collide with the
                                         root's bounding box (the reference's rough-terrain tasks create their actors with collision filter
                                         0 = self-collision on, tasks/anymal_terrain.py:282; the flat tasks with 1 = off).  A leg or arm
                                         can then no longer swing through the base.  Link-link pairs are not modelled.                        */
} b2g_sim_params;

/* per-DOF drive properties, identical for every env (reference: tasks/anymal.py:199-203,214) */
typedef struct b2g_dof_props {
    int32_t drive_mode[B2G_MAX_DOF];
    float stiffness[B2G_MAX_DOF];
    float damping[B2G_MAX_DOF];
    float effort[B2G_MAX_DOF];
    float lower[B2G_MAX_DOF];
    float upper[B2G_MAX_DOF];
    float velocity[B2G_MAX_DOF];
} b2g_dof_props;

/* heightfield terrain: int16 samples, row index along +x, column along +y
 * (reference: tasks/anymal_terrain.py:196-209,515-538,549-576). height = raw * vertical_scale;
 * world x of row i = origin_x + i * horizontal_scale. */
typedef struct b2g_heightfield {
    int32_t rows, cols;
    float horizontal_scale, vertical_scale;
    float origin_x, origin_y;
    float friction, restitution;
} b2g_heightfield;

/* tensors the sim owns; kind selects one (gym.acquire_*_tensor, reference: tasks/anymal.py:110-113,
 * tasks/useful_hound.py:440-455) */
enum b2g_tensor_kind {
    B2G_T_ROOT_STATE = 0,       /* (N,13) f32  pos3 quat4(xyzw) linvel3 angvel3, world      */
    B2G_T_DOF_STATE = 1,        /* (N*nd,2) f32 pos, vel                                     */
    B2G_T_NET_CONTACT = 2,      /* (N*nb,3) f32 world                                        */
    B2G_T_DOF_FORCE = 3,        /* (N*nd) f32                                                */
    B2G_T_RIGID_BODY_STATE = 4, /* (N*nb,13) f32                                             */
    B2G_T_DOF_TARGET = 5,       /* (N*nd) f32 position/velocity targets                      */
    B2G_T_DOF_ACTUATION = 6,    /* (N*nd) f32 efforts                                        */
    B2G_T_JACOBIAN = 7,         /* floating base (N,nb,6,6+nd), fixed base (N,nb-1,6,nd) f32; rows lin3, ang3 */
    B2G_T_MASS_MATRIX = 8,      /* (N,nd,nd) f32 (joint block, Isaac Gym convention)         */
    B2G_T_FRICTION = 9,         /* (N) f32 per-env shape friction coefficient                */
    B2G_T_ENV_SCALE = 10,       /* (N,4) f32 per-env scale of [link masses+inertias, drive stiffness, drive damping, spare]:
                                   tensorised domain randomisation (vec_task.py:610-840); ones until acquired */
    B2G_T_LINK_SCALE = 11,      /* (N,nd+1,B2G_LINK_SCALE_COLS) f32 per-LINK randomisation, link 0 = root, link 1+d = child link of
                                   DOF d: [scale of the link's mass+inertia, scale of DOF d's stiffness, scale of its damping,
                                   offset added to its lower limit, offset added to its upper limit, spare]; the scales multiply
                                   B2G_T_ENV_SCALE.  The reference randomises every body / DOF property entry on its own
                                   (utils/dr_utils.py:135-238, cfg/task/Anymal.yaml:146-170); (1,1,1,0,0,0) until acquired */
    B2G_T_COUNT
};

typedef struct b2g_tensor_desc {
    void* data;          /* device pointer */
    int32_t dtype;       /* 0 = f32, 1 = i32, 2 = i64, 3 = u8, 4 = i16 */
    int32_t ndim;
    int64_t shape[4];
    int32_t device_id;
} b2g_tensor_desc;

typedef struct b2g_sim b2g_sim;

/* ---- lifecycle: gymapi.acquire_gym / gym.create_sim / prepare_sim (vec_task.py:247,63,262) ---- */
int b2g_abi_version(void);
const char* b2g_last_error(void);
int b2g_sim_create(int device_id, const b2g_sim_params* params, b2g_sim** out);
int b2g_sim_destroy(b2g_sim* sim);
int b2g_sim_set_params(b2g_sim* sim, const b2g_sim_params* params);       /* gym.set_sim_params */
int b2g_sim_get_params(const b2g_sim* sim, b2g_sim_params* out);          /* gym.get_sim_params */
/* gym.add_ground (tasks/anymal.py:159-164): sets has_ground + friction in the params */
int b2g_sim_add_ground(b2g_sim* sim, float static_friction, float dynamic_friction, float restitution);
/* gym.add_triangle_mesh for a gridded terrain (tasks/anymal_terrain.py:196-209): host int16 samples.  Once the articulation is
 * known as well, a coarse conservative bound of the field (max height / min normal z per 8x8-sample block, dilated by the largest
 * link radius) is built for the kernels' per-link contact early-out; it never changes a result.  Environment switch
 * B2G_NO_HFC=1 (read here) leaves it out (A/B timing, bit-identity tests). */
int b2g_sim_add_heightfield(b2g_sim* sim, const b2g_heightfield* hf, const int16_t* samples_host);
/* gym.load_asset + N x (create_env, create_actor, set_actor_dof_properties) (tasks/anymal.py:205-216).
 * root_pose7 = start pose (pos3, quat xyzw) given to create_actor; env_spacing/num_per_row = create_env grid. */
int b2g_sim_add_articulation(b2g_sim* sim, const b2g_model* model, const b2g_dof_props* props, int n_envs,
                             const float* root_pose7, float env_spacing, int num_per_row);
int b2g_sim_prepare(b2g_sim* sim);                                         /* gym.prepare_sim */
int b2g_sim_set_dof_props(b2g_sim* sim, const b2g_dof_props* props);       /* gym.set_actor_dof_properties */

/* ---- tensors: gym.acquire_* (tasks/anymal.py:110-113) ---- */
int b2g_sim_tensor(b2g_sim* sim, int kind, b2g_tensor_desc* out);

/* ---- stepping: gym.simulate (vec_task.py:382) = `substeps` sub-steps of dt/substeps ---- */
int b2g_sim_simulate(b2g_sim* sim, void* stream);
/* gym.refresh_rigid_body_state_tensor / refresh_jacobian_tensors / refresh_mass_matrix_tensors
 * (tasks/useful_hound.py:729-732). Root/DOF/contact/DOF-force tensors alias live sim state, so
 * their refresh is a no-op kept for API compatibility. */
int b2g_sim_refresh(b2g_sim* sim, int kind, void* stream);
/* gym.set_*_tensor_indexed (tasks/anymal.py:291-297): copy rows idx[0..n) of `src` (full-size device
 * tensor, same layout as kind) into the sim tensor. src == the sim's own tensor is a no-op. */
int b2g_sim_set_indexed(b2g_sim* sim, int kind, const void* src_dev, const int32_t* idx_dev, int n, void* stream);
/* gym.set_dof_position_target_tensor / set_dof_actuation_force_tensor / set_actor_root_state_tensor /
 * set_dof_state_tensor (tasks/anymal.py:229, tasks/anymal_terrain.py:446,439): full copy */
int b2g_sim_set_tensor(b2g_sim* sim, int kind, const void* src_dev, void* stream);

/* forward dynamics probe used by the parity tests: qdd (N,nd) and root spatial acceleration (N,6:
 * angular3, linear3 of the root origin, world) for the current state and DOF_ACTUATION efforts, no
 * contact, no drives. */
int b2g_sim_forward_dynamics(b2g_sim* sim, float* qdd_dev, float* root_acc_dev, void* stream);

/* ---- fused task steps: pre_physics_step + simulate x k + post_physics_step in one launch ---- */

/* Anymal / Hound flat task (reference: tasks/anymal.py:226-304,311-386; tasks/hound.py same lines) */
typedef struct b2g_anymal_cfg {
    float lin_vel_scale, ang_vel_scale, dof_pos_scale, dof_vel_scale, action_scale;
    float rew_lin_vel_xy, rew_ang_vel_z, rew_torque;   /* already multiplied by dt (anymal.py:99-100) */
    float clip_obs, clip_actions;
    float cmd_x[2], cmd_y[2], cmd_yaw[2];
    float default_dof_pos[B2G_MAX_DOF];
    float init_root[13];
    int32_t base_body;
    int32_t n_knee;
    int32_t knee_bodies[8];
    int64_t max_episode_length;
    uint64_t seed;
} b2g_anymal_cfg;

typedef struct b2g_task_buffers {
    float* obs;          /* (N,num_obs) unclamped, the task's obs_buf                       */
    float* obs_clamped;  /* (N,num_obs) clamp(obs, +-clip_obs): what step() returns          */
    float* rew;          /* (N)                                                              */
    int64_t* reset;      /* (N) int64 (flat tasks) -- vec_task.py:316                        */
    int64_t* progress;   /* (N) int64                                                        */
    int64_t* timeout;    /* (N) int64 0/1                                                    */
    float* commands;     /* (N,3|4)                                                          */
    float* actions;      /* (N,na) last clamped actions (task's self.actions)                */
    const float* rand_override; /* optional (N,n_rand) uniforms in [0,1) replacing Philox (tests) */
} b2g_task_buffers;

int b2g_task_anymal_create(b2g_sim* sim, const b2g_anymal_cfg* cfg);

/* Cartpole task (reference: tasks/cartpole.py:36-196): effort = action * max_push_effort on DOF 0 (:159-163), obs =
 * [x, xdot, theta, thetadot] (:131-142), reward/reset (:180-196), reset draws (:144-158). */
typedef struct b2g_cartpole_cfg {
    float reset_dist, max_push_effort;
    float clip_obs, clip_actions;
    int64_t max_episode_length;
    uint64_t seed;
} b2g_cartpole_cfg;
int b2g_task_cartpole_create(b2g_sim* sim, const b2g_cartpole_cfg* cfg);

/* Fixed-base arm reach task Houndarm (reference: tasks/hound_arm.py:74-567), one launch per step: actions -> pose deltas
 * (* cmd_limit / action_scale, :499-501) -> operational-space torques from the CURRENT mass matrix, Jacobian row and
 * end-effector velocity (:462-493; the reference refreshes them at the end of the previous post_physics_step, after resets) ->
 * sub-steps with effort drives -> post_physics_step: progress, reset_idx (:394-459: 3 command draws, 6 joint-noise draws,
 * clamped to the limits, reset_buf = 0), observations [eef pos, eef quat, command] (:383-392), reward / reset (:550-567),
 * time-outs and clamped observations (vec_task.py:394-402).  Needs a fixed-base single chain of <= 6 DOF.
 * The same struct drives Manipulator (tasks/manipulator.py: the same task on the 7-DOF Franka arm): n DOFs <= B2G_MAX_FIXED_CHAIN_LEN,
 * six actions, null-space posture and reset around default_dof_pos (:153-155, :407-414), the last n_reset_tail joints reset to their
 * default without noise (:417, written for a gripper the asset does not have: it hits the arm's last two joints). */
typedef struct b2g_houndarm_cfg {
    float clip_obs, clip_actions, action_scale, dof_noise;
    float cmd_limit[6];
    float kp, kp_null;                 /* kd = 2 sqrt(kp) (:166-169) */
    float cmd_range[6];                /* x lo, x hi, y lo, y hi, z lo, z hi */
    float dist_scale, vel_scale;
    int32_t eef_body, jac_body;        /* API bodies: end-effector state row; row of the fixed-base Jacobian the task takes */
    int64_t max_episode_length;
    uint64_t seed;
    float default_dof_pos[8];          /* zeros for Houndarm (hound_arm.py:160-162) */
    int32_t n_reset_tail;              /* 0 for Houndarm, 2 for Manipulator */
    int32_t pad_;
} b2g_houndarm_cfg;
int b2g_task_houndarm_create(b2g_sim* sim, const b2g_houndarm_cfg* cfg);

/* Rough-terrain locomotion task: AnymalTerrain / HoundTerrain (reference: tasks/anymal_terrain.py:45-538,
 * tasks/Hound_terrain.py same lines).  One step() = `decimation` sim steps with fresh explicit PD torques
 * (:441-451) + `extra_sim_steps` with the last torques (the generic VecTask loop, vec_task.py:379-382), then
 * post_physics_step (:453-485): push, heading command, termination (:294-300 / Hound_terrain.py:304-311), 13-term reward
 * (:315-382), reset_idx with terrain curriculum (:384-435), 188 observations with the 14x10 height scan (:302-313,
 * :503-538) and uniform observation noise (:481-482).  Two launches per step (physics+reward, then reset+obs). */
#define B2G_REW_TERMS 14
typedef struct b2g_terrain_cfg {
    float lin_vel_scale, ang_vel_scale, dof_pos_scale, dof_vel_scale, height_meas_scale, action_scale;
    float kp, kd, torque_limit;
    int32_t decimation, extra_sim_steps;
    float dt;                        /* decimation * sim dt (anymal_terrain.py:94-95)                           */
    float rew[B2G_REW_TERMS];        /* x dt; order: termination, lin_vel_xy, lin_vel_z, ang_vel_z, ang_vel_xy, orient,
                                        torque, joint_acc, base_height, air_time, collision, stumble, action_rate, hip */
    float base_height_target;        /* 0.52 (anymal_terrain.py:330) / 0.48 (Hound_terrain.py:347)              */
    float clip_obs, clip_actions;
    float cmd_x[2], cmd_y[2], cmd_yaw[2];
    float default_dof_pos[B2G_MAX_DOF];
    float init_root[13];
    int32_t add_noise;
    float noise_lin_vel, noise_ang_vel, noise_gravity, noise_dof_pos, noise_dof_vel, noise_height;   /* x level x obs scale */
    int32_t base_body;
    int32_t n_knee, knee_bodies[8];
    int32_t n_feet, feet_bodies[8];
    int32_t hound_termination;       /* 1: Hound_terrain.py:304-311 (knee + extra bodies always terminate)        */
    int32_t n_term_extra, term_extra_bodies[8];
    int32_t allow_knee_contacts;
    int32_t hip_dofs[4];
    int64_t max_episode_length;
    int32_t push_interval;
    float max_episode_length_s;
    int32_t custom_origins, curriculum;
    int32_t n_hx, n_hy;              /* height-scan grid (x-major): 14 x 10                                       */
    float hx[16], hy[16];
    int32_t hs_rows, hs_cols;        /* height_samples grid incl. border                                          */
    float border_size, hscale, vscale, env_length;
    int32_t env_rows, env_cols;      /* terrain levels x terrain types                                            */
    uint64_t seed;
    /* hound + manipulator arm (reference: tasks/useful_hound.py). n_ctrl_dof = PD-controlled leg DOFs (observed, :482-497);
     * the DOFs of chain `arm_chain` are driven by the operational-space torque law (:660-691) from action columns
     * [n_ctrl_dof, n_ctrl_dof+6). arm_chain < 0: no arm (AnymalTerrain / HoundTerrain). */
    int32_t n_ctrl_dof;
    int32_t arm_chain;
    float arm_kp, arm_kp_null, arm_action_scale, arm_dof_noise;
    float arm_cmd_limit[6];
    int32_t eef_body;                /* body whose state fills obs[..+7] (:440-447)                               */
    int32_t jac_body;                /* row of the Jacobian tensor the task slices (:448-451)                     */
    int32_t refresh_eef;             /* 0: replicate the reference (end-effector state is never refreshed) */
} b2g_terrain_cfg;
/* height_samples (hs_rows*hs_cols int16) and terrain_origins (env_rows*env_cols*3 f32) are host arrays (NULL for a plane) */
int b2g_task_terrain_create(b2g_sim* sim, const b2g_terrain_cfg* cfg, const int16_t* height_samples_host, const float* terrain_origins_host);
/* the caller's step counter (common_step_counter AFTER its increment for this step, anymal_terrain.py:460) */
int b2g_task_terrain_set_step(b2g_sim* sim, int64_t common_step_counter);
/* init_done / curriculum switch of update_terrain_level (anymal_terrain.py:428) */
int b2g_task_terrain_set_init_done(b2g_sim* sim, int init_done);
/* enable != 0: the library keeps common_step_counter in device memory and advances it at the end of every
 * b2g_task_step, so consecutive steps need no host-side state and can be captured in a CUDA graph
 * (b2g_task_terrain_set_step then sets the value the NEXT step uses). */
int b2g_task_terrain_device_step(b2g_sim* sim, int enable);

/* task-generic entry points (dispatch on the task created on this sim) */
int b2g_task_step(b2g_sim* sim, const float* actions_dev, void* stream);          /* VecTask.step, one launch   */
int b2g_task_post_only(b2g_sim* sim, const float* actions_dev, void* stream);     /* post_physics_step only     */
/* parity-test entry of the hound+arm task: evaluate the operational-space torque law once on the ARM_MM / ARM_JAC / EEF_STATE
 * tensors and the current arm DOF state; the 6 torques land in the arm columns of the TORQUES tensor */
int b2g_task_osc_probe(b2g_sim* sim, const float* actions_dev, void* stream);
int b2g_task_step_host(b2g_sim* sim, const float* actions_host, float* obs_host, float* rew_host, int64_t* reset_host,
                       int64_t* timeout_host, void* stream);                      /* host buffers, synchronises */
/* allocate-and-describe the task buffers the sim owns (obs_buf, rew_buf, reset_buf, ...) */
enum b2g_task_tensor_kind {
    B2G_TT_OBS = 0, B2G_TT_OBS_CLAMPED = 1, B2G_TT_REW = 2, B2G_TT_RESET = 3, B2G_TT_PROGRESS = 4,
    B2G_TT_TIMEOUT = 5, B2G_TT_COMMANDS = 6, B2G_TT_ACTIONS = 7, B2G_TT_RAND_OVERRIDE = 8,
    /* rough-terrain tasks */
    B2G_TT_TORQUES = 9, B2G_TT_LAST_ACTIONS = 10, B2G_TT_LAST_DOF_VEL = 11, B2G_TT_FEET_AIR_TIME = 12, B2G_TT_EPISODE_SUMS = 13,
    B2G_TT_ENV_ORIGINS = 14, B2G_TT_TERRAIN_LEVELS = 15, B2G_TT_TERRAIN_TYPES = 16, B2G_TT_NOISE_OVERRIDE = 17, B2G_TT_PUSH_OVERRIDE = 18,
    B2G_TT_EXTRAS = 19, B2G_TT_MEASURED_HEIGHTS = 20,
    /* hound + arm */
    B2G_TT_ARM_MM = 21, B2G_TT_ARM_JAC = 22, B2G_TT_EEF_STATE = 23, B2G_TT_ARM_COMMANDS = 24, B2G_TT_COUNT
};
int b2g_task_tensor(b2g_sim* sim, int kind, b2g_tensor_desc* out);
/* reset_idx(all envs) as in the task constructor (tasks/anymal.py:146) + first observations */
int b2g_task_anymal_reset_all(b2g_sim* sim, void* stream);
/* VecTask.step for the flat task: one kernel launch. actions_dev: (N,12) f32 */
int b2g_task_anymal_step(b2g_sim* sim, const float* actions_dev, void* stream);
/* parity-test entry: post_physics_step only (tasks/anymal.py:231-239) on the sim tensors as they are --
 * root/DOF state, DOF forces, net contact forces, commands, progress, reset -- no physics */
int b2g_task_anymal_post_only(b2g_sim* sim, const float* actions_dev, void* stream);
/* use_rand_override != 0: reset draws come from the RAND_OVERRIDE tensor instead of Philox */
int b2g_task_set_rand_override(b2g_sim* sim, int use_rand_override);

/* host-buffer entry used by the end-to-end benchmark and non-torch callers: blocking, returns when the results are in the
 * caller's buffers.  Page-locked action buffers are read by the kernel in place, pageable ones are staged with one H2D copy.
 * Page-locked result buffers in the packed b2g_task_host_layout are written by the SMs themselves (tail of the fused step
 * kernel for the flat tasks and of k_terrain_post for the rough-terrain tasks, k_mirror_host otherwise) and the call returns when the published sequence word arrives -- no
 * copy command, no stream synchronisation; any other buffers are served with D2H copies + cudaStreamSynchronize.
 * Environment switch B2G_HOST_MIRROR=0 forces the copy path (A/B timing). */
int b2g_task_anymal_step_host(b2g_sim* sim, const float* actions_host, float* obs_host, float* rew_host,
                              int64_t* reset_host, int64_t* timeout_host, void* stream);

/* Host-buffer layout of the fast path of b2g_task_step_host: byte offsets of obs, rew, reset and time-outs inside a single
 * page-locked host allocation of total_bytes (16-byte aligned); pass obs_host = base + offsets[0], rew_host = base +
 * offsets[1], ... (pageable memory in this layout: one D2H copy; buffers laid out differently: one copy each). */
int b2g_task_host_layout(const b2g_sim* sim, int64_t* offsets /*[4]*/, int64_t* total_bytes);

/* number of kernels this library has launched since creation (bench.py's gpu_launches) */
int64_t b2g_sim_launch_count(const b2g_sim* sim);

/* Contact bookkeeping of every sub-step simulated since creation (or the last reset), so that the fixed number of contact
 * slots per chain (B2G_MAX_CONTACTS_PER_CHAIN) can never bite silently -- PhysX, which the reference uses behind
 * gym.simulate (vec_task.py:379-382), keeps every contact a body has.  out[0] = active contact points, out[1] = candidates
 * inside the contact offset that found no free slot and were dropped, out[2] = environment sub-steps with at least one
 * dropped candidate, out[3] = environment sub-steps simulated.  Synchronises the device.  reset != 0 zeroes the counters. */
int b2g_sim_contact_stats(b2g_sim* sim, int64_t* out /*[4]*/, int reset);

/* sizeof() of the public PODs (0 model, 1 sim_params, 2 dof_props, 3 heightfield, 4 tensor_desc,
 * 5 anymal_cfg, 6 cartpole_cfg, 7 terrain_cfg, 8 houndarm_cfg) so a foreign-language mirror of this header can verify its layout */
int b2g_sizeof(int which);

/* gymtorch.wrap_tensor (tasks/anymal.py:121-126): wrap a tensor description as a DLPack
 * DLManagedTensor* (v0.8 ABI, kDLCUDA, non-owning; its deleter frees only the descriptor). The
 * consumer (torch.from_dlpack on a "dltensor" capsule) gets a view that stays valid until
 * b2g_sim_destroy. */
int b2g_dlpack_from_desc(const b2g_tensor_desc* desc, void** out_managed);

/* ---------------------------------------------------------------------------------------------------------------
 * Rollout policy (next row 8(f)-1): the actor-critic MLP rl_games builds from cfg/train/AnymalPPO.yaml:10-33
 * (shared trunk of three ELU layers, mu head, value head; input normalisation cfg/train/AnymalPPO.yaml:45),
 * evaluated once per VecTask.step for every environment. One fused tcgen05 kernel: bf16 operands, fp32 accumulation
 * in tensor memory, activations stay in shared memory between layers. Weights are given in nn.Linear layout
 * (out x in, row-major fp32, device memory) and packed once per update.
 * Two kernels behind the same calls: when all weights + one 128-row activation tile fit in shared memory (units
 * [256,128,64], cfg/train/AnymalPPO.yaml) they stay resident for every tile of a CTA; otherwise (the rough-terrain
 * networks, units [512,256,128], cfg/train/AnymalTerrainPPO.yaml / UsefulHoundPPO.yaml) they stream through a
 * shared-memory ring fed by the TMA engine.
 * Limits (B2G_ERR_UNSUPPORTED otherwise): exactly three hidden layers, widths multiples of 16; first layer <= 512
 * (a multiple of 32 above 256), the others <= 256; n_obs <= 256; n_actions <= 31.
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct b2g_policy b2g_policy;
enum b2g_policy_layer { B2G_POLICY_HIDDEN0 = 0, B2G_POLICY_HIDDEN1 = 1, B2G_POLICY_HIDDEN2 = 2, B2G_POLICY_MU = 3, B2G_POLICY_VALUE = 4 };
int b2g_policy_create(int device, int n_obs, const int* units /*[3]*/, int n_actions, b2g_policy** out);
void b2g_policy_destroy(b2g_policy* policy);
/* W_dev: (out,in) f32, b_dev: (out) f32, both device pointers; the value head is (1,in) */
int b2g_policy_set_layer(b2g_policy* policy, int layer, const float* W_dev, const float* b_dev, void* stream);
/* running mean / variance of the observations (rl_games RunningMeanStd: (x-mean)/sqrt(var+eps), clamp to +-clip);
 * null pointers = identity */
int b2g_policy_set_obs_norm(b2g_policy* policy, const float* mean_dev, const float* var_dev, float eps, float clip, void* stream);
/* obs_dev (n_rows,n_obs) f32 -> mu_dev (n_rows,n_actions) f32, value_dev (n_rows) f32; one launch */
int b2g_policy_forward(b2g_policy* policy, const float* obs_dev, int n_rows, float* mu_dev, float* value_dev, void* stream);
int64_t b2g_policy_launch_count(const b2g_policy* policy);

/* ---- learner-side kernels of the PPO minibatch update.  Reference: the reference trains through rl_games'
 * a2c_continuous (train.py:200-218, cfg/train/AnymalPPO.yaml); the fork states the same update in-tree for its AMP agent
 * (learning/common_agent.py:312-400 calc_gradients: actor / critic / bound losses, clip_grad_norm_, optimizer step).  All pointers
 * are device pointers, float32 unless stated. ---- */
#define B2G_PPO_MAX_ACTIONS 24
#define B2G_ADAM_MAX_PARTIALS 512

/* Loss head of one minibatch: from the network outputs of the minibatch rows and the rollout buffers (gathered through `index`),
 *   loss = a_loss + 0.5 critic_coef c_loss - entropy_coef entropy + bounds_loss_coef b_loss
 * with a_loss = mean max(-A r, -A clip(r, 1 +- e_clip)), r = exp(old_neglogp - neglogp), c_loss = mean max((v - R)^2, (v_clip - R)^2),
 * b_loss = mean sum_k relu(mu_k - mu_bound)^2 + relu(-mu_bound - mu_k)^2, entropy of the diagonal Gaussian; and the gradients of
 * `loss` with respect to mu, value and log_std in closed form (what autograd would give, torch.max's tie rule included). */
typedef struct b2g_ppo_head_args {
    /* network outputs for the minibatch rows */
    const float* mu;             /* (n_rows, n_actions) */
    const float* value;          /* (n_rows) normalised value prediction */
    const float* log_std;        /* (n_actions) */
    /* rollout buffers, full size; row index[i] belongs to minibatch row i (null: identity) */
    const int64_t* index;        /* (n_rows) */
    const float* actions;        /* (*, n_actions) */
    const float* old_mu;         /* (*, n_actions) */
    const float* old_neglogp;    /* (*) */
    const float* advantages;     /* (*) normalised */
    const float* old_values;     /* (*) normalised */
    const float* returns;        /* (*) normalised */
    int32_t n_rows, n_actions;
    float e_clip, critic_coef, entropy_coef, bounds_loss_coef, mu_bound;
    /* outputs */
    float* grad_mu;              /* (n_rows, n_actions) d loss / d mu */
    float* grad_value;           /* (n_rows) */
    float* grad_log_std;         /* (n_actions) */
    float* out;                  /* (6): loss, a_loss, c_loss, b_loss, kl(old || new), entropy */
    float* partial;              /* workspace, b2g_ppo_head_workspace_floats(n_rows) floats */
} b2g_ppo_head_args;
int b2g_ppo_head(const b2g_ppo_head_args* args, void* stream);
int b2g_ppo_head_workspace_floats(int n_rows);

/* ---- rollout bookkeeping of the learner (what rl_games' play_steps does between two env steps; the fork states it in-tree for its AMP
 * agent, learning/common_agent.py:250-310): a handful of launches per step instead of ~65 element-wise torch kernels ---- */

/* rl_games RunningMeanStd on a (rows, cols) float32 batch: batch moments (float64 sums) merged into the running float64 (mean, var,
 * count) by the parallel-variance formula; optional float32 copies of mean and 1 / sqrt(var + eps).  partial: b2g_stat_workspace_doubles. */
int b2g_running_stat_update(const float* x, int rows, int cols, double* mean, double* var, double* count, double* partial, float* mean_f32,
                            float* inv_std_f32, float eps, void* stream);
int b2g_stat_workspace_doubles(int rows, int cols);
/* out = clamp((x - mean) * inv_std, +-clip), element-wise over (rows, cols) */
int b2g_normalize_store(const float* x, const float* mean_f32, const float* inv_std_f32, float* out, int rows, int cols, float clip, void* stream);

/* Sampling of step t: a = mu + exp(log_std) * N(0, 1) with Philox4x32-10 + Box-Muller keyed by (seed, *rollout_counter, t, env), neglogp,
 * the de-normalised value v * sqrt(var + eps) + mean, rows t of the rollout buffers, and the clamped action for the environment. */
typedef struct b2g_rollout_sample_args {
    const float* mu;              /* (N, A) policy mean of the current observations */
    const float* value;           /* (N) normalised value prediction */
    const float* log_std;         /* (A) */
    const double* value_mean;     /* running statistics of the returns (scalars, float64) */
    const double* value_var;
    float value_eps, action_clip;
    int32_t n_envs, n_actions, t;
    uint64_t seed;
    const int64_t* rollout_counter;   /* device scalar, advanced once per rollout (b2g_rollout_counter_advance) */
    float* b_actions;             /* (T, N, A) */
    float* b_mu;                  /* (T, N, A) */
    float* b_neglogp;             /* (T, N) */
    float* b_values;              /* (T, N) de-normalised */
    float* env_actions;           /* (N, A) clamp(a, +-action_clip) */
} b2g_rollout_sample_args;
int b2g_rollout_sample(const b2g_rollout_sample_args* args, void* stream);
int b2g_rollout_counter_advance(int64_t* counter, void* stream);

/* After env.step of step t: b_rewards = reward_scale * r + gamma * V_t * time_out (reward_shaper + value_bootstrap), b_dones, and the RAW
 * episode statistics (running per-env sums; `finished` += (sum of finished episodes' rewards, of their lengths, their count)). */
typedef struct b2g_rollout_post_args {
    const float* reward;          /* (N) */
    const void* done;             /* (N) int64 or 1-byte flags (flag_bytes) */
    const void* time_out;
    int32_t flag_bytes;           /* 8 or 1 */
    int32_t n_envs, t;
    float reward_scale, gamma;
    const float* b_values;        /* (T, N) */
    float* b_rewards;             /* (T, N) */
    float* b_dones;               /* (T, N) 0 / 1 */
    float* ep_reward;             /* (N) */
    float* ep_length;             /* (N) */
    double* finished;             /* (3) */
} b2g_rollout_post_args;
int b2g_rollout_post(const b2g_rollout_post_args* args, void* stream);

/* End of the rollout: GAE(gamma, tau) as rl_games' discount_values runs it (in-tree statement: learning/common_agent.py:406-418), returns =
 * advantages + values, running statistics of the returns, then f_ret / f_val normalised by them and f_adv = (A - mean A) / (std A + 1e-8)
 * with the unbiased standard deviation (torch.std).  partial: 4 * ceil(T N / 512) doubles. */
typedef struct b2g_gae_args {
    const float* rewards;         /* (T, N) */
    const float* values;          /* (T, N) */
    const float* dones;           /* (T, N) */
    const float* v_last;          /* (N) value of the observation after the last step, de-normalised */
    int32_t horizon, n_envs;
    float gamma, tau, value_eps;
    double* value_mean;           /* running statistics of the returns, updated */
    double* value_var;
    double* value_count;
    float* adv;                   /* (T, N) scratch */
    float* ret;                   /* (T, N) scratch */
    float* f_ret;                 /* (T N) */
    float* f_val;
    float* f_adv;
    double* partial;              /* 2 x b2g_stat_workspace_doubles(horizon * n_envs, 1) doubles */
} b2g_gae_args;
int b2g_gae_finish(const b2g_gae_args* args, void* stream);

/* Hidden layer of the actor-critic MLP around the library GEMM: h = elu(z + bias) in place over the GEMM output z (rows x cols,
 * row-major, cols % 4 == 0), and the backward pass of that pair fused with the bias gradient: dz = dh * elu'(z) (from the stored
 * output h), dbias = column sums of dz (deterministic two-stage sum).  workspace: b2g_mlp_elu_backward_workspace_floats floats. */
int b2g_mlp_bias_elu(float* z, const float* bias, int rows, int cols, void* h_bf16 /* optional (rows, cols) bf16 copy of the output */, void* stream);
int b2g_mlp_elu_backward(const float* dh, const float* h, float* dz /* may be null when dz_bf16 is given */, float* dbias, float* partial, int rows, int cols,
                         void* dz_bf16 /* optional bf16 copy of dz */, void* stream);
int b2g_mlp_elu_backward_workspace_floats(int rows, int cols);
/* Forward of the two output heads on the last hidden layer h (rows x hidden, row-major): mu (rows x n_actions) = h W_mu^T + b_mu, value (rows) =
 * h W_v^T + b_v, float32 accumulation, one pass over h (replaces torch.addmm x 2 of the learner's forward, learning/fused_update.py). */
int b2g_mlp_heads_forward(const float* h, const float* w_mu, const float* b_mu, const float* w_v, const float* b_v, int rows, int hidden, int n_actions,
                          float* mu, float* value, void* stream);
/* Backward of the two output heads on the last hidden layer h (rows x hidden): mu = h W_mu^T + b_mu (n_actions rows), value = h W_v^T + b_v.
 * dh = dmu W_mu + dv W_v; dw_cat ((n_actions + 1) x (hidden + 1), row-major) = [dmu | dv]^T [h | 1]: rows 0..A-1 = d W_mu | d b_mu, row A =
 * d W_v | d b_v.  One pass over the minibatch instead of two K = rows GEMMs with a dozen output rows, two bias reductions and an add. */
int b2g_mlp_heads_backward(const float* h, const float* dmu, const float* dv, const float* w_mu, const float* w_v, int rows, int hidden, int n_actions,
                           float* dh, float* dw_cat, float* partial, void* stream);
int b2g_mlp_heads_backward_workspace_floats(int rows, int hidden, int n_actions);
/* Same pass, the four parameter gradients stored where the caller keeps them (e.g. the .grad views of one flat gradient vector): no
 * packed matrix, no split-and-accumulate kernels behind it. */
int b2g_mlp_heads_backward_scatter(const float* h, const float* dmu, const float* dv, const float* w_mu, const float* w_v, int rows, int hidden, int n_actions,
                                   float* dh, float* dw_mu, float* db_mu, float* dw_v, float* db_v, float* partial, void* stream);
/* Pseudo-random permutation of 0..n-1 into out (int64): the shuffle of a mini-epoch's minibatches (the role of torch.randperm in the learner,
 * learning/ppo.py).  Keyed Feistel bijection with cycle walking, one thread per index, deterministic in (seed, counter). */
int b2g_random_permutation(int64_t* out, int n, uint64_t seed, uint64_t counter, void* stream);
/* Minibatch gather of the update: dst[r] = src[index[r]] for (rows, cols) float32 rows (cols % 4 == 0), as float32 (dst) and / or bf16
 * (dst_bf16) in one pass (torch: index_select + a cast kernel). */
int b2g_gather_rows(const float* src, const int64_t* index, int rows, int cols, float* dst, void* dst_bf16, void* stream);

/* Global-norm clipping + Adam on one flat parameter vector (torch.optim.Adam semantics: no weight decay, no amsgrad).  The effective
 * gradient is grad * grad_scale (1 / world size after an all-reduce SUM); *step is advanced by one. */
typedef struct b2g_adam_args {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int32_t n;
    const float* lr;             /* device scalar (the adaptive-KL schedule rewrites it between graph replays) */
    int64_t* step;               /* device scalar, number of updates so far */
    float beta1, beta2, eps;
    float max_grad_norm;         /* <= 0: no clipping */
    float grad_scale;
    float* partial;              /* workspace, B2G_ADAM_MAX_PARTIALS floats */
    float* out_norm;             /* (2): gradient norm before clipping, clip coefficient */
} b2g_adam_args;
int b2g_adam_clip_step(const b2g_adam_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GYM_H */
