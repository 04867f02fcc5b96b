/*
 * so100_oracle.h — CPU fp64 ORACLE for the so100 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may load this library,
 * and only as the checker / the timed CPU baseline.  The product (libso100_b200.so) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in the un-vendored third-party wheel
 * mujoco==3.3.1 (reference pyproject.toml:9, pixi.lock:1511), which is not installable here, and the reference
 * ships no tests, fixtures or golden trajectories.  This file therefore RESTATES MuJoCo's published
 * mj_step pipeline for the one so100 model (forward kinematics -> joint-space inertia -> RNE bias -> position
 * servo actuation -> friction-loss + joint-limit soft constraints solved by Newton with exact line search ->
 * semi-implicit Euler), and TRANSLITERATES the reference's task logic.  It is pinned only by analytic invariants
 * and the survey's hand-derived anchors (tests/test_oracle_*.py), and opportunistically by
 * tools/dump_mujoco_golden.py wherever MuJoCo is installed.
 *
 * The two structs below have the same layout as so100_model / so100_task_cfg in include/so100_b200.h so that one
 * ctypes definition feeds both sides; nothing else is shared with the product.
 */
#ifndef SO100_ORACLE_H
#define SO100_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NJ 6
#define ORC_MAX_START 64
#define ORC_MAX_PAD 8

typedef struct orc_model {
  int32_t struct_size, nsubstep;
  double timestep;
  double gravity[3];
  double base_pos[3], base_quat[4];
  double body_pos[ORC_NJ][3], body_quat[ORC_NJ][4], body_ipos[ORC_NJ][3], body_iquat[ORC_NJ][4];
  double body_mass[ORC_NJ], body_inertia[ORC_NJ][3];
  double jnt_axis[ORC_NJ][3], jnt_range[ORC_NJ][2], jnt_armature[ORC_NJ], jnt_frictionloss[ORC_NJ];
  double jnt_solref_limit[ORC_NJ][2], jnt_solimp_limit[ORC_NJ][5];
  double dof_solref_friction[ORC_NJ][2], dof_solimp_friction[ORC_NJ][5];
  double act_kp[ORC_NJ], act_dampratio[ORC_NJ], act_kv[ORC_NJ], act_ctrlrange[ORC_NJ][2], act_forcerange[ORC_NJ][2];
  int32_t ee_body, wrist_body, cam_body, _pad0;
  double ee_offset[3], cam_pos[3], cam_quat[4], cam_fovy_deg;
  double block_half_z, block_mass, block_friction, contact_solref[2], contact_solimp[5];
  int32_t block_ncon, _pad1;
  /* arm <-> floor contact: primitive box colliders on the jaws against the floor plane */
  int32_t n_pad, _pad2;
  int32_t pad_body[ORC_MAX_PAD];
  double pad_pos[ORC_MAX_PAD][3], pad_size[ORC_MAX_PAD][3];
  double pad_solref[2], pad_solimp[5], pad_friction;
  double floor_solref[2], floor_solimp[5], floor_friction;
} orc_model;

typedef struct orc_task_cfg {
  int32_t struct_size, task, num_envs, max_episode_steps;
  int64_t env_offset;
  uint64_t seed;
  uint32_t flags;
  int32_t n_start;
  double joint_step_scale;
  double start_positions[ORC_MAX_START][ORC_NJ];
  double rest_position[ORC_NJ], start_position05[ORC_NJ];
  double block_dist_range[2], block_theta_half, reach_threshold;
  double block_space_start[2][3], block_space_end[2][3];
  double block_speed_min, block_speed_max, ramp_seconds, cam_res_w, cam_res_h, obs_noise;
  int32_t lost_limit, _pad0;
} orc_task_cfg;

/* kinematics of one configuration (world frame) */
typedef struct orc_kin {
  double xpos[ORC_NJ][3], xmat[ORC_NJ][9], xaxis[ORC_NJ][3], xipos[ORC_NJ][3], ximat[ORC_NJ][9];
  double end_pos[3], wrist_pos[3], cam_xpos[3], cam_xmat[9];
} orc_kin;

/* full per-env state, plain data so tests can read and write it */
typedef struct orc_env_state {
  double qpos[ORC_NJ], qvel[ORC_NJ], qacc_warm[ORC_NJ], ctrl[ORC_NJ];
  double time;
  double block[3];
  double block_vz; /* qvel[8] of the reference: the free block only ever moves along z (floor contact) */
  /* stale kinematics as left by the last mj_step (all zero after mj_resetData) */
  double end_pos[3], wrist_pos[3], block_xpos[3], cam_xpos[3], cam_xmat[9];
  /* task state */
  double task_block_pos[3], last_block_pos[3]; /* Env02 self.block_pos / self.last_block_pos (persist across resets) */
  double cmd[ORC_NJ], last_angvel[ORC_NJ], target[3], target_dt, target_time, last_centre[2]; /* Env05 */
  int32_t elapsed_steps, ever_stepped, has_last_block, angvel_valid, centre_valid, miss_count;
  double ep_return;
} orc_env_state;

typedef struct orc_sim orc_sim; /* model constants + task cfg + array of env states */

int orc_sizeof_model(void);
int orc_sizeof_task_cfg(void);
int orc_sizeof_env_state(void);

orc_sim *orc_create(const orc_model *m, const orc_task_cfg *cfg);
void orc_destroy(orc_sim *s);
int orc_obs_dim(const orc_sim *s);
void orc_get_derived(const orc_sim *s, double *dof_M0, double *kv, double *invweight0);
orc_env_state *orc_state(orc_sim *s, int env);
int64_t orc_get_tick(const orc_sim *s);
void orc_set_tick(orc_sim *s, int64_t tick);

/* physics building blocks (single configuration) */
void orc_fk(const orc_sim *s, const double *qpos, orc_kin *out);
void orc_mass_matrix(const orc_sim *s, const double *qpos, double *M /*36 row-major, incl. armature*/);
void orc_bias(const orc_sim *s, const double *qpos, const double *qvel, double *bias /*6*/);
/* pad <-> floor contacts of one configuration (MuJoCo mjc_PlaneBox per pad box): returns the count and fills, per
   contact, pos[3], dist, body (arrays sized 4 * ORC_MAX_PAD); also the body's translational body_invweight0 */
int orc_contacts(const orc_sim *s, const double *qpos, double *pos /*[n][3]*/, double *dist, int *body);
void orc_body_invweight0(const orc_sim *s, double *tran /*[6]*/);
/* one mj_forward on the arm: returns qacc; optional outputs may be NULL. niter_out = Newton iterations used. */
void orc_forward(const orc_sim *s, const double *qpos, const double *qvel, const double *ctrl,
                 const double *qacc_warm, double *qacc, double *qacc_smooth, double *qfrc_constraint, int *niter_out);
/* n mj_step substeps of the free block's z coordinate (gravity + applied force fz + floor contact), in place */
void orc_block_substeps(const orc_sim *s, double *z, double *vz, double fz_applied, int n);
/* n mj_step substeps on raw arrays (qpos,qvel,warm updated in place) */
void orc_substeps(const orc_sim *s, double *qpos, double *qvel, double *qacc_warm, const double *ctrl, int n);
/* total mechanical energy (kinetic incl. armature, potential) for invariants */
void orc_energy(const orc_sim *s, const double *qpos, const double *qvel, double *kinetic, double *potential);

/* env API; obs are float32 [num_envs, obs_dim] row-major; mask NULL = all envs. nthreads<=0 -> all online cores */
int orc_hw_threads(void);
void orc_reset(orc_sim *s, const uint8_t *mask, float *obs, int nthreads);
void orc_step(orc_sim *s, const float *actions, float *obs, double *reward, uint8_t *terminated,
              uint8_t *truncated, float *terminal_obs, double *ep_return, int32_t *ep_len, int nthreads);

/* Bulk state exchange with the product's structure-of-arrays layout (include/so100_b200.h so100_state_view:
   field[k][env] at ptr[k * N + env], float32 / int32), so that a parity test can start the oracle from a state the
   CUDA path has reached (65 536 envs, decorrelated) and compare after K more steps.  qpos is qpos - qpos_comp (the
   kernel's compensated sum).  NULL = skip. */
void orc_set_state_soa(orc_sim *s, const float *qpos, const float *qvel, const float *qacc_warm, const float *qpos_comp,
                       const float *block /*[4][N]*/, const float *snap /*[12][N]*/, const float *aux /*[24][N]*/,
                       const int32_t *counters /*[4][N]*/, const float *ep_return);
/* qpos, qvel [6][N], block [4][N] (x y z vz) as doubles */
void orc_get_state_soa(const orc_sim *s, double *qpos, double *qvel, double *block);

/* RNG exposed for tests: Philox4x32-10, key=(seed lo, seed hi), counter=(global env id, tick, stream, 0).
   streams: 0 auto-reset draws, 1 task draws (Env02 relocate / Env05 retarget), 2 Env05 obs noise,
   3 draws of an explicit orc_reset call, 4 Env05 reset-obs noise (only with FRESH_FK_ON_RESET) */
void orc_philox(uint64_t seed, uint32_t env_id, uint32_t tick, uint32_t stream, uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
