/*
 * so100_oracle.c — CPU fp64 ORACLE for the so100 hot path.  TEST INFRASTRUCTURE ONLY (see so100_oracle.h).
 *
 * PARITY UNPINNED: the reference ships no tests / golden vectors, and its arithmetic lives in the un-vendored
 * wheel mujoco==3.3.1 (reference pyproject.toml:9, pixi.lock:1511).  Physics below RESTATES MuJoCo's published
 * mj_step pipeline specialised to the so100 MJCF; task logic TRANSLITERATES the reference:
 *   Env01   src/so100_mujoco_rl/envs/env01_v1.py:15-63
 *   Env02   src/so100_mujoco_rl/envs/env02_v1.py:18-81
 *   Env05   src/so100_mujoco_rl/envs/env03_v1.py:35-215 (inherited) + env05_v1.py:13-75 + env_base_02.py:85-127
 *   Env06   src/so100_mujoco_rl/envs/env06_v1.py:18-82 + env_base_06.py:144-162, :200-264
 *   reward  src/so100_mujoco_rl/envs/env_base_01.py:144-239,  obs :241-270
 *   TimeLimit / auto-reset: src/so100_mujoco_rl/__init__.py:5-45 + SB3 DummyVecEnv semantics.
 *
 * The formulation is deliberately NOT the one the CUDA kernels use: quaternion world-frame kinematics, mass matrix
 * from per-body Jacobians, world-frame Newton-Euler bias, dense Cholesky, and MuJoCo-style primal Newton with an
 * exact piecewise-quadratic line search (the kernels use rotation-matrix recursion, composite rigid bodies and a
 * projected Gauss-Seidel solve).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -pthread -shared -fPIC).
 */
#include "so100_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NJ ORC_NJ
#define MJMINVAL 1e-15

struct orc_sim {
  orc_model m;
  orc_task_cfg cfg;
  int n;
  int64_t tick;
  orc_env_state *env;
  /* derived (what MuJoCo's compiler + mj_setConst produce) */
  double base_quat[4], body_quat[NJ][4], cam_mat[9];
  double body_I[NJ][9]; /* inertia about COM, body frame */
  double dof_M0[NJ], kv[NJ], invw[NJ];
  double fy;            /* focal length in pixels, env_base_02.py:100 */
  /* pad <-> floor contact pair: mixed geom parameters (mj_contactParam) and body_invweight0 (translational) */
  double pc_solref[2], pc_solimp[5], pc_mu, body_tran[NJ];
};

/* ------------------------------------------------------------------ small math */
static void q_norm(const double *q, double *o) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) o[i] = q[i] / n;
}
static void q_mul(const double *a, const double *b, double *o) {
  double r[4];
  r[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  r[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  r[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  r[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  memcpy(o, r, sizeof r);
}
static void q_mat(const double *q, double *R) { /* row-major */
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
static void mv(const double *R, const double *v, double *o) {
  double r[3];
  for (int i = 0; i < 3; i++) r[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
  memcpy(o, r, sizeof r);
}
static void mtv(const double *R, const double *v, double *o) {
  double r[3];
  for (int i = 0; i < 3; i++) r[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
  memcpy(o, r, sizeof r);
}
static void mm(const double *A, const double *B, double *o) {
  double r[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(o, r, sizeof r);
}
static void cross(const double *a, const double *b, double *o) {
  double r[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  memcpy(o, r, sizeof r);
}
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* Cholesky of a 6x6 SPD (row-major) in place (lower); returns 0 on success */
static int chol6(double *A) {
  for (int j = 0; j < NJ; j++) {
    double d = A[j * NJ + j];
    for (int k = 0; k < j; k++) d -= A[j * NJ + k] * A[j * NJ + k];
    if (!(d > 0)) return -1;
    d = sqrt(d);
    A[j * NJ + j] = d;
    for (int i = j + 1; i < NJ; i++) {
      double s = A[i * NJ + j];
      for (int k = 0; k < j; k++) s -= A[i * NJ + k] * A[j * NJ + k];
      A[i * NJ + j] = s / d;
    }
  }
  return 0;
}
static void chol6_solve(const double *L, const double *b, double *x) {
  double y[NJ];
  for (int i = 0; i < NJ; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * NJ + k] * y[k];
    y[i] = s / L[i * NJ + i];
  }
  for (int i = NJ - 1; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < NJ; k++) s -= L[k * NJ + i] * x[k];
    x[i] = s / L[i * NJ + i];
  }
}

/* ------------------------------------------------------------------ RNG: Philox4x32-10 (Salmon et al., SC'11) */
void orc_philox(uint64_t seed, uint32_t env_id, uint32_t tick, uint32_t stream, uint32_t out[4]) {
  uint32_t c0 = env_id, c1 = tick, c2 = stream, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* 24-bit uniform in [0,1): exactly representable in fp32 and fp64, so GPU and oracle see the same draw */
static double u01(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }

/* RNG streams (counter word 2) */
enum { STREAM_RESET = 0, STREAM_TASK = 1, STREAM_NOISE = 2, STREAM_API_RESET = 3, STREAM_RESET_NOISE = 4 };

/* ------------------------------------------------------------------ kinematics */
void orc_fk(const orc_sim *s, const double *qpos, orc_kin *k) {
  const orc_model *m = &s->m;
  double ppos[3] = {m->base_pos[0], m->base_pos[1], m->base_pos[2]}, pquat[4];
  memcpy(pquat, s->base_quat, sizeof pquat);
  for (int i = 0; i < NJ; i++) {
    double R[9], off[3], quat[4], ql[4], ax[3];
    q_mat(pquat, R);
    mv(R, m->body_pos[i], off);
    for (int c = 0; c < 3; c++) k->xpos[i][c] = ppos[c] + off[c];
    q_mul(pquat, s->body_quat[i], quat);
    q_mat(quat, R);
    mv(R, m->jnt_axis[i], ax);
    memcpy(k->xaxis[i], ax, sizeof ax);
    double half = 0.5 * qpos[i], sn = sin(half);
    ql[0] = cos(half); ql[1] = m->jnt_axis[i][0] * sn; ql[2] = m->jnt_axis[i][1] * sn; ql[3] = m->jnt_axis[i][2] * sn;
    q_mul(quat, ql, quat);
    q_norm(quat, quat);
    q_mat(quat, k->xmat[i]);
    double off2[3], iq[4], Ri[9];
    mv(k->xmat[i], m->body_ipos[i], off2);
    for (int c = 0; c < 3; c++) k->xipos[i][c] = k->xpos[i][c] + off2[c];
    q_norm(m->body_iquat[i], iq);
    q_mat(iq, Ri);
    mm(k->xmat[i], Ri, k->ximat[i]);
    memcpy(ppos, k->xpos[i], sizeof ppos);
    memcpy(pquat, quat, sizeof pquat);
  }
  double t[3];
  mv(k->xmat[m->ee_body], m->ee_offset, t); /* env_base_01.py:118-127 */
  for (int c = 0; c < 3; c++) k->end_pos[c] = k->xpos[m->ee_body][c] + t[c];
  memcpy(k->wrist_pos, k->xpos[m->wrist_body], sizeof k->wrist_pos); /* env_base_01.py:114-116 */
  mv(k->xmat[m->cam_body], m->cam_pos, t);
  for (int c = 0; c < 3; c++) k->cam_xpos[c] = k->xpos[m->cam_body][c] + t[c];
  mm(k->xmat[m->cam_body], s->cam_mat, k->cam_xmat);
}

/* world inertia of body i about its COM */
static void world_inertia(const orc_sim *s, const orc_kin *k, int i, double *Iw) {
  double D[9] = {0}, T[9], Rt[9];
  D[0] = s->m.body_inertia[i][0]; D[4] = s->m.body_inertia[i][1]; D[8] = s->m.body_inertia[i][2];
  mm(k->ximat[i], D, T);
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Rt[3 * a + b] = k->ximat[i][3 * b + a];
  mm(T, Rt, Iw);
}

/* joint-space inertia from per-body Jacobians: M = sum_i m_i Jv^T Jv + Jw^T I_i Jw, + armature */
static void mass_from_kin(const orc_sim *s, const orc_kin *k, double *M) {
  memset(M, 0, sizeof(double) * NJ * NJ);
  for (int i = 0; i < NJ; i++) {
    double Iw[9], Jv[NJ][3], Jw[NJ][3], IJw[NJ][3];
    world_inertia(s, k, i, Iw);
    for (int j = 0; j <= i; j++) {
      double r[3];
      for (int c = 0; c < 3; c++) r[c] = k->xipos[i][c] - k->xpos[j][c];
      memcpy(Jw[j], k->xaxis[j], sizeof Jw[j]);
      cross(k->xaxis[j], r, Jv[j]);
      mv(Iw, Jw[j], IJw[j]);
    }
    for (int a = 0; a <= i; a++)
      for (int b = 0; b <= i; b++)
        M[a * NJ + b] += s->m.body_mass[i] * dot3(Jv[a], Jv[b]) + dot3(Jw[a], IJw[b]);
  }
  for (int j = 0; j < NJ; j++) M[j * NJ + j] += s->m.jnt_armature[j];
}
void orc_mass_matrix(const orc_sim *s, const double *qpos, double *M) {
  orc_kin k;
  orc_fk(s, qpos, &k);
  mass_from_kin(s, &k, M);
}

/* qfrc_bias = RNE(q, qd, qdd=0) incl. gravity, world frame */
static void bias_from_kin(const orc_sim *s, const orc_kin *k, const double *qvel, double *bias) {
  double w[NJ][3], al[NJ][3], a[NJ][3], F[NJ][3], N[NJ][3];
  double pw[3] = {0, 0, 0}, pal[3] = {0, 0, 0}, pa[3], ppos[3];
  for (int c = 0; c < 3; c++) { pa[c] = -s->m.gravity[c]; ppos[c] = s->m.base_pos[c]; }
  for (int i = 0; i < NJ; i++) {
    double r[3], t1[3], t2[3], zq[3];
    for (int c = 0; c < 3; c++) { r[c] = k->xpos[i][c] - ppos[c]; zq[c] = k->xaxis[i][c] * qvel[i]; }
    /* origin acceleration of body i (rigidly attached to the parent at r) */
    cross(pal, r, t1);
    cross(pw, r, t2);
    cross(pw, t2, t2);
    for (int c = 0; c < 3; c++) a[i][c] = pa[c] + t1[c] + t2[c];
    cross(pw, zq, t1);
    for (int c = 0; c < 3; c++) { w[i][c] = pw[c] + zq[c]; al[i][c] = pal[c] + t1[c]; }
    /* COM acceleration, inertial force and moment */
    double rc[3], ac[3], Iw[9], Iw_w[3], Ial[3];
    for (int c = 0; c < 3; c++) rc[c] = k->xipos[i][c] - k->xpos[i][c];
    cross(al[i], rc, t1);
    cross(w[i], rc, t2);
    cross(w[i], t2, t2);
    for (int c = 0; c < 3; c++) ac[c] = a[i][c] + t1[c] + t2[c];
    world_inertia(s, k, i, Iw);
    mv(Iw, w[i], Iw_w);
    mv(Iw, al[i], Ial);
    cross(w[i], Iw_w, t1);
    for (int c = 0; c < 3; c++) { F[i][c] = s->m.body_mass[i] * ac[c]; N[i][c] = Ial[c] + t1[c]; }
    memcpy(pw, w[i], sizeof pw); memcpy(pal, al[i], sizeof pal); memcpy(pa, a[i], sizeof pa);
    memcpy(ppos, k->xpos[i], sizeof ppos);
  }
  double f[3] = {0, 0, 0}, n[3] = {0, 0, 0}; /* wrench of the distal sub-chain about origin of body i+1 */
  for (int i = NJ - 1; i >= 0; i--) {
    double rc[3], t1[3], t2[3] = {0, 0, 0};
    for (int c = 0; c < 3; c++) rc[c] = k->xipos[i][c] - k->xpos[i][c];
    cross(rc, F[i], t1);
    if (i < NJ - 1) {
      double d[3];
      for (int c = 0; c < 3; c++) d[c] = k->xpos[i + 1][c] - k->xpos[i][c];
      cross(d, f, t2);
    }
    for (int c = 0; c < 3; c++) { n[c] = N[i][c] + t1[c] + n[c] + t2[c]; f[c] = F[i][c] + f[c]; }
    bias[i] = dot3(k->xaxis[i], n);
  }
}
void orc_bias(const orc_sim *s, const double *qpos, const double *qvel, double *bias) {
  orc_kin k;
  orc_fk(s, qpos, &k);
  bias_from_kin(s, &k, qvel, bias);
}

void orc_energy(const orc_sim *s, const double *qpos, const double *qvel, double *kinetic, double *potential) {
  orc_kin k;
  double M[NJ * NJ], T = 0, V = 0;
  orc_fk(s, qpos, &k);
  mass_from_kin(s, &k, M);
  for (int a = 0; a < NJ; a++) for (int b = 0; b < NJ; b++) T += 0.5 * qvel[a] * M[a * NJ + b] * qvel[b];
  for (int i = 0; i < NJ; i++) V -= s->m.body_mass[i] * dot3(s->m.gravity, k.xipos[i]);
  *kinetic = T; *potential = V;
}

/* ------------------------------------------------------------------ constraints + solver */
#define MAX_CON (4 * ORC_MAX_PAD)
#define MAX_ROWS (3 * NJ + 4 * MAX_CON)
typedef struct { double J[NJ]; double aref, D, R, floss; int friction; } crow;
typedef struct { double pos[3], dist; int body; } ccontact;

static double impedance(const double *solimp, double pos) { /* MuJoCo getimpedance, margin 0 */
  if (solimp[0] == solimp[1] || solimp[2] <= MJMINVAL) return 0.5 * (solimp[0] + solimp[1]);
  double x = fabs(pos / solimp[2]);
  if (x >= 1) return solimp[1];
  if (x <= 0) return solimp[0];
  double y;
  if (solimp[4] == 1) y = x;
  else if (x <= solimp[3]) y = pow(x, solimp[4]) / pow(solimp[3], solimp[4] - 1);
  else y = 1 - pow(1 - x, solimp[4]) / pow(1 - solimp[3], solimp[4] - 1);
  return solimp[0] + y * (solimp[1] - solimp[0]);
}
static void kb_from_solref(const double *solref, const double *solimp, double h, double *K, double *B) {
  double tc = solref[0], dr = solref[1], dmax = solimp[1];
  if (tc > 0) {
    if (tc < 2 * h) tc = 2 * h; /* refsafe */
    *K = 1.0 / fmax(MJMINVAL, dmax * dmax * tc * tc * dr * dr);
    *B = 2.0 / fmax(MJMINVAL, dmax * tc);
  } else { *K = -solref[0] / (dmax * dmax); *B = -solref[1] / dmax; }
}

/* Arm <-> floor contacts.  The reference MJCF gives the jaws eight PRIMITIVE box colliders
 * (so_arm100_camera.xml:60-61 class finger_collision, :108-111 fixed-jaw pads, :120-123 moving-jaw pads); the floor is
 * the plane geom of env01.xml:39, and only block <-> arm pairs are excluded (env01.xml:42-49), so pad <-> floor is live.
 * MuJoCo 3.3.1 semantics restated [3P, from knowledge of mjc_PlaneBox / mj_collideGeoms; no source in the container]:
 * for each of the box's 8 corners (bit 0/1/2 of the index = +x/+y/+z half-size) with ldist = n . (R corner) <= 0 and
 * dist = n . (centre - plane) + ldist <= margin (= 0), at most 4 per box in index order, a contact with that dist,
 * frame normal = plane normal and pos = corner - n dist / 2; it enters the constraints iff dist < margin - gap (= 0).
 * (The arm's MESH colliders are not in the checkout: DESIGN.md D2.) */
static int collide_pads(const orc_sim *s, const orc_kin *k, ccontact *con) {
  const orc_model *m = &s->m;
  int n = 0;
  if (!(s->cfg.flags & 16u)) return 0; /* SO100_FLAG_ARM_CONTACT not set: the pads are ignored, as in the product */
  for (int p = 0; p < m->n_pad && p < ORC_MAX_PAD; p++) {
    int b = m->pad_body[p], cnt = 0;
    double c[3], t[3];
    mv(k->xmat[b], m->pad_pos[p], t);
    for (int a = 0; a < 3; a++) c[a] = k->xpos[b][a] + t[a];
    double dist = c[2]; /* plane z = 0, normal +z */
    for (int i = 0; i < 8 && cnt < 4; i++) {
      double vec[3] = {(i & 1) ? m->pad_size[p][0] : -m->pad_size[p][0], (i & 2) ? m->pad_size[p][1] : -m->pad_size[p][1],
                       (i & 4) ? m->pad_size[p][2] : -m->pad_size[p][2]}, corner[3];
      mv(k->xmat[b], vec, corner);
      double ldist = corner[2];
      if (dist + ldist > 0 || ldist > 0) continue;
      cnt++;
      if (!(dist + ldist < 0)) continue; /* in the gap: detected, not instantiated */
      ccontact *q = &con[n++];
      q->dist = dist + ldist; q->body = b;
      for (int a = 0; a < 3; a++) q->pos[a] = c[a] + corner[a];
      q->pos[2] -= q->dist / 2;
    }
  }
  return n;
}
int orc_contacts(const orc_sim *s, const double *qpos, double *pos, double *dist, int *body) {
  orc_kin k;
  ccontact con[MAX_CON];
  orc_fk(s, qpos, &k);
  int n = collide_pads(s, &k, con);
  for (int i = 0; i < n; i++) { memcpy(pos + 3 * i, con[i].pos, sizeof con[i].pos); dist[i] = con[i].dist; body[i] = con[i].body; }
  return n;
}
void orc_body_invweight0(const orc_sim *s, double *tran) { memcpy(tran, s->body_tran, sizeof s->body_tran); }

static int make_rows(const orc_sim *s, const orc_kin *k, const double *qpos, const double *qvel, crow *rows) {
  const orc_model *m = &s->m;
  int n = 0;
  for (int j = 0; j < NJ; j++) { /* friction-loss rows first (MuJoCo order: equality, friction, limit, contact) */
    if (m->jnt_frictionloss[j] <= 0) continue;
    double K, B, imp = impedance(m->dof_solimp_friction[j], 0.0);
    kb_from_solref(m->dof_solref_friction[j], m->dof_solimp_friction[j], m->timestep, &K, &B);
    crow *r = &rows[n++];
    memset(r->J, 0, sizeof r->J);
    r->J[j] = 1; r->friction = 1; r->floss = m->jnt_frictionloss[j];
    r->R = fmax(MJMINVAL, (1 - imp) / imp * s->invw[j]); r->D = 1 / r->R;
    r->aref = -B * qvel[j]; /* K = 0 for friction rows */
  }
  for (int j = 0; j < NJ; j++) {
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? qpos[j] - m->jnt_range[j][0] : m->jnt_range[j][1] - qpos[j];
      if (!(dist < 0)) continue;
      double K, B, imp = impedance(m->jnt_solimp_limit[j], dist);
      kb_from_solref(m->jnt_solref_limit[j], m->jnt_solimp_limit[j], m->timestep, &K, &B);
      crow *r = &rows[n++];
      memset(r->J, 0, sizeof r->J);
      r->J[j] = side == 0 ? 1 : -1; r->friction = 0; r->floss = 0;
      r->R = fmax(MJMINVAL, (1 - imp) / imp * s->invw[j]); r->D = 1 / r->R;
      r->aref = -B * (r->J[j] * qvel[j]) - K * imp * dist;
    }
  }
  /* contact rows: condim 3, pyramidal cone (the scene's <option>; the attached model's elliptic setting does not
   * survive <attach>): J = J_n +- mu J_t1, J_n +- mu J_t2; every row pos = dist, aref = -B (J qvel) - K imp dist;
   * diagApprox = (1 + mu^2) body_invweight0_trans(body), R = 2 mu_reg^2 max(MINVAL, (1-imp)/imp diagApprox),
   * mu_reg = mu / sqrt(impratio) = mu [3P: mj_instantiateContact, mj_diagApprox, mj_makeImpedance]. */
  ccontact con[MAX_CON];
  int nc = collide_pads(s, k, con);
  for (int c = 0; c < nc; c++) {
    double J3[3][NJ] = {{0}};
    for (int j = 0; j <= con[c].body; j++) {
      double r[3], col[3];
      for (int a = 0; a < 3; a++) r[a] = con[c].pos[a] - k->xpos[j][a];
      cross(k->xaxis[j], r, col);
      for (int a = 0; a < 3; a++) J3[a][j] = col[a];
    }
    double K, B, mu = s->pc_mu, imp = impedance(s->pc_solimp, con[c].dist);
    kb_from_solref(s->pc_solref, s->pc_solimp, m->timestep, &K, &B);
    double R0 = fmax(MJMINVAL, (1 - imp) / imp * (1 + mu * mu) * s->body_tran[con[c].body]);
    double R = fmax(MJMINVAL, 2 * mu * mu * R0);
    /* tangent frame of mju_makeFrame for normal (0,0,1): t1 = +y, t2 = -x */
    for (int t = 0; t < 2; t++)
      for (int sg = 0; sg < 2; sg++) {
        crow *r = &rows[n++];
        double vel = 0;
        for (int j = 0; j < NJ; j++) {
          double jt = t == 0 ? J3[1][j] : -J3[0][j];
          r->J[j] = J3[2][j] + (sg == 0 ? mu : -mu) * jt;
          vel += r->J[j] * qvel[j];
        }
        r->friction = 0; r->floss = 0; r->R = R; r->D = 1 / R;
        r->aref = -B * vel - K * imp * con[c].dist;
      }
  }
  return n;
}
/* cost, gradient contribution and curvature of one row at residual r = J a - aref */
static void row_eval(const crow *r, double res, double *cost, double *force, double *curv) {
  if (r->friction) {
    double rf = r->R * r->floss;
    if (res <= -rf) { *cost = -0.5 * rf * r->floss - r->floss * res; *force = r->floss; *curv = 0; }
    else if (res >= rf) { *cost = -0.5 * rf * r->floss + r->floss * res; *force = -r->floss; *curv = 0; }
    else { *cost = 0.5 * r->D * res * res; *force = -r->D * res; *curv = r->D; }
  } else if (res < 0) { *cost = 0.5 * r->D * res * res; *force = -r->D * res; *curv = r->D; }
  else { *cost = 0; *force = 0; *curv = 0; }
}
static double jdot(const double *J, const double *a) {
  double v = 0;
  for (int j = 0; j < NJ; j++) v += J[j] * a[j];
  return v;
}
static double total_cost(const double *M, const double *as, const crow *rows, int nr, const double *a) {
  double c = 0, d[NJ];
  for (int i = 0; i < NJ; i++) d[i] = a[i] - as[i];
  for (int i = 0; i < NJ; i++) for (int j = 0; j < NJ; j++) c += 0.5 * d[i] * M[i * NJ + j] * d[j];
  for (int k = 0; k < nr; k++) {
    double rc, f, cv;
    row_eval(&rows[k], jdot(rows[k].J, a) - rows[k].aref, &rc, &f, &cv);
    c += rc;
  }
  return c;
}
static int cmp_d(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

static int newton_solve(const double *M, const double *as, const crow *rows, int nr, double *a) {
  int it;
  for (it = 0; it < 200; it++) {
    double g[NJ], H[NJ * NJ], p[NJ], d[NJ];
    for (int i = 0; i < NJ; i++) d[i] = a[i] - as[i];
    for (int i = 0; i < NJ; i++) { g[i] = 0; for (int j = 0; j < NJ; j++) g[i] += M[i * NJ + j] * d[j]; }
    memcpy(H, M, sizeof H);
    double fscale = 1;
    for (int k = 0; k < nr; k++) {
      double rc, f, cv;
      row_eval(&rows[k], jdot(rows[k].J, a) - rows[k].aref, &rc, &f, &cv);
      for (int i = 0; i < NJ; i++) {
        g[i] -= rows[k].J[i] * f;
        if (cv != 0) for (int j = 0; j < NJ; j++) H[i * NJ + j] += cv * rows[k].J[i] * rows[k].J[j];
      }
      fscale += fabs(f);
    }
    double gn = 0, scale = fscale;
    for (int i = 0; i < NJ; i++) {
      double f = 0;
      for (int j = 0; j < NJ; j++) f += M[i * NJ + j] * as[j];
      gn += g[i] * g[i]; scale += fabs(f);
    }
    if (sqrt(gn) <= 1e-14 * scale) break;
    if (chol6(H)) return -1;
    for (int i = 0; i < NJ; i++) g[i] = -g[i];
    chol6_solve(H, g, p);
    /* exact line search: phi'(alpha) is piecewise linear and increasing; walk its breakpoints */
    double bp[2 * MAX_ROWS + 2], r0[MAX_ROWS], sl[MAX_ROWS];
    int nb = 0;
    for (int k = 0; k < nr; k++) {
      double slope = jdot(rows[k].J, p);
      r0[k] = jdot(rows[k].J, a) - rows[k].aref; sl[k] = slope;
      if (slope == 0) continue;
      if (rows[k].friction) {
        double rf = rows[k].R * rows[k].floss, a1 = (-rf - r0[k]) / slope, a2 = (rf - r0[k]) / slope;
        if (a1 > 0) bp[nb++] = a1;
        if (a2 > 0) bp[nb++] = a2;
      } else { double a1 = -r0[k] / slope; if (a1 > 0) bp[nb++] = a1; }
    }
    qsort(bp, nb, sizeof(double), cmp_d);
    bp[nb++] = INFINITY;
    double Mp[NJ], pMp = 0, lo = 0, alpha = 0;
    for (int i = 0; i < NJ; i++) { Mp[i] = 0; for (int j = 0; j < NJ; j++) Mp[i] += M[i * NJ + j] * p[j]; pMp += p[i] * Mp[i]; }
    for (int b = 0; b < nb; b++) {
      /* derivative and curvature just inside the segment (lo, bp[b]) */
      double mid = isinf(bp[b]) ? lo + 1.0 : 0.5 * (lo + bp[b]);
      double dphi = 0, cphi = pMp;
      for (int i = 0; i < NJ; i++) dphi += Mp[i] * (a[i] + mid * p[i] - as[i]);
      for (int k = 0; k < nr; k++) {
        double rc, f, cv;
        row_eval(&rows[k], r0[k] + mid * sl[k], &rc, &f, &cv);
        dphi -= f * sl[k]; cphi += cv * sl[k] * sl[k];
      }
      double root = mid - dphi / cphi; /* phi' is linear inside the segment */
      if (root <= bp[b]) { alpha = root < lo ? lo : root; break; }
      lo = bp[b]; alpha = lo;
    }
    if (!(alpha > 0)) break;
    double step = 0, amax = 1;
    for (int i = 0; i < NJ; i++) { a[i] += alpha * p[i]; step = fmax(step, fabs(alpha * p[i])); amax = fmax(amax, fabs(a[i])); }
    if (step <= 1e-15 * amax) break; /* no representable progress left (stiff contact rows put a rounding floor under |grad|) */
  }
  return it;
}

static void forward_from_kin(const orc_sim *s, const orc_kin *k, const double *qpos, const double *qvel,
                             const double *ctrl, const double *warm, double *qacc, double *qacc_smooth,
                             double *qfrc_constraint, int *niter_out) {
  const orc_model *m = &s->m;
  double M[NJ * NJ], L[NJ * NJ], bias[NJ], fs[NJ], as[NJ];
  mass_from_kin(s, k, M);
  bias_from_kin(s, k, qvel, bias);
  for (int j = 0; j < NJ; j++) { /* position servo, SURVEY B.4 */
    double c = fmin(fmax(ctrl[j], m->act_ctrlrange[j][0]), m->act_ctrlrange[j][1]);
    double f = m->act_kp[j] * c - m->act_kp[j] * qpos[j] - s->kv[j] * qvel[j];
    f = fmin(fmax(f, m->act_forcerange[j][0]), m->act_forcerange[j][1]);
    fs[j] = f - bias[j];
  }
  memcpy(L, M, sizeof L);
  chol6(L);
  chol6_solve(L, fs, as);
  crow rows[MAX_ROWS];
  int nr = make_rows(s, k, qpos, qvel, rows);
  double a[NJ];
  if (warm && total_cost(M, as, rows, nr, warm) < total_cost(M, as, rows, nr, as)) memcpy(a, warm, sizeof a);
  else memcpy(a, as, sizeof a);
  int it = newton_solve(M, as, rows, nr, a);
  memcpy(qacc, a, sizeof a);
  if (qacc_smooth) memcpy(qacc_smooth, as, sizeof as);
  if (qfrc_constraint) {
    memset(qfrc_constraint, 0, sizeof(double) * NJ);
    for (int r = 0; r < nr; r++) {
      double rc, f, cv;
      row_eval(&rows[r], jdot(rows[r].J, a) - rows[r].aref, &rc, &f, &cv);
      for (int j = 0; j < NJ; j++) qfrc_constraint[j] += rows[r].J[j] * f;
    }
  }
  if (niter_out) *niter_out = it;
}
void orc_forward(const orc_sim *s, const double *qpos, const double *qvel, const double *ctrl,
                 const double *qacc_warm, double *qacc, double *qacc_smooth, double *qfrc_constraint, int *niter_out) {
  orc_kin k;
  orc_fk(s, qpos, &k);
  forward_from_kin(s, &k, qpos, qvel, ctrl, qacc_warm, qacc, qacc_smooth, qfrc_constraint, niter_out);
}

/* ------------------------------------------------------------------ block <-> floor contact
 * The Env01/02/06 block is a free box (env01.xml:29-34) spawned with its centre ON the floor plane (env01_v1.py:51-52,
 * env02_v1.py:61-62: z = 0.0, i.e. penetrating by its half-size), so MuJoCo's soft contact pushes it up until it rests
 * ~0.1 mm inside the plane.  Nothing else can touch it (env01.xml:42-49 excludes every arm body that can reach the
 * spawn annulus), it starts axis-aligned with zero velocity, and teleports only rewrite qpos[0:3]; by symmetry the
 * unique minimiser of MuJoCo's convex constraint problem then has no tangential or angular component, and the
 * block's 6-dof dynamics reduce EXACTLY to its z coordinate.  What MuJoCo 3.3.1 does for that coordinate [3P, restated
 * from knowledge of mjc_PlaneBox, mj_instantiateContact and mj_makeImpedance; no source in the container]:
 *   - plane-box collision makes one contact per bottom corner (4), each with dist = z - half_z, included while dist <= 0;
 *   - condim 3 + pyramidal cone (the scene's default <option>): 4 rows per contact, J = n +- mu t_k, so J_z = 1, J.v = vz;
 *   - every row: pos = dist, aref = -B vz - K imp(dist) dist with (K, B) from solref as for the limit rows;
 *   - regularisation: diagApprox = (1 + mu^2) body_invweight0_trans = (1 + mu^2)/m, R0 = (1-imp)/imp diagApprox,
 *     every pyramid row R = 2 mu_reg^2 R0, mu_reg = mu / sqrt(impratio) = mu (impratio 1);
 *   - row force = -min(0, J a - aref) / R.  With all 16 rows identical:  m a = m g + f_applied + (16/R) max(0, aref - a).
 * For mu = 1 the total stiffness 16/R equals that of 4 elliptic-cone normal rows, as MuJoCo's scaling intends. */
static double block_accel(const orc_sim *s, double z, double vz, double fz_applied) {
  const orc_model *m = &s->m;
  double a_free = m->gravity[2] + fz_applied / m->block_mass;
  double dist = z - m->block_half_z;
  if (m->block_ncon <= 0 || dist > 0) return a_free;
  double K, B, imp = impedance(m->contact_solimp, dist), mu = m->block_friction;
  kb_from_solref(m->contact_solref, m->contact_solimp, m->timestep, &K, &B);
  double R0 = fmax(MJMINVAL, (1 - imp) / imp * (1 + mu * mu) / m->block_mass);
  double R = fmax(MJMINVAL, 2 * mu * mu * R0);
  double D = 4.0 * m->block_ncon / R;
  double aref = -B * vz - K * imp * dist;
  if (a_free >= aref) return a_free; /* J a - aref >= 0: the rows carry no force */
  return (m->block_mass * a_free + D * aref) / (m->block_mass + D);
}
void orc_block_substeps(const orc_sim *s, double *z, double *vz, double fz_applied, int n) {
  for (int t = 0; t < n; t++) {
    *vz += s->m.timestep * block_accel(s, *z, *vz, fz_applied);
    *z += s->m.timestep * *vz;
  }
}

/* n x mj_step on the arm (and the free block's z when blk != NULL); if kin_last != NULL it receives the kinematics
   computed in the LAST substep and blk_xpos the block position of that same instant */
static void substeps_kin(const orc_sim *s, double *qpos, double *qvel, double *warm, const double *ctrl, int n,
                         orc_kin *kin_last, double *blk, double *blk_vz, double *blk_xpos) {
  double h = s->m.timestep;
  for (int t = 0; t < n; t++) {
    orc_kin k;
    double qacc[NJ];
    orc_fk(s, qpos, &k);
    forward_from_kin(s, &k, qpos, qvel, ctrl, warm, qacc, NULL, NULL, NULL);
    for (int j = 0; j < NJ; j++) { /* mj_Euler, SURVEY B.8 (no joint damping -> no implicit term) */
      qvel[j] += h * qacc[j];
      qpos[j] += h * qvel[j];
      warm[j] = qacc[j];
    }
    if (kin_last && t == n - 1) *kin_last = k;
    if (blk) {
      if (blk_xpos && t == n - 1) memcpy(blk_xpos, blk, 3 * sizeof(double));
      if (blk_vz) orc_block_substeps(s, &blk[2], blk_vz, 0.0, 1);
    }
  }
}
void orc_substeps(const orc_sim *s, double *qpos, double *qvel, double *qacc_warm, const double *ctrl, int n) {
  substeps_kin(s, qpos, qvel, qacc_warm, ctrl, n, NULL, NULL, NULL, NULL);
}

/* ------------------------------------------------------------------ construction */
int orc_sizeof_model(void) { return (int)sizeof(orc_model); }
int orc_sizeof_task_cfg(void) { return (int)sizeof(orc_task_cfg); }
int orc_sizeof_env_state(void) { return (int)sizeof(orc_env_state); }

orc_sim *orc_create(const orc_model *m, const orc_task_cfg *cfg) {
  if (!m || !cfg || m->struct_size != (int)sizeof(orc_model) || cfg->struct_size != (int)sizeof(orc_task_cfg)) return NULL;
  if (cfg->num_envs <= 0) return NULL;
  if (cfg->task != 1 && cfg->task != 2 && cfg->task != 5 && cfg->task != 6) return NULL;
  orc_sim *s = (orc_sim *)calloc(1, sizeof(orc_sim));
  s->m = *m; s->cfg = *cfg; s->n = cfg->num_envs;
  s->env = (orc_env_state *)calloc((size_t)s->n, sizeof(orc_env_state));
  q_norm(m->base_quat, s->base_quat);
  for (int i = 0; i < NJ; i++) q_norm(m->body_quat[i], s->body_quat[i]);
  double cq[4];
  q_norm(m->cam_quat, cq);
  q_mat(cq, s->cam_mat);
  s->fy = 0.5 * cfg->cam_res_h / tan(m->cam_fovy_deg * M_PI / 180.0 / 2);
  /* mj_setConst: dof_M0, dof_invweight0 at qpos0 = 0; kv from dampratio */
  double q0[NJ] = {0}, M[NJ * NJ], L[NJ * NJ];
  orc_mass_matrix(s, q0, M);
  memcpy(L, M, sizeof L);
  chol6(L);
  for (int j = 0; j < NJ; j++) {
    double e[NJ] = {0}, x[NJ];
    e[j] = 1;
    chol6_solve(L, e, x);
    s->invw[j] = x[j];
    s->dof_M0[j] = M[j * NJ + j];
    s->kv[j] = m->act_dampratio[j] > 0 ? m->act_dampratio[j] * 2 * sqrt(m->act_kp[j] * s->dof_M0[j]) : m->act_kv[j];
  }
  /* body_invweight0 (mj_setConst): mean diagonal of J M^-1 J^T for the translational Jacobian of the body COM at qpos0 */
  {
    orc_kin k;
    orc_fk(s, q0, &k);
    for (int b = 0; b < NJ; b++) {
      double tr = 0;
      for (int c = 0; c < 3; c++) {
        double Jr[NJ] = {0}, x[NJ];
        for (int j = 0; j <= b; j++) {
          double r[3], col[3];
          for (int a = 0; a < 3; a++) r[a] = k.xipos[b][a] - k.xpos[j][a];
          cross(k.xaxis[j], r, col);
          Jr[j] = col[c];
        }
        chol6_solve(L, Jr, x);
        for (int j = 0; j < NJ; j++) tr += Jr[j] * x[j];
      }
      s->body_tran[b] = fmax(MJMINVAL, tr / 3);
    }
  }
  /* pad <-> floor pair (mj_contactParam): equal priority and solmix -> mean solref / solimp, max friction; then the
     solimp clamps of getsolparam (d0, dmax, midpoint in [mjMINIMP, mjMAXIMP] = [1e-4, 0.9999], width >= 0, power >= 1) */
  for (int i = 0; i < 2; i++) s->pc_solref[i] = 0.5 * (m->pad_solref[i] + m->floor_solref[i]);
  for (int i = 0; i < 5; i++) s->pc_solimp[i] = 0.5 * (m->pad_solimp[i] + m->floor_solimp[i]);
  s->pc_solimp[0] = fmin(0.9999, fmax(0.0001, s->pc_solimp[0]));
  s->pc_solimp[1] = fmin(0.9999, fmax(0.0001, s->pc_solimp[1]));
  s->pc_solimp[2] = fmax(0.0, s->pc_solimp[2]);
  s->pc_solimp[3] = fmin(0.9999, fmax(0.0001, s->pc_solimp[3]));
  s->pc_solimp[4] = fmax(1.0, s->pc_solimp[4]);
  s->pc_mu = fmax(m->pad_friction, m->floor_friction);
  return s;
}
void orc_destroy(orc_sim *s) { if (s) { free(s->env); free(s); } }
int orc_obs_dim(const orc_sim *s) { return s->cfg.task == 5 ? 8 : 15; }
void orc_get_derived(const orc_sim *s, double *dof_M0, double *kv, double *invweight0) {
  for (int j = 0; j < NJ; j++) { dof_M0[j] = s->dof_M0[j]; kv[j] = s->kv[j]; invweight0[j] = s->invw[j]; }
}
orc_env_state *orc_state(orc_sim *s, int env) { return (env >= 0 && env < s->n) ? &s->env[env] : NULL; }
int64_t orc_get_tick(const orc_sim *s) { return s->tick; }
void orc_set_tick(orc_sim *s, int64_t tick) { s->tick = tick; }

/* ------------------------------------------------------------------ bulk state exchange (layout of so100_state_view) */
void orc_set_state_soa(orc_sim *s, const float *qpos, const float *qvel, const float *warm, const float *qcomp, const float *block,
                       const float *snap, const float *aux, const int32_t *cnt, const float *ep_return) {
  const int n = s->n, task = s->cfg.task;
  const double dt = s->m.nsubstep * s->m.timestep;
  for (int i = 0; i < n; i++) {
    orc_env_state *e = &s->env[i];
    for (int j = 0; j < NJ; j++) {
      if (qpos) e->qpos[j] = (double)qpos[j * n + i] - (qcomp ? (double)qcomp[j * n + i] : 0.0);
      if (qvel) e->qvel[j] = qvel[j * n + i];
      if (warm) e->qacc_warm[j] = warm[j * n + i];
    }
    if (block) { for (int k = 0; k < 3; k++) e->block[k] = block[k * n + i]; e->block_vz = task == 5 ? 0.0 : block[3 * n + i]; }
    if (snap) {
      if (task == 5) {
        for (int k = 0; k < 3; k++) e->cam_xpos[k] = snap[k * n + i];
        for (int k = 0; k < 9; k++) e->cam_xmat[k] = snap[(3 + k) * n + i];
      } else {
        for (int k = 0; k < 3; k++) { e->end_pos[k] = snap[k * n + i]; e->block_xpos[k] = snap[(4 + k) * n + i]; }
        e->wrist_pos[0] = e->wrist_pos[1] = 0.0; e->wrist_pos[2] = snap[3 * n + i]; /* only z is ever read (env_base_01.py:213) */
      }
    }
    if (aux) {
      if (task == 2 || task == 6)
        for (int k = 0; k < 3; k++) { e->task_block_pos[k] = aux[k * n + i]; e->last_block_pos[k] = aux[(3 + k) * n + i]; }
      if (task == 5) {
        for (int j = 0; j < NJ; j++) { e->cmd[j] = aux[j * n + i]; e->last_angvel[j] = aux[(6 + j) * n + i]; }
        for (int k = 0; k < 3; k++) e->target[k] = aux[(12 + k) * n + i];
        e->target_dt = aux[15 * n + i]; e->last_centre[0] = aux[16 * n + i]; e->last_centre[1] = aux[17 * n + i];
      }
    }
    if (cnt) {
      int fl = cnt[n + i];
      e->elapsed_steps = cnt[i]; e->time = e->elapsed_steps * dt;
      e->ever_stepped = fl & 1; e->has_last_block = (fl >> 1) & 1; e->angvel_valid = (fl >> 2) & 1; e->centre_valid = (fl >> 3) & 1;
      if (task == 5) { e->miss_count = cnt[2 * n + i]; e->target_time = cnt[3 * n + i] * dt; }
    }
    if (ep_return) e->ep_return = ep_return[i];
  }
}
void orc_get_state_soa(const orc_sim *s, double *qpos, double *qvel, double *block) {
  const int n = s->n;
  for (int i = 0; i < n; i++) {
    const orc_env_state *e = &s->env[i];
    for (int j = 0; j < NJ; j++) { if (qpos) qpos[j * n + i] = e->qpos[j]; if (qvel) qvel[j * n + i] = e->qvel[j]; }
    if (block) { for (int k = 0; k < 3; k++) block[k * n + i] = e->block[k]; block[3 * n + i] = e->block_vz; }
  }
}

/* ------------------------------------------------------------------ task logic */
static void draw(const orc_sim *s, int env, int stream, double u[4], uint32_t raw[4]) {
  uint32_t r[4];
  int64_t gid = s->cfg.env_offset + env;
  orc_philox(s->cfg.seed, (uint32_t)gid, (uint32_t)s->tick, (uint32_t)stream, r);
  for (int i = 0; i < 4; i++) u[i] = u01(r[i]);
  if (raw) memcpy(raw, r, sizeof r);
}
/* block (r, theta) draw of env01_v1.py:45-49 / env02_v1.py:55-59; slot 1 is the discarded theta */
static void place_block(const orc_sim *s, orc_env_state *e, const double u[4]) {
  double dist = s->cfg.block_dist_range[0] + (s->cfg.block_dist_range[1] - s->cfg.block_dist_range[0]) * u[0];
  double theta = -0.5 * M_PI + (-s->cfg.block_theta_half + 2 * s->cfg.block_theta_half * u[2]);
  e->block[0] = dist * cos(theta); e->block[1] = dist * sin(theta); e->block[2] = 0.0;
}
static void obs_env0102(const orc_env_state *e, float *obs) { /* env_base_01.py:241-270 */
  for (int j = 0; j < NJ; j++) obs[j] = (float)e->qpos[j];
  for (int c = 0; c < 3; c++) {
    obs[6 + c] = (float)(e->block_xpos[c] - e->end_pos[c]);
    obs[9 + c] = (float)e->block_xpos[c];
    obs[12 + c] = (float)e->end_pos[c];
  }
}
/* env_base_02.py:88-127 on the STALE camera pose and the fresh block position; returns 1 if detected */
static int project05(const orc_sim *s, const orc_env_state *e, double *cx, double *cy) {
  double rel[3], pc[3];
  for (int c = 0; c < 3; c++) rel[c] = e->block[c] - e->cam_xpos[c];
  mtv(e->cam_xmat, rel, pc);
  double W = s->cfg.cam_res_w, H = s->cfg.cam_res_h;
  double u = s->fy * pc[0] / pc[2] + W / 2, v = s->fy * pc[1] / pc[2] + H / 2;
  if (isnan(u) || isnan(v)) return 0;
  u = trunc(u); v = trunc(v); /* int(): toward zero, so (-1,0) passes the bounds test */
  if (u < 0 || u >= W || v < 0 || v >= H) return 0;
  *cx = (W - u) / W; *cy = (H - v) / H;
  return 1;
}
static void obs_env05(const orc_sim *s, orc_env_state *e, int env, int stream, float *obs) { /* env05_v1.py:32-75 */
  double cx = -1.0, cy = -1.0, px, py;
  if (project05(s, e, &px, &py)) {
    double u[4];
    draw(s, env, stream, u, NULL);
    cx = px + (-s->cfg.obs_noise + 2 * s->cfg.obs_noise * u[0]);
    cy = py + (-s->cfg.obs_noise + 2 * s->cfg.obs_noise * u[1]);
  }
  for (int j = 0; j < NJ; j++) obs[j] = (float)e->cmd[j];
  obs[6] = (float)cx; obs[7] = (float)cy;
}
static void snapshot(orc_env_state *e, const orc_kin *k) {
  memcpy(e->end_pos, k->end_pos, sizeof e->end_pos);
  memcpy(e->wrist_pos, k->wrist_pos, sizeof e->wrist_pos);
  memcpy(e->cam_xpos, k->cam_xpos, sizeof e->cam_xpos);
  memcpy(e->cam_xmat, k->cam_xmat, sizeof e->cam_xmat);
}

/* MujocoEnv.reset = mj_resetData + reset_model (no mj_forward: kinematics stay zero) */
static void reset_env(const orc_sim *s, orc_env_state *e, int env, int stream, float *obs) {
  const orc_task_cfg *c = &s->cfg;
  double u[4];
  uint32_t raw[4];
  memset(e->qpos, 0, sizeof e->qpos); memset(e->qvel, 0, sizeof e->qvel);
  memset(e->qacc_warm, 0, sizeof e->qacc_warm); memset(e->ctrl, 0, sizeof e->ctrl);
  e->time = 0; e->elapsed_steps = 0; e->ep_return = 0; e->block_vz = 0; /* mj_resetData zeroes qvel */
  memset(e->end_pos, 0, sizeof e->end_pos); memset(e->wrist_pos, 0, sizeof e->wrist_pos);
  memset(e->block_xpos, 0, sizeof e->block_xpos);
  memset(e->cam_xpos, 0, sizeof e->cam_xpos); memset(e->cam_xmat, 0, sizeof e->cam_xmat);
  if (c->task == 1) { /* env01_v1.py:39-63 */
    draw(s, env, stream, u, raw);
    place_block(s, e, u);
    int idx = (int)(((uint64_t)raw[3] * (uint64_t)c->n_start) >> 32);
    for (int j = 0; j < NJ - 1; j++) e->qpos[j] = c->start_positions[idx][j]; /* Jaw skipped */
  } else if (c->task == 2 || c->task == 6) { /* env02_v1.py:70-81, :52-68; env06_v1.py:53-82 is the same code */
    draw(s, env, stream, u, NULL);
    double prev[3] = {e->task_block_pos[0], e->task_block_pos[1], e->task_block_pos[2]};
    place_block(s, e, u);
    if (!e->has_last_block) memcpy(e->last_block_pos, e->block, sizeof e->last_block_pos);
    else memcpy(e->last_block_pos, prev, sizeof prev);
    memcpy(e->task_block_pos, e->block, sizeof e->task_block_pos);
    e->has_last_block = 1;
    for (int j = 0; j < NJ; j++) e->qpos[j] = c->rest_position[j];
  } else { /* env03_v1.py:203-215, :35-57 */
    for (int j = 0; j < NJ; j++) { e->qpos[j] = c->start_position05[j]; e->cmd[j] = c->start_position05[j]; }
    for (int k = 0; k < 3; k++) {
      e->target[k] = (c->block_space_start[0][k] + c->block_space_start[1][k]) / 2;
      e->block[k] = e->target[k];
    }
    e->target_dt = 0.01; e->target_time = 0.0;
    e->centre_valid = 0; e->miss_count = 0;
  }
  if (c->flags & 1u) { /* SO100_FLAG_FRESH_FK_ON_RESET: what a mj_forward in reset_model would give */
    orc_kin k;
    orc_fk(s, e->qpos, &k);
    snapshot(e, &k);
    memcpy(e->block_xpos, e->block, sizeof e->block_xpos);
  }
  if (c->task == 5) obs_env05(s, e, env, STREAM_RESET_NOISE, obs); else obs_env0102(e, obs);
}

static double joint_penalty(const orc_sim *s, const double *ang) { /* env_base_01.py:144-163 */
  double r = 0;
  for (int j = 0; j < NJ; j++) {
    double lo = s->m.jnt_range[j][0], hi = s->m.jnt_range[j][1];
    double lt = lo + 0.05 * (hi - lo), ut = hi - 0.05 * (hi - lo);
    if (ang[j] < lt) r -= (lt - ang[j]) * 10.0;
    else if (ang[j] > ut) r -= (ang[j] - ut) * 10.0;
  }
  return r;
}
static double reward_env0102(const orc_sim *s, orc_env_state *e, int in_reach) { /* env_base_01.py:180-239; env_base_06.py:200-264 */
  double reward = 0, d[3];
  for (int c = 0; c < 3; c++) d[c] = e->block_xpos[c] - e->end_pos[c];
  double distance = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (e->block_xpos[1] < -0.1) {
    double pitch = e->qpos[1];
    if (e->ever_stepped && pitch < -0.7 * M_PI) reward += (pitch + 0.7 * M_PI) * 0.7;
  }
  if (e->ever_stepped && e->end_pos[2] < 0.02) reward += (e->end_pos[2] - 0.02) * 20.0;
  if (e->ever_stepped && e->wrist_pos[2] < 0.08) {
    double w = (e->wrist_pos[2] - 0.08) * 10.0;
    reward += fmin(fmax(w, -0.8), 0.8);
  }
  reward += fmin(-distance + 0.02, 0.0) * 0.5;
  if (in_reach) { /* Env06 only: env_base_06.py:149-162, sigmoid on the normalised jaw opening */
    double jn = fmin(fmax((e->qpos[5] + 0.2) / 2.2, 0.0), 1.0);
    reward += 100.0 * (1.0 / (1.0 + exp(-10 * (jn - 0.3))));
  }
  reward += joint_penalty(s, e->qpos);
  e->ever_stepped = 1;
  return reward;
}

static void step_env(const orc_sim *s, orc_env_state *e, int env, const float *act, float *obs, double *reward,
                     uint8_t *terminated, uint8_t *truncated, float *terminal_obs, double *ep_return, int32_t *ep_len) {
  const orc_task_cfg *c = &s->cfg;
  const int od = c->task == 5 ? 8 : 15;
  double a[NJ], rew = 0;
  int term = 0;
  for (int j = 0; j < NJ; j++) {
    a[j] = (double)act[j];
    if (c->flags & 2u) a[j] = fmin(fmax(a[j], -1.0), 1.0);
  }
  orc_kin kin;
  if (c->task == 1 || c->task == 2 || c->task == 6) {
    int in_reach = 0;
    if (c->task == 6) { /* env06_v1.py:19 */
      double d[3];
      for (int k = 0; k < 3; k++) d[k] = e->block_xpos[k] - e->end_pos[k];
      in_reach = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < c->reach_threshold;
    }
    rew = reward_env0102(s, e, in_reach); /* PRE-step reward on stale kinematics */
    for (int j = 0; j < NJ; j++) e->ctrl[j] = e->qpos[j] + a[j] * c->joint_step_scale;
    if (c->task == 6 && in_reach) { /* env06_v1.py:30-38: bonus without relocation */
      double bd = 0;
      for (int k = 0; k < 3; k++) bd += (e->task_block_pos[k] - e->last_block_pos[k]) * (e->task_block_pos[k] - e->last_block_pos[k]);
      rew += sqrt(bd) * 20;
    }
    if (c->task == 2) {
      double d[3];
      for (int k = 0; k < 3; k++) d[k] = e->block_xpos[k] - e->end_pos[k];
      if (sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < c->reach_threshold) { /* env02_v1.py:29-37 */
        double bd = 0, u[4];
        for (int k = 0; k < 3; k++) bd += (e->task_block_pos[k] - e->last_block_pos[k]) * (e->task_block_pos[k] - e->last_block_pos[k]);
        rew += sqrt(bd) * 20;
        draw(s, env, STREAM_TASK, u, NULL);
        memcpy(e->last_block_pos, e->task_block_pos, sizeof e->last_block_pos);
        place_block(s, e, u);
        memcpy(e->task_block_pos, e->block, sizeof e->task_block_pos);
      }
    }
    /* SO100_FLAG_STATIC_BLOCK (8): the block is held at its spawn pose (no gravity, no contact) */
    substeps_kin(s, e->qpos, e->qvel, e->qacc_warm, e->ctrl, s->m.nsubstep, &kin, e->block,
                 (c->flags & 8u) ? NULL : &e->block_vz, e->block_xpos);
    snapshot(e, &kin);
    obs_env0102(e, obs);
  } else {
    /* env03_v1.py:124-201 */
    double time = e->elapsed_steps * (s->m.nsubstep * s->m.timestep);
    double f = fmin(time / c->ramp_seconds, 1.0);
    double smin[3], smax[3], speed;
    for (int k = 0; k < 3; k++) {
      smin[k] = c->block_space_start[0][k] + f * (c->block_space_end[0][k] - c->block_space_start[0][k]);
      smax[k] = c->block_space_start[1][k] + f * (c->block_space_end[1][k] - c->block_space_start[1][k]);
    }
    speed = f <= 0.05 ? c->block_speed_min : c->block_speed_min + (f - 0.05) * (c->block_speed_max - c->block_speed_min) / (1.0 - 0.05);
    double d[3], dist;
    for (int k = 0; k < 3; k++) d[k] = e->target[k] - e->block[k];
    dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (!(time - e->target_time < e->target_dt && dist > 0.02)) { /* :77-93 */
      double u[4];
      draw(s, env, STREAM_TASK, u, NULL);
      for (int k = 0; k < 3; k++) e->target[k] = smin[k] + (smax[k] - smin[k]) * u[k];
      e->target_dt = 1.2 + (5.1 - 1.2) * u[3];
      e->target_time = time;
    }
    for (int k = 0; k < 3; k++) d[k] = e->target[k] - e->block[k]; /* :95-122 */
    dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (dist > 0) {
      double sd = fmin(speed * s->m.timestep, dist);
      for (int k = 0; k < 3; k++) e->block[k] += d[k] / dist * sd;
    }
    double newcmd[NJ];
    for (int j = 0; j < NJ; j++) { newcmd[j] = e->cmd[j] + a[j] * c->joint_step_scale; e->ctrl[j] = newcmd[j]; }
    /* Env05's block is scripted: repositioned, its velocity zeroed and gravity cancelled every env step (:95-122) */
    substeps_kin(s, e->qpos, e->qvel, e->qacc_warm, e->ctrl, s->m.nsubstep, &kin, e->block, NULL, e->block_xpos);
    snapshot(e, &kin);
    obs_env05(s, e, env, STREAM_NOISE, obs);
    if (obs[6] == -1.0f && obs[7] == -1.0f) { /* :152-164 */
      if (e->miss_count > c->lost_limit) term = 1;
      e->miss_count += 1;
    } else { e->last_centre[0] = obs[6]; e->last_centre[1] = obs[7]; e->centre_valid = 1; e->miss_count = 0; }
    rew = 0.5;
    if (e->centre_valid) rew += -sqrt((0.5 - e->last_centre[0]) * (0.5 - e->last_centre[0]) + (0.5 - e->last_centre[1]) * (0.5 - e->last_centre[1]));
    rew += joint_penalty(s, e->cmd); /* commanded (old) angles, SURVEY Q7 */
    double pen = 0, av[NJ]; /* env_base_01.py:165-178 with timestep = 0.002 */
    for (int j = 0; j < NJ; j++) av[j] = (newcmd[j] - e->cmd[j]) / s->m.timestep;
    if (e->angvel_valid) for (int j = 0; j < NJ; j++) pen += fabs(av[j] - e->last_angvel[j]) * 0.0025;
    memcpy(e->last_angvel, av, sizeof av);
    e->angvel_valid = 1;
    rew += -pen * f;
    obs[6] *= 5; obs[7] *= 5;
    memcpy(e->cmd, newcmd, sizeof newcmd);
  }
  e->elapsed_steps += 1;
  e->time = e->elapsed_steps * (s->m.nsubstep * s->m.timestep);
  e->ep_return += rew;
  int trunc_ = e->elapsed_steps >= c->max_episode_steps;
  *reward = rew; *terminated = (uint8_t)term; *truncated = (uint8_t)(trunc_ && !term);
  if (term || trunc_) {
    if (terminal_obs) memcpy(terminal_obs, obs, sizeof(float) * od);
    if (ep_return) *ep_return = e->ep_return;
    if (ep_len) *ep_len = e->elapsed_steps;
    reset_env(s, e, env, STREAM_RESET, obs);
  }
}

/* ---- minimal pthread parallel-for over envs (this image's gcc has no libgomp) */
typedef struct {
  orc_sim *s; int lo, hi, is_step;
  const uint8_t *mask; const float *actions; float *obs; double *reward; uint8_t *terminated, *truncated;
  float *terminal_obs; double *ep_return; int32_t *ep_len;
} job_t;
static void *job_run(void *p) {
  job_t *j = (job_t *)p;
  orc_sim *s = j->s;
  const int od = orc_obs_dim(s);
  for (int i = j->lo; i < j->hi; i++) {
    if (j->is_step)
      step_env(s, &s->env[i], i, j->actions + (size_t)i * NJ, j->obs + (size_t)i * od, j->reward + i, j->terminated + i,
               j->truncated + i, j->terminal_obs ? j->terminal_obs + (size_t)i * od : NULL,
               j->ep_return ? j->ep_return + i : NULL, j->ep_len ? j->ep_len + i : NULL);
    else if (!j->mask || j->mask[i]) reset_env(s, &s->env[i], i, STREAM_API_RESET, j->obs + (size_t)i * od);
  }
  return NULL;
}
int orc_hw_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
static void run_jobs(job_t *proto, int nthreads) {
  int n = proto->s->n;
  if (nthreads <= 0) nthreads = orc_hw_threads();
  if (nthreads > n) nthreads = n;
  if (nthreads > 256) nthreads = 256;
  if (nthreads <= 1) { proto->lo = 0; proto->hi = n; job_run(proto); return; }
  pthread_t th[256];
  job_t jobs[256];
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = *proto;
    jobs[t].lo = (int)((int64_t)n * t / nthreads); jobs[t].hi = (int)((int64_t)n * (t + 1) / nthreads);
    pthread_create(&th[t], NULL, job_run, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

void orc_reset(orc_sim *s, const uint8_t *mask, float *obs, int nthreads) {
  job_t j = {0};
  j.s = s; j.is_step = 0; j.mask = mask; j.obs = obs;
  run_jobs(&j, nthreads);
}

void orc_step(orc_sim *s, const float *actions, float *obs, double *reward, uint8_t *terminated,
              uint8_t *truncated, float *terminal_obs, double *ep_return, int32_t *ep_len, int nthreads) {
  s->tick += 1;
  job_t j = {0};
  j.s = s; j.is_step = 1; j.actions = actions; j.obs = obs; j.reward = reward; j.terminated = terminated;
  j.truncated = truncated; j.terminal_obs = terminal_obs; j.ep_return = ep_return; j.ep_len = ep_len;
  run_jobs(&j, nthreads);
}
