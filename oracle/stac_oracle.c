/*
 * stac_oracle.c -- CPU restatement of the STAC fitting hot path of talmolab/stac-mjx.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under stac_mjx_b200/ may import, link or
 * execute this file; it is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker / CPU arm.
 *
 * PARITY STATUS
 *   m-phase (closed-form offsets) + hinge-chain FK : PINNED against the six
 *     known-answer tests of reference tests/unit/test_m_opt.py:72-225
 *     (tests/test_oracle_cpu.py).
 *   q-phase (FISTA projected gradient)             : PARITY UNPINNED.  The reference
 *     holds no golden qpos and its arithmetic lives in two
 *     un-vendored dependencies that cannot be imported in this image:
 *       - mujoco-mjx (unpinned in reference pyproject.toml:23-24):
 *           mujoco.mjx._src.smooth.kinematics, mujoco.mjx._src.math
 *       - jaxopt==0.8.5 (reference pyproject.toml:35):
 *           jaxopt.ProjectedGradient -> ProximalGradient (_update_accel, _ls, _error),
 *           jaxopt.projection.projection_box, base.IterativeSolver.run
 *     Their published algorithms are restated below and anchored on the reference's
 *     call sites: stac_mjx/stac_core.py:27-63 (q_loss), :66-99 (_q_opt), :102-172
 *     (_m_opt); stac_mjx/compute_stac.py:17-104 (root_optimization), :170-278
 *     (pose_optimization); stac_mjx/utils.py:129-169 (make_qs / replace_qs).
 *
 * THREE ARITHMETIC ORDERS (argument `mode`)
 *   mode 0  "mjx order":  body-by-body world-frame FK exactly in the operation order
 *           of MJX smooth.kinematics, libm sin/cos, left-to-right sums, direct
 *           Jacobian-transpose gradient.  This is the faithful restatement.
 *   mode 1  "canonical":  the same mathematics in the operation order the CUDA
 *           kernels use (parent-frame local transforms, pointer-jumping composition,
 *           32-slot butterfly sums, prefix-scan wrench gradient, polynomial sincos,
 *           explicit fused multiply-adds).  Compiled as float32 it is meant to be
 *           bit-identical to the GPU path, so solver decisions (line-search accepts,
 *           stopping iteration) coincide; tests/ quantify mode0-vs-mode1 and
 *           f32-vs-f64 spreads as the floating-point noise floor of the algorithm.
 *
 *   mode 2  "fast order": what libstacb computes by default.  For the models the register-resident
 *           hinge-tree solver serves (rodent, C. elegans, ...) q_loss / _q_opt follow the operation
 *           order of stacb_fast.cuh (fast_order.h); everything else is mode 1.
 *   Mode 0 is FROZEN: tests/test_oracle_cpu.py pins a digest of its float64 outputs, so the
 *   faithful restatement cannot drift when modes 1 / 2 follow a kernel change.
 *
 * Build: oracle/build.sh  (gcc -O2 -ffp-contract=off -mfma; REAL=float and REAL=double)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif

#define IS_F32 (sizeof(REAL) == 4)
#define R(x) ((REAL)(x))

static inline REAL r_fma(REAL a, REAL b, REAL c) { return IS_F32 ? (REAL)fmaf((float)a, (float)b, (float)c) : (REAL)fma(a, b, c); }
static inline REAL r_sqrt(REAL a) { return IS_F32 ? (REAL)sqrtf((float)a) : (REAL)sqrt(a); }
static inline REAL r_min(REAL a, REAL b) { return a < b ? a : b; }
static inline REAL r_max(REAL a, REAL b) { return a > b ? a : b; }
#define R_EPS (IS_F32 ? (REAL)FLT_EPSILON : (REAL)DBL_EPSILON)

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
#define LANES 32

typedef struct {
  int32_t nbody, nq, njnt, nsite;
  const int32_t *body_parent, *body_jntadr, *body_jntnum;
  const REAL *body_pos, *body_quat;
  const int32_t *jnt_type, *jnt_qposadr, *jnt_bodyid;
  const REAL *jnt_pos, *jnt_axis, *qpos0;
  const int32_t *site_body; /* body id of each keypoint site, keypoint order */
} omodel;

/* ------------------------------------------------------------------ */
/* primitive operations                                               */
/* ------------------------------------------------------------------ */

typedef struct { REAL x, y, z; } v3;
typedef struct { REAL w, x, y, z; } q4;

/* --- mode 0: expressions as written in mujoco.mjx._src.math (no explicit fma) --- */

static inline v3 m_rotate(v3 v, q4 q) {
  /* math.rotate: r = 2*(dot(u,v)*u) + (s*s - dot(u,u))*v + 2*s*cross(u,v) */
  REAL s = q.w;
  REAL duv = q.x * v.x + q.y * v.y + q.z * v.z;
  REAL duu = q.x * q.x + q.y * q.y + q.z * q.z;
  REAL k = s * s - duu;
  v3 c = { q.y * v.z - q.z * v.y, q.z * v.x - q.x * v.z, q.x * v.y - q.y * v.x };
  v3 r;
  r.x = R(2) * (duv * q.x) + k * v.x;
  r.y = R(2) * (duv * q.y) + k * v.y;
  r.z = R(2) * (duv * q.z) + k * v.z;
  r.x = r.x + R(2) * s * c.x;
  r.y = r.y + R(2) * s * c.y;
  r.z = r.z + R(2) * s * c.z;
  return r;
}

static inline q4 m_qmul(q4 u, q4 v) {
  /* math.quat_mul */
  q4 r;
  r.w = u.w * v.w - u.x * v.x - u.y * v.y - u.z * v.z;
  r.x = u.w * v.x + u.x * v.w + u.y * v.z - u.z * v.y;
  r.y = u.w * v.y - u.x * v.z + u.y * v.w + u.z * v.x;
  r.z = u.w * v.z + u.x * v.y - u.y * v.x + u.z * v.w;
  return r;
}

static inline q4 m_normalize4(q4 q, REAL *n_out) {
  /* math.normalize: x / (n + 1e-6 * (n == 0)) */
  REAL n = r_sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  REAL d = n + (n == R(0) ? R(1e-6) : R(0));
  q4 r = { q.w / d, q.x / d, q.y / d, q.z / d };
  if (n_out) *n_out = d;
  return r;
}

static inline q4 m_axis_angle(v3 a, REAL ang) {
  /* math.axis_angle_to_quat */
  REAL h = ang * R(0.5);
  REAL s = IS_F32 ? (REAL)sinf((float)h) : (REAL)sin(h);
  REAL c = IS_F32 ? (REAL)cosf((float)h) : (REAL)cos(h);
  q4 r = { c, a.x * s, a.y * s, a.z * s };
  return r;
}

/* --- mode 1: canonical fused forms (mirrored op-for-op by csrc/stacb_math.cuh) --- */

static inline REAL c_dot3(v3 a, v3 b) { return r_fma(a.z, b.z, r_fma(a.y, b.y, a.x * b.x)); }
static inline v3 c_cross(v3 a, v3 b) {
  v3 r = { r_fma(a.y, b.z, -(a.z * b.y)), r_fma(a.z, b.x, -(a.x * b.z)), r_fma(a.x, b.y, -(a.y * b.x)) };
  return r;
}
static inline v3 c_rotate(v3 v, q4 q) {
  v3 u = { q.x, q.y, q.z };
  REAL d = c_dot3(u, v);
  REAL uu = c_dot3(u, u);
  REAL k = r_fma(q.w, q.w, -uu);
  v3 c = c_cross(u, v);
  REAL d2 = R(2) * d, s2 = R(2) * q.w;
  v3 r = { r_fma(s2, c.x, r_fma(k, v.x, d2 * u.x)), r_fma(s2, c.y, r_fma(k, v.y, d2 * u.y)),
           r_fma(s2, c.z, r_fma(k, v.z, d2 * u.z)) };
  return r;
}
static inline q4 c_qmul(q4 u, q4 v) {
  q4 r;
  r.w = r_fma(-u.z, v.z, r_fma(-u.y, v.y, r_fma(-u.x, v.x, u.w * v.w)));
  r.x = r_fma(-u.z, v.y, r_fma(u.y, v.z, r_fma(u.x, v.w, u.w * v.x)));
  r.y = r_fma(u.z, v.x, r_fma(u.y, v.w, r_fma(-u.x, v.z, u.w * v.y)));
  r.z = r_fma(u.z, v.w, r_fma(-u.y, v.x, r_fma(u.x, v.y, u.w * v.z)));
  return r;
}
/* canonical: one division, x * (1/d); *n_out receives the RECIPROCAL 1/d (the canonical gradient multiplies by it) */
static inline q4 c_normalize4(q4 q, REAL *n_out) {
  REAL n2 = r_fma(q.z, q.z, r_fma(q.y, q.y, r_fma(q.x, q.x, q.w * q.w)));
  REAL n = r_sqrt(n2);
  REAL d = n + (n == R(0) ? R(1e-6) : R(0));
  REAL ri = R(1) / d;
  q4 r = { q.w * ri, q.x * ri, q.y * ri, q.z * ri };
  if (n_out) *n_out = ri;
  return r;
}
static inline void c_sincos(REAL x, REAL *sp, REAL *cp) {
  if (!IS_F32) { *sp = (REAL)sin(x); *cp = (REAL)cos(x); return; }
  /* 3-term Cody-Waite reduction by pi/2, Cephes-style minimax polynomials */
  float xf = (float)x;
  float j = rintf(xf * 0.636619747f);
  float r = fmaf(-j, 1.57079637e+00f, xf);
  r = fmaf(-j, -4.37113883e-08f, r);
  r = fmaf(-j, -1.71512489e-15f, r);
  float r2 = r * r;
  float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, r2, -1.6666654611e-1f);
  float s = fmaf(r * r2, ps, r);
  float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, r2, 4.166664568298827e-2f);
  float c = fmaf(r2 * r2, pc, fmaf(-0.5f, r2, 1.0f));
  int q = ((int)j) & 3;
  float so = (q & 1) ? c : s, co = (q & 1) ? s : c;
  if (q == 1 || q == 2) co = -co;
  if (q >= 2) so = -so;
  *sp = (REAL)so; *cp = (REAL)co;
}
static inline q4 c_axis_angle(v3 a, REAL ang) {
  REAL s, c;
  c_sincos(ang * R(0.5), &s, &c);
  q4 r = { c, a.x * s, a.y * s, a.z * s };
  return r;
}

static inline v3 ld3(const REAL *p) { v3 r = { p[0], p[1], p[2] }; return r; }
static inline q4 ld4(const REAL *p) { q4 r = { p[0], p[1], p[2], p[3] }; return r; }
static inline void st3(REAL *p, v3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
static inline void st4(REAL *p, q4 q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }
static inline v3 add3(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static inline v3 sub3(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }

/* ------------------------------------------------------------------ */
/* schedule derived from the tree (shared by both modes)              */
/* ------------------------------------------------------------------ */

typedef struct {
  int nbody, njnt, K;
  int *depth;          /* [nbody] */
  int *subsize;        /* [nbody] size of DFS subtree (ids are pre-order) */
  int nact, *act;      /* active bodies (ancestors-or-self of site bodies), ascending */
  int nfull, *full;    /* bodies 1..nbody-1 */
  int rounds_act, rounds_full;
  int *anc;            /* [rounds_full][nbody]: ancestor at distance 2^r, -1 if none (or world) */
  int *site_order;     /* sorted position -> keypoint index (sorted by body id, then index) */
  int *jnt_s, *jnt_e;  /* [njnt] range of sorted site positions under the joint's body */
  int spl;             /* sorted sites per lane = ceil(K/32) */
  int cbody;           /* reference body for torques: first active body */
} osched;

static int ceil_log2(int x) { int r = 0; while ((1 << r) < x) r++; return r; }

static osched *sched_create(const omodel *m) {
  osched *s = (osched *)calloc(1, sizeof(osched));
  int nb = m->nbody, K = m->nsite;
  s->nbody = nb; s->njnt = m->njnt; s->K = K;
  s->depth = (int *)calloc(nb, sizeof(int));
  s->subsize = (int *)calloc(nb, sizeof(int));
  for (int b = 1; b < nb; b++) s->depth[b] = s->depth[m->body_parent[b]] + 1;
  for (int b = nb - 1; b >= 0; b--) { s->subsize[b] += 1; if (b > 0) s->subsize[m->body_parent[b]] += s->subsize[b]; }
  char *isact = (char *)calloc(nb, 1);
  for (int k = 0; k < K; k++) { int b = m->site_body[k]; while (b != 0 && !isact[b]) { isact[b] = 1; b = m->body_parent[b]; } }
  s->act = (int *)calloc(nb, sizeof(int)); s->full = (int *)calloc(nb, sizeof(int));
  int da = 1, df = 1;
  for (int b = 1; b < nb; b++) {
    s->full[s->nfull++] = b; if (s->depth[b] > df) df = s->depth[b];
    if (isact[b]) { s->act[s->nact++] = b; if (s->depth[b] > da) da = s->depth[b]; }
  }
  free(isact);
  s->rounds_act = ceil_log2(da); s->rounds_full = ceil_log2(df);
  int nr = s->rounds_full > 0 ? s->rounds_full : 1;
  s->anc = (int *)calloc((size_t)nr * nb, sizeof(int));
  for (int b = 0; b < nb; b++) { int p = b > 0 ? m->body_parent[b] : 0; s->anc[b] = (b > 0 && p != 0) ? p : -1; }
  for (int r = 1; r < nr; r++)
    for (int b = 0; b < nb; b++) { int a = s->anc[(r - 1) * nb + b]; s->anc[r * nb + b] = (a >= 0) ? s->anc[(r - 1) * nb + a] : -1; }
  /* sites sorted by (body id, keypoint index): stable insertion sort */
  s->site_order = (int *)calloc(K > 0 ? K : 1, sizeof(int));
  for (int k = 0; k < K; k++) {
    int i = k;
    while (i > 0 && m->site_body[s->site_order[i - 1]] > m->site_body[k]) { s->site_order[i] = s->site_order[i - 1]; i--; }
    s->site_order[i] = k;
  }
  s->jnt_s = (int *)calloc(m->njnt > 0 ? m->njnt : 1, sizeof(int)); s->jnt_e = (int *)calloc(m->njnt > 0 ? m->njnt : 1, sizeof(int));
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], lo = b, hi = b + s->subsize[b];
    int a = 0; while (a < K && m->site_body[s->site_order[a]] < lo) a++;
    int e = a; while (e < K && m->site_body[s->site_order[e]] < hi) e++;
    s->jnt_s[j] = a; s->jnt_e[j] = e;
  }
  s->spl = (K + LANES - 1) / LANES; if (s->spl < 1) s->spl = 1;
  s->cbody = s->nact > 0 ? s->act[0] : 0;
  return s;
}

static void sched_destroy(osched *s) {
  if (!s) return;
  free(s->depth); free(s->subsize); free(s->act); free(s->full); free(s->anc); free(s->site_order); free(s->jnt_s); free(s->jnt_e); free(s);
}

/* ------------------------------------------------------------------ */
/* workspace                                                          */
/* ------------------------------------------------------------------ */

typedef struct {
  REAL *P, *Q;         /* world pose per body [nbody*3], [nbody*4] */
  REAL *P2, *Q2;       /* double buffer for pointer jumping */
  REAL *anchor, *axis; /* per joint [njnt*3]: world (mode 0) or parent-frame (mode 1) */
  REAL *qn;            /* qpos after FK (normalised quaternions) [nq] */
  REAL *S;             /* site world positions, keypoint order [K*3] */
  REAL *res;           /* masked residuals [K*3] */
  REAL *W;             /* wrench scratch [ (K+1) * 6 ] */
  REAL *qfull;         /* [nq] */
  REAL *fnorm;         /* per-joint quaternion normalisation: divisor (mode 0) or its reciprocal (mode 1) [njnt] */
} owork;

static owork *work_create(const omodel *m) {
  owork *w = (owork *)calloc(1, sizeof(owork));
  int nb = m->nbody, nj = m->njnt > 0 ? m->njnt : 1, K = m->nsite > 0 ? m->nsite : 1;
  w->P = (REAL *)calloc(nb * 3, sizeof(REAL)); w->Q = (REAL *)calloc(nb * 4, sizeof(REAL));
  w->P2 = (REAL *)calloc(nb * 3, sizeof(REAL)); w->Q2 = (REAL *)calloc(nb * 4, sizeof(REAL));
  w->anchor = (REAL *)calloc(nj * 3, sizeof(REAL)); w->axis = (REAL *)calloc(nj * 3, sizeof(REAL));
  w->qn = (REAL *)calloc(m->nq > 0 ? m->nq : 1, sizeof(REAL));
  w->S = (REAL *)calloc(K * 3, sizeof(REAL)); w->res = (REAL *)calloc(K * 3, sizeof(REAL));
  w->W = (REAL *)calloc((K + LANES + 1) * 6, sizeof(REAL));
  w->qfull = (REAL *)calloc(m->nq > 0 ? m->nq : 1, sizeof(REAL));
  w->fnorm = (REAL *)calloc(nj, sizeof(REAL));
  return w;
}
static void work_destroy(owork *w) {
  if (!w) return;
  free(w->P); free(w->Q); free(w->P2); free(w->Q2); free(w->anchor); free(w->axis); free(w->qn); free(w->S); free(w->res); free(w->W); free(w->qfull); free(w->fnorm); free(w);
}

/* ------------------------------------------------------------------ */
/* forward kinematics                                                 */
/* ------------------------------------------------------------------ */

/* mode 0: MJX smooth.kinematics over the listed bodies (ascending ids, parents first).
 * anchor/axis are the world-frame xanchor/xaxis MJX stores. */
static void fk_mjx(const omodel *m, const int *set, int nset, const REAL *q, owork *w) {
  memcpy(w->qn, q, sizeof(REAL) * m->nq);
  w->P[0] = w->P[1] = w->P[2] = R(0); w->Q[0] = R(1); w->Q[1] = w->Q[2] = w->Q[3] = R(0);
  for (int i = 0; i < nset; i++) {
    int b = set[i], p = m->body_parent[b];
    v3 pos = ld3(m->body_pos + 3 * b); q4 quat = ld4(m->body_quat + 4 * b);
    /* the scan's root level has no carry: top-level bodies keep their own pos/quat */
    if (p != 0) { q4 pq = ld4(w->Q + 4 * p); pos = add3(ld3(w->P + 3 * p), m_rotate(pos, pq)); quat = m_qmul(pq, quat); }
    for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
      int j = m->body_jntadr[b] + jj, adr = m->jnt_qposadr[j], t = m->jnt_type[j];
      v3 jpos = ld3(m->jnt_pos + 3 * j), jaxis = ld3(m->jnt_axis + 3 * j);
      v3 anchor, axis;
      if (t == JNT_FREE) { anchor = ld3(q + adr); axis.x = R(0); axis.y = R(0); axis.z = R(1); }
      else { anchor = add3(m_rotate(jpos, quat), pos); axis = m_rotate(jaxis, quat); }
      st3(w->anchor + 3 * j, anchor); st3(w->axis + 3 * j, axis);
      if (t == JNT_FREE) {
        pos = ld3(q + adr);
        quat = m_normalize4(ld4(q + adr + 3), &w->fnorm[j]);
        st4(w->qn + adr + 3, quat);
      } else if (t == JNT_BALL) {
        q4 ql = m_normalize4(ld4(q + adr), &w->fnorm[j]);
        st4(w->qn + adr, ql);
        quat = m_qmul(quat, ql);
        pos = sub3(anchor, m_rotate(jpos, quat));
      } else if (t == JNT_HINGE) {
        q4 ql = m_axis_angle(jaxis, q[adr] - m->qpos0[adr]);
        quat = m_qmul(quat, ql);
        pos = sub3(anchor, m_rotate(jpos, quat));
      } else { /* slide */
        REAL d = q[adr] - m->qpos0[adr];
        pos.x = pos.x + axis.x * d; pos.y = pos.y + axis.y * d; pos.z = pos.z + axis.z * d;
      }
    }
    st3(w->P + 3 * b, pos); st4(w->Q + 4 * b, quat);
  }
}

/* mode 1: parent-frame local transforms, then pointer-jumping composition.
 * anchor/axis are stored in the PARENT frame of the joint's body. */
static void fk_canon(const omodel *m, const osched *s, const int *set, int nset, int rounds, const REAL *q, owork *w) {
  int nb = m->nbody;
  memcpy(w->qn, q, sizeof(REAL) * m->nq);
  w->P[0] = w->P[1] = w->P[2] = R(0); w->Q[0] = R(1); w->Q[1] = w->Q[2] = w->Q[3] = R(0);
  for (int i = 0; i < nset; i++) {
    int b = set[i];
    v3 pos = ld3(m->body_pos + 3 * b); q4 quat = ld4(m->body_quat + 4 * b);
    for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
      int j = m->body_jntadr[b] + jj, adr = m->jnt_qposadr[j], t = m->jnt_type[j];
      v3 jpos = ld3(m->jnt_pos + 3 * j), jaxis = ld3(m->jnt_axis + 3 * j);
      v3 anchor, axis;
      if (t == JNT_FREE) { anchor = ld3(q + adr); axis.x = R(0); axis.y = R(0); axis.z = R(1); }
      else { anchor = add3(c_rotate(jpos, quat), pos); axis = c_rotate(jaxis, quat); }
      st3(w->anchor + 3 * j, anchor); st3(w->axis + 3 * j, axis);
      if (t == JNT_FREE) {
        pos = ld3(q + adr);
        quat = c_normalize4(ld4(q + adr + 3), &w->fnorm[j]);
        st4(w->qn + adr + 3, quat);
      } else if (t == JNT_BALL) {
        q4 ql = c_normalize4(ld4(q + adr), &w->fnorm[j]);
        st4(w->qn + adr, ql);
        quat = c_qmul(quat, ql);
        pos = sub3(anchor, c_rotate(jpos, quat));
      } else if (t == JNT_HINGE) {
        q4 ql = c_axis_angle(jaxis, q[adr] - m->qpos0[adr]);
        quat = c_qmul(quat, ql);
        pos = sub3(anchor, c_rotate(jpos, quat));
      } else {
        REAL d = q[adr] - m->qpos0[adr];
        pos.x = r_fma(axis.x, d, pos.x); pos.y = r_fma(axis.y, d, pos.y); pos.z = r_fma(axis.z, d, pos.z);
      }
    }
    st3(w->P + 3 * b, pos); st4(w->Q + 4 * b, quat);
  }
  for (int r = 0; r < rounds; r++) {
    const int *anc = s->anc + (size_t)r * nb;
    for (int i = 0; i < nset; i++) {
      int b = set[i], a = anc[b];
      if (a >= 0) {
        q4 qa = ld4(w->Q + 4 * a);
        st3(w->P2 + 3 * b, add3(ld3(w->P + 3 * a), c_rotate(ld3(w->P + 3 * b), qa)));
        st4(w->Q2 + 4 * b, c_qmul(qa, ld4(w->Q + 4 * b)));
      } else {
        memcpy(w->P2 + 3 * b, w->P + 3 * b, 3 * sizeof(REAL)); memcpy(w->Q2 + 4 * b, w->Q + 4 * b, 4 * sizeof(REAL));
      }
    }
    for (int i = 0; i < nset; i++) {
      int b = set[i];
      memcpy(w->P + 3 * b, w->P2 + 3 * b, 3 * sizeof(REAL)); memcpy(w->Q + 4 * b, w->Q2 + 4 * b, 4 * sizeof(REAL));
    }
  }
}

static void sites_eval(const omodel *m, int mode, const REAL *site_pos, owork *w) {
  for (int k = 0; k < m->nsite; k++) {
    int b = m->site_body[k];
    v3 off = ld3(site_pos + 3 * k); q4 qb = ld4(w->Q + 4 * b);
    v3 r = mode ? c_rotate(off, qb) : m_rotate(off, qb);
    st3(w->S + 3 * k, add3(ld3(w->P + 3 * b), r));
  }
}

/* 32-slot butterfly: slot l combines with slot l^off for off = 16,8,4,2,1 */
static REAL butterfly32(REAL *p) {
  REAL t[LANES];
  for (int off = 16; off >= 1; off >>= 1) {
    for (int l = 0; l < LANES; l++) t[l] = p[l] + p[l ^ off];
    memcpy(p, t, sizeof(t));
  }
  return p[0];
}

/* ------------------------------------------------------------------ */
/* loss and gradient                                                  */
/* ------------------------------------------------------------------ */

/* make_qs (utils.py:129-144): (1 - mask) * q0 + mask * q  ==  mask ? q : q0 for finite values */
static void make_qs(int nq, const REAL *q0, const uint8_t *mask, const REAL *q, REAL *out) {
  for (int i = 0; i < nq; i++) out[i] = mask[i] ? q[i] : q0[i];
}

/* q_loss (stac_core.py:27-63).  If grad != NULL also d loss / d q (masked by qmask). */
static REAL loss_eval(const omodel *m, const osched *s, int mode, owork *w, const REAL *q, const REAL *q0, const uint8_t *qmask,
                      const REAL *kp, const REAL *kpmask /*3K of 0/1*/, const REAL *site_pos, REAL *grad) {
  int K = m->nsite, nq = m->nq;
  make_qs(nq, q0, qmask, q, w->qfull);
  if (mode) fk_canon(m, s, s->act, s->nact, s->rounds_act, w->qfull, w); else fk_mjx(m, s->act, s->nact, w->qfull, w);
  sites_eval(m, mode, site_pos, w);
  for (int c = 0; c < 3 * K; c++) w->res[c] = (kp[c] - w->S[c]) * kpmask[c];
  REAL loss;
  if (!mode) {
    loss = R(0);
    for (int c = 0; c < 3 * K; c++) loss = loss + w->res[c] * w->res[c];
  } else {
    REAL part[LANES];
    for (int l = 0; l < LANES; l++) {
      REAL acc = R(0);
      for (int i = 0; i < s->spl; i++) {
        int pos = l * s->spl + i;
        if (pos < K) {
          const REAL *r = w->res + 3 * s->site_order[pos];
          REAL e = r_fma(r[2], r[2], r_fma(r[1], r[1], r[0] * r[0]));
          acc = (i == 0) ? e : acc + e;
        }
      }
      part[l] = acc;
    }
    loss = butterfly32(part);
  }
  if (!grad) return loss;

  for (int i = 0; i < nq; i++) grad[i] = R(0);
  if (!mode) {
    /* direct Jacobian transpose with MJX's world xanchor / xaxis */
    for (int ia = 0; ia < s->nact; ia++) {
      int b = s->act[ia];
      for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
        int j = m->body_jntadr[b] + jj, adr = m->jnt_qposadr[j], t = m->jnt_type[j];
        v3 F = { 0, 0, 0 }, T = { 0, 0, 0 };
        v3 ref = (t == JNT_FREE) ? ld3(w->P + 3 * b) : ld3(w->anchor + 3 * j);
        for (int k = 0; k < K; k++) {
          int sb = m->site_body[k];
          if (sb < b || sb >= b + s->subsize[b]) continue;
          v3 f = { R(-2) * (kpmask[3 * k] * w->res[3 * k]), R(-2) * (kpmask[3 * k + 1] * w->res[3 * k + 1]), R(-2) * (kpmask[3 * k + 2] * w->res[3 * k + 2]) };
          v3 d = sub3(ld3(w->S + 3 * k), ref);
          F = add3(F, f);
          v3 c = { d.y * f.z - d.z * f.y, d.z * f.x - d.x * f.z, d.x * f.y - d.y * f.x };
          T = add3(T, c);
        }
        if (t == JNT_HINGE) { v3 a = ld3(w->axis + 3 * j); grad[adr] = a.x * T.x + a.y * T.y + a.z * T.z; }
        else if (t == JNT_SLIDE) { v3 a = ld3(w->axis + 3 * j); grad[adr] = a.x * F.x + a.y * F.y + a.z * F.z; }
        else if (t == JNT_FREE) {
          grad[adr] = F.x; grad[adr + 1] = F.y; grad[adr + 2] = F.z;
          q4 qh = ld4(w->Q + 4 * b); q4 tq = { 0, T.x, T.y, T.z };
          q4 h = m_qmul(tq, qh); h.w *= R(2); h.x *= R(2); h.y *= R(2); h.z *= R(2);
          REAL pr = qh.w * h.w + qh.x * h.x + qh.y * h.y + qh.z * h.z, n = w->fnorm[j];
          grad[adr + 3] = (h.w - qh.w * pr) / n; grad[adr + 4] = (h.x - qh.x * pr) / n;
          grad[adr + 5] = (h.y - qh.y * pr) / n; grad[adr + 6] = (h.z - qh.z * pr) / n;
        } else { /* ball: right-multiplied local rotation; supported when it is the body's last joint */
          q4 qb = ld4(w->Q + 4 * b); q4 qc = { qb.w, -qb.x, -qb.y, -qb.z };
          v3 tl = m_rotate(T, qc);
          q4 ql = ld4(w->qn + adr); q4 tq = { 0, tl.x, tl.y, tl.z };
          q4 h = m_qmul(ql, tq); h.w *= R(2); h.x *= R(2); h.y *= R(2); h.z *= R(2);
          REAL pr = ql.w * h.w + ql.x * h.x + ql.y * h.y + ql.z * h.z, n = w->fnorm[j];
          grad[adr] = (h.w - ql.w * pr) / n; grad[adr + 1] = (h.x - ql.x * pr) / n;
          grad[adr + 2] = (h.y - ql.y * pr) / n; grad[adr + 3] = (h.z - ql.z * pr) / n;
        }
      }
    }
  } else {
    /* prefix-scan wrench about c = world position of the first active body */
    v3 cref = ld3(w->P + 3 * s->cbody);
    int spl = s->spl;
    REAL *I = w->W; /* inclusive prefix, [K][6] */
    REAL tot[LANES][6];
    for (int l = 0; l < LANES; l++) {
      REAL acc[6] = { 0, 0, 0, 0, 0, 0 };
      for (int i = 0; i < spl; i++) {
        int pos = l * spl + i;
        if (pos >= K) break;
        int k = s->site_order[pos];
        v3 f = { R(-2) * (kpmask[3 * k] * w->res[3 * k]), R(-2) * (kpmask[3 * k + 1] * w->res[3 * k + 1]), R(-2) * (kpmask[3 * k + 2] * w->res[3 * k + 2]) };
        v3 d = sub3(ld3(w->S + 3 * k), cref);
        v3 tq = c_cross(d, f);
        REAL v[6] = { f.x, f.y, f.z, tq.x, tq.y, tq.z };
        for (int e = 0; e < 6; e++) { acc[e] = (i == 0) ? v[e] : acc[e] + v[e]; I[6 * pos + e] = acc[e]; }
      }
      for (int e = 0; e < 6; e++) tot[l][e] = acc[e];
    }
    /* Hillis-Steele inclusive scan over lane totals */
    for (int off = 1; off < LANES; off <<= 1) {
      REAL t2[LANES][6];
      for (int l = 0; l < LANES; l++) for (int e = 0; e < 6; e++) t2[l][e] = (l >= off) ? tot[l][e] + tot[l - off][e] : tot[l][e];
      memcpy(tot, t2, sizeof(tot));
    }
    for (int l = 1; l < LANES; l++)
      for (int i = 0; i < spl; i++) { int pos = l * spl + i; if (pos >= K) break; for (int e = 0; e < 6; e++) I[6 * pos + e] = tot[l - 1][e] + I[6 * pos + e]; }
    for (int ia = 0; ia < s->nact; ia++) {
      int b = s->act[ia], p = m->body_parent[b];
      for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
        int j = m->body_jntadr[b] + jj, adr = m->jnt_qposadr[j], t = m->jnt_type[j];
        int a = s->jnt_s[j], e = s->jnt_e[j];
        if (e <= a) continue;
        REAL wr[6];
        for (int c = 0; c < 6; c++) wr[c] = (a > 0) ? I[6 * (e - 1) + c] - I[6 * (a - 1) + c] : I[6 * (e - 1) + c];
        v3 F = { wr[0], wr[1], wr[2] }, T = { wr[3], wr[4], wr[5] };
        if (t == JNT_FREE) {
          v3 d = sub3(ld3(w->P + 3 * b), cref); v3 cr = c_cross(d, F); v3 Tp = sub3(T, cr);
          grad[adr] = F.x; grad[adr + 1] = F.y; grad[adr + 2] = F.z;
          q4 qh = ld4(w->Q + 4 * b); q4 tq = { 0, Tp.x, Tp.y, Tp.z };
          q4 h = c_qmul(tq, qh); h.w *= R(2); h.x *= R(2); h.y *= R(2); h.z *= R(2);
          REAL pr = r_fma(qh.z, h.z, r_fma(qh.y, h.y, r_fma(qh.x, h.x, qh.w * h.w))), n = w->fnorm[j];
          grad[adr + 3] = r_fma(-qh.w, pr, h.w) * n; grad[adr + 4] = r_fma(-qh.x, pr, h.x) * n;
          grad[adr + 5] = r_fma(-qh.y, pr, h.y) * n; grad[adr + 6] = r_fma(-qh.z, pr, h.z) * n;
          continue;
        }
        /* hinge / slide: bring the subtree wrench into the PARENT frame once (the parent-frame anchor / axis of
           every joint of the body are then used directly); world body 0 is the identity pose */
        q4 pq = ld4(w->Q + 4 * p); v3 pp = ld3(w->P + 3 * p);
        q4 pc = { pq.w, -pq.x, -pq.y, -pq.z };
        v3 T0 = sub3(T, c_cross(sub3(pp, cref), F));
        v3 Fp = c_rotate(F, pc), Tp = c_rotate(T0, pc);
        v3 al = ld3(w->anchor + 3 * j), xl = ld3(w->axis + 3 * j);
        if (t == JNT_SLIDE) { grad[adr] = c_dot3(xl, Fp); continue; }
        if (t == JNT_HINGE) { grad[adr] = c_dot3(xl, sub3(Tp, c_cross(al, Fp))); continue; }
        /* ball: world-frame anchor, torque about it */
        v3 A = add3(pp, c_rotate(al, pq));
        v3 d = sub3(A, cref); v3 cr = c_cross(d, F); v3 Ta = sub3(T, cr);
        /* ball */
        q4 qb = ld4(w->Q + 4 * b); q4 qc = { qb.w, -qb.x, -qb.y, -qb.z };
        v3 tl = c_rotate(Ta, qc);
        q4 ql = ld4(w->qn + adr); q4 tq = { 0, tl.x, tl.y, tl.z };
        q4 h = c_qmul(ql, tq); h.w *= R(2); h.x *= R(2); h.y *= R(2); h.z *= R(2);
        REAL pr = r_fma(ql.z, h.z, r_fma(ql.y, h.y, r_fma(ql.x, h.x, ql.w * h.w))), n = w->fnorm[j];
        grad[adr] = r_fma(-ql.w, pr, h.w) * n; grad[adr + 1] = r_fma(-ql.x, pr, h.x) * n;
        grad[adr + 2] = r_fma(-ql.y, pr, h.y) * n; grad[adr + 3] = r_fma(-ql.z, pr, h.z) * n;
      }
    }
  }
  for (int i = 0; i < nq; i++) if (!qmask[i]) grad[i] = R(0);
  return loss;
}

/* ------------------------------------------------------------------ */
/* jaxopt 0.8.5 ProjectedGradient(fun=q_loss, projection=projection_box,   */
/* maxiter, tol) with defaults stepsize=0 (backtracking), maxls=15,        */
/* acceleration=True, decrease_factor=0.5                                   */
/* ------------------------------------------------------------------ */

static REAL vec_dot(int n, const REAL *a, const REAL *b, int mode) {
  if (!mode) { REAL s = R(0); for (int i = 0; i < n; i++) s = s + a[i] * b[i]; return s; }
  REAL part[LANES];
  for (int l = 0; l < LANES; l++) {
    REAL acc = R(0); int first = 1;
    for (int i = l; i < n; i += LANES) { acc = first ? a[i] * b[i] : r_fma(a[i], b[i], acc); first = 0; }
    part[l] = acc;
  }
  return butterfly32(part);
}

static inline REAL clipr(REAL x, REAL lo, REAL hi) { return r_min(r_max(x, lo), hi); }

typedef struct { REAL error; int iters; int ls_evals; } solve_info;

/* diagnostic: histogram of the accepted line-search candidate index (rodent session: 17.5 % / 66.6 % / 15.6 % at 0 / 1 / 2) */
static long long g_ls_hist[17];
void oracle_ls_hist(long long *out, int reset) { for (int i = 0; i < 17; i++) { out[i] = g_ls_hist[i]; if (reset) g_ls_hist[i] = 0; } }
/* diagnostic: the accepted candidate index of every iteration, in order (single-threaded runs only); -1 marks the start of a solve */
#define LS_TRACE_MAX (1 << 22)
static signed char *g_ls_trace; static long long g_ls_trace_n;
static void ls_trace_push(int v) {
  if (!g_ls_trace) return;
  if (g_ls_trace_n < LS_TRACE_MAX) g_ls_trace[g_ls_trace_n++] = (signed char)v;
}
long long oracle_ls_trace(signed char *out, long long cap, int enable) {
  long long n = g_ls_trace_n < cap ? g_ls_trace_n : cap;
  if (out && g_ls_trace) memcpy(out, g_ls_trace, (size_t)n);
  if (enable && !g_ls_trace) g_ls_trace = (signed char *)malloc(LS_TRACE_MAX);
  if (!enable && g_ls_trace) { free(g_ls_trace); g_ls_trace = NULL; }
  g_ls_trace_n = 0;
  return n;
}

static solve_info q_opt(const omodel *m, const osched *s, int mode, owork *w, const REAL *q0, const REAL *lb, const REAL *ub,
                        const uint8_t *qmask, const REAL *kp, const REAL *kpmask, const REAL *site_pos,
                        REAL tol, int maxiter, int maxls, REAL *params) {
  int nq = m->nq;
  REAL *buf = (REAL *)calloc((size_t)nq * 6, sizeof(REAL));
  REAL *x = buf, *y = buf + nq, *g = buf + 2 * nq, *xn = buf + 3 * nq, *d = buf + 4 * nq, *gn = buf + 5 * nq;
  memcpy(x, q0, sizeof(REAL) * nq); memcpy(y, q0, sizeof(REAL) * nq);
  REAL t = R(1), step = R(1), err = (REAL)INFINITY;
  solve_info info = { err, 0, 0 };
  if (maxiter <= 0) { memcpy(params, x, sizeof(REAL) * nq); free(buf); return info; }
  do {
    REAL fy = loss_eval(m, s, mode, w, y, q0, qmask, kp, kpmask, site_pos, g);
    REAL st = step;
    int halvings = 0;
    for (;;) {
      for (int i = 0; i < nq; i++) xn[i] = clipr(mode ? r_fma(-st, g[i], y[i]) : y[i] + (-st) * g[i], lb[i], ub[i]);
      REAL fn = loss_eval(m, s, mode, w, xn, q0, qmask, kp, kpmask, site_pos, NULL);
      info.ls_evals++;
      for (int i = 0; i < nq; i++) d[i] = xn[i] - y[i];
      REAL sq = vec_dot(nq, d, d, mode), dg = vec_dot(nq, d, g, mode);
      REAL dec = st * (fn - fy);
      REAL cond = mode ? r_fma(st, dg, R(0.5) * sq) : st * dg + R(0.5) * sq;
      if (!(dec > cond + R_EPS) || halvings >= maxls) break;
      st = st * R(0.5); halvings++;
    }
#ifdef _OPENMP
#pragma omp atomic
#endif
    g_ls_hist[halvings > 16 ? 16 : halvings]++;
    if (info.iters == 0) ls_trace_push(-1);
    ls_trace_push(halvings);
    step = (st <= R(1e-6)) ? R(1) : st / R(0.5);
    REAL tn = R(0.5) * (R(1) + r_sqrt(mode ? r_fma(R(4) * t, t, R(1)) : R(1) + R(4) * t * t));
    REAL beta = (t - R(1)) / tn;
    for (int i = 0; i < nq; i++) { REAL df = xn[i] - x[i]; y[i] = mode ? r_fma(beta, df, xn[i]) : xn[i] + beta * df; }
    loss_eval(m, s, mode, w, xn, q0, qmask, kp, kpmask, site_pos, gn);
    for (int i = 0; i < nq; i++) d[i] = clipr(xn[i] - gn[i], lb[i], ub[i]) - xn[i];
    err = r_sqrt(vec_dot(nq, d, d, mode));
    memcpy(x, xn, sizeof(REAL) * nq);
    t = tn; info.iters++;
  } while (err > tol && info.iters < maxiter);
  info.error = err;
  memcpy(params, x, sizeof(REAL) * nq);
  free(buf);
  return info;
}

#include "fast_order.h"

/* mode 2 = what libstacb computes by default: fast order where the register-resident solver serves the model, canonical
 * order (mode 1) everywhere else (other models; FK outputs and m-phase statistics of every model) */
static solve_info q_opt_any(const omodel *m, const osched *s, const ofast *F, int mode, owork *w, const REAL *q0, const REAL *lb, const REAL *ub,
                            const uint8_t *qmask, const REAL *kp, const REAL *kpmask, const REAL *site_pos, REAL tol, int maxiter, int maxls,
                            REAL *params) {
  if (mode == 2 && F) return fast_q_opt(m, F, q0, lb, ub, qmask, kp, kpmask, site_pos, tol, maxiter, maxls, params);
  return q_opt(m, s, mode, w, q0, lb, ub, qmask, kp, kpmask, site_pos, tol, maxiter, maxls, params);
}

/* 1 when mode 2 differs from mode 1 for this model (the register-resident solver serves it) */
int oracle_fast_path(const omodel *m) {
  osched *s = sched_create(m);
  ofast *F = fast_create(m, s);
  int r = F != NULL;
  fast_destroy(F); sched_destroy(s);
  return r;
}

/* ------------------------------------------------------------------ */
/* exported entry points (ctypes)                                     */
/* ------------------------------------------------------------------ */

/* utils.kinematics on all bodies + site_xpos for the keypoint sites. */
int oracle_fk(const omodel *m, int mode, const REAL *qpos, const REAL *site_pos, REAL *qpos_out, REAL *xpos, REAL *xquat, REAL *site_xpos) {
  osched *s = sched_create(m); owork *w = work_create(m);
  if (mode) fk_canon(m, s, s->full, s->nfull, s->rounds_full, qpos, w); else fk_mjx(m, s->full, s->nfull, qpos, w);
  sites_eval(m, mode, site_pos, w);
  if (qpos_out) memcpy(qpos_out, w->qn, sizeof(REAL) * m->nq);
  if (xpos) memcpy(xpos, w->P, sizeof(REAL) * 3 * m->nbody);
  if (xquat) memcpy(xquat, w->Q, sizeof(REAL) * 4 * m->nbody);
  if (site_xpos) memcpy(site_xpos, w->S, sizeof(REAL) * 3 * m->nsite);
  work_destroy(w); sched_destroy(s);
  return 0;
}

int oracle_loss_grad(const omodel *m, int mode, const REAL *q, const REAL *q0, const uint8_t *qmask, const REAL *kp, const uint8_t *kpmask,
                     const REAL *site_pos, REAL *loss, REAL *grad) {
  osched *s = sched_create(m); owork *w = work_create(m);
  REAL *km = (REAL *)calloc(3 * m->nsite + 1, sizeof(REAL));
  for (int c = 0; c < 3 * m->nsite; c++) km[c] = kpmask[c] ? R(1) : R(0);
  ofast *F = mode == 2 ? fast_create(m, s) : NULL;
  if (F) *loss = fast_loss_eval(m, F, q, q0, qmask, kp, km, site_pos, grad);
  else *loss = loss_eval(m, s, mode, w, q, q0, qmask, kp, km, site_pos, grad);
  fast_destroy(F);
  free(km); work_destroy(w); sched_destroy(s);
  return 0;
}

int oracle_q_opt(const omodel *m, int mode, const REAL *q0, const REAL *lb, const REAL *ub, const uint8_t *qmask, const REAL *kp,
                 const uint8_t *kpmask, const REAL *site_pos, REAL tol, int maxiter, int maxls, REAL *params, REAL *error,
                 int32_t *iters, int32_t *ls_evals) {
  osched *s = sched_create(m); owork *w = work_create(m);
  REAL *km = (REAL *)calloc(3 * m->nsite + 1, sizeof(REAL));
  for (int c = 0; c < 3 * m->nsite; c++) km[c] = kpmask[c] ? R(1) : R(0);
  ofast *F = mode == 2 ? fast_create(m, s) : NULL;
  solve_info si = q_opt_any(m, s, F, mode, w, q0, lb, ub, qmask, kp, km, site_pos, tol, maxiter, maxls, params);
  *error = si.error; *iters = si.iters; *ls_evals = si.ls_evals;
  fast_destroy(F);
  free(km); work_destroy(w); sched_destroy(s);
  return 0;
}

/* replace_qs (utils.py:147-169): set qpos, run kinematics (which normalises quaternions in qpos). */
static void replace_qs(const omodel *m, const osched *s, int mode, owork *w, const REAL *q, REAL *qpos) {
  if (mode) fk_canon(m, s, s->full, s->nfull, s->rounds_full, q, w); else fk_mjx(m, s->full, s->nfull, q, w);
  memcpy(qpos, w->qn, sizeof(REAL) * m->nq);
}

/*
 * One clip: optional root_optimization on frame 0 (compute_stac.py:17-104) followed by
 * pose_optimization over all frames (compute_stac.py:170-278).
 *   qpos_io : in  = mjx_data.qpos on entry (qpos0 for ik_only; last pose of the previous pass in fit_offsets)
 *             out = mjx_data.qpos after the last frame
 *   root_dims : 7 (free root) or 4 (slide root), compute_stac.py:51-54
 *   iters/ls  : [F][1+P] per solve; root_stats[4] = iters, ls of the two root solves
 */
int oracle_pose_clip(const omodel *m, int mode, const REAL *kp, int F, REAL *qpos_io, const REAL *site_pos, const REAL *lb, const REAL *ub,
                     const uint8_t *part_masks, int P, int do_root, int root_kp_idx, const uint8_t *trunk_kps, int root_dims,
                     REAL tol, int maxiter, int maxls,
                     REAL *qpos_out, REAL *xpos_out, REAL *xquat_out, REAL *sites_out, REAL *err_out, int32_t *iters_out, int32_t *ls_out,
                     int32_t *root_stats) {
  int nq = m->nq, K = m->nsite, nb = m->nbody;
  osched *s = sched_create(m); owork *w = work_create(m);
  ofast *Fz = mode == 2 ? fast_create(m, s) : NULL;
  REAL *ones = (REAL *)calloc(3 * K + 1, sizeof(REAL)), *trunk = (REAL *)calloc(3 * K + 1, sizeof(REAL));
  uint8_t *allq = (uint8_t *)calloc(nq + 1, 1), *rootq = (uint8_t *)calloc(nq + 1, 1);
  REAL *q0 = (REAL *)calloc(nq, sizeof(REAL)), *par = (REAL *)calloc(nq, sizeof(REAL)), *qm = (REAL *)calloc(nq, sizeof(REAL));
  for (int c = 0; c < 3 * K; c++) { ones[c] = R(1); trunk[c] = (trunk_kps && trunk_kps[c / 3]) ? R(1) : R(0); }
  for (int i = 0; i < nq; i++) { allq[i] = 1; rootq[i] = (i < root_dims); }
  REAL *qpos = qpos_io;
  if (do_root) {
    for (int rep = 0; rep < 2; rep++) {
      memcpy(q0, qpos, sizeof(REAL) * nq);
      for (int c = 0; c < 3; c++) q0[c] = kp[3 * root_kp_idx + c];
      solve_info si = q_opt_any(m, s, Fz, mode, w, q0, lb, ub, rootq, kp, trunk, site_pos, tol, maxiter, maxls, par);
      make_qs(nq, q0, rootq, par, qm);
      replace_qs(m, s, mode, w, qm, qpos);
      if (root_stats) { root_stats[2 * rep] = si.iters; root_stats[2 * rep + 1] = si.ls_evals; }
    }
  }
  for (int f = 0; f < F; f++) {
    const REAL *kpf = kp + (size_t)f * 3 * K;
    solve_info si;
    memcpy(q0, qpos, sizeof(REAL) * nq);
    si = q_opt_any(m, s, Fz, mode, w, q0, lb, ub, allq, kpf, ones, site_pos, tol, maxiter, maxls, par);
    replace_qs(m, s, mode, w, par, qpos);
    if (iters_out) { iters_out[(size_t)f * (1 + P)] = si.iters; ls_out[(size_t)f * (1 + P)] = si.ls_evals; }
    for (int p = 0; p < P; p++) {
      const uint8_t *pm = part_masks + (size_t)p * nq;
      memcpy(q0, qpos, sizeof(REAL) * nq);
      si = q_opt_any(m, s, Fz, mode, w, q0, lb, ub, pm, kpf, ones, site_pos, tol, maxiter, maxls, par);
      make_qs(nq, q0, pm, par, qm);
      replace_qs(m, s, mode, w, qm, qpos);
      if (iters_out) { iters_out[(size_t)f * (1 + P) + 1 + p] = si.iters; ls_out[(size_t)f * (1 + P) + 1 + p] = si.ls_evals; }
    }
    /* w holds the full-body FK of the frame's final qpos */
    sites_eval(m, mode, site_pos, w);
    if (qpos_out) memcpy(qpos_out + (size_t)f * nq, qpos, sizeof(REAL) * nq);
    if (xpos_out) memcpy(xpos_out + (size_t)f * nb * 3, w->P, sizeof(REAL) * nb * 3);
    if (xquat_out) memcpy(xquat_out + (size_t)f * nb * 4, w->Q, sizeof(REAL) * nb * 4);
    if (sites_out) memcpy(sites_out + (size_t)f * K * 3, w->S, sizeof(REAL) * K * 3);
    if (err_out) err_out[f] = si.error;
  }
  free(ones); free(trunk); free(allq); free(rootq); free(q0); free(par); free(qm);
  fast_destroy(Fz);
  work_destroy(w); sched_destroy(s);
  return 0;
}

/* Several independent clips (the reference's vmap over clips, stac.py:405-440), OpenMP over clips. */
int oracle_pose_clips(const omodel *m, int mode, const REAL *kp, int C, int F, const REAL *qpos_init, const REAL *site_pos, const REAL *lb,
                      const REAL *ub, const uint8_t *part_masks, int P, int do_root, int root_kp_idx, const uint8_t *trunk_kps, int root_dims,
                      REAL tol, int maxiter, int maxls, REAL *qpos_out, REAL *xpos_out, REAL *xquat_out, REAL *sites_out, REAL *err_out,
                      int32_t *iters_out, int32_t *ls_out, int32_t *root_stats, int nthreads) {
  int nq = m->nq, K = m->nsite, nb = m->nbody;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int c = 0; c < C; c++) {
    REAL *qp = (REAL *)malloc(sizeof(REAL) * nq);
    memcpy(qp, qpos_init + (size_t)c * nq, sizeof(REAL) * nq);
    oracle_pose_clip(m, mode, kp + (size_t)c * F * 3 * K, F, qp, site_pos, lb, ub, part_masks, P, do_root, root_kp_idx, trunk_kps, root_dims,
                     tol, maxiter, maxls, qpos_out ? qpos_out + (size_t)c * F * nq : NULL, xpos_out ? xpos_out + (size_t)c * F * nb * 3 : NULL,
                     xquat_out ? xquat_out + (size_t)c * F * nb * 4 : NULL, sites_out ? sites_out + (size_t)c * F * K * 3 : NULL,
                     err_out ? err_out + (size_t)c * F : NULL, iters_out ? iters_out + (size_t)c * F * (1 + P) : NULL,
                     ls_out ? ls_out + (size_t)c * F * (1 + P) : NULL, root_stats ? root_stats + 4 * c : NULL);
    free(qp);
  }
  return 0;
}

/* mode >= 1 sums the frames in the order of the CUDA kernel: chunks of 8 consecutive frames in frame order, then the chunk
 * partials in chunk order (csrc/stacb_device.cuh: m_phase_kernel).
 * m != NULL: the data term of the objective at the offsets m instead,  z2_out = sum_t sum_k |y - p - R m|^2  (s_out unused). */
#define MCH 8
static int m_phase(const omodel *m, int mode, const REAL *kp, const REAL *q, const REAL *moff, int T, REAL *s_out, REAL *z2_out) {
  int K = m->nsite, nq = m->nq;
  osched *s = sched_create(m); owork *w = work_create(m);
  REAL *tot = (REAL *)calloc(3 * K + 1, sizeof(REAL)), *part = (REAL *)calloc(3 * K + 1, sizeof(REAL));
  for (int t = 0; t < T; t++) {
    const REAL *qt = q + (size_t)t * nq, *yt = kp + (size_t)t * 3 * K;
    int first = mode ? (t % MCH == 0) : (t == 0);
    if (mode) fk_canon(m, s, s->act, s->nact, s->rounds_act, qt, w); else fk_mjx(m, s->act, s->nact, qt, w);
    REAL zz[LANES]; for (int l = 0; l < LANES; l++) zz[l] = R(0);
    REAL zf = R(0);
    for (int pos = 0; pos < K; pos++) {
      int k = mode ? s->site_order[pos] : pos, b = m->site_body[k];
      q4 qb = ld4(w->Q + 4 * b); v3 p = ld3(w->P + 3 * b);
      v3 z = { yt[3 * k] - p.x, yt[3 * k + 1] - p.y, yt[3 * k + 2] - p.z };
      if (moff) {
        v3 mk = ld3(moff + 3 * k);
        v3 r = sub3(z, mode ? c_rotate(mk, qb) : m_rotate(mk, qb));
        z = r;
      } else {
        /* math.quat_to_mat */
        REAL ww = qb.w * qb.w, xx = qb.x * qb.x, yy = qb.y * qb.y, zq = qb.z * qb.z;
        REAL xy = qb.x * qb.y, xz = qb.x * qb.z, yz = qb.y * qb.z, wx = qb.w * qb.x, wy = qb.w * qb.y, wz = qb.w * qb.z;
        REAL M[3][3] = { { ww + xx - yy - zq, R(2) * (xy - wz), R(2) * (xz + wy) },
                         { R(2) * (xy + wz), ww - xx + yy - zq, R(2) * (yz - wx) },
                         { R(2) * (xz - wy), R(2) * (yz + wx), ww - xx - yy + zq } };
        REAL zv[3] = { z.x, z.y, z.z };
        for (int i = 0; i < 3; i++) {
          REAL c = mode ? r_fma(M[2][i], zv[2], r_fma(M[1][i], zv[1], M[0][i] * zv[0])) : M[0][i] * zv[0] + M[1][i] * zv[1] + M[2][i] * zv[2];
          part[3 * k + i] = first ? c : part[3 * k + i] + c;
        }
      }
      if (mode) {
        REAL e = r_fma(z.z, z.z, r_fma(z.y, z.y, z.x * z.x));
        int l = pos / s->spl, i = pos % s->spl;
        zz[l] = (i == 0) ? e : zz[l] + e;
      } else part[3 * K] = part[3 * K] + (z.x * z.x + z.y * z.y + z.z * z.z);  /* mode 0: one running sum over sites and frames */
    }
    if (mode) { zf = butterfly32(zz); part[3 * K] = first ? zf : part[3 * K] + zf; }
    if (mode && (t % MCH == MCH - 1 || t == T - 1)) {  /* chunk complete */
      int c0 = (t / MCH) == 0;
      for (int j = 0; j <= 3 * K; j++) tot[j] = c0 ? part[j] : tot[j] + part[j];
    }
  }
  if (!mode) memcpy(tot, part, sizeof(REAL) * (3 * K + 1));
  if (s_out && !moff) memcpy(s_out, tot, sizeof(REAL) * 3 * K);
  *z2_out = tot[3 * K];
  free(tot); free(part);
  work_destroy(w); sched_destroy(s);
  return 0;
}

/*
 * _m_opt sufficient statistics (stac_core.py:146-159): for T frames,
 *   s[k,i] = sum_t sum_j R_tk[j,i] * (y_tk[j] - p_tk[j]),  z2 = sum_t sum_k |y_tk - p_tk|^2
 * with p/R the world position / rotation matrix of each keypoint site's body.
 */
int oracle_m_stats(const omodel *m, int mode, const REAL *kp, const REAL *q, int T, REAL *s_out, REAL *z2_out) {
  return m_phase(m, mode, kp, q, NULL, T, s_out, z2_out);
}

/* data term of the m-phase objective at the offsets moff [K,3], from the residuals */
int oracle_m_residual(const omodel *m, int mode, const REAL *kp, const REAL *q, const REAL *moff, int T, REAL *out) {
  return m_phase(m, mode, kp, q, moff, T, NULL, out);
}

int oracle_real_size(void) { return (int)sizeof(REAL); }
