/*
 * fast_order.h -- mode 2 ("fast order") of the oracle: the arithmetic of the register-resident hinge-tree solver
 * (stac_mjx_b200/csrc/stacb_fast.cuh), restated lane by lane in plain C.  Included by stac_oracle.c.
 *
 * TEST INFRASTRUCTURE ONLY (see stac_oracle.c).  Same mathematics as modes 0 / 1 (reference stac_mjx/stac_core.py:27-99
 * over MJX kinematics and jaxopt ProjectedGradient); only the float32 operation order differs:
 *   - 32 "lanes": lane e < nact is active body e (ascending body id), lane 31 and every lane >= nact is an identity element;
 *     lane p < K is the marker site at sorted position p; solver slot j < JM of lane e is hinge j of body e, slot JM of
 *     lanes 0..6 the free joint; every other qpos address is passive (zero gradient);
 *   - first hinge of a body folded with the body's constant pose (quat = Qc cos h + Qs sin h, pos = A + B cos t + C sin t),
 *     further hinges composed in the parent frame, world quaternions by pointer jumping, ONE rotation of each local offset by
 *     the parent's world quaternion, world positions by pointer jumping with plain additions;
 *   - rotate(v, q) = v + 2 (s t + u x t), t = u x v;
 *   - one site per lane (site p on lane p + 1, so the inclusive wrench scan doubles as the exclusive one): loss by the 32-lane
 *     butterfly, wrench prefix by Hillis-Steele over lanes;
 *   - sin / cos of the hinge half-angles up to a common sign (reduction by pi, no quadrant logic);
 *   - solver reductions: per-lane fma chain over the slots, then the butterfly; the passive coordinates enter the first
 *     line search through their own lane-dealt sum.
 * (JM, RT) is the kernel variant the dispatcher picks (first fit of (1,5,1), (3,4,3)): surplus slots / rounds are no-ops
 * but are executed here as well so that even signed zeros agree.
 */

#define FJ 3 /* slots a variant can have */
#define FNS (FJ + 1)
#define FMAXL 256 /* lanes of the widest variant: 8 warps */

typedef struct {
  int JM, RT, n, K, has_free, free_e, free_adr, folded;
  int W, NL, idl;                        /* warps per evaluation, lanes = 32 W, identity lane = NL - 1 */
  v3 srel_p[FMAXL]; q4 srel_q[FMAXL];    /* folded models: pose of the site's own body in its element's frame */
  q4 Qc[FMAXL], Qs[FMAXL];
  v3 A[FMAXL], B[FMAXL], C[FMAXL], anc0[FMAXL], ax0[FMAXL];
  v3 jax[FMAXL][FJ], jpp[FMAXL][FJ], jcx[FMAXL][FJ], jps[FMAXL][FJ];
  REAL ref[FMAXL][FJ];
  int hinge[FMAXL][FJ], pfree[FMAXL];
  int src[FMAXL][8], par[FMAXL], sa[FMAXL], se[FMAXL];
  int sk[FMAXL], seb[FMAXL];             /* sites: keypoint index (-1 none), lane of the body */
  int valid[FMAXL][FNS], adr[FMAXL][FNS]; /* solver slots */
  int npassive, *passive;
} ofast;

static inline v3 f_rotq(v3 v, q4 q) {
  v3 u = { q.x, q.y, q.z };
  v3 t = c_cross(u, v);
  v3 c = c_cross(u, t);
  v3 w = { r_fma(q.w, t.x, c.x), r_fma(q.w, t.y, c.y), r_fma(q.w, t.z, c.z) };
  v3 r = { r_fma(R(2), w.x, v.x), r_fma(R(2), w.y, v.y), r_fma(R(2), w.z, v.z) };
  return r;
}

/* sum over the NL = 32 W lanes of an evaluation: the 32-lane butterfly inside every warp, then the warp sums added in warp order
 * (W = 1: the butterfly alone) */
static REAL f_cta_sum(int W, REAL *part) {
  REAL tot = R(0);
  for (int w = 0; w < W; w++) { REAL b = butterfly32(part + LANES * w); tot = (w == 0) ? b : tot + b; }
  return tot;
}

/* (-1)^j (sin x, cos x), j = rint(x / pi): mirrors sincos_pi of csrc/stacb_math.cuh (the common sign cancels in FK) */
static inline void f_sincos(REAL x, REAL *sp, REAL *cp) {
  if (!IS_F32) { *sp = (REAL)sin(x); *cp = (REAL)cos(x); return; }
  float xf = (float)x;
  float j = rintf(xf * 0.318309873f);
  float r = fmaf(-j, 3.14159274e+00f, xf);
  r = fmaf(-j, -8.74227766e-08f, r);
  r = fmaf(-j, -3.43024902e-15f, r);
  float r2 = r * r;
  float ps = fmaf(r2, 2.6056311526190257e-06f, -0.0001980953966267407f);
  ps = fmaf(ps, r2, 0.008333065547049046f);
  ps = fmaf(ps, r2, -0.16666659712791443f);
  *sp = (REAL)fmaf(r * r2, ps, r);
  float pc = fmaf(r2, -2.619248391511064e-07f, 2.4769240553723648e-05f);
  pc = fmaf(pc, r2, -0.0013888567918911576f);
  pc = fmaf(pc, r2, 0.041666656732559204f);
  *cp = (REAL)fmaf(r2 * r2, pc, fmaf(-0.5f, r2, 1.0f));
}

/* normalize4_nr of csrc/stacb_math.cuh: integer-seeded reciprocal square root, two Newton steps and one in residual form */
static inline q4 f_normalize4(q4 q, REAL *rinv_out) {
  if (!IS_F32) {
    REAL n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    REAL r = n2 > 0 ? R(1) / (REAL)sqrt((double)n2) : R(0);
    q4 o = { q.w * r, q.x * r, q.y * r, q.z * r };
    *rinv_out = r;
    return o;
  }
  float w = (float)q.w, x = (float)q.x, y = (float)q.y, z = (float)q.z;
  float n2 = fmaf(z, z, fmaf(y, y, fmaf(x, x, w * w)));
  float h = 0.5f * n2;
  int32_t bits; memcpy(&bits, &n2, 4);
  bits = 0x5f3759df - (bits >> 1);
  float r; memcpy(&r, &bits, 4);
  for (int i = 0; i < 2; i++) r = r * fmaf(-(h * r), r, 1.5f);
  float e = fmaf(-(n2 * r), r, 1.0f);
  r = fmaf(0.5f * r, e, r);
  q4 o = { (REAL)(w * r), (REAL)(x * r), (REAL)(y * r), (REAL)(z * r) };
  *rinv_out = (REAL)r;
  return o;
}

static void fast_destroy(ofast *F) { if (F) { free(F->passive); free(F); } }

/* pose of body b in the frame of its ancestor `top` (0 = world): constant transforms composed from the top down
 * (mirrors rel_pose of csrc/stacb_abi.cu operation for operation) */
static void f_rel_pose(const omodel *m, int top, int b, v3 *pos, q4 *quat) {
  int chain[512], n = 0;
  for (int c = b; c != top; c = m->body_parent[c]) chain[n++] = c;
  pos->x = pos->y = pos->z = R(0); quat->w = R(1); quat->x = quat->y = quat->z = R(0);
  for (int i = n - 1; i >= 0; i--) {
    int c = chain[i];
    if (i == n - 1) { *pos = ld3(m->body_pos + 3 * c); *quat = ld4(m->body_quat + 4 * c); }
    else {
      *pos = add3(*pos, f_rotq(ld3(m->body_pos + 3 * c), *quat));
      *quat = c_qmul(*quat, ld4(m->body_quat + 4 * c));
    }
  }
}

/* NULL when the register-resident path does not serve the model (mirrors stacb_tree_create / fits_fast).
 * Elements = the active bodies when they fit a warp; otherwise the JOINTED active bodies, every jointless (welded) active body
 * folded into its nearest jointed ancestor with a constant relative pose (fruitfly). */
static ofast *fast_create(const omodel *m, const osched *s) {
  int nb = m->nbody, K = m->nsite;
  if (nb > 512) return NULL;
  int fold = s->nact > 31;
  int *loc = (int *)calloc(nb, sizeof(int)), *el = (int *)calloc(nb, sizeof(int)), *epar = (int *)calloc(nb, sizeof(int)), *edep = (int *)calloc(nb, sizeof(int));
  int n_el = 0;
  for (int b = 0; b < nb; b++) loc[b] = -1;
  for (int e = 0; e < s->nact; e++) { int b = s->act[e]; if (!fold || m->body_jntnum[b] > 0) { loc[b] = n_el; el[n_el++] = b; } }
#define OWNER(bb, out) { int q_ = (bb); while (q_ != 0 && loc[q_] < 0) q_ = m->body_parent[q_]; (out) = q_; }
  int free_e = -1, free_j = -1, any_other = 0, jm = 1, maxd = 1;
  for (int e = 0; e < s->nact; e++) {  /* joint census over ALL active bodies, as the library does */
    int b = s->act[e];
    if (m->body_jntnum[b] > jm) jm = m->body_jntnum[b];
    for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
      int j = m->body_jntadr[b] + jj, t = m->jnt_type[j];
      if (t == JNT_FREE && free_e < 0 && jj == 0) { free_e = loc[b]; free_j = j; }
      else if (t != JNT_HINGE) any_other = 1;
    }
  }
  for (int e = 0; e < n_el; e++) {
    int pb; OWNER(m->body_parent[el[e]], pb);
    epar[e] = pb != 0 ? loc[pb] : -1;
    edep[e] = epar[e] >= 0 ? edep[epar[e]] + 1 : 1;
    if (edep[e] > maxd) maxd = edep[e];
  }
  int rounds = ceil_log2(maxd);
  int nquat = 0;
  for (int j = 0; j < m->njnt; j++) if (m->jnt_type[j] == JNT_FREE || m->jnt_type[j] == JNT_BALL) nquat++;
  int JM = 0, RT = 0, bplf = (s->nfull + 31) / 32, W = 1;
  if (n_el <= 31) {  /* one warp: the kernel variant is the first fit, surplus slots / rounds are executed as no-ops */
    if (jm <= 1 && rounds <= 5 && bplf <= 1) { JM = 1; RT = 5; }
    else if (jm <= 2 && rounds <= 4 && bplf <= 3) { JM = 2; RT = 4; }
    else if (jm <= 3 && rounds <= 4 && bplf <= 3) { JM = 3; RT = 4; }
    if (K > 31) JM = 0;
  } else {           /* wide variant: W = 2, 4, 6 or 8 warps per evaluation, one hinge per element, exact round count */
    W = 2 * ((n_el + 1 + 63) / 64);
    if (W <= 8 && jm <= 1 && rounds <= 8 && bplf <= 8 && K <= 32 * W - 1 && free_e < 32) { JM = 1; RT = rounds; }
  }
  int site_ok = 1;
  for (int k = 0; k < K; k++) { int ob; OWNER(m->site_body[k], ob); if (ob == 0) site_ok = 0; }
  if (n_el < 1 || any_other || nquat != (free_j >= 0 ? 1 : 0) || (free_j >= 0 && free_e < 0) || JM == 0 || !site_ok) {
    free(loc); free(el); free(epar); free(edep);
    return NULL;
  }
  /* ancestor tables of the element tree */
  int nr = rounds > 0 ? rounds : 1;
  int *eanc = (int *)calloc((size_t)nr * n_el, sizeof(int));
  for (int e = 0; e < n_el; e++) eanc[e] = epar[e];
  for (int r = 1; r < nr; r++)
    for (int e = 0; e < n_el; e++) { int a = eanc[(size_t)(r - 1) * n_el + e]; eanc[(size_t)r * n_el + e] = a >= 0 ? eanc[(size_t)(r - 1) * n_el + a] : -1; }
  ofast *F = (ofast *)calloc(1, sizeof(ofast));
  F->JM = JM; F->RT = RT; F->n = n_el; F->K = K; F->has_free = free_e >= 0;
  F->free_adr = free_e >= 0 ? m->jnt_qposadr[free_j] : 0;
  F->free_e = free_e >= 0 ? free_e : 0;
  F->folded = fold;
  F->W = W; F->NL = LANES * W; F->idl = F->NL - 1;
  char *covered = (char *)calloc(m->nq + 1, 1);
  for (int l = 0; l < F->NL; l++) {
    int on = l < n_el, b = on ? el[l] : 0;
    v3 bpos = { 0, 0, 0 }; q4 bquat = { 1, 0, 0, 0 };
    int nj = 0;
    if (on) {
      nj = m->body_jntnum[b];
      if (fold) { int pb; OWNER(m->body_parent[b], pb); f_rel_pose(m, pb, b, &bpos, &bquat); }
      else { bpos = ld3(m->body_pos + 3 * b); bquat = ld4(m->body_quat + 4 * b); }
    }
    F->pfree[l] = on && free_e >= 0 && l == free_e;
    v3 a0 = { 0, 0, 0 }, p0 = { 0, 0, 0 };
    for (int jj = 0; jj < JM; jj++) {
      int j = on && jj < nj ? m->body_jntadr[b] + jj : -1;
      int h = j >= 0 && m->jnt_type[j] == JNT_HINGE;
      F->hinge[l][jj] = h;
      F->valid[l][jj] = h;
      F->adr[l][jj] = h ? m->jnt_qposadr[j] : 0;
      F->ref[l][jj] = h ? m->qpos0[m->jnt_qposadr[j]] : R(0);
      if (h) covered[m->jnt_qposadr[j]] = 1;
      v3 z = { 0, 0, 0 };
      v3 a = h ? ld3(m->jnt_axis + 3 * j) : z, jp = h ? ld3(m->jnt_pos + 3 * j) : z;
      if (jj == 0) { a0 = a; p0 = jp; }
      REAL da = c_dot3(a, jp);
      v3 ada = { a.x * da, a.y * da, a.z * da };
      F->jax[l][jj] = a; F->jps[l][jj] = jp; F->jpp[l][jj] = sub3(jp, ada); F->jcx[l][jj] = c_cross(a, jp);
    }
    F->valid[l][JM] = F->has_free && l < 7;
    F->adr[l][JM] = F->valid[l][JM] ? F->free_adr + l : 0;
    if (F->valid[l][JM]) covered[F->free_adr + l] = 1;
    F->Qc[l] = bquat;
    q4 qa = { 0, a0.x, a0.y, a0.z };
    F->Qs[l] = c_qmul(bquat, qa);
    v3 rb = f_rotq(F->jpp[l][0], bquat), rc = f_rotq(F->jcx[l][0], bquat);
    F->A[l] = add3(bpos, rb);
    F->B[l].x = -rb.x; F->B[l].y = -rb.y; F->B[l].z = -rb.z;
    F->C[l].x = -rc.x; F->C[l].y = -rc.y; F->C[l].z = -rc.z;
    F->anc0[l] = add3(f_rotq(p0, bquat), bpos);
    F->ax0[l] = f_rotq(a0, bquat);
    for (int r = 0; r < RT; r++) {
      int a = l >= n_el ? l : F->idl;
      if (on && r < rounds) { int t = eanc[(size_t)r * n_el + l]; if (t >= 0) a = t; }
      F->src[l][r] = a;
    }
    F->par[l] = on ? (epar[l] >= 0 ? epar[l] : F->idl) : l;
    int j0 = on && nj > 0 ? m->body_jntadr[b] : -1;
    int live = j0 >= 0 && s->jnt_e[j0] > s->jnt_s[j0];
    F->sa[l] = live ? s->jnt_s[j0] : 0; F->se[l] = live ? s->jnt_e[j0] : 0;
    F->sk[l] = -1; F->seb[l] = F->idl;
    F->srel_p[l].x = F->srel_p[l].y = F->srel_p[l].z = R(0);
    F->srel_q[l].w = R(1); F->srel_q[l].x = F->srel_q[l].y = F->srel_q[l].z = R(0);
    if (l >= 1 && l <= K) {  /* site p on lane p + 1 */
      int k = s->site_order[l - 1], ob;
      OWNER(m->site_body[k], ob);
      F->sk[l] = k; F->seb[l] = loc[ob];
      if (fold) f_rel_pose(m, ob, m->site_body[k], &F->srel_p[l], &F->srel_q[l]);
    }
  }
#undef OWNER
  F->passive = (int *)calloc(m->nq + 1, sizeof(int));
  for (int i = 0; i < m->nq; i++) if (!covered[i]) F->passive[F->npassive++] = i;
  free(covered); free(loc); free(el); free(epar); free(edep); free(eanc);
  if (F->npassive > 0 && F->passive[0] < 3) { fast_destroy(F); return NULL; }
  return F;
}

typedef struct {
  v3 P[FMAXL]; q4 Q[FMAXL], Qp[FMAXL];
  v3 lp[FMAXL][FJ]; q4 lq[FMAXL][FJ]; /* parent-frame pose of the body before hinge slot j >= 1 */
  v3 s[FMAXL], res[FMAXL];
  v3 fpos; q4 fq; REAL frinv;
} ffwd;

/* per-site data in lane order */
typedef struct { v3 off[FMAXL], kp[FMAXL], km[FMAXL]; } fsites;

static void fast_sites(const ofast *F, const REAL *site_pos, const REAL *kp, const REAL *kpmask, fsites *st) {
  memset(st, 0, sizeof(*st));
  for (int l = 0; l < F->NL; l++) {
    int k = F->sk[l];
    if (k < 0) continue;
    if (site_pos) st->off[l] = ld3(site_pos + 3 * k);
    if (F->folded) st->off[l] = add3(F->srel_p[l], f_rotq(st->off[l], F->srel_q[l]));
    if (kp) st->kp[l] = ld3(kp + 3 * k);
    if (kpmask) st->km[l] = ld3(kpmask + 3 * k);
  }
}

static REAL fast_fwd(const ofast *F, const fsites *st, REAL pt[FMAXL][FNS], ffwd *S) {
  const int JM = F->JM, RT = F->RT;
  if (F->has_free) {
    S->fpos.x = pt[0][JM]; S->fpos.y = pt[1][JM]; S->fpos.z = pt[2][JM];
    q4 raw = { pt[3][JM], pt[4][JM], pt[5][JM], pt[6][JM] };
    S->fq = f_normalize4(raw, &S->frinv);
  } else {
    S->fpos.x = S->fpos.y = S->fpos.z = R(0); S->fq.w = R(1); S->fq.x = S->fq.y = S->fq.z = R(0); S->frinv = R(1);
  }
  q4 Q[FMAXL], Qn[FMAXL];
  v3 lp[FMAXL];
  for (int l = 0; l < F->NL; l++) {
    REAL sh[FJ], ch[FJ];
    for (int j = 0; j < JM; j++) f_sincos((pt[l][j] - F->ref[l][j]) * R(0.5), &sh[j], &ch[j]);
    REAL ct = r_fma(ch[0], ch[0], -(sh[0] * sh[0])), sn = R(2) * (sh[0] * ch[0]);
    q4 quat = { r_fma(F->Qs[l].w, sh[0], F->Qc[l].w * ch[0]), r_fma(F->Qs[l].x, sh[0], F->Qc[l].x * ch[0]),
                r_fma(F->Qs[l].y, sh[0], F->Qc[l].y * ch[0]), r_fma(F->Qs[l].z, sh[0], F->Qc[l].z * ch[0]) };
    v3 pos = { r_fma(F->C[l].x, sn, r_fma(F->B[l].x, ct, F->A[l].x)), r_fma(F->C[l].y, sn, r_fma(F->B[l].y, ct, F->A[l].y)),
               r_fma(F->C[l].z, sn, r_fma(F->B[l].z, ct, F->A[l].z)) };
    if (F->pfree[l]) { pos = S->fpos; quat = S->fq; }
    for (int j = 1; j < JM; j++) {
      S->lp[l][j] = pos; S->lq[l][j] = quat;
      ct = r_fma(ch[j], ch[j], -(sh[j] * sh[j]));
      sn = R(2) * (sh[j] * ch[j]);
      REAL om = R(1) - ct;
      v3 pl = { r_fma(-F->jcx[l][j].x, sn, F->jpp[l][j].x * om), r_fma(-F->jcx[l][j].y, sn, F->jpp[l][j].y * om),
                r_fma(-F->jcx[l][j].z, sn, F->jpp[l][j].z * om) };
      pos = add3(pos, f_rotq(pl, quat));
      q4 ql = { ch[j], F->jax[l][j].x * sh[j], F->jax[l][j].y * sh[j], F->jax[l][j].z * sh[j] };
      quat = c_qmul(quat, ql);
    }
    Q[l] = quat; lp[l] = pos;
  }
  v3 v[FMAXL], vn[FMAXL];
  if (F->W == 1) {  /* one warp: quaternions first, then ONE rotation per element and additive jumping of the positions */
    for (int r = 0; r < RT; r++) {
      for (int l = 0; l < F->NL; l++) Qn[l] = c_qmul(Q[F->src[l][r]], Q[l]);
      memcpy(Q, Qn, sizeof(Q));
    }
    for (int l = 0; l < F->NL; l++) { S->Qp[l] = Q[F->par[l]]; v[l] = f_rotq(lp[l], S->Qp[l]); }
    for (int r = 0; r < RT; r++) {
      for (int l = 0; l < F->NL; l++) vn[l] = add3(v[F->src[l][r]], v[l]);
      memcpy(v, vn, sizeof(v));
    }
  } else {          /* several warps: position and quaternion jump together, one exchange per round (stacb_wide.cuh) */
    for (int l = 0; l < F->NL; l++) v[l] = lp[l];
    for (int r = 0; r < RT; r++) {
      for (int l = 0; l < F->NL; l++) {
        int a = F->src[l][r];
        vn[l] = add3(v[a], f_rotq(v[l], Q[a]));
        Qn[l] = c_qmul(Q[a], Q[l]);
      }
      memcpy(Q, Qn, sizeof(Q)); memcpy(v, vn, sizeof(v));
    }
    for (int l = 0; l < F->NL; l++) S->Qp[l] = Q[F->par[l]];
  }
  REAL part[FMAXL];
  for (int l = 0; l < F->NL; l++) {
    S->P[l] = v[l]; S->Q[l] = Q[l];
  }
  for (int l = 0; l < F->NL; l++) {
    v3 pb = v[F->seb[l]]; q4 qb = Q[F->seb[l]];
    S->s[l] = add3(pb, f_rotq(st->off[l], qb));
    S->res[l].x = (st->kp[l].x - S->s[l].x) * st->km[l].x;
    S->res[l].y = (st->kp[l].y - S->s[l].y) * st->km[l].y;
    S->res[l].z = (st->kp[l].z - S->s[l].z) * st->km[l].z;
    part[l] = r_fma(S->res[l].z, S->res[l].z, r_fma(S->res[l].y, S->res[l].y, S->res[l].x * S->res[l].x));
  }
  return f_cta_sum(F->W, part);
}

static void fast_bwd(const ofast *F, const ffwd *S, int free_wanted, REAL g[FMAXL][FNS]) {
  const int JM = F->JM;
  v3 c = S->P[0];
  REAL w[FMAXL][6], t2[FMAXL][6], wrl[FMAXL][6];
  for (int l = 0; l < F->NL; l++) {
    v3 f = { R(-2) * S->res[l].x, R(-2) * S->res[l].y, R(-2) * S->res[l].z };
    v3 tq = c_cross(sub3(S->s[l], c), f);
    w[l][0] = f.x; w[l][1] = f.y; w[l][2] = f.z; w[l][3] = tq.x; w[l][4] = tq.y; w[l][5] = tq.z;
  }
  for (int off = 1; off < LANES; off <<= 1) {  /* Hillis-Steele inside every warp */
    for (int l = 0; l < F->NL; l++) for (int i = 0; i < 6; i++) t2[l][i] = ((l % LANES) >= off) ? w[l][i] + w[l - off][i] : w[l][i];
    memcpy(w, t2, sizeof(w));
  }
  if (F->W > 1) {  /* warp w >= 1 adds the totals of the warps before it, summed in warp order */
    REAL offs[6], tot[8][6];
    for (int ww = 0; ww < F->W; ww++) for (int i = 0; i < 6; i++) tot[ww][i] = w[LANES * ww + LANES - 1][i];
    for (int ww = 1; ww < F->W; ww++) {
      for (int i = 0; i < 6; i++) offs[i] = (ww == 1) ? tot[0][i] : offs[i] + tot[ww - 1][i];
      for (int l = LANES * ww; l < LANES * (ww + 1); l++) for (int i = 0; i < 6; i++) w[l][i] = w[l][i] + offs[i];
    }
  }
  /* site p sits on lane p + 1: the inclusive scan at lane l is the sum over the sites p < l */
  for (int l = 0; l < F->NL; l++) {
    REAL *wr = wrl[l];
    for (int i = 0; i < 6; i++) wr[i] = w[F->se[l]][i] - w[F->sa[l]][i];
    v3 Fo = { wr[0], wr[1], wr[2] }, Tq = { wr[3], wr[4], wr[5] };
    v3 pp = S->P[F->par[l]];
    q4 pc = { S->Qp[l].w, -S->Qp[l].x, -S->Qp[l].y, -S->Qp[l].z };
    v3 T0 = sub3(Tq, c_cross(sub3(pp, c), Fo));
    v3 Fp = f_rotq(Fo, pc), Tp = f_rotq(T0, pc);
    g[l][0] = c_dot3(F->ax0[l], sub3(Tp, c_cross(F->anc0[l], Fp)));
    for (int j = 1; j < JM; j++) {
      v3 anc = add3(S->lp[l][j], f_rotq(F->jps[l][j], S->lq[l][j])), ax = f_rotq(F->jax[l][j], S->lq[l][j]);
      g[l][j] = c_dot3(ax, sub3(Tp, c_cross(anc, Fp)));
    }
    g[l][JM] = R(0);
  }
  if (free_wanted) {
    const REAL *wf = wrl[F->free_e];
    v3 Ff = { wf[0], wf[1], wf[2] }, Tw = { wf[3], wf[4], wf[5] };
    v3 Tf = sub3(Tw, c_cross(sub3(S->fpos, c), Ff));
    q4 qh = S->fq, tq = { 0, Tf.x, Tf.y, Tf.z };
    q4 h = c_qmul(tq, qh); h.w *= R(2); h.x *= R(2); h.y *= R(2); h.z *= R(2);
    REAL pr = r_fma(qh.z, h.z, r_fma(qh.y, h.y, r_fma(qh.x, h.x, qh.w * h.w))), n = S->frinv;
    REAL g7[7] = { Ff.x, Ff.y, Ff.z, r_fma(-qh.w, pr, h.w) * n, r_fma(-qh.x, pr, h.x) * n, r_fma(-qh.y, pr, h.y) * n, r_fma(-qh.z, pr, h.z) * n };
    for (int l = 0; l < 7; l++) g[l][JM] = g7[l];
  }
}

/* slot layout <-> qpos layout */
static void fast_gather(const ofast *F, const REAL *q, REAL out[FMAXL][FNS]) {
  for (int l = 0; l < F->NL; l++) for (int m = 0; m <= F->JM; m++) out[l][m] = F->valid[l][m] ? q[F->adr[l][m]] : R(0);
}
static void fast_bits(const ofast *F, const uint8_t *qmask, int bits[FMAXL][FNS]) {
  for (int l = 0; l < F->NL; l++) for (int m = 0; m <= F->JM; m++) bits[l][m] = F->valid[l][m] && qmask[F->adr[l][m]];
}
static int fast_free_wanted(const ofast *F, int bits[FMAXL][FNS]) {
  if (!F->has_free) return 0;
  for (int l = 0; l < 7; l++) if (bits[l][F->JM]) return 1;
  return 0;
}
static REAL fast_lane_dot(int NS, const REAL *a, const REAL *b) {
  REAL acc = a[0] * b[0];
  for (int m = 1; m < NS; m++) acc = r_fma(a[m], b[m], acc);
  return acc;
}
static REAL fast_dot(const ofast *F, REAL a[FMAXL][FNS], REAL b[FMAXL][FNS]) {
  REAL part[FMAXL];
  for (int l = 0; l < F->NL; l++) part[l] = fast_lane_dot(F->JM + 1, a[l], b[l]);
  return f_cta_sum(F->W, part);
}

/* q_loss and its gradient in fast order; q, q0, grad in the qpos layout */
static REAL fast_loss_eval(const omodel *m, const ofast *F, const REAL *q, const REAL *q0, const uint8_t *qmask, const REAL *kp,
                           const REAL *kpmask, const REAL *site_pos, REAL *grad) {
  fsites st; ffwd S;
  fast_sites(F, site_pos, kp, kpmask, &st);
  REAL qs[FMAXL][FNS], q0s[FMAXL][FNS], pt[FMAXL][FNS], g[FMAXL][FNS];
  int bits[FMAXL][FNS];
  fast_gather(F, q, qs); fast_gather(F, q0, q0s); fast_bits(F, qmask, bits);
  for (int l = 0; l < F->NL; l++) for (int mm = 0; mm <= F->JM; mm++) pt[l][mm] = bits[l][mm] ? qs[l][mm] : q0s[l][mm];
  REAL loss = fast_fwd(F, &st, pt, &S);
  if (grad) {
    fast_bwd(F, &S, fast_free_wanted(F, bits), g);
    for (int i = 0; i < m->nq; i++) grad[i] = R(0);
    for (int l = 0; l < F->NL; l++) for (int mm = 0; mm <= F->JM; mm++) if (bits[l][mm]) grad[F->adr[l][mm]] = g[l][mm];
  }
  return loss;
}

/* jaxopt ProjectedGradient in fast order (stacb_fast.cuh: solve); q0, params in the qpos layout */
static solve_info fast_q_opt(const omodel *m, const ofast *F, const REAL *q0, const REAL *lb, const REAL *ub, const uint8_t *qmask,
                             const REAL *kp, const REAL *kpmask, const REAL *site_pos, REAL tol, int maxiter, int maxls, REAL *params) {
  const int NS = F->JM + 1;
  fsites st; ffwd S;
  fast_sites(F, site_pos, kp, kpmask, &st);
  REAL q0s[FMAXL][FNS], x[FMAXL][FNS], y[FMAXL][FNS], g[FMAXL][FNS], xn[FMAXL][FNS], d[FMAXL][FNS], gt[FMAXL][FNS], pt[FMAXL][FNS];
  REAL lbs[FMAXL][FNS], ubs[FMAXL][FNS], gm[FMAXL][FNS], dn[FMAXL][FNS];
  int bits[FMAXL][FNS];
  fast_gather(F, q0, q0s); fast_gather(F, lb, lbs); fast_gather(F, ub, ubs); fast_bits(F, qmask, bits);
  const int fw = fast_free_wanted(F, bits);
  /* frozen slots (valid, not optimised): zero gradient; they stay at q0 inside the solve (unbounded effective box) and the squared
     length of the reference's one move to clip(q0) enters the first line search through sqn (solve_setup of stacb_fast.cuh) */
  for (int l = 0; l < F->NL; l++) for (int mm = 0; mm < NS; mm++) {
    int frozen = F->valid[l][mm] && !bits[l][mm];
    gm[l][mm] = bits[l][mm] ? R(1) : R(0);
    dn[l][mm] = frozen ? clipr(q0s[l][mm], lbs[l][mm], ubs[l][mm]) - q0s[l][mm] : R(0);
    if (frozen) { lbs[l][mm] = -(REAL)INFINITY; ubs[l][mm] = (REAL)INFINITY; }
  }
  const REAL sqn = fast_dot(F, dn, dn);
  /* passive coordinates: squared length of the move to clip(q0), dealt out over the lanes */
  REAL sqp;
  {
    REAL part[FMAXL];
    for (int l = 0; l < F->NL; l++) {
      REAL acc = R(0); int first = 1;
      for (int i = l; i < F->npassive; i += F->NL) {
        int p = F->passive[i];
        REAL dd = clipr(q0[p], lb[p], ub[p]) - q0[p];
        acc = first ? dd * dd : r_fma(dd, dd, acc);
        first = 0;
      }
      part[l] = acc;
    }
    sqp = sqn + f_cta_sum(F->W, part);
  }
  memcpy(params, q0, sizeof(REAL) * m->nq);
  for (int l = 0; l < F->NL; l++) for (int mm = 0; mm < NS; mm++) { x[l][mm] = q0s[l][mm]; y[l][mm] = x[l][mm]; }
  REAL t = R(1), step = R(1), err = (REAL)INFINITY;
  solve_info info = { err, 0, 0 };
  if (maxiter <= 0) return info;
  do {
    memcpy(pt, y, sizeof(pt));
    REAL fy = fast_fwd(F, &st, pt, &S);
    fast_bwd(F, &S, fw, g);
    for (int l = 0; l < F->NL; l++) for (int mm = 0; mm < NS; mm++) g[l][mm] = g[l][mm] * gm[l][mm];
    REAL stp = step;
    int halv = 0;
    for (;;) {
      for (int l = 0; l < F->NL; l++) for (int mm = 0; mm < NS; mm++) {
        xn[l][mm] = clipr(r_fma(-stp, g[l][mm], y[l][mm]), lbs[l][mm], ubs[l][mm]);
        d[l][mm] = xn[l][mm] - y[l][mm];
      }
      REAL sq = fast_dot(F, d, d), dg = fast_dot(F, d, g);
      sq = sq + sqp;
      memcpy(pt, xn, sizeof(pt));
      REAL fn = fast_fwd(F, &st, pt, &S);
      info.ls_evals++;
      REAL dec = stp * (fn - fy);
      REAL cond = r_fma(stp, dg, R(0.5) * sq);
      if (!(dec > cond + R_EPS) || halv >= maxls) break;
      stp = stp * R(0.5); halv++;
    }
    if (info.iters == 0) ls_trace_push(-1);
    ls_trace_push(halv);
    /* S holds the state of the accepted candidate: gradient at x+ */
    fast_bwd(F, &S, fw, gt);
    for (int l = 0; l < F->NL; l++) for (int mm = 0; mm < NS; mm++) gt[l][mm] = gt[l][mm] * gm[l][mm];
    step = (stp <= R(1e-6)) ? R(1) : stp / R(0.5);
    REAL tn = R(0.5) * (R(1) + r_sqrt(r_fma(R(4) * t, t, R(1))));
    REAL beta = (t - R(1)) / tn;
    for (int l = 0; l < F->NL; l++) for (int mm = 0; mm < NS; mm++) {
      y[l][mm] = r_fma(beta, xn[l][mm] - x[l][mm], xn[l][mm]);
      d[l][mm] = clipr(xn[l][mm] - gt[l][mm], lbs[l][mm], ubs[l][mm]) - xn[l][mm];
      x[l][mm] = xn[l][mm];
    }
    err = r_sqrt(fast_dot(F, d, d));
    t = tn; info.iters++;
    sqp = R(0);
  } while (err > tol && info.iters < maxiter);
  info.error = err;
  for (int i = 0; i < F->npassive; i++) { int p = F->passive[i]; params[p] = clipr(q0[p], lb[p], ub[p]); }
  for (int l = 0; l < F->NL; l++) for (int mm = 0; mm < NS; mm++) if (F->valid[l][mm]) {
    int a = F->adr[l][mm];
    params[a] = bits[l][mm] ? x[l][mm] : clipr(x[l][mm], lb[a], ub[a]);  /* the reference's iterate of a frozen coordinate is clip(q0) */
  }
  return info;
}
