// stacb_fast.cuh -- the register-resident solver for hinge trees (sm_100a): the hot path of libstacb.so.
//
// Applies to models whose ACTIVE subtree (ancestors of the marker bodies) reduces to at most 31 ELEMENTS and at most 31 marker
// sites, with hinge joints only plus (optionally) one free joint on a top-level body.  Elements are the active bodies themselves
// when they fit (rodent / rat23: 31, C. elegans: 25); otherwise the active bodies that carry joints, every jointless (welded) active
// body folded into its nearest jointed ancestor by stacb_tree_create (fruitfly: 49-57 active bodies -> 25 elements).  Everything
// else takes the general kernels of stacb_device.cuh.
//
// Mapping (fixed for the whole kernel, all state in registers, no shared memory inside an evaluation):
//   lane e  <->  element e (body id order, parents first); lane 31 (and every lane >= n) is an IDENTITY
//               element, so "no ancestor" needs no select: composing with the identity is exact;
//   lane p+1 <-> marker site at sorted position p (sites sorted by body id, so a subtree is a contiguous range; lane 0 carries
//               a zero wrench, which makes the inclusive prefix scan over the lanes the exclusive one of the sites);
//   solver  <->  slot j < JM of a lane is the angle of hinge j of the lane's body, slot JM of lanes 0..6 holds the seven
//               coordinates of the free joint.  Coordinates of the model that are not in this layout ("passive": hinges
//               outside the active subtree) have an identically zero gradient and never move (see solve).
//
// Canonical arithmetic of this path ("fast order"; the CPU oracle's mode 2 mirrors it op for op):
//   * sin / cos of the hinge half-angles up to a common sign (reduction by pi, no quadrant logic: sincos_pi);
//   * first hinge of a body folded with the body's constant pose: quat = Qc cos(h) + Qs sin(h), pos = A + B cos(t) + C sin(t),
//     h = (q - ref)/2, cos t = cos^2 h - sin^2 h, sin t = 2 sin h cos h  (Qc, Qs, A, B, C derived once per lane);
//   * further hinges of the same body composed in the parent frame: quat' = quat * ql, pos' = pos + R(quat)(jperp (1 - cos t) - (a x jpos) sin t);
//   * world quaternions by pointer jumping (round r composes with the ancestor at distance 2^r), then ONE rotation of every local
//     offset by its parent's world quaternion and pointer jumping with plain additions for the world positions;
//   * rotate(v, q) = v + 2 (s t + u x t), t = u x v;
//   * residuals, loss butterfly, wrench prefix scan and parent-frame Jacobian transpose as in DESIGN.md section 4.
// Replaces reference stac_mjx/stac_core.py:27-99 and the MJX / jaxopt code under it (see stacb_device.cuh for the citations).
#pragma once
#include "stacb_device.cuh"

namespace stacb {
namespace fast {

constexpr unsigned FULL = 0xffffffffu;
constexpr int IDL = 31;  // identity lane

// v + 2 (s (u x v) + u x (u x v)): rotation of v by the (unit) quaternion q = (s, u)
__device__ __forceinline__ V3 rotq(V3 v, Q4 q) {
  const V3 u = mk3(q.x, q.y, q.z);
  const V3 t = cross3(u, v);
  const V3 c = cross3(u, t);
  const V3 w = mk3(fmaf(q.w, t.x, c.x), fmaf(q.w, t.y, c.y), fmaf(q.w, t.z, c.z));
  return mk3(fmaf(2.0f, w.x, v.x), fmaf(2.0f, w.y, v.y), fmaf(2.0f, w.z, v.z));
}
__device__ __forceinline__ Q4 conj4(Q4 q) { return mk4(q.w, -q.x, -q.y, -q.z); }

template <int JM, int RT>
struct LaneC {
  Q4 Qc, Qs;            // first hinge folded with the body's constant pose
  V3 A, B, C;
  V3 anc[JM], ax[JM];   // slot 0: constant parent-frame anchor / axis of the first hinge (slots >= 1 are derived in the reverse sweep)
  V3 jax[JM], jpp[JM], jcx[JM], jps[JM];  // slots >= 1: axis, jpos - a (a.jpos), a x jpos, jpos (body frame)
  float ref[JM];
  int adr[JM];          // qpos address of the hinge in slot j (0 when the slot is empty)
  bool hinge[JM];
  bool pfree;           // the lane's body carries the primary free joint
  int src[RT];          // lane of the ancestor at distance 2^r (IDL when there is none)
  int par;              // lane of the parent (IDL for a top-level body)
  int sa, se;           // sorted-site range below the body (0, 0 when nothing to differentiate)
};

struct SiteC {
  int k;       // keypoint index of the site at sorted position `lane`, -1 if none
  int eb;      // lane of the site's body
  int ef;      // full-set element of the site's body (per-frame outputs)
  V3 off, kp, km;
};

template <int JM>
struct Fwd {
  V3 P; Q4 Q;          // world pose of the lane's body
  Q4 Qp;               // world quaternion of its parent
  V3 lp[JM]; Q4 lq[JM]; // parent-frame pose of the body BEFORE hinge slot j >= 1 is applied (the reverse sweep derives that hinge's
                        // anchor and axis from it; slot 0's are lane constants)
  V3 s, res;           // marker site position and masked residual (site lanes)
  V3 fpos; Q4 fq; float frinv;  // free joint: position, normalised quaternion, reciprocal of the normalisation divisor
};

template <int JM, int RT>
__device__ __forceinline__ void lane_init(LaneC<JM, RT> &L, const DevTree &T, int lane, int idl = IDL) {
  const DevSet &S = T.fs;
  const bool on = lane < S.n;
  BodyConst b;
  load_body(b, S.rec + (size_t)(on ? lane : 0) * REC);
  if (!on) { b.nj = 0; b.parent = -1; b.pos = mk3(0.f, 0.f, 0.f); b.quat = mk4(1.f, 0.f, 0.f, 0.f); }
  L.pfree = on && lane == T.fs_free_e;
  V3 a0 = mk3(0.f, 0.f, 0.f), p0 = mk3(0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < JM; j++) {
    const bool h = on && j < b.nj && b.jtype[j] == STACB_JNT_HINGE;
    L.hinge[j] = h;
    L.adr[j] = h ? b.jadr[j] : 0;
    L.ref[j] = h ? b.jref[j] : 0.f;
    const V3 a = h ? b.jaxis[j] : mk3(0.f, 0.f, 0.f), jp = h ? b.jpos[j] : mk3(0.f, 0.f, 0.f);
    if (j == 0) { a0 = a; p0 = jp; }
    const float da = dot3(a, jp);
    L.jax[j] = a;
    L.jps[j] = jp;
    L.jpp[j] = sub3(jp, mk3(a.x * da, a.y * da, a.z * da));
    L.jcx[j] = cross3(a, jp);
    L.anc[j] = mk3(0.f, 0.f, 0.f);
    L.ax[j] = mk3(0.f, 0.f, 0.f);
  }
  L.Qc = b.quat;
  L.Qs = qmul(b.quat, mk4(0.f, a0.x, a0.y, a0.z));
  const V3 rb = rotq(L.jpp[0], b.quat), rc = rotq(L.jcx[0], b.quat);
  L.A = add3(b.pos, rb);
  L.B = mk3(-rb.x, -rb.y, -rb.z);
  L.C = mk3(-rc.x, -rc.y, -rc.z);
  L.anc[0] = add3(rotq(p0, b.quat), b.pos);
  L.ax[0] = rotq(a0, b.quat);
#pragma unroll
  for (int r = 0; r < RT; r++) {
    int a = lane >= S.n ? lane : idl;
    if (on && r < S.rounds) { const int t = __ldg(S.anc + r * S.n + lane); if (t >= 0) a = t; }
    L.src[r] = a;
  }
  L.par = on ? (b.parent >= 0 ? b.parent : idl) : lane;
  const bool live = on && b.nj > 0 && b.jse[0] > b.jsa[0];
  L.sa = live ? b.jsa[0] : 0;
  L.se = live ? b.jse[0] : 0;
}

__device__ __forceinline__ void site_init(SiteC &st, const DevTree &T, int lane, const float *__restrict__ site_pos, int idl = IDL) {
  st.k = -1; st.eb = idl; st.ef = 0;
  st.off = mk3(0.f, 0.f, 0.f); st.kp = mk3(0.f, 0.f, 0.f); st.km = mk3(0.f, 0.f, 0.f);
  if (lane >= 1 && lane <= T.K) {
    st.k = __ldg(T.site_order + lane - 1);
    st.eb = __ldg(T.site_efs + lane - 1);
    st.ef = __ldg(T.site_efull + lane - 1);
    if (site_pos) st.off = mk3(__ldg(site_pos + 3 * st.k), __ldg(site_pos + 3 * st.k + 1), __ldg(site_pos + 3 * st.k + 2));
    if (T.site_rel) {  // the site's body is welded to the element: express the offset in the element's frame (uniform branch)
      const float *r = T.site_rel + 7 * (lane - 1);
      st.off = add3(mk3(__ldg(r), __ldg(r + 1), __ldg(r + 2)), rotq(st.off, mk4(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5), __ldg(r + 6))));
    }
  }
}
__device__ __forceinline__ void site_load_kp(SiteC &st, const float *__restrict__ kp) {
  if (st.k >= 0) st.kp = mk3(kp[3 * st.k], kp[3 * st.k + 1], kp[3 * st.k + 2]);
}
__device__ __forceinline__ void site_mask_u8(SiteC &st, const uint8_t *__restrict__ m3 /*[3K] or null = ones*/) {
  if (st.k >= 0) st.km = m3 ? mk3(m3[3 * st.k] ? 1.f : 0.f, m3[3 * st.k + 1] ? 1.f : 0.f, m3[3 * st.k + 2] ? 1.f : 0.f) : mk3(1.f, 1.f, 1.f);
}
__device__ __forceinline__ void site_mask_kp(SiteC &st, const uint8_t *__restrict__ per_kp /*[K] or null = ones*/) {
  if (st.k >= 0) { const float m = (!per_kp || per_kp[st.k]) ? 1.f : 0.f; st.km = mk3(m, m, m); }
}

// Forward half of q_loss (stac_core.py:27-63) at the point `pt` (solver layout, already merged with q0 by make_qs).
// RT = pointer-jumping rounds of the kernel variant (>= ceil(log2(depth)); surplus rounds compose with the identity).
template <int JM, int RT>
__device__ __forceinline__ float eval_fwd(const LaneC<JM, RT> &L, const SiteC &st, bool has_free, const float (&pt)[JM + 1], Fwd<JM> &S) {
  // free joint: every lane normalises the same seven values (broadcast from lanes 0..6)
  if (has_free) {
    float fr[7];
#pragma unroll
    for (int i = 0; i < 7; i++) fr[i] = __shfl_sync(FULL, pt[JM], i);
    S.fpos = mk3(fr[0], fr[1], fr[2]);
    S.fq = normalize4_nr(mk4(fr[3], fr[4], fr[5], fr[6]), &S.frinv);
  } else {
    S.fpos = mk3(0.f, 0.f, 0.f); S.fq = mk4(1.f, 0.f, 0.f, 0.f); S.frinv = 1.f;
  }
  float sh[JM], ch[JM];
#pragma unroll
  for (int j = 0; j < JM; j++) sincos_pi((pt[j] - L.ref[j]) * 0.5f, &sh[j], &ch[j]);
  // first hinge, folded with the body's constant pose
  float ct = fmaf(ch[0], ch[0], -(sh[0] * sh[0])), sn = 2.0f * (sh[0] * ch[0]);
  Q4 quat = mk4(fmaf(L.Qs.w, sh[0], L.Qc.w * ch[0]), fmaf(L.Qs.x, sh[0], L.Qc.x * ch[0]), fmaf(L.Qs.y, sh[0], L.Qc.y * ch[0]),
                fmaf(L.Qs.z, sh[0], L.Qc.z * ch[0]));
  V3 pos = mk3(fmaf(L.C.x, sn, fmaf(L.B.x, ct, L.A.x)), fmaf(L.C.y, sn, fmaf(L.B.y, ct, L.A.y)), fmaf(L.C.z, sn, fmaf(L.B.z, ct, L.A.z)));
  pos = sel3(L.pfree, S.fpos, pos);
  quat = sel4(L.pfree, S.fq, quat);
#pragma unroll
  for (int j = 1; j < JM; j++) {  // further hinges of the same body
    S.lp[j] = pos; S.lq[j] = quat;
    ct = fmaf(ch[j], ch[j], -(sh[j] * sh[j]));
    sn = 2.0f * (sh[j] * ch[j]);
    const float om = 1.0f - ct;
    const V3 pl = mk3(fmaf(-L.jcx[j].x, sn, L.jpp[j].x * om), fmaf(-L.jcx[j].y, sn, L.jpp[j].y * om), fmaf(-L.jcx[j].z, sn, L.jpp[j].z * om));
    pos = add3(pos, rotq(pl, quat));
    quat = qmul(quat, mk4(ch[j], L.jax[j].x * sh[j], L.jax[j].y * sh[j], L.jax[j].z * sh[j]));
  }
  // world quaternions: pointer jumping
  Q4 Q = quat;
#pragma unroll
  for (int r = 0; r < RT; r++) Q = qmul(shfl4(Q, L.src[r]), Q);
  // world positions: rotate the local offset by the parent's world quaternion, then pointer jumping with additions
  S.Qp = shfl4(Q, L.par);
  V3 v = rotq(pos, S.Qp);
#pragma unroll
  for (int r = 0; r < RT; r++) v = add3(shfl3(v, L.src[r]), v);
  S.P = v;
  S.Q = Q;
  // marker sites, masked residuals, loss
  const V3 pb = shfl3(v, st.eb);
  const Q4 qb = shfl4(Q, st.eb);
  S.s = add3(pb, rotq(st.off, qb));
  S.res = mk3((st.kp.x - S.s.x) * st.km.x, (st.kp.y - S.s.y) * st.km.y, (st.kp.z - S.s.z) * st.km.z);
  return warp_sum(fmaf(S.res.z, S.res.z, fmaf(S.res.y, S.res.y, S.res.x * S.res.x)));
}

// Reverse half: d loss / d (solver slots) at the point of the last eval_fwd (its state S is still live).
//   free_wanted: some coordinate of the free joint is optimised (uniform); free_e: lane of the free joint's body.
template <int JM, int RT>
__device__ __forceinline__ void eval_bwd(const LaneC<JM, RT> &L, const Fwd<JM> &S, int lane, bool free_wanted, int free_e, float (&g)[JM + 1]) {
  const V3 c = shfl3(S.P, 0);
  const V3 f = mk3(-2.0f * S.res.x, -2.0f * S.res.y, -2.0f * S.res.z);
  const V3 tq = cross3(sub3(S.s, c), f);
  float w[6] = {f.x, f.y, f.z, tq.x, tq.y, tq.z};
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const float up = __shfl_up_sync(FULL, w[i], off);
      if (lane >= off) w[i] = w[i] + up;
    }
  }
  // sites sit one lane up (site p on lane p + 1, lane 0 carries a zero wrench): the inclusive scan at lane l is the sum over
  // the sites p < l, so a subtree's range [sa, se) is a difference of two scan values without an exclusive shift
  float wr[6];
#pragma unroll
  for (int i = 0; i < 6; i++) wr[i] = __shfl_sync(FULL, w[i], L.se) - __shfl_sync(FULL, w[i], L.sa);
  const V3 F = mk3(wr[0], wr[1], wr[2]), Tq = mk3(wr[3], wr[4], wr[5]);
  // the subtree wrench in the parent frame of the body, then every hinge slot of the body
  const V3 pp = shfl3(S.P, L.par);
  const Q4 pc = conj4(S.Qp);
  const V3 T0 = sub3(Tq, cross3(sub3(pp, c), F));
  const V3 Fp = rotq(F, pc), Tp = rotq(T0, pc);
  g[0] = dot3(L.ax[0], sub3(Tp, cross3(L.anc[0], Fp)));
#pragma unroll
  for (int j = 1; j < JM; j++) {  // anchor / axis of hinge slot j in the parent frame, from the pose the forward pass kept
    const V3 anc = add3(S.lp[j], rotq(L.jps[j], S.lq[j])), ax = rotq(L.jax[j], S.lq[j]);
    g[j] = dot3(ax, sub3(Tp, cross3(anc, Fp)));
  }
  g[JM] = 0.f;
  if (free_wanted) {  // uniform
    float wf[6];  // the free joint's subtree is its body's subtree: that lane's wrench
#pragma unroll
    for (int i = 0; i < 6; i++) wf[i] = __shfl_sync(FULL, wr[i], free_e);
    const V3 Ff = mk3(wf[0], wf[1], wf[2]);
    const V3 Tf = sub3(mk3(wf[3], wf[4], wf[5]), cross3(sub3(S.fpos, c), Ff));
    float g4[4];
    quat_grad_left(S.fq, Tf, S.frinv, g4);
    float v = Ff.x;
    v = lane == 1 ? Ff.y : v; v = lane == 2 ? Ff.z : v; v = lane == 3 ? g4[0] : v;
    v = lane == 4 ? g4[1] : v; v = lane == 5 ? g4[2] : v; v = lane == 6 ? g4[3] : v;
    g[JM] = lane < 7 ? v : 0.f;
  }
}

// Solver-side lane state.  Slots without a coordinate carry lb = ub = 0 and a masked (zero) gradient, so every solver
// vector stays exactly zero there without a select.
template <int NS>
struct Slots {
  float lb[NS], ub[NS];
  bool valid[NS];
};

template <int NS>
__device__ __forceinline__ float lane_dot(const float (&a)[NS], const float (&b)[NS]) {
  float acc = a[0] * b[0];
#pragma unroll
  for (int m = 1; m < NS; m++) acc = fmaf(a[m], b[m], acc);
  return acc;
}

__device__ __forceinline__ void warp_sum2(float &a, float &b) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const float ta = __shfl_xor_sync(FULL, a, off), tb = __shfl_xor_sync(FULL, b, off);
    a = a + ta;
    b = b + tb;
  }
}

// FISTA momentum beta_k = (t_k - 1) / t_{k+1}, t_0 = 1, t_{k+1} = (1 + sqrt(1 + 4 t_k^2)) / 2 depends on the iteration number
// only: the first BT values are tabulated once per CTA in shared memory ([BT] = t_BT for the iterations beyond the table).
constexpr int BT = 512;
__device__ __forceinline__ void beta_table_init(float *tbl) {  // one thread; callers synchronise the CTA afterwards
  float t = 1.0f;
  for (int k = 0; k < BT; k++) {
    const float tn = 0.5f * (1.0f + sqrtf(fmaf(4.0f * t, t, 1.0f)));
    tbl[k] = (t - 1.0f) / tn;
    t = tn;
  }
  tbl[BT] = t;
}
// momentum of iteration `it`; `t` is only carried beyond the table
__device__ __forceinline__ float beta_of(const float *tbl, int it, float &t) {
  if (it < BT) return tbl[it];
  if (it == BT) t = tbl[BT];
  const float tn = 0.5f * (1.0f + sqrtf(fmaf(4.0f * t, t, 1.0f)));
  const float beta = (t - 1.0f) / tn;
  t = tn;
  return beta;
}

// Per-kernel constants every evaluation needs (registers).
struct Uni {
  int lane, free_e;
  bool has_free;
  float tol;
  int maxiter, maxls;
  const float *betas;  // shared-memory momentum table (beta_table_init)
};

template <int JM>
__device__ __forceinline__ bool free_wanted_of(const Uni &u, unsigned maskbits) {
  return u.has_free && __any_sync(FULL, u.lane < 7 && ((maskbits >> JM) & 1u));
}

// Per-solve view of the slots.  A coordinate the solve does not optimise ("frozen": valid, mask bit clear) has zero gradient: the
// reference's iteration moves it to clip(q0) once and make_qs discards it afterwards.  Here it simply stays at q0 -- unbounded
// effective box, gradient multiplied by gm = 0, so no select is needed anywhere in the iteration -- and the squared length of that
// one move, which enters the first line search, is returned (0 whenever q0 is inside the box).
template <int NS>
struct SolveC { float lb[NS], ub[NS], gm[NS]; };

template <int NS>
__device__ __forceinline__ float solve_setup(const Slots<NS> &co, const float (&q0)[NS], unsigned maskbits, SolveC<NS> &sc) {
  const float inf = __int_as_float(0x7f800000);
  float dn[NS];
#pragma unroll
  for (int m = 0; m < NS; m++) {
    const bool bit = (maskbits >> m) & 1u, frozen = co.valid[m] && !bit;
    sc.gm[m] = bit ? 1.0f : 0.0f;
    sc.lb[m] = frozen ? -inf : co.lb[m];
    sc.ub[m] = frozen ? inf : co.ub[m];
    dn[m] = frozen ? clipm(q0[m], co.lb[m], co.ub[m]) - q0[m] : 0.0f;
  }
  return warp_sum(lane_dot<NS>(dn, dn));
}

// ------------------------------------------------------------------------------------------
// jaxopt 0.8.5 ProjectedGradient.run (ProximalGradient._update_accel / _ls / _error with the box projection), one warp.
// Two-state machine so the kernel holds ONE forward and ONE reverse evaluation:
//   state Y : the point is the FISTA extrapolation y -> loss + gradient, then the first line-search candidate
//   state LS: the point is a candidate x+ -> loss; rejected: halve the step; accepted: gradient at x+ from the state of
//             this same evaluation (the reference recomputes FK there: same values), error, next y.
//   sqp: squared distance the projection moves the PASSIVE coordinates in the first iteration (they have zero gradient:
//        x+ = clip(q0) from then on; zero whenever q0 is inside the box); the frozen slots' share is added here (solve_setup).
// ------------------------------------------------------------------------------------------
template <int JM, int RT>
__device__ __forceinline__ SolveOut solve(const LaneC<JM, RT> &L, const SiteC &st, const Uni &u, const Slots<JM + 1> &co, const float (&q0)[JM + 1],
                                          unsigned maskbits, float sqp, float (&x)[JM + 1]) {
  constexpr int NS = JM + 1;
  float y[NS], g[NS], xn[NS], d[NS], gt[NS];
#pragma unroll
  for (int m = 0; m < NS; m++) { x[m] = q0[m]; y[m] = x[m]; xn[m] = x[m]; g[m] = 0.f; gt[m] = 0.f; }
  float t = 1.0f, step = 1.0f, stp = 1.0f, fy = 0.f, sq = 0.f, dg = 0.f;
  int halv = 0;
  bool in_ls = false;
  SolveOut out;
  out.iters = 0; out.ls = 0; out.bad = false; out.err = __int_as_float(0x7f800000);
  if (u.maxiter <= 0) return out;
  const bool fw = free_wanted_of<JM>(u, maskbits);
  SolveC<NS> sc;
  sqp = solve_setup<NS>(co, q0, maskbits, sc) + sqp;
  Fwd<JM> S;
  for (;;) {
    float pt[NS];
#pragma unroll
    for (int m = 0; m < NS; m++) pt[m] = in_ls ? xn[m] : y[m];
    const float f = eval_fwd<JM, RT>(L, st, u.has_free, pt, S);
    bool rejected = false;
    if (in_ls) {
      out.ls++;
      if (!(f - f == 0.0f)) out.bad = true;
      const float dec = stp * (f - fy);
      const float cond = fmaf(stp, dg, 0.5f * sq);
      rejected = (dec > cond + 1.1920929e-07f) && (halv < u.maxls);
    }
    if (!rejected) {
      eval_bwd<JM, RT>(L, S, u.lane, fw, u.free_e, gt);
#pragma unroll
      for (int m = 0; m < NS; m++) gt[m] = gt[m] * sc.gm[m];
      if (in_ls) {  // accepted x+ = xn
        step = (stp <= 1e-6f) ? 1.0f : stp / 0.5f;
        const float beta = beta_of(u.betas, out.iters, t);
#pragma unroll
        for (int m = 0; m < NS; m++) {
          y[m] = fmaf(beta, xn[m] - x[m], xn[m]);
          d[m] = clipm(xn[m] - gt[m], sc.lb[m], sc.ub[m]) - xn[m];
          x[m] = xn[m];
        }
        out.err = sqrtf(warp_sum(lane_dot<NS>(d, d)));
        out.iters++;
        sqp = 0.f;
        if (!(out.err > u.tol && out.iters < u.maxiter)) break;
        in_ls = false;
        continue;
      }
      fy = f;
#pragma unroll
      for (int m = 0; m < NS; m++) g[m] = gt[m];
      stp = step;
      halv = 0;
      in_ls = true;
    } else {
      stp = stp * 0.5f;
      halv++;
    }
#pragma unroll
    for (int m = 0; m < NS; m++) {
      xn[m] = clipm(fmaf(-stp, g[m], y[m]), sc.lb[m], sc.ub[m]);
      d[m] = xn[m] - y[m];
    }
    sq = lane_dot<NS>(d, d);
    dg = lane_dot<NS>(d, g);
    warp_sum2(sq, dg);
    sq = sq + sqp;
  }
  return out;
}

// ------------------------------------------------------------------------------------------
// Latency mode: 2 * NC warps cooperate on ONE chain.  Per FISTA iteration warp j < NC evaluates the line-search candidate
// x+_j (step s / 2^j) and the gradient there (stopping criterion), warp NC + j evaluates the extrapolation y'(x+_j) and the
// gradient there (next iteration); ONE block barrier per round, results exchanged through a double-buffered shared area.
// The accepted candidate is the first one the sequential line search would accept and every evaluation is the same
// arithmetic as in `solve`, so results are bit-identical to the one-warp solver.
// ------------------------------------------------------------------------------------------
template <int NS, int NC>
struct Xchg {  // shared memory, [parity] double-buffered
  float g[2][NC][32 * NS];  // Y warp j: grad f(y'_j)
  float e[2][NC][32];       // X warp j: per-lane partial of |clip(x+_j - grad f(x+_j)) - x+_j|^2 (stopping criterion)
  float fx[2][NC], fy[2][NC], sq[2][NC], dg[2][NC];  // f(x+_j) (X warp), f(y'_j) and the two line-search sums of candidate j (Y warp)
};

// The work of a round is split so that both roles carry about the same instruction count (the barrier waits for the slowest):
//   X warp j: forward + reverse at x+_j, per-lane residual partials;
//   Y warp j: the line-search sums |x+_j - y|^2, <x+_j - y, g> of candidate j, forward + reverse at y'(x+_j).
// After the barrier every warp evaluates the accept conditions of all candidates from the exchanged scalars (same operations as the
// sequential search).  The stopping criterion of an accepted iteration is reduced (butterfly + sqrt) while the NEXT round is being
// evaluated and tested just before that round's barrier: when it fires, the round in flight is discarded -- nothing it computed has
// touched the solver state yet -- so the result is the sequential solver's, one speculative round per converged solve later.
template <int JM, int RT, int NC>
__device__ __forceinline__ SolveOut solve_coop(const LaneC<JM, RT> &L, const SiteC &st, const Uni &u, const Slots<JM + 1> &co, const float (&q0)[JM + 1],
                                               unsigned maskbits, float sqp, float (&x)[JM + 1], Xchg<JM + 1, NC> *xc, int w, int &par) {
  constexpr int NS = JM + 1;
  float y[NS], g[NS], gt[NS], xj[NS], pt[NS], d[NS];
#pragma unroll
  for (int m = 0; m < NS; m++) { x[m] = q0[m]; y[m] = x[m]; g[m] = 0.f; }
  SolveOut out;
  out.iters = 0; out.ls = 0; out.bad = false; out.err = __int_as_float(0x7f800000);
  if (u.maxiter <= 0) return out;
  const bool fw = free_wanted_of<JM>(u, maskbits);
  const int j = w % NC;
  const bool isY = w >= NC;
  SolveC<NS> sc;
  sqp = solve_setup<NS>(co, q0, maskbits, sc) + sqp;
  Fwd<JM> S;
  // f(y0), grad f(y0): every warp computes them (identical values)
  float fy = eval_fwd<JM, RT>(L, st, u.has_free, y, S);
  eval_bwd<JM, RT>(L, S, u.lane, fw, u.free_e, g);
#pragma unroll
  for (int m = 0; m < NS; m++) g[m] = g[m] * sc.gm[m];
  float t = 1.0f, stp = 1.0f;
  float beta = beta_of(u.betas, 0, t);
  int base = 0;
  bool pending = false;  // the last accepted iteration's stopping criterion has not been tested yet
  float e_prev = 0.f;
  for (;;) {
    float sj = stp;
#pragma unroll
    for (int i = 0; i < NC - 1; i++) sj = (i < j) ? sj * 0.5f : sj;
#pragma unroll
    for (int m = 0; m < NS; m++) {
      xj[m] = clipm(fmaf(-sj, g[m], y[m]), sc.lb[m], sc.ub[m]);
      d[m] = xj[m] - y[m];
      pt[m] = isY ? fmaf(beta, xj[m] - x[m], xj[m]) : xj[m];
    }
    if (isY) {
      float sq = lane_dot<NS>(d, d), dg = lane_dot<NS>(d, g);
      warp_sum2(sq, dg);
      sq = sq + sqp;
      if (u.lane == 0) { xc->sq[par][j] = sq; xc->dg[par][j] = dg; }
    }
    const float f = eval_fwd<JM, RT>(L, st, u.has_free, pt, S);
    eval_bwd<JM, RT>(L, S, u.lane, fw, u.free_e, gt);
#pragma unroll
    for (int m = 0; m < NS; m++) gt[m] = gt[m] * sc.gm[m];
    if (!isY) {
#pragma unroll
      for (int m = 0; m < NS; m++) d[m] = clipm(xj[m] - gt[m], sc.lb[m], sc.ub[m]) - xj[m];
      xc->e[par][j][u.lane] = lane_dot<NS>(d, d);
      if (u.lane == 0) xc->fx[par][j] = f;
    } else {
#pragma unroll
      for (int m = 0; m < NS; m++) xc->g[par][j][32 * m + u.lane] = gt[m];
      if (u.lane == 0) xc->fy[par][j] = f;
    }
    if (pending) {  // uniform over the CTA: every warp reduces the same partials in the same order
      out.err = sqrtf(warp_sum(e_prev));
      pending = false;
      if (!(out.err > u.tol)) break;  // converged one round ago: the round in flight is dropped
    }
    __syncthreads();
    const int rp = par;
    par ^= 1;
    int k = -1;
    bool nonfinite = false;
#pragma unroll
    for (int i = NC - 1; i >= 0; i--) {  // accept conditions of the sequential line search, candidate i = step / 2^i
      float si = stp;
#pragma unroll
      for (int h = 0; h < NC - 1; h++) si = (h < i) ? si * 0.5f : si;
      const float fi = xc->fx[rp][i];
      const float dec = si * (fi - fy);
      const float cond = fmaf(si, xc->dg[rp][i], 0.5f * xc->sq[rp][i]);
      const bool rejected = (dec > cond + 1.1920929e-07f) && (base + i < u.maxls);
      k = rejected ? k : i;
      nonfinite = (!rejected) ? !(fi - fi == 0.0f) : (nonfinite || !(fi - fi == 0.0f));  // candidates 0..k (all when none is accepted)
    }
    if (nonfinite) out.bad = true;
    if (k < 0) {  // every candidate rejected: next NC step sizes
      base += NC;
#pragma unroll
      for (int i = 0; i < NC; i++) stp = stp * 0.5f;
      continue;
    }
    out.ls += base + k + 1;  // evaluations the sequential line search performs
    float sk = stp;
#pragma unroll
    for (int i = 0; i < NC - 1; i++) sk = (i < k) ? sk * 0.5f : sk;
#pragma unroll
    for (int m = 0; m < NS; m++) {
      const float xk = clipm(fmaf(-sk, g[m], y[m]), sc.lb[m], sc.ub[m]);
      y[m] = fmaf(beta, xk - x[m], xk);
      x[m] = xk;
    }
#pragma unroll
    for (int m = 0; m < NS; m++) g[m] = xc->g[rp][k][32 * m + u.lane];
    fy = xc->fy[rp][k];
    e_prev = xc->e[rp][k][u.lane];
    stp = (sk <= 1e-6f) ? 1.0f : sk / 0.5f;
    out.iters++;
    sqp = 0.f;
    base = 0;
    if (!(out.iters < u.maxiter)) {  // iteration cap: no further round, reduce the criterion now
      out.err = sqrtf(warp_sum(e_prev));
      break;
    }
    beta = beta_of(u.betas, out.iters, t);
    pending = true;  // (a run-time choice between testing now and one round late was measured: it costs more than the round it saves)
  }
  return out;
}


// ------------------------------------------------------------------------------------------
// Pair mode: TWO warps cooperate on one chain -- for chain counts between the latency regime (few chains: 2 * NC warps each) and
// the throughput regime (one warp each).  Per FISTA iteration
//   phase A: warp 0 / warp 1 evaluate the line-search candidates x+_a (step s) and x+_b (step s / 2) at the same time (forward
//            only); none accepted: next two step sizes;
//   phase B: the warp that evaluated the accepted candidate turns its forward state into grad f(x+) (stopping criterion) while
//            its partner evaluates the extrapolation y' and grad f(y') for the next iteration.
// Critical path 2 forward + 1 reverse evaluations instead of 3 + 2, at the instruction budget of the one-warp solver (no
// speculation on y').  Same arithmetic, same accepted candidate: bit-identical to `solve`.
// ------------------------------------------------------------------------------------------
template <int NS>
struct XchgPair {  // shared memory
  float g[32 * NS];
  float acc[2][2], nf[2][2];  // [parity][warp]: accepted flag, non-finite loss flag of the two candidates
  float err, fy;
};

template <int JM, int RT>
__device__ __forceinline__ SolveOut solve_pair(const LaneC<JM, RT> &L, const SiteC &st, const Uni &u, const Slots<JM + 1> &co, const float (&q0)[JM + 1],
                                               unsigned maskbits, float sqp, float (&x)[JM + 1], XchgPair<JM + 1> *xc, int w, int &par) {
  constexpr int NS = JM + 1;
  float y[NS], g[NS], gt[NS], xj[NS], d[NS];
#pragma unroll
  for (int m = 0; m < NS; m++) { x[m] = q0[m]; y[m] = x[m]; g[m] = 0.f; }
  SolveOut out;
  out.iters = 0; out.ls = 0; out.bad = false; out.err = __int_as_float(0x7f800000);
  if (u.maxiter <= 0) return out;
  const bool fw = free_wanted_of<JM>(u, maskbits);
  SolveC<NS> sc;
  sqp = solve_setup<NS>(co, q0, maskbits, sc) + sqp;
  Fwd<JM> S;
  float fy = eval_fwd<JM, RT>(L, st, u.has_free, y, S);  // f(y0), grad f(y0): both warps (identical values)
  eval_bwd<JM, RT>(L, S, u.lane, fw, u.free_e, g);
#pragma unroll
  for (int m = 0; m < NS; m++) g[m] = g[m] * sc.gm[m];
  float t = 1.0f, stp = 1.0f;
  int base = 0;
  for (;;) {
    // phase A: this warp's line-search candidate
    const float sj = w ? stp * 0.5f : stp;
    float sq, dg;
#pragma unroll
    for (int m = 0; m < NS; m++) {
      xj[m] = clipm(fmaf(-sj, g[m], y[m]), sc.lb[m], sc.ub[m]);
      d[m] = xj[m] - y[m];
    }
    sq = lane_dot<NS>(d, d);
    dg = lane_dot<NS>(d, g);
    warp_sum2(sq, dg);
    sq = sq + sqp;
    const float f = eval_fwd<JM, RT>(L, st, u.has_free, xj, S);
    const float dec = sj * (f - fy);
    const float cond = fmaf(sj, dg, 0.5f * sq);
    const bool rejected = (dec > cond + 1.1920929e-07f) && (base + w < u.maxls);
    if (u.lane == 0) {
      xc->acc[par][w] = rejected ? 0.f : 1.f;
      xc->nf[par][w] = (f - f == 0.0f) ? 0.f : 1.f;
    }
    __syncthreads();
    const bool a0 = xc->acc[par][0] != 0.f, a1 = xc->acc[par][1] != 0.f;
    const int k = a0 ? 0 : (a1 ? 1 : -1);
    if (xc->nf[par][0] != 0.f || (k != 0 && xc->nf[par][1] != 0.f)) out.bad = true;  // candidates the sequential search evaluates
    par ^= 1;
    if (k < 0) {  // both rejected: next two step sizes
      base += 2;
      stp = stp * 0.25f;
      continue;
    }
    out.ls += base + k + 1;
    const float sk = k ? stp * 0.5f : stp;
    const float beta = beta_of(u.betas, out.iters, t);
    // phase B
    if (w == k) {  // gradient at the accepted x+ (this warp's forward state) -> unit-step fixed-point residual
      eval_bwd<JM, RT>(L, S, u.lane, fw, u.free_e, gt);
#pragma unroll
      for (int m = 0; m < NS; m++) {
        gt[m] = gt[m] * sc.gm[m];
        d[m] = clipm(xj[m] - gt[m], sc.lb[m], sc.ub[m]) - xj[m];
      }
      const float err = sqrtf(warp_sum(lane_dot<NS>(d, d)));
      if (u.lane == 0) xc->err = err;
#pragma unroll
      for (int m = 0; m < NS; m++) { y[m] = fmaf(beta, xj[m] - x[m], xj[m]); x[m] = xj[m]; }
    } else {  // the partner: x+_k, the extrapolation y', f(y') and grad f(y')
#pragma unroll
      for (int m = 0; m < NS; m++) {
        const float xk = clipm(fmaf(-sk, g[m], y[m]), sc.lb[m], sc.ub[m]);
        y[m] = fmaf(beta, xk - x[m], xk);
        x[m] = xk;
      }
      const float fyn = eval_fwd<JM, RT>(L, st, u.has_free, y, S);
      eval_bwd<JM, RT>(L, S, u.lane, fw, u.free_e, gt);
#pragma unroll
      for (int m = 0; m < NS; m++) xc->g[32 * m + u.lane] = gt[m] * sc.gm[m];
      if (u.lane == 0) xc->fy = fyn;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NS; m++) g[m] = xc->g[32 * m + u.lane];
    fy = xc->fy;
    out.err = xc->err;
    stp = (sk <= 1e-6f) ? 1.0f : sk / 0.5f;
    out.iters++;
    sqp = 0.f;
    base = 0;
    // no third barrier: the next writes of g / fy / err come after the NEXT phase-A barrier, which a warp reaches only once it has
    // read this iteration's values; the accept flags alternate between two parities
    if (!(out.err > u.tol && out.iters < u.maxiter)) break;
  }
  return out;
}

// ------------------------------------------------------------------------------------------
// solver slots <-> qpos addresses
// ------------------------------------------------------------------------------------------
template <int JM>
struct SlotAdr { int adr[JM + 1]; };

template <int JM, int RT>
__device__ __forceinline__ void slots_init(Slots<JM + 1> &co, SlotAdr<JM> &sa, const LaneC<JM, RT> &L, const DevTree &T, int lane,
                                           const float *__restrict__ lb, const float *__restrict__ ub) {
#pragma unroll
  for (int j = 0; j < JM; j++) { co.valid[j] = L.hinge[j]; sa.adr[j] = L.adr[j]; }
  co.valid[JM] = T.fs_free_e >= 0 && lane < 7;
  sa.adr[JM] = co.valid[JM] ? T.free_adr + lane : 0;
#pragma unroll
  for (int m = 0; m <= JM; m++) {
    co.lb[m] = (co.valid[m] && lb) ? lb[sa.adr[m]] : 0.f;
    co.ub[m] = (co.valid[m] && ub) ? ub[sa.adr[m]] : 0.f;
  }
}

// which qpos addresses a solve optimises: an explicit u8 mask, or every address below `lim`
struct MaskSpec { const uint8_t *m; int lim; };
__device__ __forceinline__ bool mask_has(const MaskSpec &ms, int adr) { return ms.m ? ms.m[adr] != 0 : adr < ms.lim; }

template <int JM>
__device__ __forceinline__ unsigned slot_bits(const Slots<JM + 1> &co, const SlotAdr<JM> &sa, const MaskSpec &ms) {
  unsigned b = 0;
#pragma unroll
  for (int m = 0; m <= JM; m++)
    if (co.valid[m] && mask_has(ms, sa.adr[m])) b |= 1u << m;
  return b;
}

template <int JM>
__device__ __forceinline__ void slots_gather(const Slots<JM + 1> &co, const SlotAdr<JM> &sa, const float *src, float (&q)[JM + 1]) {
#pragma unroll
  for (int m = 0; m <= JM; m++) q[m] = co.valid[m] ? src[sa.adr[m]] : 0.f;
}
template <int JM>
__device__ __forceinline__ void slots_scatter(const Slots<JM + 1> &co, const SlotAdr<JM> &sa, const float (&q)[JM + 1], float *dst) {
#pragma unroll
  for (int m = 0; m <= JM; m++)
    if (co.valid[m]) dst[sa.adr[m]] = q[m];
}

// Passive coordinates (hinges outside the active subtree): zero gradient, so a solve moves them to clip(q0) in its first
// iteration and never again.  Squared length of that move (enters the first line search), lanes deal the list out.
__device__ __forceinline__ float passive_sq(const DevTree &T, int lane, const float *q, const float *__restrict__ lb, const float *__restrict__ ub) {
  float acc = 0.f;
  bool first = true;
  for (int i = lane; i < T.npassive; i += 32) {
    const int p = __ldg(T.passive + i);
    const float v = q[p];
    const float d = clipm(v, lb[p], ub[p]) - v;
    acc = first ? d * d : fmaf(d, d, acc);
    first = false;
  }
  return warp_sum(acc);
}

// replace_qs as far as the free joint's quaternion is concerned (MJX kinematics normalises it in qpos), in the slot layout
template <int JM>
__device__ __forceinline__ void normalize_free(const Uni &u, float (&q)[JM + 1]) {
  if (!u.has_free) return;
  float fr[4];
#pragma unroll
  for (int i = 0; i < 4; i++) fr[i] = __shfl_sync(FULL, q[JM], 3 + i);
  float d;
  const Q4 qn = normalize4(mk4(fr[0], fr[1], fr[2], fr[3]), &d);
  float v = q[JM];
  v = u.lane == 3 ? qn.w : v; v = u.lane == 4 ? qn.x : v; v = u.lane == 5 ? qn.y : v; v = u.lane == 6 ? qn.z : v;
  q[JM] = v;
}

// Full-model FK of the qpos vector in ch.qbuf (normalised in place) and the per-frame outputs: the general cold path.
template <int NBF>
__device__ __forceinline__ void outputs_from_qbuf(const Chain &ch, const SiteC &st, const float *__restrict__ site_pos, float *qpos_o, float *xpos_o,
                                                  float *xquat_o, float *sites_o) {
  V3 P[NBF];
  Q4 Q[NBF];
  fk_cold<NBF>(ch, ch.T.full, P, Q);
  if (qpos_o)
    for (int i = ch.lane; i < ch.T.nq; i += 32) qpos_o[i] = ch.qbuf[i];
  const DevSet &S = ch.T.full;
#pragma unroll
  for (int i = 0; i < NBF; i++) {
    const int e = ch.lane + 32 * i;
    if (e < S.n) {
      const int b = __ldg(S.rec + (size_t)e * REC + R_BODY);
      if (xpos_o) { xpos_o[3 * b] = P[i].x; xpos_o[3 * b + 1] = P[i].y; xpos_o[3 * b + 2] = P[i].z; }
      if (xquat_o) { xquat_o[4 * b] = Q[i].w; xquat_o[4 * b + 1] = Q[i].x; xquat_o[4 * b + 2] = Q[i].y; xquat_o[4 * b + 3] = Q[i].z; }
    }
  }
  if (ch.lane == 0) {
    if (xpos_o) { xpos_o[0] = 0.f; xpos_o[1] = 0.f; xpos_o[2] = 0.f; }
    if (xquat_o) { xquat_o[0] = 1.f; xquat_o[1] = 0.f; xquat_o[2] = 0.f; xquat_o[3] = 0.f; }
  }
  if (sites_o && st.k >= 0) {  // the site's own body and its original offset (st.off may be expressed in an element's frame)
    const float *o = ch.PQ + 7 * st.ef;
    const V3 off = mk3(__ldg(site_pos + 3 * st.k), __ldg(site_pos + 3 * st.k + 1), __ldg(site_pos + 3 * st.k + 2));
    const V3 s = add3(lds3(o), rotate(off, lds4(o + 3)));
    sites_o[3 * st.k] = s.x; sites_o[3 * st.k + 1] = s.y; sites_o[3 * st.k + 2] = s.z;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

// NC == 0: throughput mode, one warp per chain, four chains per CTA (MINB CTAs per SM requested from the compiler);
// NC == 1: pair mode, two warps per chain, one chain per 64-thread CTA (solve_pair);
// NC >= 2: latency mode, 2 * NC warps cooperate on one chain (solve_coop).
template <int JM, int RT, int NBF, int NC, int MINB>
__global__ void __launch_bounds__(NC ? 64 * NC : 128, MINB) fast_pose_kernel(DevTree T, PoseArgs a) {
  constexpr bool COOP = NC > 0;
  constexpr int NS = JM + 1;
  extern __shared__ float smem[];
  __shared__ int s_chain;
  __shared__ float s_beta[BT + 1];
  if (threadIdx.x == 0) beta_table_init(s_beta);
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int area = 2 * T.nqp + 7 * T.pqn;  // qbuf [nqp], (unused) [nqp], PQ [7 pqn]
  Chain ch(T, smem + (size_t)(COOP ? 0 : wib) * area, nullptr, lane, 0, 1, 0);
  Xchg<NS, (NC > 0 ? NC : 1)> *xc = reinterpret_cast<Xchg<NS, (NC > 0 ? NC : 1)> *>(smem + area);
  XchgPair<NS> *xp = reinterpret_cast<XchgPair<NS> *>(smem + area);
  const bool writer = !COOP || wib == 0;
  LaneC<JM, RT> L;
  lane_init<JM, RT>(L, T, lane);
  SiteC st;
  site_init(st, T, lane, a.site_pos);
  Slots<NS> co;
  SlotAdr<JM> sa;
  slots_init<JM, RT>(co, sa, L, T, lane, a.lb, a.ub);
  Uni u;
  u.lane = lane; u.free_e = T.fs_free_e >= 0 ? T.fs_free_e : 0; u.has_free = T.fs_free_e >= 0;
  u.tol = a.tol; u.maxiter = a.maxiter; u.maxls = a.maxls; u.betas = s_beta;
  const int nq = T.nq, K = T.K, nb = T.nbody, S1 = 1 + a.P, npassive = T.npassive;
  const MaskSpec full_ms = {nullptr, nq}, root_ms = {nullptr, a.root_dims};
  const unsigned full_bits = slot_bits<JM>(co, sa, full_ms), root_bits = slot_bits<JM>(co, sa, root_ms);
  const int n_root = a.do_root ? 2 : 0;           // two root solves on frame 0 (compute_stac.py:64-98)
  const int n_pose = (a.do_root == 2) ? 0 : a.F;  // do_root == 2: root optimisation only
  const int n_stage = n_root + n_pose * S1;
  int par = 0;

  for (;;) {
    int c = 0;
    if (COOP) {
      if (threadIdx.x == 0) s_chain = atomicAdd(a.counter, 1);
      __syncthreads();
      c = s_chain;
    } else {
      if (lane == 0) c = atomicAdd(a.counter, 1);
      c = __shfl_sync(FULL, c, 0);
    }
    if (c >= a.C) break;
    if (writer)
      for (int i = lane; i < nq; i += 32) ch.qbuf[i] = a.qpos_io[(size_t)c * nq + i];
    if (COOP) __syncthreads(); else __syncwarp();
    float q[NS], q0[NS], x[NS];
    slots_gather<JM>(co, sa, ch.qbuf, q);
    bool bad = false;
    const float *kpc = a.kp + (size_t)c * a.kp_stride;
    for (int sidx = 0; sidx < n_stage; sidx++) {
      const bool is_root = sidx < n_root;
      const int f = is_root ? 0 : (sidx - n_root) / S1;   // frame
      const int sg = is_root ? 0 : (sidx - n_root) % S1;  // 0: whole body, 1..P: INDIVIDUAL_PART_OPTIMIZATION masks
      unsigned bits;
      MaskSpec ms;
      if (is_root) {
        if (sidx == 0) { site_load_kp(st, kpc); site_mask_kp(st, a.trunk_kps); }
        bits = root_bits;
        ms = root_ms;
      } else {
        if (sg == 0) { site_load_kp(st, kpc + (size_t)f * 3 * K); if (f == 0) site_mask_kp(st, nullptr); }
        ms.m = (sg == 0) ? nullptr : a.part_masks + (size_t)(sg - 1) * nq;
        ms.lim = nq;
        bits = (sg == 0) ? full_bits : slot_bits<JM>(co, sa, ms);
      }
#pragma unroll
      for (int m = 0; m < NS; m++) {
        q0[m] = q[m];
        if (is_root && co.valid[m] && sa.adr[m] < 3) q0[m] = kpc[3 * a.root_kp_idx + sa.adr[m]];  // re-seed the root translation
      }
      const float sqp = npassive ? passive_sq(T, lane, ch.qbuf, a.lb, a.ub) : 0.f;
      SolveOut so;
      if constexpr (NC == 1) so = solve_pair<JM, RT>(L, st, u, co, q0, bits, sqp, x, xp, wib, par);
      else if constexpr (COOP) so = solve_coop<JM, RT, NC>(L, st, u, co, q0, bits, sqp, x, xc, wib, par);
      else so = solve<JM, RT>(L, st, u, co, q0, bits, sqp, x);
#pragma unroll
      for (int m = 0; m < NS; m++) q[m] = ((bits >> m) & 1u) ? x[m] : q0[m];  // utils.make_qs
      bad |= so.bad;
      if (npassive && u.maxiter > 0) {  // the passive coordinates the solve covers end at clip(q0)
        if (writer)
          for (int i = lane; i < npassive; i += 32) {
            const int p = __ldg(T.passive + i);
            if (mask_has(ms, p)) ch.qbuf[p] = clipm(ch.qbuf[p], a.lb[p], a.ub[p]);
          }
        if (COOP) __syncthreads(); else __syncwarp();
      }
      if (is_root) {
        if (a.root_stats && lane == 0 && writer) { a.root_stats[4 * c + 2 * sidx] = so.iters; a.root_stats[4 * c + 2 * sidx + 1] = so.ls; }
      } else {
        const size_t fi = (size_t)c * a.F + f;
        if (a.iters && lane == 0 && writer) { a.iters[fi * S1 + sg] = so.iters; a.ls_evals[fi * S1 + sg] = so.ls; }
        if (sg == a.P && writer) {  // last solve of the frame: full-model FK of the raw solution -> outputs
          slots_scatter<JM>(co, sa, q, ch.qbuf);
          __syncwarp();
          outputs_from_qbuf<NBF>(ch, st, a.site_pos, a.qpos ? a.qpos + fi * nq : nullptr, a.xpos ? a.xpos + fi * nb * 3 : nullptr,
                                 a.xquat ? a.xquat + fi * nb * 4 : nullptr, a.sites ? a.sites + fi * K * 3 : nullptr);
          if (a.err && lane == 0) a.err[fi] = so.err;
        }
      }
      normalize_free<JM>(u, q);  // replace_qs: kinematics normalises the quaternion (same arithmetic as the cold FK above)
    }
    if (writer) {
      slots_scatter<JM>(co, sa, q, ch.qbuf);
      __syncwarp();
      for (int i = lane; i < nq; i += 32) a.qpos_io[(size_t)c * nq + i] = ch.qbuf[i];
      if (a.status && lane == 0) a.status[c] = bad ? 1 : 0;
    }
    if (COOP) __syncthreads(); else __syncwarp();
  }
}

// B independent items: q_loss + gradient (mode 1) or one FISTA solve (mode 2); one warp per item.
template <int JM, int RT>
__global__ void __launch_bounds__(128) fast_batch_kernel(DevTree T, BatchArgs a) {
  constexpr int NS = JM + 1;
  __shared__ float s_beta[BT + 1];
  if (a.mode == 2) {
    if (threadIdx.x == 0) beta_table_init(s_beta);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  LaneC<JM, RT> L;
  lane_init<JM, RT>(L, T, lane);
  SiteC st;
  site_init(st, T, lane, a.site_pos);
  site_mask_u8(st, a.kp_mask);
  Slots<NS> co;
  SlotAdr<JM> sa;
  slots_init<JM, RT>(co, sa, L, T, lane, a.lb, a.ub);
  Uni u;
  u.lane = lane; u.free_e = T.fs_free_e >= 0 ? T.fs_free_e : 0; u.has_free = T.fs_free_e >= 0;
  u.tol = a.tol; u.maxiter = a.maxiter; u.maxls = a.maxls; u.betas = s_beta;
  const int nq = T.nq, K = T.K;
  const MaskSpec ms = {a.q_mask, nq};
  const unsigned bits = slot_bits<JM>(co, sa, ms);
  for (int b = blockIdx.x * wpb + wib; b < a.B; b += gridDim.x * wpb) {
    float q[NS], q0[NS];
    const float *qb = a.q + (size_t)b * nq;
    slots_gather<JM>(co, sa, qb, q);
    if (a.q0) slots_gather<JM>(co, sa, a.q0 + (size_t)b * nq, q0);
    else {
#pragma unroll
      for (int m = 0; m < NS; m++) q0[m] = q[m];
    }
    site_load_kp(st, a.kp + (size_t)b * 3 * K);
    if (a.mode == 1) {
      float pt[NS], g[NS];
#pragma unroll
      for (int m = 0; m < NS; m++) pt[m] = ((bits >> m) & 1u) ? q[m] : q0[m];
      Fwd<JM> S;
      const float loss = eval_fwd<JM, RT>(L, st, u.has_free, pt, S);
      if (lane == 0) a.out_a[b] = loss;
      if (a.out_b) {
        eval_bwd<JM, RT>(L, S, lane, free_wanted_of<JM>(u, bits), u.free_e, g);
        float *go = a.out_b + (size_t)b * nq;
        for (int i = lane; i < nq; i += 32) go[i] = 0.f;
        __syncwarp();
#pragma unroll
        for (int m = 0; m < NS; m++)
          if (co.valid[m] && ((bits >> m) & 1u)) go[sa.adr[m]] = g[m];
      }
    } else {
      float x[NS];
      const float sqp = T.npassive ? passive_sq(T, lane, qb, a.lb, a.ub) : 0.f;
      const SolveOut so = solve<JM, RT>(L, st, u, co, q, bits, sqp, x);
      if (u.maxiter > 0) {  // res.params of the coordinates the solve does not optimise: the reference's iterate sits at clip(q0)
#pragma unroll
        for (int m = 0; m < NS; m++)
          if (co.valid[m] && !((bits >> m) & 1u)) x[m] = clipm(x[m], co.lb[m], co.ub[m]);
      }
      float *po = a.out_a + (size_t)b * nq;
      for (int i = lane; i < T.npassive; i += 32) {
        const int p = __ldg(T.passive + i);
        po[p] = u.maxiter > 0 ? clipm(qb[p], a.lb[p], a.ub[p]) : qb[p];
      }
      slots_scatter<JM>(co, sa, x, po);
      if (lane == 0) { a.out_b[b] = so.err; a.iters[b] = so.iters; a.ls_evals[b] = so.ls; }
    }
    __syncwarp();
  }
}

}  // namespace fast
}  // namespace stacb
