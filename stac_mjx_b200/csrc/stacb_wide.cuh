// stacb_wide.cuh -- the register-resident solver on W = 2..8 warps per evaluation (sm_100a), for hinge trees whose element set
// (stacb_fast.cuh) does not fit one warp: the mouse model of the reference (181 jointed active bodies, depth 85, 34 keypoints).
//
// Same mapping and the same arithmetic as stacb_fast.cuh with NL = 32 W lanes -- thread t <-> element t, thread NL - 1 the identity
// element, thread p + 1 <-> marker site p, one hinge per element (JM = 1) in solver slot 0, the free joint in slot 1 of threads 0..6 --
// except that what a single warp exchanges by shuffles goes through shared memory here:
//   * pointer jumping: double-buffered pose arrays, one CTA barrier per round (exact round count, no surplus rounds), position and
//     quaternion together (P <- P_a + R(Q_a) P, Q <- Q_a Q) so that a round costs ONE barrier;
//   * sums over the lanes: the 32-lane butterfly inside every warp, then the warp sums added in warp order;
//   * wrench prefix: Hillis-Steele inside every warp, then warp w >= 1 adds the totals of the warps before it (summed in warp order);
// The CPU oracle's mode 2 restates exactly this order (W = 1 reduces to stacb_fast.cuh).  One chain per CTA, frames sequential:
// replaces reference stac_mjx/stac_core.py:27-99 for wide trees.  G = 1: one group of W warps runs the sequential line search of
// `fast::solve`; G = 2 (few chains): two groups of W warps each share one chain the way the two warps of `fast::solve_pair` do --
// both line-search candidates at once, then grad f(x+) and the next extrapolation at once -- every evaluation inside its own group
// (named barrier 1 + group, own WX area), the groups meeting at CTA barriers only to exchange results.  Bit-identical either way.
#pragma once
#include "stacb_fast.cuh"

namespace stacb {
namespace wide {

using namespace fast;

constexpr int WRT = 8;  // pointer-jumping rounds a lane can hold (tree depth <= 256); only `rounds` of them are executed
constexpr int NS = 2;   // solver slots per thread: the element's hinge, a free-joint coordinate (threads 0..6)

template <int W>
struct WX {  // shared memory of one chain (16-byte aligned: poses and prefixes move as 128-bit shared-memory accesses)
  float4 q[2][32 * W];     // pointer jumping of the world quaternions (double buffer; w, x, y, z); q[qb] holds them after the forward pass
  float4 v[2][32 * W];     // ... of the world positions (x, y, z, unused); v[vb]
  float4 wpre[32 * W][2];  // inclusive wrench prefix by lane (six values, two unused)
  float4 red[2][3][2];     // warp partials of the sums over all lanes ([parity][value][warp / 4].{x,y,z,w} = warps 0..7)
  float4 wsum[W][2];       // warp totals of the wrench scan (six values, two unused)
};
__device__ __forceinline__ Q4 q4_of(const float4 a) { return mk4(a.x, a.y, a.z, a.w); }
__device__ __forceinline__ V3 v3_of(const float4 a) { return mk3(a.x, a.y, a.z); }

struct UniW {
  int t, lane, warp, grp, rounds, free_sa, free_se;  // t / warp: thread and warp index inside the group
  bool has_free;
  float tol;
  int maxiter, maxls;
  const float *betas;
};

// barrier of the W warps that share one evaluation: the CTA barrier for G = 1, named barrier 1 + group for G = 2
template <int W, int G>
__device__ __forceinline__ void gsync(const UniW &u) {
  if constexpr (G == 1) __syncthreads();
  else if (u.grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * W) : "memory");  // (immediate ids: only three barriers are reserved)
  else asm volatile("bar.sync 2, %0;" ::"n"(32 * W) : "memory");
}

// sums over all lanes of up to three values at once: one barrier
template <int W, int NV, int G = 1>
__device__ __forceinline__ void cta_sum(float (&val)[NV], const UniW &u, WX<W> *X, int &par) {
#pragma unroll
  for (int i = 0; i < NV; i++) val[i] = warp_sum(val[i]);
  if (u.lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) reinterpret_cast<float *>(&X->red[par][i][0])[u.warp] = val[i];
  }
  gsync<W, G>(u);
#pragma unroll
  for (int i = 0; i < NV; i++) {  // the warp partials added in warp order
    const float4 lo = X->red[par][i][0];
    float tot = lo.x + lo.y;
    if constexpr (W > 2) tot = (tot + lo.z) + lo.w;
    if constexpr (W > 4) {
      const float4 hi = X->red[par][i][1];
      tot = (tot + hi.x) + hi.y;
      if constexpr (W > 6) tot = (tot + hi.z) + hi.w;
    }
    val[i] = tot;
  }
  par ^= 1;
}

struct FwdW {
  V3 P; Q4 Q, Qp;
  V3 s, res;
  V3 fpos; Q4 fq; float frinv;
  int qb, vb;  // which halves of the double buffers hold the world poses
};

template <int W, int G = 1>
__device__ __forceinline__ float eval_fwd(const LaneC<1, WRT> &L, const SiteC &st, const UniW &u, const float (&pt)[NS], FwdW &S, WX<W> *X, int &par) {
  const int t = u.t;
  // free joint: its seven coordinates and its element live in warp 0 (stacb_tree_create guarantees it), which broadcasts them by
  // shuffles exactly as the one-warp path does; the other warps never touch them
  S.fpos = mk3(0.f, 0.f, 0.f); S.fq = mk4(1.f, 0.f, 0.f, 0.f); S.frinv = 1.f;
  if (u.has_free && u.warp == 0) {  // warp-uniform
    float fr[7];
#pragma unroll
    for (int i = 0; i < 7; i++) fr[i] = __shfl_sync(FULL, pt[1], i);
    S.fpos = mk3(fr[0], fr[1], fr[2]);
    S.fq = normalize4_nr(mk4(fr[3], fr[4], fr[5], fr[6]), &S.frinv);
  }
  float sh, ch;
  sincos_pi((pt[0] - L.ref[0]) * 0.5f, &sh, &ch);
  const float ct = fmaf(ch, ch, -(sh * sh)), sn = 2.0f * (sh * ch);
  Q4 Q = mk4(fmaf(L.Qs.w, sh, L.Qc.w * ch), fmaf(L.Qs.x, sh, L.Qc.x * ch), fmaf(L.Qs.y, sh, L.Qc.y * ch), fmaf(L.Qs.z, sh, L.Qc.z * ch));
  V3 v = mk3(fmaf(L.C.x, sn, fmaf(L.B.x, ct, L.A.x)), fmaf(L.C.y, sn, fmaf(L.B.y, ct, L.A.y)), fmaf(L.C.z, sn, fmaf(L.B.z, ct, L.A.z)));
  v = sel3(L.pfree, S.fpos, v);
  Q = sel4(L.pfree, S.fq, Q);
  // world poses: pointer jumping of position and quaternion TOGETHER through shared memory, one barrier per round
  // (P <- P_a + R(Q_a) P, Q <- Q_a Q with the ancestor a at distance 2^r; the one-warp path jumps them separately by shuffles)
  int b = 0;
  X->q[0][t] = make_float4(Q.w, Q.x, Q.y, Q.z);
  X->v[0][t] = make_float4(v.x, v.y, v.z, 0.f);
  gsync<W, G>(u);
#pragma unroll
  for (int r = 0; r < WRT; r++) {
    if (r < u.rounds) {  // uniform; round r reads half r & 1 of the double buffers and writes the other
      const Q4 Qa = q4_of(X->q[r & 1][L.src[r]]);
      const V3 Pa = v3_of(X->v[r & 1][L.src[r]]);
      v = add3(Pa, rotq(v, Qa));
      Q = qmul(Qa, Q);
      X->q[(r & 1) ^ 1][t] = make_float4(Q.w, Q.x, Q.y, Q.z);
      X->v[(r & 1) ^ 1][t] = make_float4(v.x, v.y, v.z, 0.f);
      gsync<W, G>(u);
    }
  }
  b = u.rounds & 1;
  S.Qp = q4_of(X->q[b][L.par]);
  S.Q = Q;
  S.P = v;
  S.qb = b;
  S.vb = b;
  const int c = b;
  // marker sites, masked residuals, loss
  S.s = add3(v3_of(X->v[c][st.eb]), rotq(st.off, q4_of(X->q[b][st.eb])));
  S.res = mk3((st.kp.x - S.s.x) * st.km.x, (st.kp.y - S.s.y) * st.km.y, (st.kp.z - S.s.z) * st.km.z);
  float e[1] = {fmaf(S.res.z, S.res.z, fmaf(S.res.y, S.res.y, S.res.x * S.res.x))};
  cta_sum<W, 1, G>(e, u, X, par);
  return e[0];
}

template <int W, int G = 1>
__device__ __forceinline__ void eval_bwd(const LaneC<1, WRT> &L, const FwdW &S, const UniW &u, bool free_wanted, float (&g)[NS], WX<W> *X) {
  const V3 c = v3_of(X->v[S.vb][0]);
  const V3 f = mk3(-2.0f * S.res.x, -2.0f * S.res.y, -2.0f * S.res.z);
  const V3 tq = cross3(sub3(S.s, c), f);
  float w[6] = {f.x, f.y, f.z, tq.x, tq.y, tq.z};
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const float up = __shfl_up_sync(FULL, w[i], off);
      if (u.lane >= off) w[i] = w[i] + up;
    }
  }
  if (u.lane == 31) {
    X->wsum[u.warp][0] = make_float4(w[0], w[1], w[2], w[3]);
    X->wsum[u.warp][1] = make_float4(w[4], w[5], 0.f, 0.f);
  }
  gsync<W, G>(u);
  if (u.warp > 0) {  // warp-uniform
    float off[6];
    {
      const float4 a = X->wsum[0][0], c2 = X->wsum[0][1];
      off[0] = a.x; off[1] = a.y; off[2] = a.z; off[3] = a.w; off[4] = c2.x; off[5] = c2.y;
    }
    for (int ww = 1; ww < u.warp; ww++) {
      const float4 a = X->wsum[ww][0], c2 = X->wsum[ww][1];
      off[0] = off[0] + a.x; off[1] = off[1] + a.y; off[2] = off[2] + a.z; off[3] = off[3] + a.w; off[4] = off[4] + c2.x; off[5] = off[5] + c2.y;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) w[i] = w[i] + off[i];
  }
  X->wpre[u.t][0] = make_float4(w[0], w[1], w[2], w[3]);
  X->wpre[u.t][1] = make_float4(w[4], w[5], 0.f, 0.f);
  gsync<W, G>(u);
  float wr[6];
  {
    const float4 e0 = X->wpre[L.se][0], e1 = X->wpre[L.se][1], a0 = X->wpre[L.sa][0], a1 = X->wpre[L.sa][1];
    wr[0] = e0.x - a0.x; wr[1] = e0.y - a0.y; wr[2] = e0.z - a0.z; wr[3] = e0.w - a0.w; wr[4] = e1.x - a1.x; wr[5] = e1.y - a1.y;
  }
  const V3 F = mk3(wr[0], wr[1], wr[2]), Tq = mk3(wr[3], wr[4], wr[5]);
  const V3 pp = v3_of(X->v[S.vb][L.par]);
  const Q4 pc = conj4(S.Qp);
  const V3 T0 = sub3(Tq, cross3(sub3(pp, c), F));
  const V3 Fp = rotq(F, pc), Tp = rotq(T0, pc);
  g[0] = dot3(L.ax[0], sub3(Tp, cross3(L.anc[0], Fp)));
  g[1] = 0.f;
  if (free_wanted) {  // uniform; the free joint's subtree is its body's subtree: the same two prefix entries
    float wf[6];
    {
      const float4 e0 = X->wpre[u.free_se][0], e1 = X->wpre[u.free_se][1], a0 = X->wpre[u.free_sa][0], a1 = X->wpre[u.free_sa][1];
      wf[0] = e0.x - a0.x; wf[1] = e0.y - a0.y; wf[2] = e0.z - a0.z; wf[3] = e0.w - a0.w; wf[4] = e1.x - a1.x; wf[5] = e1.y - a1.y;
    }
    const V3 Ff = mk3(wf[0], wf[1], wf[2]);
    const V3 Tf = sub3(mk3(wf[3], wf[4], wf[5]), cross3(sub3(S.fpos, c), Ff));
    float g4[4];
    quat_grad_left(S.fq, Tf, S.frinv, g4);
    float v = Ff.x;
    v = u.t == 1 ? Ff.y : v; v = u.t == 2 ? Ff.z : v; v = u.t == 3 ? g4[0] : v;
    v = u.t == 4 ? g4[1] : v; v = u.t == 5 ? g4[2] : v; v = u.t == 6 ? g4[3] : v;
    g[1] = u.t < 7 ? v : 0.f;
  }
}

// fast::solve on W warps (same two-state machine, same arithmetic; the sums over the lanes are CTA sums)
template <int W>
__device__ __forceinline__ SolveOut solve(const LaneC<1, WRT> &L, const SiteC &st, const UniW &u, const Slots<NS> &co, const float (&q0)[NS],
                                          unsigned maskbits, float sqp, float (&x)[NS], WX<W> *X, int &par) {
  float y[NS], g[NS], xn[NS], d[NS], gt[NS];
#pragma unroll
  for (int m = 0; m < NS; m++) { x[m] = q0[m]; y[m] = x[m]; xn[m] = x[m]; g[m] = 0.f; gt[m] = 0.f; }
  float t = 1.0f, step = 1.0f, stp = 1.0f, fy = 0.f, sq = 0.f, dg = 0.f;
  int halv = 0;
  bool in_ls = false;
  SolveOut out;
  out.iters = 0; out.ls = 0; out.bad = false; out.err = __int_as_float(0x7f800000);
  if (u.maxiter <= 0) return out;
  const bool fw = u.has_free && __syncthreads_or(u.t < 7 && ((maskbits >> 1) & 1u));
  SolveC<NS> sc;
  {
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
    float v[1] = {fast::lane_dot<NS>(dn, dn)};
    cta_sum<W, 1>(v, u, X, par);
    sqp = v[0] + sqp;
  }
  FwdW S;
  for (;;) {
    float pt[NS];
#pragma unroll
    for (int m = 0; m < NS; m++) pt[m] = in_ls ? xn[m] : y[m];
    const float f = eval_fwd<W>(L, st, u, pt, S, X, par);
    bool rejected = false;
    if (in_ls) {
      out.ls++;
      if (!(f - f == 0.0f)) out.bad = true;
      const float dec = stp * (f - fy);
      const float cond = fmaf(stp, dg, 0.5f * sq);
      rejected = (dec > cond + 1.1920929e-07f) && (halv < u.maxls);
    }
    if (!rejected) {  // uniform over the CTA: f, fy, sq, dg are CTA sums
      eval_bwd<W>(L, S, u, fw, gt, X);
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
        float v[1] = {fast::lane_dot<NS>(d, d)};
        cta_sum<W, 1>(v, u, X, par);
        out.err = sqrtf(v[0]);
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
    float v2[2] = {fast::lane_dot<NS>(d, d), fast::lane_dot<NS>(d, g)};
    cta_sum<W, 2>(v2, u, X, par);
    sq = v2[0] + sqp;
    dg = v2[1];
  }
  return out;
}

// fast::solve_pair on two groups of W warps (group g evaluates what warp g evaluates there; same arithmetic, same accepted candidate)
template <int W>
struct XPW {  // shared memory
  float g[32 * W * NS];
  float acc[2][2], nf[2][2];  // [parity][group]: accepted flag, non-finite loss flag of the two candidates
  float err, fy;
};

template <int W>
__device__ __forceinline__ SolveOut solve_pair_w(const LaneC<1, WRT> &L, const SiteC &st, const UniW &u, const Slots<NS> &co, const float (&q0)[NS],
                                                 unsigned maskbits, float sqp, float (&x)[NS], WX<W> *X, XPW<W> *xc, int &par, int &xpar) {
  constexpr int G = 2;
  float y[NS], g[NS], gt[NS], xj[NS], d[NS];
#pragma unroll
  for (int m = 0; m < NS; m++) { x[m] = q0[m]; y[m] = x[m]; g[m] = 0.f; }
  SolveOut out;
  out.iters = 0; out.ls = 0; out.bad = false; out.err = __int_as_float(0x7f800000);
  if (u.maxiter <= 0) return out;
  const bool fw = u.has_free && __syncthreads_or(u.t < 7 && ((maskbits >> 1) & 1u));
  SolveC<NS> sc;
  {
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
    float v[1] = {fast::lane_dot<NS>(dn, dn)};
    cta_sum<W, 1, G>(v, u, X, par);
    sqp = v[0] + sqp;
  }
  FwdW S;
  float fy = eval_fwd<W, G>(L, st, u, y, S, X, par);  // f(y0), grad f(y0): both groups (identical values)
  eval_bwd<W, G>(L, S, u, fw, g, X);
#pragma unroll
  for (int m = 0; m < NS; m++) g[m] = g[m] * sc.gm[m];
  float t = 1.0f, stp = 1.0f;
  int base = 0;
  const int w = u.grp;
  for (;;) {
    // phase A: this group's line-search candidate
    const float sj = w ? stp * 0.5f : stp;
#pragma unroll
    for (int m = 0; m < NS; m++) {
      xj[m] = clipm(fmaf(-sj, g[m], y[m]), sc.lb[m], sc.ub[m]);
      d[m] = xj[m] - y[m];
    }
    float v2[2] = {fast::lane_dot<NS>(d, d), fast::lane_dot<NS>(d, g)};
    cta_sum<W, 2, G>(v2, u, X, par);
    const float sq = v2[0] + sqp, dg = v2[1];
    const float f = eval_fwd<W, G>(L, st, u, xj, S, X, par);
    const float dec = sj * (f - fy);
    const float cond = fmaf(sj, dg, 0.5f * sq);
    const bool rejected = (dec > cond + 1.1920929e-07f) && (base + w < u.maxls);
    if (u.t == 0) {
      xc->acc[xpar][w] = rejected ? 0.f : 1.f;
      xc->nf[xpar][w] = (f - f == 0.0f) ? 0.f : 1.f;
    }
    __syncthreads();
    const bool a0 = xc->acc[xpar][0] != 0.f, a1 = xc->acc[xpar][1] != 0.f;
    const int k = a0 ? 0 : (a1 ? 1 : -1);
    if (xc->nf[xpar][0] != 0.f || (k != 0 && xc->nf[xpar][1] != 0.f)) out.bad = true;  // candidates the sequential search evaluates
    xpar ^= 1;
    if (k < 0) {  // both rejected: next two step sizes
      base += 2;
      stp = stp * 0.25f;
      continue;
    }
    out.ls += base + k + 1;
    const float sk = k ? stp * 0.5f : stp;
    const float beta = beta_of(u.betas, out.iters, t);
    // phase B
    if (w == k) {  // gradient at the accepted x+ (this group's forward state) -> unit-step fixed-point residual
      eval_bwd<W, G>(L, S, u, fw, gt, X);
#pragma unroll
      for (int m = 0; m < NS; m++) {
        gt[m] = gt[m] * sc.gm[m];
        d[m] = clipm(xj[m] - gt[m], sc.lb[m], sc.ub[m]) - xj[m];
      }
      float v[1] = {fast::lane_dot<NS>(d, d)};
      cta_sum<W, 1, G>(v, u, X, par);
      if (u.t == 0) xc->err = sqrtf(v[0]);
#pragma unroll
      for (int m = 0; m < NS; m++) { y[m] = fmaf(beta, xj[m] - x[m], xj[m]); x[m] = xj[m]; }
    } else {  // the partner: x+_k, the extrapolation y', f(y') and grad f(y')
#pragma unroll
      for (int m = 0; m < NS; m++) {
        const float xk = clipm(fmaf(-sk, g[m], y[m]), sc.lb[m], sc.ub[m]);
        y[m] = fmaf(beta, xk - x[m], xk);
        x[m] = xk;
      }
      const float fyn = eval_fwd<W, G>(L, st, u, y, S, X, par);
      eval_bwd<W, G>(L, S, u, fw, gt, X);
#pragma unroll
      for (int m = 0; m < NS; m++) xc->g[32 * W * m + u.t] = gt[m] * sc.gm[m];
      if (u.t == 0) xc->fy = fyn;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NS; m++) g[m] = xc->g[32 * W * m + u.t];
    fy = xc->fy;
    out.err = xc->err;
    stp = (sk <= 1e-6f) ? 1.0f : sk / 0.5f;
    out.iters++;
    sqp = 0.f;
    base = 0;
    // no third barrier: g / fy / err are written again only after the NEXT phase-A barrier, which a thread reaches only once it has
    // read this iteration's values; the accept flags alternate between two parities
    if (!(out.err > u.tol && out.iters < u.maxiter)) break;
  }
  return out;
}

__device__ __forceinline__ void slots_init_w(Slots<NS> &co, SlotAdr<1> &sa, const LaneC<1, WRT> &L, const DevTree &T, int t, const float *__restrict__ lb,
                                             const float *__restrict__ ub) {
  co.valid[0] = L.hinge[0]; sa.adr[0] = L.adr[0];
  co.valid[1] = T.fs_free_e >= 0 && t < 7;
  sa.adr[1] = co.valid[1] ? T.free_adr + t : 0;
#pragma unroll
  for (int m = 0; m < NS; m++) {
    co.lb[m] = (co.valid[m] && lb) ? lb[sa.adr[m]] : 0.f;
    co.ub[m] = (co.valid[m] && ub) ? ub[sa.adr[m]] : 0.f;
  }
}

template <int W, int G = 1>
__device__ __forceinline__ float passive_sq_w(const DevTree &T, const UniW &u, const float *q, const float *__restrict__ lb, const float *__restrict__ ub,
                                              WX<W> *X, int &par) {
  float acc = 0.f;
  bool first = true;
  for (int i = u.t; i < T.npassive; i += 32 * W) {
    const int p = __ldg(T.passive + i);
    const float v = q[p];
    const float d = clipm(v, lb[p], ub[p]) - v;
    acc = first ? d * d : fmaf(d, d, acc);
    first = false;
  }
  float v[1] = {acc};
  cta_sum<W, 1, G>(v, u, X, par);
  return v[0];
}

// One chain per CTA of G groups of W warps; frames sequential; the stage loop of fast::fast_pose_kernel.
// MINB: CTAs per SM the registers are capped for (2 once there are more chains than SMs and two CTAs fit the register file)
// G = 2: pair mode (solve_pair_w) -- both groups carry the whole solver state; group 0 stages and writes.
template <int W, int NBF, int MINB, int G>
__global__ void __launch_bounds__(32 * W * G, MINB) wide_pose_kernel(DevTree T, PoseArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_chain;
  __shared__ float s_beta[BT + 1];
  if (threadIdx.x == 0) beta_table_init(s_beta);
  __syncthreads();
  constexpr int NT = 32 * W * G;  // threads of the CTA (staging loops); a group has 32 W
  const int tc = threadIdx.x, grp = tc / (32 * W), t = tc - grp * (32 * W), lane = t & 31, warp = t >> 5, idl = 32 * W - 1;
  const int area = 2 * T.nqp + 7 * T.pqn;  // qbuf [nqp], (unused) [nqp], PQ [7 pqn]: the cold full-model FK of the outputs (warp 0)
  Chain ch(T, smem, nullptr, lane, 0, 1, 0);
  const int xoff = (area + 3) & ~3;  // the exchange areas start on a 16-byte boundary
  WX<W> *X = reinterpret_cast<WX<W> *>(smem + xoff) + grp;
  XPW<W> *xc = reinterpret_cast<XPW<W> *>(reinterpret_cast<WX<W> *>(smem + xoff) + G);
  const bool writer = warp == 0 && grp == 0;
  LaneC<1, WRT> L;
  lane_init<1, WRT>(L, T, t, idl);
  SiteC st;
  site_init(st, T, t, a.site_pos, idl);
  Slots<NS> co;
  SlotAdr<1> sa;
  slots_init_w(co, sa, L, T, t, a.lb, a.ub);
  UniW u;
  u.t = t; u.lane = lane; u.warp = warp; u.grp = grp; u.rounds = T.fs.rounds; u.free_sa = T.free_sa; u.free_se = T.free_se; u.has_free = T.fs_free_e >= 0;
  u.tol = a.tol; u.maxiter = a.maxiter; u.maxls = a.maxls; u.betas = s_beta;
  const int nq = T.nq, K = T.K, nb = T.nbody, S1 = 1 + a.P, npassive = T.npassive;
  const MaskSpec full_ms = {nullptr, nq}, root_ms = {nullptr, a.root_dims};
  const unsigned full_bits = slot_bits<1>(co, sa, full_ms), root_bits = slot_bits<1>(co, sa, root_ms);
  const int n_root = a.do_root ? 2 : 0;
  const int n_pose = (a.do_root == 2) ? 0 : a.F;
  const int n_stage = n_root + n_pose * S1;
  int par = 0, xpar = 0;
  for (;;) {
    if (threadIdx.x == 0) s_chain = atomicAdd(a.counter, 1);
    __syncthreads();
    const int c = s_chain;
    if (c >= a.C) break;
    for (int i = tc; i < nq; i += NT) ch.qbuf[i] = a.qpos_io[(size_t)c * nq + i];
    __syncthreads();
    float q[NS], q0[NS], x[NS];
    slots_gather<1>(co, sa, ch.qbuf, q);
    bool bad = false;
    const float *kpc = a.kp + (size_t)c * a.kp_stride;
    for (int sidx = 0; sidx < n_stage; sidx++) {
      const bool is_root = sidx < n_root;
      const int f = is_root ? 0 : (sidx - n_root) / S1;
      const int sg = is_root ? 0 : (sidx - n_root) % S1;
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
        bits = (sg == 0) ? full_bits : slot_bits<1>(co, sa, ms);
      }
#pragma unroll
      for (int m = 0; m < NS; m++) {
        q0[m] = q[m];
        if (is_root && co.valid[m] && sa.adr[m] < 3) q0[m] = kpc[3 * a.root_kp_idx + sa.adr[m]];
      }
      const float sqp = npassive ? passive_sq_w<W, G>(T, u, ch.qbuf, a.lb, a.ub, X, par) : 0.f;
      SolveOut so;
      if constexpr (G == 2) so = solve_pair_w<W>(L, st, u, co, q0, bits, sqp, x, X, xc, par, xpar);
      else so = solve<W>(L, st, u, co, q0, bits, sqp, x, X, par);
#pragma unroll
      for (int m = 0; m < NS; m++) q[m] = ((bits >> m) & 1u) ? x[m] : q0[m];
      bad |= so.bad;
      if (npassive && u.maxiter > 0) {
        for (int i = tc; i < npassive; i += NT) {
          const int p = __ldg(T.passive + i);
          if (mask_has(ms, p)) ch.qbuf[p] = clipm(ch.qbuf[p], a.lb[p], a.ub[p]);
        }
        __syncthreads();
      }
      if (is_root) {
        if (a.root_stats && tc == 0) { a.root_stats[4 * c + 2 * sidx] = so.iters; a.root_stats[4 * c + 2 * sidx + 1] = so.ls; }
      } else {
        const size_t fi = (size_t)c * a.F + f;
        if (a.iters && tc == 0) { a.iters[fi * S1 + sg] = so.iters; a.ls_evals[fi * S1 + sg] = so.ls; }
        if (sg == a.P) {  // last solve of the frame: full-model FK of the raw solution -> outputs (warp 0)
          if (grp == 0) slots_scatter<1>(co, sa, q, ch.qbuf);
          __syncthreads();
          if (writer) {
            outputs_from_qbuf<NBF>(ch, st, a.site_pos, a.qpos ? a.qpos + fi * nq : nullptr, a.xpos ? a.xpos + fi * nb * 3 : nullptr,
                                   a.xquat ? a.xquat + fi * nb * 4 : nullptr, nullptr);
            if (a.err && lane == 0) a.err[fi] = so.err;
          }
          __syncthreads();
          if (a.sites && st.k >= 0 && grp == 0) {  // every site thread writes its own site from the cold FK's poses (same arithmetic as outputs_from_qbuf)
            const float *o = ch.PQ + 7 * st.ef;
            const V3 off = mk3(__ldg(a.site_pos + 3 * st.k), __ldg(a.site_pos + 3 * st.k + 1), __ldg(a.site_pos + 3 * st.k + 2));
            const V3 sp = add3(lds3(o), rotate(off, lds4(o + 3)));
            float *so3 = a.sites + fi * K * 3 + 3 * st.k;
            so3[0] = sp.x; so3[1] = sp.y; so3[2] = sp.z;
          }
          __syncthreads();
        }
      }
      if (warp == 0) normalize_free<1>(Uni{lane, 0, u.has_free, 0.f, 0, 0, nullptr}, q);  // threads 0..6 hold the free joint
    }
    if (grp == 0) slots_scatter<1>(co, sa, q, ch.qbuf);
    __syncthreads();
    for (int i = tc; i < nq; i += NT) a.qpos_io[(size_t)c * nq + i] = ch.qbuf[i];
    if (a.status && tc == 0) a.status[c] = bad ? 1 : 0;
    __syncthreads();
  }
}

// B independent items: q_loss + gradient (mode 1) or one FISTA solve (mode 2); one CTA of W warps per item.
template <int W>
__global__ void __launch_bounds__(32 * W, 1) wide_batch_kernel(DevTree T, BatchArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_beta[BT + 1];
  if (threadIdx.x == 0) beta_table_init(s_beta);
  __syncthreads();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, idl = 32 * W - 1;
  WX<W> *X = reinterpret_cast<WX<W> *>(smem);
  LaneC<1, WRT> L;
  lane_init<1, WRT>(L, T, t, idl);
  SiteC st;
  site_init(st, T, t, a.site_pos, idl);
  site_mask_u8(st, a.kp_mask);
  Slots<NS> co;
  SlotAdr<1> sa;
  slots_init_w(co, sa, L, T, t, a.lb, a.ub);
  UniW u;
  u.t = t; u.lane = lane; u.warp = warp; u.grp = 0; u.rounds = T.fs.rounds; u.free_sa = T.free_sa; u.free_se = T.free_se; u.has_free = T.fs_free_e >= 0;
  u.tol = a.tol; u.maxiter = a.maxiter; u.maxls = a.maxls; u.betas = s_beta;
  const int nq = T.nq, K = T.K;
  const MaskSpec ms = {a.q_mask, nq};
  const unsigned bits = slot_bits<1>(co, sa, ms);
  int par = 0;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    float q[NS], q0[NS];
    const float *qb = a.q + (size_t)b * nq;
    slots_gather<1>(co, sa, qb, q);
    if (a.q0) slots_gather<1>(co, sa, a.q0 + (size_t)b * nq, q0);
    else {
#pragma unroll
      for (int m = 0; m < NS; m++) q0[m] = q[m];
    }
    site_load_kp(st, a.kp + (size_t)b * 3 * K);
    if (a.mode == 1) {
      float pt[NS], g[NS];
#pragma unroll
      for (int m = 0; m < NS; m++) pt[m] = ((bits >> m) & 1u) ? q[m] : q0[m];
      FwdW S;
      const float loss = eval_fwd<W>(L, st, u, pt, S, X, par);
      if (t == 0) a.out_a[b] = loss;
      if (a.out_b) {
        const bool fw = u.has_free && __syncthreads_or(t < 7 && ((bits >> 1) & 1u));
        eval_bwd<W>(L, S, u, fw, g, X);
        float *go = a.out_b + (size_t)b * nq;
        for (int i = t; i < nq; i += 32 * W) go[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int m = 0; m < NS; m++)
          if (co.valid[m] && ((bits >> m) & 1u)) go[sa.adr[m]] = g[m];
      }
    } else {
      float x[NS];
      const float sqp = T.npassive ? passive_sq_w<W>(T, u, qb, a.lb, a.ub, X, par) : 0.f;
      const SolveOut so = solve<W>(L, st, u, co, q, bits, sqp, x, X, par);
      if (u.maxiter > 0) {
#pragma unroll
        for (int m = 0; m < NS; m++)
          if (co.valid[m] && !((bits >> m) & 1u)) x[m] = clipm(x[m], co.lb[m], co.ub[m]);
      }
      float *po = a.out_a + (size_t)b * nq;
      for (int i = t; i < T.npassive; i += 32 * W) {
        const int p = __ldg(T.passive + i);
        po[p] = u.maxiter > 0 ? clipm(qb[p], a.lb[p], a.ub[p]) : qb[p];
      }
      slots_scatter<1>(co, sa, x, po);
      if (t == 0) { a.out_b[b] = so.err; a.iters[b] = so.iters; a.ls_evals[b] = so.ls; }
    }
    __syncthreads();
  }
}

}  // namespace wide
}  // namespace stacb
