// stacb_device.cuh -- fused STAC solver for sm_100a: device code and kernels (C ABI and dispatch in stacb_abi.cu).
//
// One chain = one clip (frames strictly sequential, warm-started) or one independent solve.  A warp evaluates
//   FK (parent-frame local transforms, pointer-jumping composition) -> marker sites -> masked squared residual
//   -> analytic gradient (prefix-scan subtree wrench, rotated into each body's parent frame, Jacobian transpose per joint)
// with lane <-> body / site / coordinate mappings fixed for the whole kernel and the tree records in registers, and drives
// the FISTA projected-gradient solver with backtracking (jaxopt 0.8.5 ProjectedGradient semantics).
//   throughput modes: one warp per chain (solve);  latency mode: four cooperating warps per chain (solve4).
// Replaces reference stac_mjx/stac_core.py:27-99 (q_loss/_q_opt), compute_stac.py:17-104,170-278 (root/pose optimisation
// loops), stac_core.py:146-159 (m-phase statistics) and the MJX / jaxopt code underneath them.  Compiled with -fmad=false:
// every fused multiply-add is explicit, so the arithmetic is the canonical order of DESIGN.md section 4 and is reproduced
// bit for bit by the CPU oracle (which is never linked or included here).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "stacb.h"
#include "stacb_math.cuh"

namespace stacb {

constexpr int JMAX = 3;                     // joints per body the descriptor can hold
#ifndef V_JM
#define V_JM 3
#endif
constexpr int JM = V_JM;                    // joint slots the hot path of this kernel variant processes (<= JMAX)
constexpr int REC = 10 + 12 * JMAX + 1;     // words per body record (odd stride)
constexpr int R_POS = 0, R_QUAT = 3, R_NJNT = 7, R_PARENT = 8, R_BODY = 9, R_JNT = 10;
constexpr int J_TYPE = 0, J_ADR = 1, J_POS = 2, J_AXIS = 5, J_REF = 8, J_SA = 9, J_SE = 10, J_STRIDE = 12;

struct DevSet {
  int n, rounds;
  const int *rec;  // [n][REC]
  const int *anc;  // [rounds][n] set-local ancestor index at distance 2^r, -1 if none
};

struct DevTree {
  int nbody, nq, njnt, K, spl;
  DevSet act, full;
  const int *site_order;  // [K] sorted position -> keypoint index
  const int *site_eact;   // [K] by sorted position: active-set index of the site's body
  const int *site_efull;  // [K] by sorted position: full-set index of the site's body
  int nqp, pqn, npre;     // per-chain shared memory: qbuf[nqp] gbuf[nqp] PQ[pqn*7] Ipre[npre*6]
  int nquat;              // free / ball joints of the whole model
  int free_e, free_adr, free_sa, free_se;  // primary free joint: active-set element (-1 if none), qpos address, site range
  int any_other;          // some active joint is neither a hinge nor the primary free joint (ball / slide / extra free)
  const int *quat_adr;    // [nquat] qpos address of each quaternion
  int npassive;           // register-resident path (stacb_fast.cuh): qpos addresses outside its solver slots
  const int *passive;     // [npassive] ascending
  // element set of the register-resident path: the active bodies, or -- when those do not fit a warp -- the active bodies that
  // carry joints, every jointless (welded) active body folded into its nearest jointed ancestor with a constant relative pose
  DevSet fs;
  int fs_free_e;          // element of the primary free joint (-1 if none)
  const int *site_efs;    // [K] by sorted position: element the site rides on
  const float *site_rel;  // [K][7] by sorted position: pose (pos, quat) of the site's own body in that element's frame; null when nothing was folded
};

__host__ __device__ inline int role_smem_floats(const DevTree &T) { return 2 * T.nqp + 7 * T.pqn; }  // qbuf, gbuf, PQ
__host__ __device__ inline int warp_smem_floats(const DevTree &T) { return 6 * T.npre; }               // Ipre
__host__ __device__ inline int chain_smem_floats(const DevTree &T) { return role_smem_floats(T) + warp_smem_floats(T); }

// Per-evaluation scratch in shared memory. One "role" = one evaluation of the loss; it is carried out by G member warps
// (G = 1 everywhere except the grouped latency mode for wide trees, where the bodies of the tree are dealt out over
// G = 3 warps).  qbuf / gbuf / PQ are shared by the members of a role, Ipre is private to each warp.
struct Chain {
  const DevTree &T;
  int lane;
  int g, G, bar;  // member index, members per role, named barrier of the role
  float *qbuf, *gbuf, *PQ, *Ipre;
  __device__ Chain(const DevTree &t, float *role_base, float *warp_base, int ln, int g_, int G_, int bar_)
      : T(t), lane(ln), g(g_), G(G_), bar(bar_) {
    qbuf = role_base;
    gbuf = qbuf + t.nqp;
    PQ = gbuf + t.nqp;
    Ipre = warp_base;
  }
  // barrier among the members of the role (a warp-level fence when the role is a single warp)
  __device__ __forceinline__ void sync() const {
    if (G == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(32 * G) : "memory");
  }
  // active-set element handled by this lane in slot i
  __device__ __forceinline__ int elem(int i) const { return lane + 32 * (g + G * i); }
};

__device__ __forceinline__ float ldf(const int *p) { return __int_as_float(__ldg(p)); }
__device__ __forceinline__ V3 ldv3(const int *p) { return mk3(ldf(p), ldf(p + 1), ldf(p + 2)); }
__device__ __forceinline__ V3 lds3(const float *p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ Q4 lds4(const float *p) { return mk4(p[0], p[1], p[2], p[3]); }
__device__ __forceinline__ void sts7(float *o, V3 p, Q4 q) { o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = q.w; o[4] = q.x; o[5] = q.y; o[6] = q.z; }
__device__ __forceinline__ V3 shfl3(V3 v, int src) {
  return mk3(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src));
}
__device__ __forceinline__ Q4 shfl4(Q4 q, int src) {
  return mk4(__shfl_sync(0xffffffffu, q.w, src), __shfl_sync(0xffffffffu, q.x, src), __shfl_sync(0xffffffffu, q.y, src),
             __shfl_sync(0xffffffffu, q.z, src));
}

// One body's constants (a record of the tree descriptor). On the hot path these live in registers for
// the whole kernel (the lane <-> body mapping never changes); cold paths load them on demand.
struct BodyConst {
  V3 pos; Q4 quat;
  int nj, parent, body;
  int jtype[JMAX], jadr[JMAX], jsa[JMAX], jse[JMAX];
  V3 jpos[JMAX], jaxis[JMAX];
  float jref[JMAX];
};

__device__ __forceinline__ void load_body(BodyConst &b, const int *__restrict__ r) {
  b.pos = ldv3(r + R_POS);
  b.quat = mk4(ldf(r + R_QUAT), ldf(r + R_QUAT + 1), ldf(r + R_QUAT + 2), ldf(r + R_QUAT + 3));
  b.nj = __ldg(r + R_NJNT); b.parent = __ldg(r + R_PARENT); b.body = __ldg(r + R_BODY);
#pragma unroll
  for (int jj = 0; jj < JMAX; jj++) {
    const int *jr = r + R_JNT + J_STRIDE * jj;
    b.jtype[jj] = __ldg(jr + J_TYPE); b.jadr[jj] = __ldg(jr + J_ADR); b.jsa[jj] = __ldg(jr + J_SA); b.jse[jj] = __ldg(jr + J_SE);
    b.jpos[jj] = ldv3(jr + J_POS); b.jaxis[jj] = ldv3(jr + J_AXIS); b.jref[jj] = ldf(jr + J_REF);
  }
}

constexpr int RMAX = 8;  // pointer-jumping rounds kept in registers (tree depth <= 256)

// Register-resident view of the ACTIVE body set for one lane.
template <int NB>
struct Hot {
  BodyConst bc[NB];
  unsigned long long anc_lo[NB], anc_hi[NB];  // ancestor at distance 2^r, 16 bits per round (0xffff = none)
  bool on[NB];
  bool hinge[NB][JM];  // slot holds a hinge joint
  bool other[NB][JM];  // slot holds a joint handled by the rare (divergent) path
  int hadr[NB][JM];    // qpos address of the hinge (0 when the slot is not a hinge: always a valid address)
  bool pfree[NB];        // this body carries the primary free joint
};

template <int NB>
__device__ __forceinline__ void hot_init(Hot<NB> &H, const Chain &ch) {
  const DevTree &T = ch.T;
  const DevSet &S = T.act;
#pragma unroll
  for (int i = 0; i < NB; i++) {
    const int e = ch.elem(i);
    H.on[i] = e < S.n;
    load_body(H.bc[i], S.rec + (size_t)(H.on[i] ? e : 0) * REC);
    if (!H.on[i]) { H.bc[i].nj = 0; H.bc[i].parent = -1; }
    H.pfree[i] = H.on[i] && e == T.free_e;
#pragma unroll
    for (int jj = 0; jj < JM; jj++) {
      const bool has = jj < H.bc[i].nj;
      H.hinge[i][jj] = has && H.bc[i].jtype[jj] == STACB_JNT_HINGE;
      H.other[i][jj] = has && !H.hinge[i][jj] && !(H.pfree[i] && jj == 0);
      H.hadr[i][jj] = H.hinge[i][jj] ? H.bc[i].jadr[jj] : 0;
      if (!H.hinge[i][jj]) { H.bc[i].jref[jj] = 0.f; }
    }
    H.anc_lo[i] = ~0ull; H.anc_hi[i] = ~0ull;
#pragma unroll
    for (int r = 0; r < RMAX; r++) {
      const int a = (H.on[i] && r < S.rounds) ? __ldg(S.anc + r * S.n + e) : -1;
      const unsigned long long v = (unsigned long long)(a & 0xffff) << (16 * (r & 3));
      const unsigned long long mask = ~(0xffffull << (16 * (r & 3)));
      if (r < 4) H.anc_lo[i] = (H.anc_lo[i] & mask) | v; else H.anc_hi[i] = (H.anc_hi[i] & mask) | v;
    }
  }
}

template <int NB>
__device__ __forceinline__ int hot_anc(const Hot<NB> &H, int i, int r) {
  const unsigned long long p = (r < 4) ? H.anc_lo[i] : H.anc_hi[i];
  const int a = (int)((p >> (16 * (r & 3))) & 0xffffull);
  return a == 0xffff ? -1 : a;
}

// Per-lane results of the local phase that the reverse sweep needs.
template <int NB>
struct Keep {
  V3 anchor[NB][JMAX];  // joint anchor in the parent frame of its body
  V3 axis[NB][JMAX];    // joint axis in the parent frame of its body
  float fnorm[NB];      // reciprocal of the divisor of the quaternion normalisation (free / ball joint)
};

template <int NB>
struct FkState {
  V3 P[NB];
  Q4 Q[NB];
  Keep<NB> keep;
  float fd;  // reciprocal divisor of the primary free joint's quaternion normalisation (uniform across the warp)
};

// MJX smooth.kinematics per-body step, evaluated in the parent's frame (canonical order).
template <int NB, bool KEEP>
__device__ __forceinline__ void fk_local(const BodyConst &b, float *qbuf, V3 &pos, Q4 &quat, Keep<NB> *keep, int slot) {
  pos = b.pos;
  quat = b.quat;
  // half-angle sines / cosines of all hinges first: independent chains the scheduler can interleave
  float sn[JMAX], cs[JMAX];
#pragma unroll
  for (int jj = 0; jj < JMAX; jj++) {
    sn[jj] = 0.f; cs[jj] = 1.f;
    if (jj < b.nj && b.jtype[jj] == STACB_JNT_HINGE) sincos_canon((qbuf[b.jadr[jj]] - b.jref[jj]) * 0.5f, &sn[jj], &cs[jj]);
  }
#pragma unroll
  for (int jj = 0; jj < JMAX; jj++) {
    if (jj < b.nj) {
      const int type = b.jtype[jj], adr = b.jadr[jj];
      const V3 jpos = b.jpos[jj], jaxis = b.jaxis[jj];
      if (type == STACB_JNT_HINGE) {
        const V3 anchor = add3(rotate(jpos, quat), pos);
        if (KEEP) { keep->anchor[slot][jj] = anchor; keep->axis[slot][jj] = rotate(jaxis, quat); }
        quat = qmul(quat, mk4(cs[jj], jaxis.x * sn[jj], jaxis.y * sn[jj], jaxis.z * sn[jj]));
        pos = sub3(anchor, rotate(jpos, quat));
      } else if (type == STACB_JNT_FREE) {
        float d;
        pos = lds3(qbuf + adr);
        if (KEEP) { keep->anchor[slot][jj] = pos; keep->axis[slot][jj] = mk3(0.f, 0.f, 1.f); }
        quat = normalize4(lds4(qbuf + adr + 3), &d);
        qbuf[adr + 3] = quat.w; qbuf[adr + 4] = quat.x; qbuf[adr + 5] = quat.y; qbuf[adr + 6] = quat.z;
        if (KEEP) keep->fnorm[slot] = d;
      } else if (type == STACB_JNT_BALL) {
        float d;
        const V3 anchor = add3(rotate(jpos, quat), pos);
        if (KEEP) { keep->anchor[slot][jj] = anchor; keep->axis[slot][jj] = rotate(jaxis, quat); }
        const Q4 ql = normalize4(lds4(qbuf + adr), &d);
        qbuf[adr] = ql.w; qbuf[adr + 1] = ql.x; qbuf[adr + 2] = ql.y; qbuf[adr + 3] = ql.z;
        if (KEEP) keep->fnorm[slot] = d;
        quat = qmul(quat, ql);
        pos = sub3(anchor, rotate(jpos, quat));
      } else {  // slide
        const V3 axis = rotate(jaxis, quat);
        if (KEEP) { keep->anchor[slot][jj] = add3(rotate(jpos, quat), pos); keep->axis[slot][jj] = axis; }
        const float d = qbuf[adr] - b.jref[jj];
        pos = mk3(fmaf(axis.x, d, pos.x), fmaf(axis.y, d, pos.y), fmaf(axis.z, d, pos.z));
      }
    }
  }
}

__device__ __forceinline__ V3 sel3(bool c, V3 a, V3 b) { return mk3(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z); }
__device__ __forceinline__ Q4 sel4(bool c, Q4 a, Q4 b) { return mk4(c ? a.w : b.w, c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z); }

// Hot-path FK over the active set. Same arithmetic as fk_local, laid out for the scheduler:
//  * the primary free joint is normalised by every lane (uniform code, no divergence) and selected by its body's lane;
//  * the hinge slots are straight-line code applied through selects (a lane without a hinge in a slot computes on
//    zeros and discards the result), so the three slots and the bookkeeping rotations interleave freely;
//  * ball / slide / additional free joints take a divergent path that is skipped warp-uniformly when the model has none.
// With one body per lane the pointer-jumping rounds exchange poses by warp shuffles, otherwise through PQ in shared memory.
template <int NB, bool KEEP>
__device__ __forceinline__ void fk_hot(const Chain &ch, const Hot<NB> &H, FkState<NB> &S) {
  V3 fpos = mk3(0.f, 0.f, 0.f);
  Q4 fq = mk4(1.f, 0.f, 0.f, 0.f);
  float fd = 1.f;
  if (ch.T.free_e >= 0) {  // uniform
    const int fa = ch.T.free_adr;
    fpos = lds3(ch.qbuf + fa);
    fq = normalize4(lds4(ch.qbuf + fa + 3), &fd);
  }
  S.fd = fd;
#pragma unroll
  for (int i = 0; i < NB; i++) {
    const BodyConst &b = H.bc[i];
    V3 pos = sel3(H.pfree[i], fpos, b.pos);
    Q4 quat = sel4(H.pfree[i], fq, b.quat);
    if (KEEP) S.keep.fnorm[i] = fd;
    float sn[JM], cs[JM];
#pragma unroll
    for (int jj = 0; jj < JM; jj++) sincos_canon((ch.qbuf[H.hadr[i][jj]] - b.jref[jj]) * 0.5f, &sn[jj], &cs[jj]);
#pragma unroll
    for (int jj = 0; jj < JM; jj++) {
      const V3 jpos = b.jpos[jj], jaxis = b.jaxis[jj];
      const V3 anchor = add3(rotate(jpos, quat), pos);
      if (KEEP) { S.keep.anchor[i][jj] = anchor; S.keep.axis[i][jj] = rotate(jaxis, quat); }
      const Q4 qn = qmul(quat, mk4(cs[jj], jaxis.x * sn[jj], jaxis.y * sn[jj], jaxis.z * sn[jj]));
      const V3 pn = sub3(anchor, rotate(jpos, qn));
      pos = sel3(H.hinge[i][jj], pn, pos);
      quat = sel4(H.hinge[i][jj], qn, quat);
      if (__builtin_expect(ch.T.any_other != 0, 0)) {  // uniform, rare
        if (H.other[i][jj]) {
          const int type = b.jtype[jj], adr = b.jadr[jj];
          if (type == STACB_JNT_FREE) {
            float d;
            pos = lds3(ch.qbuf + adr);
            if (KEEP) { S.keep.anchor[i][jj] = pos; S.keep.axis[i][jj] = mk3(0.f, 0.f, 1.f); }
            quat = normalize4(lds4(ch.qbuf + adr + 3), &d);
            ch.qbuf[adr + 3] = quat.w; ch.qbuf[adr + 4] = quat.x; ch.qbuf[adr + 5] = quat.y; ch.qbuf[adr + 6] = quat.z;
            if (KEEP) S.keep.fnorm[i] = d;
          } else if (type == STACB_JNT_BALL) {
            float d;
            const Q4 ql = normalize4(lds4(ch.qbuf + adr), &d);
            ch.qbuf[adr] = ql.w; ch.qbuf[adr + 1] = ql.x; ch.qbuf[adr + 2] = ql.y; ch.qbuf[adr + 3] = ql.z;
            if (KEEP) S.keep.fnorm[i] = d;
            quat = qmul(quat, ql);
            pos = sub3(anchor, rotate(jpos, quat));
          } else {  // slide
            const V3 axis = rotate(jaxis, quat);
            const float d = ch.qbuf[adr] - ldf(ch.T.act.rec + (size_t)ch.elem(i) * REC + R_JNT + J_STRIDE * jj + J_REF);
            pos = mk3(fmaf(axis.x, d, pos.x), fmaf(axis.y, d, pos.y), fmaf(axis.z, d, pos.z));
          }
        }
      }
    }
    S.P[i] = pos;
    S.Q[i] = quat;
    if (ch.T.free_e >= 0) ch.sync();  // every lane of the role has read the raw quaternion before its owner overwrites it
    if (H.pfree[i]) {  // MJX writes the normalised quaternion back into qpos
      const int fa = ch.T.free_adr;
      ch.qbuf[fa + 3] = fq.w; ch.qbuf[fa + 4] = fq.x; ch.qbuf[fa + 5] = fq.y; ch.qbuf[fa + 6] = fq.z;
    }
  }
  const int rounds = ch.T.act.rounds;
  if constexpr (NB == 1) {
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
      const int a = hot_anc<NB>(H, 0, r);
      const int src = (a >= 0) ? a : ch.lane;
      const V3 pa = shfl3(S.P[0], src);
      const Q4 qa = shfl4(S.Q[0], src);
      S.P[0] = sel3(a >= 0, add3(pa, rotate(S.P[0], qa)), S.P[0]);
      S.Q[0] = sel4(a >= 0, qmul(qa, S.Q[0]), S.Q[0]);
    }
  } else {
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
#pragma unroll
      for (int i = 0; i < NB; i++)
        if (H.on[i]) sts7(ch.PQ + 7 * ch.elem(i), S.P[i], S.Q[i]);
      ch.sync();
#pragma unroll
      for (int i = 0; i < NB; i++) {
        const int a = hot_anc<NB>(H, i, r);
        const float *o = ch.PQ + 7 * (a >= 0 ? a : 0);
        const V3 pa = lds3(o);
        const Q4 qa = lds4(o + 3);
        S.P[i] = sel3(a >= 0, add3(pa, rotate(S.P[i], qa)), S.P[i]);
        S.Q[i] = sel4(a >= 0, qmul(qa, S.Q[i]), S.Q[i]);
      }
      ch.sync();
    }
#pragma unroll
    for (int i = 0; i < NB; i++)
      if (H.on[i]) sts7(ch.PQ + 7 * ch.elem(i), S.P[i], S.Q[i]);
    ch.sync();
  }
}

// World pose of active-set element e. Warp-collective when NB == 1 (every lane must call it).
template <int NB>
__device__ __forceinline__ void gather_pose(const Chain &ch, const FkState<NB> &S, int e, V3 &p, Q4 &q) {
  if constexpr (NB == 1) {
    p = shfl3(S.P[0], e);
    q = shfl4(S.Q[0], e);
  } else {
    const float *o = ch.PQ + 7 * e;
    p = lds3(o);
    q = lds4(o + 3);
  }
}

// Cold-path FK over any body set with constants read from global memory; results in PQ (shared memory).
template <int NB>
__device__ __forceinline__ void fk_cold(const Chain &ch, const DevSet &S, V3 (&P)[NB], Q4 (&Q)[NB]) {
#pragma unroll
  for (int i = 0; i < NB; i++) {
    const int e = ch.lane + 32 * i;
    if (e < S.n) {
      BodyConst b;
      load_body(b, S.rec + (size_t)e * REC);
      fk_local<NB, false>(b, ch.qbuf, P[i], Q[i], nullptr, i);
    }
  }
  for (int r = 0; r < S.rounds; r++) {
#pragma unroll
    for (int i = 0; i < NB; i++) {
      const int e = ch.lane + 32 * i;
      if (e < S.n) sts7(ch.PQ + 7 * e, P[i], Q[i]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NB; i++) {
      const int e = ch.lane + 32 * i;
      if (e < S.n) {
        const int a = __ldg(S.anc + r * S.n + e);
        if (a >= 0) {
          const float *o = ch.PQ + 7 * a;
          const V3 pa = lds3(o);
          const Q4 qa = lds4(o + 3);
          P[i] = add3(pa, rotate(P[i], qa));
          Q[i] = qmul(qa, Q[i]);
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < NB; i++) {
    const int e = ch.lane + 32 * i;
    if (e < S.n) sts7(ch.PQ + 7 * e, P[i], Q[i]);
  }
  __syncwarp();
}

// Per-lane marker data: the lane owns sorted site positions lane*spl + i.
template <int SPL>
struct Sites {
  int k[SPL];      // keypoint index, -1 if the slot is empty
  int eact[SPL], efull[SPL];
  V3 off[SPL];     // site offset in the body frame (site_pos)
  V3 kp[SPL];      // observed keypoint
  V3 km[SPL];      // 0/1 mask per coordinate
};

template <int SPL>
struct SiteVals {
  V3 s[SPL];    // marker site world position
  V3 res[SPL];  // masked residual
};

template <int SPL>
__device__ __forceinline__ void sites_init(const Chain &ch, Sites<SPL> &st, const float *__restrict__ site_pos) {
#pragma unroll
  for (int i = 0; i < SPL; i++) {
    const int pos = ch.lane * ch.T.spl + i;
    st.k[i] = -1; st.eact[i] = 0; st.efull[i] = 0;
    st.off[i] = mk3(0.f, 0.f, 0.f); st.kp[i] = mk3(0.f, 0.f, 0.f); st.km[i] = mk3(0.f, 0.f, 0.f);
    if (i < ch.T.spl && pos < ch.T.K) {
      const int k = __ldg(ch.T.site_order + pos);
      st.k[i] = k;
      st.eact[i] = __ldg(ch.T.site_eact + pos);
      st.efull[i] = __ldg(ch.T.site_efull + pos);
      if (site_pos) st.off[i] = mk3(__ldg(site_pos + 3 * k), __ldg(site_pos + 3 * k + 1), __ldg(site_pos + 3 * k + 2));
    }
  }
}

template <int SPL>
__device__ __forceinline__ void sites_load_kp(Sites<SPL> &st, const float *__restrict__ kp) {
#pragma unroll
  for (int i = 0; i < SPL; i++)
    if (st.k[i] >= 0) st.kp[i] = mk3(kp[3 * st.k[i]], kp[3 * st.k[i] + 1], kp[3 * st.k[i] + 2]);
}

template <int SPL>
__device__ __forceinline__ void sites_mask_u8(Sites<SPL> &st, const uint8_t *__restrict__ kp_mask /*[3K] or null = ones*/) {
#pragma unroll
  for (int i = 0; i < SPL; i++)
    if (st.k[i] >= 0) {
      const int k = st.k[i];
      st.km[i] = kp_mask ? mk3(kp_mask[3 * k] ? 1.f : 0.f, kp_mask[3 * k + 1] ? 1.f : 0.f, kp_mask[3 * k + 2] ? 1.f : 0.f)
                         : mk3(1.f, 1.f, 1.f);
    }
}

template <int SPL>
__device__ __forceinline__ void sites_mask_kp(Sites<SPL> &st, const uint8_t *__restrict__ per_kp /*[K] or null = ones*/) {
#pragma unroll
  for (int i = 0; i < SPL; i++)
    if (st.k[i] >= 0) {
      const float m = (!per_kp || per_kp[st.k[i]]) ? 1.f : 0.f;
      st.km[i] = mk3(m, m, m);
    }
}

// Marker sites from the active-set poses, masked residuals and the loss (q_loss, stac_core.py:57-63).
template <int NB, int SPL>
__device__ __forceinline__ float sites_loss(const Chain &ch, const FkState<NB> &S, const Sites<SPL> &st, SiteVals<SPL> &sv) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < SPL; i++) {
    V3 pb;
    Q4 qb;
    gather_pose<NB>(ch, S, st.k[i] >= 0 ? st.eact[i] : 0, pb, qb);
    sv.s[i] = mk3(0.f, 0.f, 0.f); sv.res[i] = mk3(0.f, 0.f, 0.f);
    if (st.k[i] >= 0) {
      const V3 s = add3(pb, rotate(st.off[i], qb));
      const V3 res = mk3((st.kp[i].x - s.x) * st.km[i].x, (st.kp[i].y - s.y) * st.km[i].y, (st.kp[i].z - s.z) * st.km[i].z);
      sv.s[i] = s; sv.res[i] = res;
      const float e = fmaf(res.z, res.z, fmaf(res.y, res.y, res.x * res.x));
      acc = (i == 0) ? e : acc + e;
    }
  }
  return warp_sum(acc);
}

// Inclusive prefix (over sites sorted by body) of the per-site wrench (force, torque about cref) -> Ipre.
template <int SPL>
__device__ __forceinline__ void wrench_prefix(const Chain &ch, const Sites<SPL> &st, const SiteVals<SPL> &sv, V3 cref) {
  float w[SPL][6];
  float run[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < SPL; i++) {
    if (st.k[i] >= 0) {
      const V3 res = sv.res[i];
      const V3 f = mk3(-2.0f * (st.km[i].x * res.x), -2.0f * (st.km[i].y * res.y), -2.0f * (st.km[i].z * res.z));
      const V3 tq = cross3(sub3(sv.s[i], cref), f);
      const float v[6] = {f.x, f.y, f.z, tq.x, tq.y, tq.z};
#pragma unroll
      for (int c = 0; c < 6; c++) { run[c] = (i == 0) ? v[c] : run[c] + v[c]; w[i][c] = run[c]; }
    } else {
#pragma unroll
      for (int c = 0; c < 6; c++) w[i][c] = 0.f;
    }
  }
  float tot[6];
#pragma unroll
  for (int c = 0; c < 6; c++) tot[c] = run[c];
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
    for (int c = 0; c < 6; c++) {
      const float up = __shfl_up_sync(0xffffffffu, tot[c], off);
      if (ch.lane >= off) tot[c] = tot[c] + up;
    }
  }
#pragma unroll
  for (int c = 0; c < 6; c++) {
    const float ex = __shfl_up_sync(0xffffffffu, tot[c], 1);
#pragma unroll
    for (int i = 0; i < SPL; i++) {
      if (st.k[i] >= 0) {
        const float val = (ch.lane >= 1) ? ex + w[i][c] : w[i][c];
        ch.Ipre[6 * (ch.lane * ch.T.spl + i) + c] = val;
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void wrench_range(const Chain &ch, int sa, int se, V3 &F, V3 &Tq) {
  // subtree wrench = difference of inclusive prefixes; callers pass se > sa (empty ranges are clamped and discarded)
  float wr[6];
#pragma unroll
  for (int c = 0; c < 6; c++) {
    const float hi = ch.Ipre[6 * (se - 1) + c];
    const float lo = ch.Ipre[6 * (sa > 0 ? sa - 1 : 0) + c];
    wr[c] = (sa > 0) ? hi - lo : hi;
  }
  F = mk3(wr[0], wr[1], wr[2]);
  Tq = mk3(wr[3], wr[4], wr[5]);
}

// d loss / d (raw quaternion) of a normalised quaternion acted on by world torque `tau` (left-multiplied rotation);
// n is the RECIPROCAL of the normalisation divisor (normalize4)
__device__ __forceinline__ void quat_grad_left(Q4 qh, V3 tau, float n, float *g4) {
  Q4 h = qmul(mk4(0.f, tau.x, tau.y, tau.z), qh);
  h.w *= 2.0f; h.x *= 2.0f; h.y *= 2.0f; h.z *= 2.0f;
  const float pr = fmaf(qh.z, h.z, fmaf(qh.y, h.y, fmaf(qh.x, h.x, qh.w * h.w)));
  g4[0] = fmaf(-qh.w, pr, h.w) * n; g4[1] = fmaf(-qh.x, pr, h.x) * n;
  g4[2] = fmaf(-qh.y, pr, h.y) * n; g4[3] = fmaf(-qh.z, pr, h.z) * n;
}

// Reverse sweep: each body lane turns the subtree wrench of its joints into d loss / d qpos (into gbuf).
// Hinges are straight-line code with predicated stores; the primary free joint is evaluated by every lane
// (uniform) and stored by lane 0; other joint types take a divergent path skipped when the model has none.
template <int NB>
__device__ __forceinline__ void joint_grads(const Chain &ch, const Hot<NB> &H, const FkState<NB> &S, V3 cref, bool free_wanted) {
  if (free_wanted && ch.T.free_e >= 0 && ch.T.free_se > ch.T.free_sa) {  // uniform
    V3 F, Tq, fp;
    Q4 fq;
    wrench_range(ch, ch.T.free_sa, ch.T.free_se, F, Tq);
    gather_pose<NB>(ch, S, ch.T.free_e, fp, fq);
    const float fn = S.fd;
    const V3 Tp = sub3(Tq, cross3(sub3(fp, cref), F));
    float g4[4];
    quat_grad_left(fq, Tp, fn, g4);
    if (ch.lane == 0 && ch.g == 0) {
      float *g = ch.gbuf + ch.T.free_adr;
      g[0] = F.x; g[1] = F.y; g[2] = F.z; g[3] = g4[0]; g[4] = g4[1]; g[5] = g4[2]; g[6] = g4[3];
    }
  }
#pragma unroll
  for (int i = 0; i < NB; i++) {
    const BodyConst &b = H.bc[i];
    V3 pp;
    Q4 pq;
    gather_pose<NB>(ch, S, b.parent >= 0 ? b.parent : 0, pp, pq);
    const bool has_par = b.parent >= 0;
    pp = sel3(has_par, pp, mk3(0.f, 0.f, 0.f));            // top-level body: the world frame
    pq = sel4(has_par, pq, mk4(1.f, 0.f, 0.f, 0.f));
    // every joint of a body shares the body's subtree: one wrench, brought into the parent frame once
    const int sa = b.jsa[0], se = b.jse[0];
    const bool live = b.nj > 0 && se > sa;
    V3 F, Tq;
    wrench_range(ch, live ? sa : 0, live ? se : 1, F, Tq);
    const Q4 pc = mk4(pq.w, -pq.x, -pq.y, -pq.z);
    const V3 T0 = sub3(Tq, cross3(sub3(pp, cref), F));
    const V3 Fp = rotate(F, pc), Tp = rotate(T0, pc);
#pragma unroll
    for (int jj = 0; jj < JM; jj++) {
      const V3 al = S.keep.anchor[i][jj], xl = S.keep.axis[i][jj];
      if (H.hinge[i][jj] && live) ch.gbuf[b.jadr[jj]] = dot3(xl, sub3(Tp, cross3(al, Fp)));
      if (__builtin_expect(ch.T.any_other != 0, 0)) {  // uniform, rare
        if (H.other[i][jj] && live) {
          const int type = b.jtype[jj], adr = b.jadr[jj];
          if (type == STACB_JNT_SLIDE) {
            ch.gbuf[adr] = dot3(xl, Fp);
          } else if (type == STACB_JNT_BALL) {  // right-multiplied local rotation; last joint of its body
            const V3 A = add3(pp, rotate(al, pq));
            const V3 Ta = sub3(Tq, cross3(sub3(A, cref), F));
            const Q4 qb = S.Q[i];
            const V3 tl = rotate(Ta, mk4(qb.w, -qb.x, -qb.y, -qb.z));
            const Q4 ql = lds4(ch.qbuf + adr);
            Q4 h = qmul(ql, mk4(0.f, tl.x, tl.y, tl.z));
            h.w *= 2.0f; h.x *= 2.0f; h.y *= 2.0f; h.z *= 2.0f;
            const float pr = fmaf(ql.z, h.z, fmaf(ql.y, h.y, fmaf(ql.x, h.x, ql.w * h.w))), n = S.keep.fnorm[i];
            ch.gbuf[adr] = fmaf(-ql.w, pr, h.w) * n; ch.gbuf[adr + 1] = fmaf(-ql.x, pr, h.x) * n;
            ch.gbuf[adr + 2] = fmaf(-ql.y, pr, h.y) * n; ch.gbuf[adr + 3] = fmaf(-ql.z, pr, h.z) * n;
          } else {  // an additional free joint
            const V3 Tf = sub3(Tq, cross3(sub3(S.P[i], cref), F));
            float g4[4];
            quat_grad_left(S.Q[i], Tf, S.keep.fnorm[i], g4);
            float *g = ch.gbuf + adr;
            g[0] = F.x; g[1] = F.y; g[2] = F.z; g[3] = g4[0]; g[4] = g4[1]; g[5] = g4[2]; g[6] = g4[3];
          }
        }
      }
    }
  }
  ch.sync();
}

// Solver-side per-lane state: coordinate i = lane + 32*m lives in slot m.
template <int CPL>
struct Coords {
  float lb[CPL], ub[CPL];
  bool valid[CPL];
};

// Forward half of q_loss at `pt` (stac_core.py:27-63): FK state, marker residuals, loss.
// mask bit m set <=> slot m is optimised; elsewhere q0 is used (utils.make_qs).
template <int CPL, int NB, int SPL>
__device__ __forceinline__ float eval_fwd(const Chain &ch, const Coords<CPL> &co, const Hot<NB> &H, const float (&pt)[CPL],
                                          const float (&q0)[CPL], unsigned maskbits, const Sites<SPL> &st, FkState<NB> &S,
                                          SiteVals<SPL> &sv) {
  if (ch.g == 0) {
#pragma unroll
    for (int m = 0; m < CPL; m++)
      if (co.valid[m]) ch.qbuf[ch.lane + 32 * m] = ((maskbits >> m) & 1u) ? pt[m] : q0[m];
  }
  ch.sync();
  fk_hot<NB, true>(ch, H, S);
  return sites_loss<NB, SPL>(ch, S, st, sv);
}

// Reverse half: gradient of the loss at the point of the last eval_fwd (its state is still live).
template <int CPL, int NB, int SPL>
__device__ __forceinline__ void eval_bwd(const Chain &ch, const Coords<CPL> &co, const Hot<NB> &H, unsigned maskbits,
                                         const Sites<SPL> &st, const FkState<NB> &S, const SiteVals<SPL> &sv, float (&g)[CPL]) {
  V3 cref;
  Q4 cq;
  gather_pose<NB>(ch, S, 0, cref, cq);
  wrench_prefix<SPL>(ch, st, sv, cref);
  // the free-joint gradient is skipped when none of its coordinates is optimised (every part solve): the masked entries
  // of the gradient are zero by definition (stac_core.py:52, make_qs), so nothing observable changes
  bool mine = false;
#pragma unroll
  for (int m = 0; m < CPL; m++) {
    const int i = ch.lane + 32 * m;
    mine |= ((maskbits >> m) & 1u) && i >= ch.T.free_adr && i < ch.T.free_adr + 7;
  }
  const bool free_wanted = __any_sync(0xffffffffu, mine);
  joint_grads<NB>(ch, H, S, cref, free_wanted);
#pragma unroll
  for (int m = 0; m < CPL; m++) g[m] = (co.valid[m] && ((maskbits >> m) & 1u)) ? ch.gbuf[ch.lane + 32 * m] : 0.f;
}

template <int CPL>
__device__ __forceinline__ float lane_dot(const float (&a)[CPL], const float (&b)[CPL]) {
  float acc = a[0] * b[0];
#pragma unroll
  for (int m = 1; m < CPL; m++) acc = fmaf(a[m], b[m], acc);
  return acc;
}

struct SolveOut { float err; int iters, ls; bool bad; };

// jaxopt 0.8.5 ProjectedGradient.run (ProximalGradient._update_accel/_ls/_error, box projection), written as a
// two-state machine so the kernel contains ONE forward and ONE reverse evaluation (instruction-cache footprint):
//   state Y : the point is the FISTA extrapolation y -> loss + gradient, then the first line-search candidate
//   state LS: the point is a candidate x+ -> loss; rejected: halve the step and retry; accepted: gradient at x+ from
//             the state of this same evaluation (the reference recomputes FK there: same values), error, next y.
template <int CPL, int NB, int SPL>
__device__ __forceinline__ SolveOut solve(const Chain &ch, const Coords<CPL> &co, const Hot<NB> &H, const float (&q0)[CPL],
                                          unsigned maskbits, const Sites<SPL> &st, float tol, int maxiter, int maxls, float (&x)[CPL]) {
  float y[CPL], g[CPL], xn[CPL], d[CPL], gt[CPL];
#pragma unroll
  for (int m = 0; m < CPL; m++) { x[m] = co.valid[m] ? q0[m] : 0.f; y[m] = x[m]; xn[m] = x[m]; g[m] = 0.f; gt[m] = 0.f; }
  float t = 1.0f, step = 1.0f, stp = 1.0f, fy = 0.f, sq = 0.f, dg = 0.f;
  int halv = 0;
  bool in_ls = false;
  SolveOut out;
  out.iters = 0; out.ls = 0; out.bad = false; out.err = __int_as_float(0x7f800000);
  if (maxiter <= 0) return out;
  FkState<NB> S;
  SiteVals<SPL> sv;
  for (;;) {
    float pt[CPL];
#pragma unroll
    for (int m = 0; m < CPL; m++) pt[m] = in_ls ? xn[m] : y[m];
    const float f = eval_fwd<CPL, NB, SPL>(ch, co, H, pt, q0, maskbits, st, S, sv);
    bool rejected = false;
    if (in_ls) {
      out.ls++;
      if (!(f - f == 0.0f)) out.bad = true;
      const float dec = stp * (f - fy);
      const float cond = fmaf(stp, dg, 0.5f * sq);
      rejected = (dec > cond + 1.1920929e-07f) && (halv < maxls);
    }
    if (!rejected) {
      eval_bwd<CPL, NB, SPL>(ch, co, H, maskbits, st, S, sv, gt);
      if (in_ls) {  // accepted x+ = xn: FISTA update, error at x+ (unit-step fixed-point residual)
        step = (stp <= 1e-6f) ? 1.0f : stp / 0.5f;
        const float tn = 0.5f * (1.0f + sqrtf(fmaf(4.0f * t, t, 1.0f)));
        const float beta = (t - 1.0f) / tn;
#pragma unroll
        for (int m = 0; m < CPL; m++) {
          y[m] = fmaf(beta, xn[m] - x[m], xn[m]);
          d[m] = co.valid[m] ? clipf(xn[m] - gt[m], co.lb[m], co.ub[m]) - xn[m] : 0.f;
          x[m] = xn[m];
        }
        out.err = sqrtf(warp_sum(lane_dot<CPL>(d, d)));
        t = tn;
        out.iters++;
        if (!(out.err > tol && out.iters < maxiter)) break;
        in_ls = false;
        continue;
      }
      // y was evaluated: start the line search from the carried step size
      fy = f;
#pragma unroll
      for (int m = 0; m < CPL; m++) g[m] = gt[m];
      stp = step;
      halv = 0;
      in_ls = true;
    } else {
      stp = stp * 0.5f;
      halv++;
    }
    // next candidate x+ = clip(y - stp g) and the two line-search reductions (interleaved butterflies)
#pragma unroll
    for (int m = 0; m < CPL; m++) {
      xn[m] = co.valid[m] ? clipf(fmaf(-stp, g[m], y[m]), co.lb[m], co.ub[m]) : 0.f;
      d[m] = xn[m] - y[m];
    }
    sq = lane_dot<CPL>(d, d);
    dg = lane_dot<CPL>(d, g);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const float a = __shfl_xor_sync(0xffffffffu, sq, off), b2 = __shfl_xor_sync(0xffffffffu, dg, off);
      sq = sq + a;
      dg = dg + b2;
    }
  }
  return out;
}

// ------------------------------------------------------------------------------------------
// Latency mode: four warps (one per SM sub-partition) cooperate on ONE chain.
// Per FISTA iteration the reference's sequence  f(x+_0), f(x+_1), ..., grad f(x+), f(y'), grad f(y')  is
// evaluated speculatively in parallel:
//   warp 0: f at candidate x+_a (step s)        warp 2: f at y'(x+_a) = x+_a + beta (x+_a - x)
//   warp 1: f at candidate x+_b (step s/2)      warp 3: f at y'(x+_b)
// then the warp of the first accepted candidate k computes grad f(x+_k) (stopping criterion) while warp 2+k
// computes grad f(y') for the next iteration.  Every evaluation is the same arithmetic as in the one-warp solver,
// and the accepted candidate is chosen exactly as the sequential line search would, so results are bit-identical;
// only the critical path per iteration shrinks from 3 forward + 2 reverse evaluations to 1 + 1.
// ------------------------------------------------------------------------------------------
struct Coop {
  float *g;            // [nqp] broadcast of grad f(y')
  volatile float *sc;  // [0..1] accept flags, [2] err, [3] f(y'), [4] sticky non-finite flag
  int w;               // warp role 0..3
};

template <int CPL, int NB, int SPL>
__device__ __forceinline__ SolveOut solve4(const Chain &ch, const Coop &cp, const Coords<CPL> &co, const Hot<NB> &H, const float (&q0)[CPL],
                                           unsigned maskbits, const Sites<SPL> &st, float tol, int maxiter, int maxls, float (&x)[CPL]) {
  float y[CPL], g[CPL], gt[CPL], xa[CPL], xb[CPL], pt[CPL], d[CPL];
#pragma unroll
  for (int m = 0; m < CPL; m++) { x[m] = co.valid[m] ? q0[m] : 0.f; y[m] = x[m]; g[m] = 0.f; gt[m] = 0.f; }
  float t = 1.0f, step = 1.0f;
  SolveOut out;
  out.iters = 0; out.ls = 0; out.bad = false; out.err = __int_as_float(0x7f800000);
  if (maxiter <= 0) return out;
  FkState<NB> S;
  SiteVals<SPL> sv;
  const int w = cp.w;
  // f(y0), grad f(y0): every warp computes them (identical values, no broadcast needed)
  float fy = eval_fwd<CPL, NB, SPL>(ch, co, H, y, q0, maskbits, st, S, sv);
  eval_bwd<CPL, NB, SPL>(ch, co, H, maskbits, st, S, sv, g);
  int base = 0;
  float stp_a = step;
  // FISTA momentum of the current iteration; for the following iterations the two roles that idle during the reverse
  // phase compute it (same operations) and hand it over through cp.sc[5..6]
  float tn = 0.5f * (1.0f + sqrtf(fmaf(4.0f * t, t, 1.0f)));
  float beta = (t - 1.0f) / tn;
  for (;;) {
    const float stp_b = stp_a * 0.5f;
    const float stp_m = (w & 1) ? stp_b : stp_a;
    float sq = 0.f, dg = 0.f;
#pragma unroll
    for (int m = 0; m < CPL; m++) {
      xa[m] = co.valid[m] ? clipf(fmaf(-stp_a, g[m], y[m]), co.lb[m], co.ub[m]) : 0.f;
      xb[m] = co.valid[m] ? clipf(fmaf(-stp_b, g[m], y[m]), co.lb[m], co.ub[m]) : 0.f;
      const float mine = (w & 1) ? xb[m] : xa[m];
      pt[m] = (w < 2) ? mine : fmaf(beta, mine - x[m], mine);
      d[m] = mine - y[m];
    }
    if (w < 2) {
      sq = lane_dot<CPL>(d, d);
      dg = lane_dot<CPL>(d, g);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const float a = __shfl_xor_sync(0xffffffffu, sq, off), b2 = __shfl_xor_sync(0xffffffffu, dg, off);
        sq = sq + a;
        dg = dg + b2;
      }
    }
    const float f = eval_fwd<CPL, NB, SPL>(ch, co, H, pt, q0, maskbits, st, S, sv);
    if (w < 2) {
      const float dec = stp_m * (f - fy);
      const float cond = fmaf(stp_m, dg, 0.5f * sq);
      const bool rejected = (dec > cond + 1.1920929e-07f) && (base + w < maxls);
      if (ch.lane == 0 && ch.g == 0) {
        cp.sc[w] = rejected ? 0.f : 1.f;
        if (!(f - f == 0.0f)) cp.sc[4] = 1.f;
      }
    }
    __syncthreads();
    const bool a0 = cp.sc[0] != 0.f, a1 = cp.sc[1] != 0.f;
    if (!a0 && !a1) {  // both candidates rejected: next pair of step sizes
      base += 2;
      stp_a = stp_b * 0.5f;
      __syncthreads();
      continue;
    }
    const int k = a0 ? 0 : 1;
    out.ls += base + k + 1;  // evaluations the sequential line search performs
    const float stp_k = k ? stp_b : stp_a;
    if (w == k || w == 2 + k) {
      eval_bwd<CPL, NB, SPL>(ch, co, H, maskbits, st, S, sv, gt);
      if (w == k) {  // gradient at x+ -> unit-step fixed-point residual
#pragma unroll
        for (int m = 0; m < CPL; m++) {
          const float xk = k ? xb[m] : xa[m];
          d[m] = co.valid[m] ? clipf(xk - gt[m], co.lb[m], co.ub[m]) - xk : 0.f;
        }
        const float err = sqrtf(warp_sum(lane_dot<CPL>(d, d)));
        if (ch.lane == 0 && ch.g == 0) cp.sc[2] = err;
      } else {  // gradient at y' -> next iteration
        if (ch.g == 0) {
#pragma unroll
          for (int m = 0; m < CPL; m++)
            if (co.valid[m]) cp.g[ch.lane + 32 * m] = gt[m];
          if (ch.lane == 0) cp.sc[3] = f;
        }
      }
    } else if (w == 1 - k) {  // an idle role: next iteration's momentum, t_next = tn
      const float tn2 = 0.5f * (1.0f + sqrtf(fmaf(4.0f * tn, tn, 1.0f)));
      const float beta2 = (tn - 1.0f) / tn2;
      if (ch.lane == 0 && ch.g == 0) { cp.sc[5] = tn2; cp.sc[6] = beta2; }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < CPL; m++) {
      const float xk = k ? xb[m] : xa[m];
      y[m] = fmaf(beta, xk - x[m], xk);
      x[m] = xk;
      g[m] = co.valid[m] ? cp.g[ch.lane + 32 * m] : 0.f;
    }
    fy = cp.sc[3];
    out.err = cp.sc[2];
    t = tn;
    tn = cp.sc[5];
    beta = cp.sc[6];
    step = (stp_k <= 1e-6f) ? 1.0f : stp_k / 0.5f;
    out.iters++;
    if (!(out.err > tol && out.iters < maxiter)) break;
    base = 0;
    stp_a = step;
  }
  __syncthreads();
  return out;
}

// replace_qs (utils.py:147-169) as far as qpos is concerned: kinematics normalises free / ball
// quaternions in place. q holds the full qpos vector of the chain.
template <int CPL>
__device__ __forceinline__ void normalize_qpos(const Chain &ch, const Coords<CPL> &co, float (&q)[CPL]) {
  if (ch.T.nquat == 0) return;
  if (ch.g == 0) {
#pragma unroll
    for (int m = 0; m < CPL; m++)
      if (co.valid[m]) ch.qbuf[ch.lane + 32 * m] = q[m];
    __syncwarp();
    for (int j = ch.lane; j < ch.T.nquat; j += 32) {
      const int a = __ldg(ch.T.quat_adr + j);
      float dd;
      const Q4 qn = normalize4(lds4(ch.qbuf + a), &dd);
      ch.qbuf[a] = qn.w; ch.qbuf[a + 1] = qn.x; ch.qbuf[a + 2] = qn.y; ch.qbuf[a + 3] = qn.z;
    }
  }
  ch.sync();
#pragma unroll
  for (int m = 0; m < CPL; m++)
    if (co.valid[m]) q[m] = ch.qbuf[ch.lane + 32 * m];
  ch.sync();
}

template <int CPL>
__device__ __forceinline__ void coords_init(const Chain &ch, Coords<CPL> &co, const float *__restrict__ lb, const float *__restrict__ ub) {
#pragma unroll
  for (int m = 0; m < CPL; m++) {
    const int i = ch.lane + 32 * m;
    co.valid[m] = i < ch.T.nq;
    co.lb[m] = co.valid[m] && lb ? lb[i] : 0.f;
    co.ub[m] = co.valid[m] && ub ? ub[i] : 0.f;
  }
}

template <int CPL>
__device__ __forceinline__ unsigned mask_bits_u8(const Chain &ch, const Coords<CPL> &co, const uint8_t *__restrict__ mask /*[nq] or null=all*/) {
  unsigned b = 0;
#pragma unroll
  for (int m = 0; m < CPL; m++)
    if (co.valid[m] && (!mask || mask[ch.lane + 32 * m])) b |= 1u << m;
  return b;
}

// Full-model FK of q (normalised in place) and the per-frame outputs.
template <int CPL, int NBF, int SPL>
__device__ __forceinline__ void full_outputs(const Chain &ch, const Coords<CPL> &co, float (&q)[CPL], const Sites<SPL> &st,
                                             float *qpos_o, float *xpos_o, float *xquat_o, float *sites_o) {
#pragma unroll
  for (int m = 0; m < CPL; m++)
    if (co.valid[m]) ch.qbuf[ch.lane + 32 * m] = q[m];
  __syncwarp();
  V3 P[NBF];
  Q4 Q[NBF];
  fk_cold<NBF>(ch, ch.T.full, P, Q);
#pragma unroll
  for (int m = 0; m < CPL; m++)
    if (co.valid[m]) {
      q[m] = ch.qbuf[ch.lane + 32 * m];
      if (qpos_o) qpos_o[ch.lane + 32 * m] = q[m];
    }
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
  if (sites_o) {
#pragma unroll
    for (int i = 0; i < SPL; i++)
      if (st.k[i] >= 0) {
        const float *o = ch.PQ + 7 * st.efull[i];
        const V3 s = add3(lds3(o), rotate(st.off[i], lds4(o + 3)));
        sites_o[3 * st.k[i]] = s.x; sites_o[3 * st.k[i] + 1] = s.y; sites_o[3 * st.k[i] + 2] = s.z;
      }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

struct PoseArgs {
  const float *kp; float *qpos_io; const float *site_pos, *lb, *ub; const uint8_t *part_masks; int P;
  int do_root, root_kp_idx; const uint8_t *trunk_kps; int root_dims; float tol; int maxiter, maxls;
  float *qpos, *xpos, *xquat, *sites, *err; int32_t *iters, *ls_evals, *root_stats, *status; int C, F;
  int *counter;
  long long kp_stride;  // floats between the first frames of consecutive clips (F * 3K for packed clips; less when clips overlap)
};

// MODE 0: throughput (one warp per chain, 4 chains per CTA); MODE 1: latency (4 cooperating warps per chain);
// MODE 2: dense throughput (registers capped at 128 so 16 warps fit per SM; pays off from ~16 chains per SM);
// MODE 3: grouped latency for wide trees: the 4 roles of latency mode, each carried out by GRP = 3 member warps that
//         deal the bodies of the tree out among themselves (NB is then the number of bodies per lane of ONE member).
constexpr int GRP = 3;
// JMV = JM of the translation unit: part of the kernel's symbol, so variants that differ only in joint slots cannot clash at link time
template <int CPL, int NB, int NBF, int SPL, int MODE, int JMV = JM>
__global__ void __launch_bounds__(MODE == 3 ? 128 * GRP : 128, MODE == 2 ? 4 : 1) pose_clips_kernel(DevTree T, PoseArgs a) {
  constexpr bool COOP = MODE == 1 || MODE == 3;
  constexpr int G = MODE == 3 ? GRP : 1;
  extern __shared__ float smem[];
  __shared__ int s_chain;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int role = wib / G, gmem = wib % G;  // 4 roles per CTA: chains (throughput modes) or speculative evaluations (latency modes)
  float *warp_area = smem + (size_t)4 * role_smem_floats(T);
  Chain ch(T, smem + (size_t)role * role_smem_floats(T), warp_area + (size_t)wib * warp_smem_floats(T), lane, gmem, G, 1 + role);
  Coop cp;
  cp.g = warp_area + (size_t)4 * G * warp_smem_floats(T);
  cp.sc = cp.g + T.nqp;
  cp.w = role;
  const bool writer = !COOP || wib == 0;  // latency modes: every warp holds identical solver state, warp 0 writes the outputs
  Coords<CPL> co;
  coords_init<CPL>(ch, co, a.lb, a.ub);
  Sites<SPL> st;
  sites_init<SPL>(ch, st, a.site_pos);
  Hot<NB> H;
  hot_init<NB>(H, ch);
  const int nq = T.nq, K = T.K, nb = T.nbody, S1 = 1 + a.P;
  const unsigned full_bits = mask_bits_u8<CPL>(ch, co, nullptr);
  unsigned root_bits = 0;
#pragma unroll
  for (int m = 0; m < CPL; m++)
    if (co.valid[m] && lane + 32 * m < a.root_dims) root_bits |= 1u << m;
  const int n_root = a.do_root ? 2 : 0;                 // two root solves on frame 0 (compute_stac.py:64-98)
  const int n_pose = (a.do_root == 2) ? 0 : a.F;        // do_root == 2: root optimisation only
  const int n_stage = n_root + n_pose * S1;             // every stage is one FISTA solve: ONE call site below

  for (;;) {
    int c = 0;
    if (COOP) {
      if (threadIdx.x == 0) { s_chain = atomicAdd(a.counter, 1); cp.sc[4] = 0.f; }
      __syncthreads();
      c = s_chain;
      __syncthreads();
    } else {
      if (lane == 0) c = atomicAdd(a.counter, 1);
      c = __shfl_sync(0xffffffffu, c, 0);
    }
    if (c >= a.C) break;
    if (ch.g == 0) {
      for (int m = 0; m < CPL; m++)
        if (co.valid[m]) ch.gbuf[lane + 32 * m] = 0.f;
    }
    ch.sync();
    float q[CPL], q0[CPL], x[CPL];
#pragma unroll
    for (int m = 0; m < CPL; m++) q[m] = co.valid[m] ? a.qpos_io[(size_t)c * nq + lane + 32 * m] : 0.f;
    bool bad = false;
    const float *kpc = a.kp + (size_t)c * a.kp_stride;
    for (int sidx = 0; sidx < n_stage; sidx++) {
      const bool is_root = sidx < n_root;
      const int f = is_root ? 0 : (sidx - n_root) / S1;   // frame
      const int sg = is_root ? 0 : (sidx - n_root) % S1;  // 0: whole body, 1..P: INDIVIDUAL_PART_OPTIMIZATION masks
      unsigned bits;
      if (is_root) {
        if (sidx == 0) { sites_load_kp<SPL>(st, kpc); sites_mask_kp<SPL>(st, a.trunk_kps); }
        bits = root_bits;
      } else {
        if (sg == 0) { sites_load_kp<SPL>(st, kpc + (size_t)f * 3 * K); if (f == 0) sites_mask_kp<SPL>(st, nullptr); }
        bits = (sg == 0) ? full_bits : mask_bits_u8<CPL>(ch, co, a.part_masks + (size_t)(sg - 1) * nq);
      }
#pragma unroll
      for (int m = 0; m < CPL; m++) {
        q0[m] = q[m];
        const int i = lane + 32 * m;
        if (is_root && i < 3) q0[m] = kpc[3 * a.root_kp_idx + i];  // re-seed the root translation from the root keypoint
      }
      SolveOut so;
      if constexpr (COOP) so = solve4<CPL, NB, SPL>(ch, cp, co, H, q0, bits, st, a.tol, a.maxiter, a.maxls, x);
      else so = solve<CPL, NB, SPL>(ch, co, H, q0, bits, st, a.tol, a.maxiter, a.maxls, x);
#pragma unroll
      for (int m = 0; m < CPL; m++) q[m] = ((bits >> m) & 1u) ? x[m] : q0[m];  // utils.make_qs
      bad |= so.bad;
      if (is_root) {
        if (a.root_stats && lane == 0 && writer) { a.root_stats[4 * c + 2 * sidx] = so.iters; a.root_stats[4 * c + 2 * sidx + 1] = so.ls; }
        normalize_qpos<CPL>(ch, co, q);
      } else {
        const size_t fi = (size_t)c * a.F + f;
        if (a.iters && lane == 0 && writer) { a.iters[fi * S1 + sg] = so.iters; a.ls_evals[fi * S1 + sg] = so.ls; }
        // replace_qs: kinematics normalises the quaternions; the last stage's FK also yields the frame outputs
        if (sg < a.P || (role != 0 && COOP)) {
          normalize_qpos<CPL>(ch, co, q);
        } else {  // the role that writes: its member 0 runs the full-model FK (which also normalises), the others pick q up
          if (writer) {
            full_outputs<CPL, NBF, SPL>(ch, co, q, st, a.qpos ? a.qpos + fi * nq : nullptr, a.xpos ? a.xpos + fi * nb * 3 : nullptr,
                                        a.xquat ? a.xquat + fi * nb * 4 : nullptr, a.sites ? a.sites + fi * K * 3 : nullptr);
            if (a.err && lane == 0) a.err[fi] = so.err;
          }
          if (G > 1) {
            ch.sync();
            if (ch.g != 0) {
#pragma unroll
              for (int m = 0; m < CPL; m++)
                if (co.valid[m]) q[m] = ch.qbuf[lane + 32 * m];
            }
            ch.sync();
          }
        }
      }
    }
    if (COOP && threadIdx.x == 0) { if (cp.sc[4] != 0.f) bad = true; }  // thread 0 clears, reads and reports the flag: program order
    if (writer) {
#pragma unroll
      for (int m = 0; m < CPL; m++)
        if (co.valid[m]) a.qpos_io[(size_t)c * nq + lane + 32 * m] = q[m];
      if (a.status && lane == 0) a.status[c] = bad ? 1 : 0;
    }
  }
}

struct BatchArgs {
  const float *q, *q0, *kp, *site_pos, *lb, *ub; const uint8_t *q_mask, *kp_mask; float tol; int maxiter, maxls;
  float *out_a, *out_b, *out_c, *out_d; int32_t *iters, *ls_evals; int B; int mode;  // 0 fk, 1 loss_grad, 2 q_opt
};

template <int CPL, int NB, int NBF, int SPL, int JMV = JM>
__global__ void __launch_bounds__(128) batch_kernel(DevTree T, BatchArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wpb0 = blockDim.x >> 5;
  Chain ch(T, smem + (size_t)wib * role_smem_floats(T), smem + (size_t)wpb0 * role_smem_floats(T) + (size_t)wib * warp_smem_floats(T), lane, 0, 1,
           1 + wib);
  Coords<CPL> co;
  coords_init<CPL>(ch, co, a.lb, a.ub);
  Sites<SPL> st;
  sites_init<SPL>(ch, st, a.site_pos);
  Hot<NB> H;
  hot_init<NB>(H, ch);
  const int nq = T.nq, K = T.K, nb = T.nbody;
  const int wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + wib; b < a.B; b += gridDim.x * wpb) {
    float q[CPL], q0[CPL], g[CPL];
#pragma unroll
    for (int m = 0; m < CPL; m++) {
      q[m] = co.valid[m] ? a.q[(size_t)b * nq + lane + 32 * m] : 0.f;
      q0[m] = (co.valid[m] && a.q0) ? a.q0[(size_t)b * nq + lane + 32 * m] : q[m];
      g[m] = 0.f;
      if (co.valid[m]) ch.gbuf[lane + 32 * m] = 0.f;
    }
    __syncwarp();
    if (a.mode == 0) {
      full_outputs<CPL, NBF, SPL>(ch, co, q, st, a.out_a ? a.out_a + (size_t)b * nq : nullptr, a.out_b ? a.out_b + (size_t)b * nb * 3 : nullptr,
                                  a.out_c ? a.out_c + (size_t)b * nb * 4 : nullptr, a.out_d ? a.out_d + (size_t)b * K * 3 : nullptr);
    } else if (a.mode == 1) {
      sites_load_kp<SPL>(st, a.kp + (size_t)b * 3 * K);
      sites_mask_u8<SPL>(st, a.kp_mask);
      const unsigned bits = mask_bits_u8<CPL>(ch, co, a.q_mask);
      FkState<NB> S;
      SiteVals<SPL> sv;
      const float loss = eval_fwd<CPL, NB, SPL>(ch, co, H, q, q0, bits, st, S, sv);
      if (a.out_b) eval_bwd<CPL, NB, SPL>(ch, co, H, bits, st, S, sv, g);
      if (lane == 0) a.out_a[b] = loss;
      if (a.out_b) {
#pragma unroll
        for (int m = 0; m < CPL; m++)
          if (co.valid[m]) a.out_b[(size_t)b * nq + lane + 32 * m] = g[m];
      }
    } else {
      sites_load_kp<SPL>(st, a.kp + (size_t)b * 3 * K);
      sites_mask_u8<SPL>(st, a.kp_mask);
      const unsigned bits = mask_bits_u8<CPL>(ch, co, a.q_mask);
      float x[CPL];
      const SolveOut so = solve<CPL, NB, SPL>(ch, co, H, q, bits, st, a.tol, a.maxiter, a.maxls, x);
#pragma unroll
      for (int m = 0; m < CPL; m++)
        if (co.valid[m]) a.out_a[(size_t)b * nq + lane + 32 * m] = x[m];
      if (lane == 0) { a.out_b[b] = so.err; a.iters[b] = so.iters; a.ls_evals[b] = so.ls; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// m-phase (stac_core.py:146-165).  mode 0: sufficient statistics  s[k] = sum_t R_tk^T (y_tk - p_tk),  z2 = sum_t sum_k |y_tk - p_tk|^2;
// mode 1: data term of the objective at given offsets m,  sum_t sum_k |y_tk - p_tk - R_tk m_k|^2  (no cancellation).
// Fixed summation order: a warp sums a chunk of MCH consecutive frames in frame order, the chunk partials are summed in chunk order
// by the last CTA to finish (ticket), so the result is reproducible and lands in ONE contiguous buffer { s[3K], z2, (float)T }
// that a multi-GPU fit all-reduces in place right behind this kernel on the same stream.
// ------------------------------------------------------------------------------------------
constexpr int MCH = 8;
struct MArgs { const float *kp, *q, *m; float *scratch, *out; int T, mode; int *ticket; };

template <int CPL, int NB, int NBF, int SPL, int JMV = JM>
__global__ void __launch_bounds__(128) m_phase_kernel(DevTree T, MArgs a) {
  extern __shared__ float smem[];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  Chain ch(T, smem + (size_t)wib * role_smem_floats(T), smem + (size_t)wpb * role_smem_floats(T) + (size_t)wib * warp_smem_floats(T), lane, 0, 1, 1 + wib);
  Coords<CPL> co;
  coords_init<CPL>(ch, co, nullptr, nullptr);
  Sites<SPL> st;
  sites_init<SPL>(ch, st, a.mode == 1 ? a.m : nullptr);  // mode 1: st.off carries the offsets m
  Hot<NB> H;
  hot_init<NB>(H, ch);
  const int nq = T.nq, K = T.K;
  const int n3 = a.mode == 0 ? 3 * K + 1 : 1;
  const int nchunk = (a.T + MCH - 1) / MCH;
  for (int c = blockIdx.x * wpb + wib; c < nchunk; c += gridDim.x * wpb) {
    V3 acc3[SPL];
    float accz = 0.f;
#pragma unroll
    for (int i = 0; i < SPL; i++) acc3[i] = mk3(0.f, 0.f, 0.f);
    const int t1 = min(a.T, (c + 1) * MCH);
    for (int t = c * MCH; t < t1; t++) {
      const bool first = t == c * MCH;
      sites_load_kp<SPL>(st, a.kp + (size_t)t * 3 * K);
#pragma unroll
      for (int m = 0; m < CPL; m++)
        if (co.valid[m]) ch.qbuf[lane + 32 * m] = a.q[(size_t)t * nq + lane + 32 * m];
      __syncwarp();
      FkState<NB> S;
      fk_hot<NB, false>(ch, H, S);
      float e_lane = 0.f;
#pragma unroll
      for (int i = 0; i < SPL; i++) {
        V3 p;
        Q4 qb;
        gather_pose<NB>(ch, S, st.k[i] >= 0 ? st.eact[i] : 0, p, qb);
        if (st.k[i] >= 0) {
          const V3 z = mk3(st.kp[i].x - p.x, st.kp[i].y - p.y, st.kp[i].z - p.z);
          float e;
          if (a.mode == 0) {
            // math.quat_to_mat
            const float ww = qb.w * qb.w, xx = qb.x * qb.x, yy = qb.y * qb.y, zz = qb.z * qb.z;
            const float xy = qb.x * qb.y, xz = qb.x * qb.z, yz = qb.y * qb.z, wx = qb.w * qb.x, wy = qb.w * qb.y, wz = qb.w * qb.z;
            const float M00 = ww + xx - yy - zz, M01 = 2.0f * (xy - wz), M02 = 2.0f * (xz + wy);
            const float M10 = 2.0f * (xy + wz), M11 = ww - xx + yy - zz, M12 = 2.0f * (yz - wx);
            const float M20 = 2.0f * (xz - wy), M21 = 2.0f * (yz + wx), M22 = ww - xx - yy + zz;
            const V3 cv = mk3(fmaf(M20, z.z, fmaf(M10, z.y, M00 * z.x)), fmaf(M21, z.z, fmaf(M11, z.y, M01 * z.x)),
                              fmaf(M22, z.z, fmaf(M12, z.y, M02 * z.x)));
            acc3[i] = first ? cv : add3(acc3[i], cv);
            e = fmaf(z.z, z.z, fmaf(z.y, z.y, z.x * z.x));
          } else {
            const V3 r = sub3(z, rotate(st.off[i], qb));
            e = fmaf(r.z, r.z, fmaf(r.y, r.y, r.x * r.x));
          }
          e_lane = (i == 0) ? e : e_lane + e;
        }
      }
      const float ef = warp_sum(e_lane);
      accz = first ? ef : accz + ef;
      __syncwarp();
    }
    float *o = a.scratch + (size_t)c * n3;
    if (a.mode == 0) {
#pragma unroll
      for (int i = 0; i < SPL; i++)
        if (st.k[i] >= 0) { o[3 * st.k[i]] = acc3[i].x; o[3 * st.k[i] + 1] = acc3[i].y; o[3 * st.k[i] + 2] = acc3[i].z; }
    }
    if (lane == 0) o[n3 - 1] = accz;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int j = threadIdx.x; j < n3; j += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < nchunk; c++) {
      const float v = __ldcg(a.scratch + (size_t)c * n3 + j);
      acc = (c == 0) ? v : acc + v;
    }
    a.out[j] = acc;
  }
  if (a.mode == 0 && threadIdx.x == 0) a.out[n3] = (float)a.T;
}

}  // namespace stacb

