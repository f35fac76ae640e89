// stacb_abi.cu -- host side of libstacb.so: schedule derivation, variant dispatch, C ABI (include/stacb.h).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>

#include "stacb.h"
#include "stacb_device.cuh"
#include "stacb_variants.h"

namespace stacb {

__global__ void fma_peak_kernel(float *out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = 1.0f + 1e-3f * (float)(threadIdx.x + i);
  const float m = 0.9999f, c = 1e-4f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


}  // namespace stacb

// ------------------------------------------------------------------------------------------
// host side: schedule derivation + C ABI
// ------------------------------------------------------------------------------------------

using namespace stacb;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
namespace stacb {
int post_fail(int code, const char *msg) { return fail(code, msg); }
}  // namespace stacb
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(STACB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// RAII: entry points run on the handle's device and leave the caller's current device as they found it
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct stacb_tree {
  int device;
  DevTree T;
  int cpl, bpl_act, bpl_full, jm_act;
  bool fast_ok = false;                  // the register-resident hinge-tree solver applies (stacb_fast.cuh)
  int fast_rounds = 0;                   // pointer-jumping rounds of its element set
  int wide_W = 0;                        // > 0: the element set needs this many warps per evaluation (stacb_wide.cuh)
  std::atomic<int> mode{-1};             // scheduling override of stacb_pose_clips (stacb_tree_set_mode)
  std::atomic<int> path{0};              // 0 = register-resident solver where it applies, 1 = general kernels only
  std::vector<void *> allocs;
  int *counter;                          // pool of work counters: one per in-flight launch (launches may overlap on different streams)
  mutable std::atomic<unsigned> next{0};
};

constexpr int kCounterPool = 64;

static int ceil_log2(int x) { int r = 0; while ((1 << r) < x) r++; return r; }

template <class V>
static int upload(stacb_tree *t, const std::vector<V> &h, const V **out) {
  void *d = nullptr;
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(V);
  CUDA_TRY(cudaMalloc(&d, bytes));
  t->allocs.push_back(d);
  if (!h.empty()) CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice));
  *out = (const V *)d;
  return STACB_OK;
}

static int f2i(float f) { int i; memcpy(&i, &f, 4); return i; }

// Body-set records (parents precede children because body ids are DFS pre-order).
static void build_set(const stacb_tree_desc &m, const std::vector<int> &set, const std::vector<int> &subsize, const std::vector<int> &order,
                      std::vector<int> &rec, std::vector<int> &anc, int &rounds) {
  const int n = (int)set.size();
  std::vector<int> loc(m.nbody, -1), depth(m.nbody, 0);
  for (int b = 1; b < m.nbody; b++) depth[b] = depth[m.body_parent[b]] + 1;
  int maxd = 1;
  for (int e = 0; e < n; e++) { loc[set[e]] = e; maxd = std::max(maxd, depth[set[e]]); }
  rounds = ceil_log2(maxd);
  rec.assign((size_t)n * REC, 0);
  for (int e = 0; e < n; e++) {
    const int b = set[e];
    int *r = rec.data() + (size_t)e * REC;
    for (int c = 0; c < 3; c++) r[R_POS + c] = f2i(m.body_pos[3 * b + c]);
    for (int c = 0; c < 4; c++) r[R_QUAT + c] = f2i(m.body_quat[4 * b + c]);
    r[R_NJNT] = m.body_jntnum[b];
    const int p = m.body_parent[b];
    r[R_PARENT] = (p != 0) ? loc[p] : -1;
    r[R_BODY] = b;
    for (int jj = 0; jj < m.body_jntnum[b]; jj++) {
      const int j = m.body_jntadr[b] + jj;
      int *jr = r + R_JNT + J_STRIDE * jj;
      jr[J_TYPE] = m.jnt_type[j];
      jr[J_ADR] = m.jnt_qposadr[j];
      for (int c = 0; c < 3; c++) { jr[J_POS + c] = f2i(m.jnt_pos[3 * j + c]); jr[J_AXIS + c] = f2i(m.jnt_axis[3 * j + c]); }
      const int t = m.jnt_type[j];
      jr[J_REF] = f2i((t == STACB_JNT_HINGE || t == STACB_JNT_SLIDE) ? m.qpos0[m.jnt_qposadr[j]] : 0.f);
      // sorted-site range under this body
      const int lo = b, hi = b + subsize[b], K = m.nsite;
      int a = 0;
      while (a < K && m.site_body[order[a]] < lo) a++;
      int e2 = a;
      while (e2 < K && m.site_body[order[e2]] < hi) e2++;
      jr[J_SA] = a;
      jr[J_SE] = e2;
    }
  }
  const int nr = std::max(rounds, 1);
  anc.assign((size_t)nr * n, -1);
  for (int e = 0; e < n; e++) anc[e] = rec[(size_t)e * REC + R_PARENT];
  for (int r = 1; r < nr; r++)
    for (int e = 0; e < n; e++) {
      const int a = anc[(size_t)(r - 1) * n + e];
      anc[(size_t)r * n + e] = (a >= 0) ? anc[(size_t)(r - 1) * n + a] : -1;
    }
}

// ---- folding of jointless (welded) active bodies into their nearest jointed ancestor: float32 helpers written with explicit fused
// multiply-adds in the operation order the CPU oracle uses for the same folding, so both sides derive identical constants ----
static void h_cross(const float *a, const float *b, float *r) {
  r[0] = fmaf(a[1], b[2], -(a[2] * b[1])); r[1] = fmaf(a[2], b[0], -(a[0] * b[2])); r[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}
static void h_rotq(const float *v, const float *q /*w,x,y,z*/, float *out) {
  float t[3], c[3];
  h_cross(q + 1, v, t);
  h_cross(q + 1, t, c);
  for (int i = 0; i < 3; i++) out[i] = fmaf(2.0f, fmaf(q[0], t[i], c[i]), v[i]);
}
static void h_qmul(const float *u, const float *v, float *r) {
  r[0] = fmaf(-u[3], v[3], fmaf(-u[2], v[2], fmaf(-u[1], v[1], u[0] * v[0])));
  r[1] = fmaf(-u[3], v[2], fmaf(u[2], v[3], fmaf(u[1], v[0], u[0] * v[1])));
  r[2] = fmaf(u[3], v[1], fmaf(u[2], v[0], fmaf(-u[1], v[3], u[0] * v[2])));
  r[3] = fmaf(u[3], v[0], fmaf(-u[2], v[1], fmaf(u[1], v[2], u[0] * v[3])));
}
// pose of body `b` in the frame of its ancestor `top` (0 = world): constant transforms composed from the top down
static void rel_pose(const stacb_tree_desc &m, int top, int b, float *pos, float *quat) {
  std::vector<int> chain;
  for (int c = b; c != top; c = m.body_parent[c]) chain.push_back(c);
  pos[0] = pos[1] = pos[2] = 0.f; quat[0] = 1.f; quat[1] = quat[2] = quat[3] = 0.f;
  bool first = true;
  for (int i = (int)chain.size() - 1; i >= 0; i--) {
    const int c = chain[i];
    if (first) {
      for (int k = 0; k < 3; k++) pos[k] = m.body_pos[3 * c + k];
      for (int k = 0; k < 4; k++) quat[k] = m.body_quat[4 * c + k];
      first = false;
    } else {
      float r[3], q2[4];
      h_rotq(m.body_pos + 3 * c, quat, r);
      for (int k = 0; k < 3; k++) pos[k] = pos[k] + r[k];
      h_qmul(quat, m.body_quat + 4 * c, q2);
      for (int k = 0; k < 4; k++) quat[k] = q2[k];
    }
  }
}

// Element set with the jointless active bodies folded away.  Returns false when a keypoint site has no jointed ancestor-or-self.
static bool build_folded_set(const stacb_tree_desc &m, const std::vector<int> &act, const std::vector<int> &subsize, const std::vector<int> &order,
                             std::vector<int> &rec, std::vector<int> &anc, int &rounds, std::vector<int> &site_el, std::vector<float> &site_rel,
                             int &n_el) {
  const int K = m.nsite;
  std::vector<int> el;  // element -> body
  std::vector<int> loc(m.nbody, -1);
  for (int b : act)
    if (m.body_jntnum[b] > 0) { loc[b] = (int)el.size(); el.push_back(b); }
  n_el = (int)el.size();
  auto owner = [&](int b) { while (b != 0 && loc[b] < 0) b = m.body_parent[b]; return b; };  // nearest jointed ancestor-or-self (0: none)
  std::vector<int> edepth(n_el, 1);
  rec.assign((size_t)std::max(n_el, 1) * REC, 0);
  int maxd = 1;
  for (int e = 0; e < n_el; e++) {
    const int b = el[e];
    const int pb = owner(m.body_parent[b]);
    int *r = rec.data() + (size_t)e * REC;
    float pos[3], quat[4];
    rel_pose(m, pb, b, pos, quat);
    for (int c = 0; c < 3; c++) r[R_POS + c] = f2i(pos[c]);
    for (int c = 0; c < 4; c++) r[R_QUAT + c] = f2i(quat[c]);
    r[R_NJNT] = m.body_jntnum[b];
    r[R_PARENT] = pb != 0 ? loc[pb] : -1;
    r[R_BODY] = b;
    edepth[e] = pb != 0 ? edepth[loc[pb]] + 1 : 1;
    maxd = std::max(maxd, edepth[e]);
    for (int jj = 0; jj < m.body_jntnum[b]; jj++) {
      const int j = m.body_jntadr[b] + jj;
      int *jr = r + R_JNT + J_STRIDE * jj;
      jr[J_TYPE] = m.jnt_type[j];
      jr[J_ADR] = m.jnt_qposadr[j];
      for (int c = 0; c < 3; c++) { jr[J_POS + c] = f2i(m.jnt_pos[3 * j + c]); jr[J_AXIS + c] = f2i(m.jnt_axis[3 * j + c]); }
      const int t = m.jnt_type[j];
      jr[J_REF] = f2i((t == STACB_JNT_HINGE || t == STACB_JNT_SLIDE) ? m.qpos0[m.jnt_qposadr[j]] : 0.f);
      const int lo = b, hi = b + subsize[b];
      int a = 0;
      while (a < K && m.site_body[order[a]] < lo) a++;
      int e2 = a;
      while (e2 < K && m.site_body[order[e2]] < hi) e2++;
      jr[J_SA] = a;
      jr[J_SE] = e2;
    }
  }
  rounds = ceil_log2(maxd);
  const int nr = std::max(rounds, 1);
  anc.assign((size_t)nr * std::max(n_el, 1), -1);
  for (int e = 0; e < n_el; e++) anc[e] = rec[(size_t)e * REC + R_PARENT];
  for (int r = 1; r < nr; r++)
    for (int e = 0; e < n_el; e++) {
      const int a = anc[(size_t)(r - 1) * n_el + e];
      anc[(size_t)r * n_el + e] = (a >= 0) ? anc[(size_t)(r - 1) * n_el + a] : -1;
    }
  site_el.assign(K, 0);
  site_rel.assign((size_t)K * 7, 0.f);
  for (int p = 0; p < K; p++) {
    const int sb = m.site_body[order[p]], ob = owner(sb);
    if (ob == 0) return false;
    site_el[p] = loc[ob];
    rel_pose(m, ob, sb, site_rel.data() + 7 * p, site_rel.data() + 7 * p + 3);
  }
  return true;
}

extern "C" int stacb_tree_create(const stacb_tree_desc *d, int device, stacb_tree **out) {
  if (!d || !out) return fail(STACB_E_INVALID, "null argument");
  const stacb_tree_desc &m = *d;
  if (m.nbody < 2 || m.nq < 1 || m.njnt < 0 || m.nsite < 1) return fail(STACB_E_INVALID, "empty model");
  if (!m.body_parent || !m.body_jntadr || !m.body_jntnum || !m.body_pos || !m.body_quat || !m.qpos0 || !m.site_body ||
      (m.njnt > 0 && (!m.jnt_type || !m.jnt_qposadr || !m.jnt_pos || !m.jnt_axis)))
    return fail(STACB_E_INVALID, "null array in the model description");
  {  // every joint belongs to exactly one body's [jntadr, jntadr + jntnum) range and its coordinates lie inside qpos
    static const int width[4] = {7, 4, 1, 1};
    int owned = 0;
    for (int b = 0; b < m.nbody; b++) {
      const int n = m.body_jntnum[b];
      if (n < 0 || (n > 0 && (m.body_jntadr[b] < 0 || m.body_jntadr[b] + n > m.njnt)))
        return fail(STACB_E_INVALID, "body_jntadr / body_jntnum outside the joint arrays");
      owned += n;
    }
    if (owned != m.njnt || m.body_jntnum[0] != 0) return fail(STACB_E_INVALID, "joints must be owned by non-world bodies, each by exactly one");
    for (int j = 0; j < m.njnt; j++) {
      const int t = m.jnt_type[j];
      if (t < 0 || t > 3) return fail(STACB_E_INVALID, "unknown joint type");
      if (m.jnt_qposadr[j] < 0 || m.jnt_qposadr[j] + width[t] > m.nq) return fail(STACB_E_INVALID, "jnt_qposadr outside qpos");
    }
  }
  std::vector<int> path_stack{0};  // depth-first pre-order <=> every body's parent is on the current root-to-node path
  for (int b = 1; b < m.nbody; b++) {
    if (m.body_parent[b] < 0 || m.body_parent[b] >= b) return fail(STACB_E_INVALID, "body ids must be in depth-first pre-order");
    while (!path_stack.empty() && path_stack.back() != m.body_parent[b]) path_stack.pop_back();
    if (path_stack.empty()) return fail(STACB_E_INVALID, "body ids must be in depth-first pre-order (subtrees must be contiguous id ranges)");
    path_stack.push_back(b);
    if (m.body_jntnum[b] > JMAX) return fail(STACB_E_UNSUPPORTED, "more than 3 joints on one body");
    for (int jj = 0; jj < m.body_jntnum[b]; jj++) {
      const int j = m.body_jntadr[b] + jj, t = m.jnt_type[j];
      if (t == STACB_JNT_FREE && (m.body_parent[b] != 0 || m.body_jntnum[b] != 1))
        return fail(STACB_E_INVALID, "a free joint must be the only joint of a top-level body");
      if (t == STACB_JNT_BALL && jj != m.body_jntnum[b] - 1) return fail(STACB_E_UNSUPPORTED, "a ball joint must be the last joint of its body");
      if (t < 0 || t > 3) return fail(STACB_E_INVALID, "unknown joint type");
    }
  }
  for (int k = 0; k < m.nsite; k++)
    if (m.site_body[k] <= 0 || m.site_body[k] >= m.nbody) return fail(STACB_E_INVALID, "keypoint site must be attached to a non-world body");
  DeviceGuard guard(device);
  CUDA_TRY(guard.err);
  stacb_tree *t = new stacb_tree();
  t->device = device;
  const int nb = m.nbody, K = m.nsite;
  std::vector<int> subsize(nb, 0);
  for (int b = nb - 1; b >= 0; b--) { subsize[b] += 1; if (b > 0) subsize[m.body_parent[b]] += subsize[b]; }
  std::vector<char> isact(nb, 0);
  for (int k = 0; k < K; k++) { int b = m.site_body[k]; while (b != 0 && !isact[b]) { isact[b] = 1; b = m.body_parent[b]; } }
  std::vector<int> act, full;
  for (int b = 1; b < nb; b++) { full.push_back(b); if (isact[b]) act.push_back(b); }
  std::vector<int> order(K);
  for (int k = 0; k < K; k++) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return m.site_body[a] < m.site_body[b]; });
  std::vector<int> rec_a, anc_a, rec_f, anc_f;
  int ra, rf;
  build_set(m, act, subsize, order, rec_a, anc_a, ra);
  build_set(m, full, subsize, order, rec_f, anc_f, rf);
  std::vector<int> loc_a(nb, -1), loc_f(nb, -1), se_a(K), se_f(K);
  for (size_t e = 0; e < act.size(); e++) loc_a[act[e]] = (int)e;
  for (size_t e = 0; e < full.size(); e++) loc_f[full[e]] = (int)e;
  for (int p = 0; p < K; p++) { se_a[p] = loc_a[m.site_body[order[p]]]; se_f[p] = loc_f[m.site_body[order[p]]]; }
  DevTree &T = t->T;
  if (ra > RMAX || rf > RMAX) { delete t; return fail(STACB_E_UNSUPPORTED, "tree deeper than 256 levels"); }
  T.nbody = nb; T.nq = m.nq; T.njnt = m.njnt; T.K = K; T.spl = std::max(1, (K + 31) / 32);
  T.act.n = (int)act.size(); T.act.rounds = ra;
  T.full.n = (int)full.size(); T.full.rounds = rf;
  int rc;
  if ((rc = upload(t, rec_a, &T.act.rec)) || (rc = upload(t, anc_a, &T.act.anc)) || (rc = upload(t, rec_f, &T.full.rec)) ||
      (rc = upload(t, anc_f, &T.full.anc)) || (rc = upload(t, order, &T.site_order)) || (rc = upload(t, se_a, &T.site_eact)) ||
      (rc = upload(t, se_f, &T.site_efull))) {
    stacb_tree_destroy(t);
    return rc;
  }
  std::vector<int> qadr;
  for (int j = 0; j < m.njnt; j++) {
    if (m.jnt_type[j] == STACB_JNT_FREE) qadr.push_back(m.jnt_qposadr[j] + 3);
    if (m.jnt_type[j] == STACB_JNT_BALL) qadr.push_back(m.jnt_qposadr[j]);
  }
  T.nquat = (int)qadr.size();
  // primary free joint (handled warp-uniformly by the hot path) and the rare-joint flag
  T.free_e = -1; T.free_adr = 0; T.free_sa = 0; T.free_se = 0; T.any_other = 0;
  for (size_t e = 0; e < act.size(); e++) {
    const int *r = rec_a.data() + e * REC;
    for (int jj = 0; jj < r[R_NJNT]; jj++) {
      const int *jr = r + R_JNT + J_STRIDE * jj;
      if (jr[J_TYPE] == STACB_JNT_FREE && T.free_e < 0 && jj == 0) {
        T.free_e = (int)e; T.free_adr = jr[J_ADR]; T.free_sa = jr[J_SA]; T.free_se = jr[J_SE];
      } else if (jr[J_TYPE] != STACB_JNT_HINGE) {
        T.any_other = 1;
      }
    }
  }
  if ((rc = upload(t, qadr, &T.quat_adr))) {
    stacb_tree_destroy(t);
    return rc;
  }
  t->jm_act = 1;
  for (size_t e = 0; e < act.size(); e++) t->jm_act = std::max(t->jm_act, rec_a[e * REC + R_NJNT]);
  t->cpl = (m.nq + 31) / 32; t->bpl_act = (T.act.n + 31) / 32; t->bpl_full = (T.full.n + 31) / 32;
  T.nqp = 32 * t->cpl; T.pqn = std::max(T.act.n, T.full.n); T.npre = 32 * T.spl;
  // Register-resident solver: active subtree of at most 31 bodies with hinges only plus (optionally) the primary free joint,
  // at most 31 sites, no quaternion joint anywhere else.  Its solver slots cover the active hinges and the free joint; every
  // other qpos address is "passive" (zero gradient).
  {
    std::vector<char> covered(m.nq, 0);
    for (size_t e = 0; e < act.size(); e++) {
      const int *r = rec_a.data() + e * REC;
      for (int jj = 0; jj < r[R_NJNT]; jj++) {
        const int *jr = r + R_JNT + J_STRIDE * jj;
        if (jr[J_TYPE] == STACB_JNT_HINGE) covered[jr[J_ADR]] = 1;
      }
    }
    if (T.free_e >= 0)
      for (int i = 0; i < 7; i++) covered[T.free_adr + i] = 1;
    std::vector<int> passive;
    for (int i = 0; i < m.nq; i++)
      if (!covered[i]) passive.push_back(i);
    T.npassive = (int)passive.size();
    if ((rc = upload(t, passive, &T.passive))) {
      stacb_tree_destroy(t);
      return rc;
    }
    // element set: the active bodies themselves when they fit a warp (rodent, C. elegans: records shared with the general path),
    // otherwise the jointed active bodies with the welded ones folded in (fruitfly: 49-57 active bodies, 25 of them jointed)
    T.fs = T.act; T.fs_free_e = T.free_e; T.site_efs = T.site_eact; T.site_rel = nullptr;
    t->fast_rounds = T.act.rounds;
    bool fits_warp = T.act.n <= 31;
    if (!fits_warp) {
      std::vector<int> rec_e, anc_e, site_el;
      std::vector<float> site_rel;
      int re = 0, n_el = 0;
      if (build_folded_set(m, act, subsize, order, rec_e, anc_e, re, site_el, site_rel, n_el) && n_el <= 255 && n_el >= 1 && re <= RMAX) {
        T.fs.n = n_el; T.fs.rounds = re;
        const float *rel_dev = nullptr;
        if ((rc = upload(t, rec_e, &T.fs.rec)) || (rc = upload(t, anc_e, &T.fs.anc)) || (rc = upload(t, site_el, &T.site_efs)) ||
            (rc = upload(t, site_rel, &rel_dev))) {
          stacb_tree_destroy(t);
          return rc;
        }
        T.site_rel = rel_dev;
        T.fs_free_e = -1;
        for (int e = 0; e < n_el; e++)
          if (rec_e[(size_t)e * REC + R_BODY] == (T.free_e >= 0 ? rec_a[(size_t)T.free_e * REC + R_BODY] : -1)) T.fs_free_e = e;
        t->fast_rounds = re;
        fits_warp = (T.free_e < 0 || T.fs_free_e >= 0) && n_el <= 31;
        if (n_el > 31 && (T.free_e < 0 || (T.fs_free_e >= 0 && T.fs_free_e < 32))) t->wide_W = 2 * ((n_el + 1 + 63) / 64);  // 2, 4, 6 or 8 warps
      }
    }
    const bool common = !T.any_other && T.nquat == (T.free_e >= 0 ? 1 : 0) && (passive.empty() || passive[0] >= 3);
    t->fast_ok = fits_warp && K <= 31 && common;
    if (!(t->wide_W >= 2 && t->wide_W <= 8 && common && K <= 32 * t->wide_W - 1)) t->wide_W = 0;
  }
  void *cnt = nullptr;
  if (cudaMalloc(&cnt, kCounterPool * sizeof(int)) != cudaSuccess) { stacb_tree_destroy(t); return fail(STACB_E_CUDA, "cudaMalloc(counter)"); }
  t->allocs.push_back(cnt);
  t->counter = (int *)cnt;
  *out = t;
  return STACB_OK;
}

extern "C" void stacb_tree_destroy(stacb_tree *t) {
  if (!t) return;
  for (void *p : t->allocs) cudaFree(p);
  delete t;
}

extern "C" int stacb_tree_smem_per_chain(const stacb_tree *t) { return t ? chain_smem_floats(t->T) * 4 : 0; }

namespace stacb {
#define X(c, n, f, p, j)                                                                                                    \
  cudaError_t launch_pose_##c##_##n##_##f##_##p##_##j(const DevTree &, const PoseArgs &, int, int, size_t, int, cudaStream_t); \
  cudaError_t launch_batch_##c##_##n##_##f##_##p##_##j(const DevTree &, const BatchArgs &, int, int, size_t, cudaStream_t); \
  cudaError_t launch_mphase_##c##_##n##_##f##_##p##_##j(const DevTree &, const MArgs &, int, int, size_t, cudaStream_t);
STACB_VARIANTS(X)
#undef X
#define X(j, r, f)                                                                                                   \
  cudaError_t launch_fast_pose_##j##_##r##_##f(const DevTree &, const PoseArgs &, int, int, size_t, int, cudaStream_t); \
  cudaError_t launch_fast_batch_##j##_##r##_##f(const DevTree &, const BatchArgs &, int, int, cudaStream_t);
STACB_FAST_VARIANTS(X)
#undef X
#define STACB_WIDE_WARPS(X) X(2) X(4) X(6) X(8)
#define X(w)                                                                                            \
  cudaError_t launch_wide_pose_##w(const DevTree &, const PoseArgs &, int, size_t, int, cudaStream_t);   \
  cudaError_t launch_wide_batch_##w(const DevTree &, const BatchArgs &, int, cudaStream_t);
STACB_WIDE_WARPS(X)
#undef X
}  // namespace stacb

static bool fits_fast(const stacb_tree *t, int jm, int rt, int nbf) {
  return t->fast_ok && t->path.load() == 0 && t->jm_act <= jm && t->fast_rounds <= rt && t->bpl_full <= nbf;
}

static bool fits_wide(const stacb_tree *t) {
  return t->wide_W >= 2 && t->path.load() == 0 && t->jm_act <= 1 && t->fast_rounds <= 8 && t->bpl_full <= 8;
}

static bool fits(const stacb_tree *t, int cpl, int nb, int nbf, int spl, int jm) {
  return t->cpl <= cpl && t->bpl_act <= nb && t->bpl_full <= nbf && t->T.spl <= spl && t->jm_act <= jm;
}

static int run_pose(const stacb_tree *t, PoseArgs a, cudaStream_t s) {
  DeviceGuard guard(t->device);
  CUDA_TRY(guard.err);
  // each launch owns a work counter; the pool bounds the number of launches of one handle that may be in flight at once
  a.counter = t->counter + (t->next.fetch_add(1) % kCounterPool);
  CUDA_TRY(cudaMemsetAsync(a.counter, 0, sizeof(int), s));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device);
  const int g_force_mode = t->mode.load();
  if (fits_wide(t)) {  // wide hinge trees: W warps per chain, one chain per CTA (the scheduling mode does not apply)
    const size_t area = ((size_t)2 * t->T.nqp + 7 * (size_t)t->T.pqn) * 4;
    const int grid = std::min(a.C, sms * 8);
    // more chains than SMs: registers capped for two CTAs per SM.  Pair mode (two groups of W warps per chain, modes 1 / 4) is chosen
    // automatically for W = 2 only: measured on B200 (tools/wide_pair_probe.py, profiles/mode_sweep_r2f_wide_pair.txt) it is 1.25 x
    // faster at W = 2, even at W = 4 and slower from W = 6 (the mouse: 689 vs 637 ms) -- both groups share one SM's issue slots and
    // 384+ threads cap the registers at 168
    const int wsched = g_force_mode < 0 ? (a.C <= sms ? (t->wide_W == 2 ? 2 : 0) : 1) : (g_force_mode == 4 || g_force_mode == 1 ? 2 : (g_force_mode == 2 ? 1 : 0));
#define X(w) \
  if (t->wide_W == w) { CUDA_TRY(launch_wide_pose_##w(t->T, a, grid, area, wsched, s)); return STACB_OK; }
    STACB_WIDE_WARPS(X)
#undef X
  }
  {
    // register-resident solver: latency mode (2 * NC warps per chain) for few chains, throughput mode (one warp per chain)
    // for many, dense throughput mode (registers capped for 16 warps per SM) from 16 chains per SM
    // measured on B200 (profiles/mode_sweep_r2a.txt, mode_sweep_r2c.txt): latency mode wins up to one chain per SM, one warp per
    // chain while 8 warps per SM hold every chain, the dense variant (16 warps per SM) beyond that
    // and the pair mode (two warps per chain) in between, from more than one chain per SM up to four pairs per SM
    int sched = (a.C <= sms) ? 1 : (a.C <= 4 * sms ? 4 : (a.C > 8 * sms ? 2 : 0));
    if (g_force_mode >= 0) sched = g_force_mode;
    const int nc = sched == 1 ? 2 : (sched == 3 ? 3 : (sched == 4 ? 1 : 0));
    const size_t area = ((size_t)2 * t->T.nqp + 7 * (size_t)t->T.pqn) * 4;
    const int block = nc ? 64 * nc : 128;
    const size_t smem = nc ? area + 8192 : 4 * area;  // latency mode: + the exchange area of solve_coop (< 8 KB for 4 slots, NC <= 3)
    // persistent CTAs pull chains from the work counter; an unbalanced tail wave is kept on purpose: it runs at lower occupancy and
    // therefore faster per chain than a balanced one would (profiles/mode_sweep_r2c.txt)
    const int grid = nc ? std::min(a.C, sms * 8) : std::min((a.C + 3) / 4, sms * 16);
#define X(j, r, f) \
  if (fits_fast(t, j, r, f)) { CUDA_TRY(launch_fast_pose_##j##_##r##_##f(t->T, a, grid, block, smem, sched, s)); return STACB_OK; }
    STACB_FAST_VARIANTS(X)
#undef X
  }
  // Few chains: latency mode, one CTA of four cooperating warps per chain (speculative line search, see solve4).
  // Many chains: throughput mode, one warp per chain, four chains per CTA.
  const size_t chain_bytes = (size_t)chain_smem_floats(t->T) * 4;
  const size_t coop_extra = ((size_t)t->T.nqp + 8) * 4;
  const size_t coop_bytes = 4 * chain_bytes + coop_extra;
  const size_t grouped_bytes = (4 * (size_t)role_smem_floats(t->T) + 4 * GRP * (size_t)warp_smem_floats(t->T)) * 4 + coop_extra;
  int coop = (a.C <= 2 * sms) ? 1 : (a.C >= 16 * sms ? 2 : 0);
  if (g_force_mode >= 0) coop = g_force_mode;
  if (coop == 1 && g_force_mode < 0 && t->bpl_act > 1 && a.C <= sms) coop = 3;  // wide tree, few chains: grouped latency mode
  if (coop == 3 && (t->bpl_act <= 1 || grouped_bytes > 200 * 1024)) coop = 1;
  if (coop == 1 && coop_bytes > 200 * 1024) coop = 0;
  const int wpb = coop == 3 ? 4 * GRP : 4;
  const int grid = (coop == 1 || coop == 3) ? std::min(a.C, sms * 8) : std::min((a.C + wpb - 1) / wpb, sms * 16);
  const size_t smem = coop == 3 ? grouped_bytes : (coop == 1 ? coop_bytes : wpb * chain_bytes);
  if (smem > 200 * 1024) return fail(STACB_E_UNSUPPORTED, "model needs too much shared memory per chain");
#define X(c, n, f, p, j) \
  if (fits(t, c, n, f, p, j)) { CUDA_TRY(launch_pose_##c##_##n##_##f##_##p##_##j(t->T, a, grid, 32 * wpb, smem, coop, s)); return STACB_OK; }
  STACB_VARIANTS(X)
#undef X
  return fail(STACB_E_UNSUPPORTED, "model larger than every compiled kernel variant");
}

static int run_batch(const stacb_tree *t, const BatchArgs &a, cudaStream_t s) {
  DeviceGuard guard(t->device);
  CUDA_TRY(guard.err);
  if (a.B <= 0) return STACB_OK;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device);
  const int wpb = (a.B <= 2 * sms) ? 1 : 4;
  const int grid = std::min((a.B + wpb - 1) / wpb, sms * 16);
  if ((a.mode == 1 || a.mode == 2) && fits_wide(t)) {
    const int wgrid = std::min(a.B, sms * 8);
#define X(w) \
  if (t->wide_W == w) { CUDA_TRY(launch_wide_batch_##w(t->T, a, wgrid, s)); return STACB_OK; }
    STACB_WIDE_WARPS(X)
#undef X
  }
  if (a.mode == 1 || a.mode == 2) {  // q_loss / _q_opt share the arithmetic of the pose kernel that serves this model
#define X(j, r, f) \
  if (fits_fast(t, j, r, f)) { CUDA_TRY(launch_fast_batch_##j##_##r##_##f(t->T, a, grid, 32 * wpb, s)); return STACB_OK; }
    STACB_FAST_VARIANTS(X)
#undef X
  }
  const size_t smem = (size_t)wpb * chain_smem_floats(t->T) * 4;
  if (smem > 200 * 1024) return fail(STACB_E_UNSUPPORTED, "model needs too much shared memory per chain");
#define X(c, n, f, p, j) \
  if (fits(t, c, n, f, p, j)) { CUDA_TRY(launch_batch_##c##_##n##_##f##_##p##_##j(t->T, a, grid, 32 * wpb, smem, s)); return STACB_OK; }
  STACB_VARIANTS(X)
#undef X
  return fail(STACB_E_UNSUPPORTED, "model larger than every compiled kernel variant");
}

extern "C" int stacb_fk(const stacb_tree *t, const float *qpos, const float *site_pos, float *qpos_out, float *xpos, float *xquat,
                        float *site_xpos, int B, void *stream) {
  if (!t || !qpos || !site_pos || B < 0) return fail(STACB_E_INVALID, "stacb_fk: bad argument");
  BatchArgs a{};
  a.q = qpos; a.site_pos = site_pos; a.out_a = qpos_out; a.out_b = xpos; a.out_c = xquat; a.out_d = site_xpos; a.B = B; a.mode = 0;
  return run_batch(t, a, (cudaStream_t)stream);
}

extern "C" int stacb_loss_grad(const stacb_tree *t, const float *q, const float *q0, const float *kp, const uint8_t *q_mask,
                               const uint8_t *kp_mask, const float *site_pos, float *loss, float *grad, int B, void *stream) {
  if (!t || !q || !kp || !site_pos || !loss || B < 0) return fail(STACB_E_INVALID, "stacb_loss_grad: bad argument");
  BatchArgs a{};
  a.q = q; a.q0 = q0; a.kp = kp; a.q_mask = q_mask; a.kp_mask = kp_mask; a.site_pos = site_pos; a.out_a = loss; a.out_b = grad; a.B = B; a.mode = 1;
  return run_batch(t, a, (cudaStream_t)stream);
}

extern "C" int stacb_q_opt(const stacb_tree *t, const float *q0, const float *kp, const uint8_t *q_mask, const uint8_t *kp_mask,
                           const float *site_pos, const float *lb, const float *ub, float tol, int maxiter, int maxls, float *params,
                           float *error, int32_t *iters, int32_t *ls_evals, int B, void *stream) {
  if (!t || !q0 || !kp || !site_pos || !lb || !ub || !params || !error || !iters || !ls_evals || B < 0)
    return fail(STACB_E_INVALID, "stacb_q_opt: bad argument");
  BatchArgs a{};
  a.q = q0; a.kp = kp; a.q_mask = q_mask; a.kp_mask = kp_mask; a.site_pos = site_pos; a.lb = lb; a.ub = ub; a.tol = tol; a.maxiter = maxiter;
  a.maxls = maxls; a.out_a = params; a.out_b = error; a.iters = iters; a.ls_evals = ls_evals; a.B = B; a.mode = 2;
  return run_batch(t, a, (cudaStream_t)stream);
}

extern "C" int stacb_pose_session(const stacb_tree *t, const float *kp, int clip_stride, float *qpos_io, const float *site_pos, const float *lb,
                                  const float *ub, const uint8_t *part_masks, int P, int do_root, int root_kp_idx, const uint8_t *trunk_kps,
                                  int root_dims, float tol, int maxiter, int maxls, float *qpos, float *xpos, float *xquat, float *sites,
                                  float *err, int32_t *iters, int32_t *ls_evals, int32_t *root_stats, int32_t *status, int C, int F, void *stream) {
  if (!t || !kp || !qpos_io || !site_pos || !lb || !ub || C < 0 || F < 0 || P < 0 || clip_stride < 0)
    return fail(STACB_E_INVALID, "stacb_pose_clips: bad argument");
  if (P > 0 && !part_masks) return fail(STACB_E_INVALID, "stacb_pose_clips: part_masks is null");
  if (do_root && (root_kp_idx < 0 || root_kp_idx >= t->T.K || F < 1)) return fail(STACB_E_INVALID, "stacb_pose_clips: bad root keypoint");
  if ((iters == nullptr) != (ls_evals == nullptr)) return fail(STACB_E_INVALID, "stacb_pose_clips: iters and ls_evals go together");
  if (C == 0) return STACB_OK;
  PoseArgs a{};
  a.kp = kp; a.qpos_io = qpos_io; a.site_pos = site_pos; a.lb = lb; a.ub = ub; a.part_masks = part_masks; a.P = P; a.do_root = do_root;
  a.root_kp_idx = root_kp_idx; a.trunk_kps = trunk_kps; a.root_dims = root_dims; a.tol = tol; a.maxiter = maxiter; a.maxls = maxls;
  a.qpos = qpos; a.xpos = xpos; a.xquat = xquat; a.sites = sites; a.err = err; a.iters = iters; a.ls_evals = ls_evals;
  a.root_stats = root_stats; a.status = status; a.C = C; a.F = F; a.counter = t->counter;
  a.kp_stride = (long long)clip_stride * 3 * t->T.K;
  return run_pose(t, a, (cudaStream_t)stream);
}

extern "C" int stacb_pose_clips(const stacb_tree *t, const float *kp, float *qpos_io, const float *site_pos, const float *lb, const float *ub,
                                const uint8_t *part_masks, int P, int do_root, int root_kp_idx, const uint8_t *trunk_kps, int root_dims,
                                float tol, int maxiter, int maxls, float *qpos, float *xpos, float *xquat, float *sites, float *err,
                                int32_t *iters, int32_t *ls_evals, int32_t *root_stats, int32_t *status, int C, int F, void *stream) {
  return stacb_pose_session(t, kp, F, qpos_io, site_pos, lb, ub, part_masks, P, do_root, root_kp_idx, trunk_kps, root_dims, tol, maxiter, maxls,
                            qpos, xpos, xquat, sites, err, iters, ls_evals, root_stats, status, C, F, stream);
}

static int run_mphase(const stacb_tree *t, MArgs a, cudaStream_t s) {
  DeviceGuard guard(t->device);
  CUDA_TRY(guard.err);
  a.ticket = t->counter + (t->next.fetch_add(1) % kCounterPool);
  CUDA_TRY(cudaMemsetAsync(a.ticket, 0, sizeof(int), s));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device);
  const int nchunk = (a.T + MCH - 1) / MCH;
  const int wpb = 4;
  const int grid = std::max(1, std::min((nchunk + wpb - 1) / wpb, sms * 4));
  const size_t smem = (size_t)wpb * chain_smem_floats(t->T) * 4;
  if (smem > 200 * 1024) return fail(STACB_E_UNSUPPORTED, "model needs too much shared memory per chain");
#define X(c, n, f, p, j) \
  if (fits(t, c, n, f, p, j)) { CUDA_TRY(launch_mphase_##c##_##n##_##f##_##p##_##j(t->T, a, grid, 32 * wpb, smem, s)); return STACB_OK; }
  STACB_VARIANTS(X)
#undef X
  return fail(STACB_E_UNSUPPORTED, "model larger than every compiled kernel variant");
}

extern "C" int stacb_m_stats(const stacb_tree *t, const float *kp, const float *q, float *scratch, float *out, int T, void *stream) {
  if (!t || !kp || !q || !scratch || !out || T < 0) return fail(STACB_E_INVALID, "stacb_m_stats: bad argument");
  MArgs a{};
  a.kp = kp; a.q = q; a.m = nullptr; a.scratch = scratch; a.out = out; a.T = T; a.mode = 0;
  return run_mphase(t, a, (cudaStream_t)stream);
}

extern "C" int stacb_m_residual(const stacb_tree *t, const float *kp, const float *q, const float *m, float *scratch, float *out, int T,
                                void *stream) {
  if (!t || !kp || !q || !m || !scratch || !out || T < 0) return fail(STACB_E_INVALID, "stacb_m_residual: bad argument");
  MArgs a{};
  a.kp = kp; a.q = q; a.m = m; a.scratch = scratch; a.out = out; a.T = T; a.mode = 1;
  return run_mphase(t, a, (cudaStream_t)stream);
}

extern "C" int stacb_m_scratch_floats(const stacb_tree *t, int T) {
  if (!t || T < 0) return 0;
  return std::max(1, (T + MCH - 1) / MCH) * (3 * t->T.K + 1);
}

extern "C" int stacb_fma_peak(float *out, int blocks, int threads, int iters, void *stream) {
  if (!out || blocks <= 0 || threads <= 0 || iters < 0) return fail(STACB_E_INVALID, "stacb_fma_peak: bad argument");
  fma_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters);
  CUDA_TRY(cudaGetLastError());
  return STACB_OK;
}

extern "C" int stacb_tree_set_mode(stacb_tree *t, int mode) {
  if (!t || mode < -1 || mode > 4)
    return fail(STACB_E_INVALID, "stacb_tree_set_mode: mode must be -1 (auto), 0 (throughput), 1 (latency), 2 (dense throughput), 3 (wide latency) or 4 (pair)");
  t->mode.store(mode);
  return STACB_OK;
}

extern "C" int stacb_tree_set_path(stacb_tree *t, int path) {
  if (!t || path < 0 || path > 1) return fail(STACB_E_INVALID, "stacb_tree_set_path: path must be 0 (auto) or 1 (general kernels only)");
  t->path.store(path);
  return STACB_OK;
}

extern "C" int stacb_tree_path(const stacb_tree *t) {
  if (!t) return -1;
  if (fits_wide(t)) return 1;
#define X(j, r, f) \
  if (fits_fast(t, j, r, f)) return 1;
  STACB_FAST_VARIANTS(X)
#undef X
  return 0;
}

extern "C" const char *stacb_last_error(void) { return g_err.c_str(); }
extern "C" int stacb_version(void) { return STACB_VERSION; }
