/*
 * stacb.h -- C ABI of the B200-native STAC fitting hot path (libstacb.so).
 *
 * Drop-in boundary for talmolab/stac-mjx's solver path.  The reference has no
 * native plugin interface (it is pure Python over MJX + jaxopt); the seam this
 * library replaces is the duck-typed solver object and its three drivers:
 *
 *   stac_mjx/stac_core.py:175-275   class StacCore { q_opt, m_opt }
 *   stac_mjx/stac_core.py:27-63     q_loss            -> stacb_loss_grad
 *   stac_mjx/stac_core.py:66-99     _q_opt            -> stacb_q_opt
 *   stac_mjx/stac_core.py:102-172   _m_opt            -> stacb_m_stats (+ closed form on the host)
 *   stac_mjx/utils.py:49-74,147-169 kinematics / replace_qs -> stacb_fk
 *   stac_mjx/compute_stac.py:17-104   root_optimization  \
 *   stac_mjx/compute_stac.py:170-278  pose_optimization   > stacb_pose_clips (fused, one launch per pass)
 *   stac_mjx/stac.py:405-440          the two jax.vmap over clips /
 *
 * Conventions
 *   - plain C types only; every array argument is a DEVICE pointer unless marked host;
 *     row-major, contiguous, float32 / int32 / uint8, 4-byte aligned.
 *   - the caller owns every buffer; entry points never allocate (after tree creation),
 *     never synchronise, and enqueue on the caller's stream (`stream` is a cudaStream_t
 *     passed as void*; NULL = legacy default stream).
 *   - return value: 0 on success, negative on error (STACB_E_*); stacb_last_error()
 *     returns a thread-local message for the last failing call.
 *   - no process-global state: compute entry points take a const handle and are re-entrant (up to 64 launches of one
 *     handle may be in flight at once); the two scheduling knobs (stacb_tree_set_mode / _set_path) live in the handle;
 *     the caller's current CUDA device is restored before every entry point returns.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * All arithmetic is float32 in a fixed ("canonical") operation order documented in
 * DESIGN.md, so results are reproducible bit-for-bit across launches, grid shapes and GPUs.
 */
#ifndef STACB_H_
#define STACB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STACB_VERSION 200

#define STACB_OK 0
#define STACB_E_INVALID (-1)     /* bad argument / unsupported model feature */
#define STACB_E_CUDA (-2)        /* CUDA runtime error (see stacb_last_error) */
#define STACB_E_UNSUPPORTED (-3) /* model too large for the compiled kernel variants */

/* mujoco.mjtJoint */
#define STACB_JNT_FREE 0
#define STACB_JNT_BALL 1
#define STACB_JNT_SLIDE 2
#define STACB_JNT_HINGE 3

/* Host-side model description: the MjModel fields the path reads
 * (reference stac.py:113-133,219-235; MJX smooth.kinematics).  All HOST pointers. */
typedef struct {
  int32_t nbody, nq, njnt, nsite;  /* nsite = K keypoint sites */
  const int32_t *body_parent;      /* [nbody], body 0 = world */
  const int32_t *body_jntadr;      /* [nbody], -1 if none */
  const int32_t *body_jntnum;      /* [nbody] */
  const float *body_pos;           /* [nbody,3] (already scaled, rescale.py:24-25) */
  const float *body_quat;          /* [nbody,4] w,x,y,z */
  const int32_t *jnt_type;         /* [njnt] */
  const int32_t *jnt_qposadr;      /* [njnt] */
  const int32_t *jnt_bodyid;       /* [njnt] */
  const float *jnt_pos;            /* [njnt,3] */
  const float *jnt_axis;           /* [njnt,3] unit */
  const float *qpos0;              /* [nq] */
  const int32_t *site_body;        /* [K] body id of each keypoint site, keypoint order
                                      (= mj_model.site_bodyid[site_idxs], stac_core.py:146) */
} stacb_tree_desc;

typedef struct stacb_tree stacb_tree;

/* Parse the tree once, derive the execution schedule and copy it to `device`. */
int stacb_tree_create(const stacb_tree_desc *host_desc, int device, stacb_tree **out);
void stacb_tree_destroy(stacb_tree *tree);

/* Bytes of dynamic shared memory per chain (warp) the solver kernel uses; informational. */
int stacb_tree_smem_per_chain(const stacb_tree *tree);

/* utils.kinematics + get_site_xpos for B independent qpos.
 *   qpos [B,nq]  site_pos [K,3]
 *   qpos_out [B,nq] (free/ball quaternions normalised, as MJX writes them back)  may be NULL
 *   xpos [B,nbody,3]  xquat [B,nbody,4]  site_xpos [B,K,3]                         may be NULL */
int stacb_fk(const stacb_tree *tree, const float *qpos, const float *site_pos, float *qpos_out, float *xpos,
             float *xquat, float *site_xpos, int B, void *stream);

/* stac_core.q_loss and d loss / d q for B independent (q, q0, kp) triples.
 *   q, q0 [B,nq]  kp [B,3K]  q_mask [nq] u8  kp_mask [3K] u8  site_pos [K,3]
 *   loss [B]  grad [B,nq] (NULL to skip the reverse sweep) */
int stacb_loss_grad(const stacb_tree *tree, const float *q, const float *q0, const float *kp, const uint8_t *q_mask,
                    const uint8_t *kp_mask, const float *site_pos, float *loss, float *grad, int B, void *stream);

/* stac_core._q_opt: B independent box-constrained FISTA solves (jaxopt 0.8.5 ProjectedGradient
 * defaults: backtracking line search, maxls halvings, decrease_factor 0.5, acceleration).
 *   q0 [B,nq]  kp [B,3K]  q_mask [nq]  kp_mask [3K]  lb, ub [nq]
 *   params [B,nq] (= res.params)  error [B] (= res.state.error)  iters [B]  ls_evals [B] */
int stacb_q_opt(const stacb_tree *tree, const float *q0, const float *kp, const uint8_t *q_mask, const uint8_t *kp_mask,
                const float *site_pos, const float *lb, const float *ub, float tol, int maxiter, int maxls, float *params,
                float *error, int32_t *iters, int32_t *ls_evals, int B, void *stream);

/* Fused root_optimization (frame 0, optional) + pose_optimization over C independent clips of F
 * frames: one persistent kernel, one warp per clip chain, frames strictly sequential inside.
 *   kp [C,F,3K]
 *   qpos_io [C,nq]   in: mjx_data.qpos on entry of each clip;  out: qpos after the last frame
 *   part_masks [P,nq] u8 (INDIVIDUAL_PART_OPTIMIZATION masks, stac.py:161-183), P may be 0
 *   do_root: 0 = pose only, 1 = root optimisation on frame 0 then pose, 2 = root optimisation only
 *   root_kp_idx, trunk_kps [K] u8, root_dims (7 free / 4 slide; compute_stac.py:51-54)
 *   qpos [C,F,nq]  xpos [C,F,nbody,3]  xquat [C,F,nbody,4]  sites [C,F,K,3]  err [C,F]
 *   iters, ls_evals [C,F,1+P]  root_stats [C,4] = {iters,ls} of the two root solves
 *   status [C]: 0 ok, 1 = a non-finite loss was seen in the clip
 * Any output pointer except qpos may be NULL. */
int stacb_pose_clips(const stacb_tree *tree, const float *kp, float *qpos_io, const float *site_pos, const float *lb,
                     const float *ub, const uint8_t *part_masks, int P, int do_root, int root_kp_idx,
                     const uint8_t *trunk_kps, int root_dims, float tol, int maxiter, int maxls, float *qpos, float *xpos,
                     float *xquat, float *sites, float *err, int32_t *iters, int32_t *ls_evals, int32_t *root_stats,
                     int32_t *status, int C, int F, void *stream);

/* The same launch on clips that OVERLAP in memory: clip c starts `clip_stride` frames after clip c-1 in one session buffer
 *   kp [(C-1)*clip_stride + F, 3K]
 * so the 10-frame look-ahead of the reference's `continuous` batching (utils.batch_kp_data, stac_mjx/utils.py:350-389: windows
 * of n_frames_per_clip + 10 frames every n_frames_per_clip frames) is read in place by the kernel instead of being
 * materialised per clip on the host (clip_stride = n_frames_per_clip, F = n_frames_per_clip + 10; the caller appends the
 * reference's 10 wrap-padded frames after the last clip).  clip_stride = F is stacb_pose_clips. */
int stacb_pose_session(const stacb_tree *tree, const float *kp, int clip_stride, float *qpos_io, const float *site_pos,
                       const float *lb, const float *ub, const uint8_t *part_masks, int P, int do_root, int root_kp_idx,
                       const uint8_t *trunk_kps, int root_dims, float tol, int maxiter, int maxls, float *qpos, float *xpos,
                       float *xquat, float *sites, float *err, int32_t *iters, int32_t *ls_evals, int32_t *root_stats,
                       int32_t *status, int C, int F, void *stream);

/* _m_opt sufficient statistics over T frames (stac_core.py:146-159):
 *   s[k,i] = sum_t sum_j R_tk[j,i] (y_tk[j] - p_tk[j]),   z2 = sum_t sum_k |y_tk - p_tk|^2.
 *   kp [T,3K]  q [T,nq]  scratch [stacb_m_scratch_floats(tree, T)]
 *   out [3K+2] = { s[K,3], z2, (float)T }: ONE contiguous buffer, so the ranks of a multi-GPU fit all-reduce it in place
 *   right behind this call on the same stream (3K+2 floats over NVLink) and apply the closed form redundantly.
 * One kernel launch; frames are summed in a fixed order (chunks of 8 consecutive frames in frame order, then the chunk
 * partials in chunk order): deterministic and independent of the grid. */
int stacb_m_stats(const stacb_tree *tree, const float *kp, const float *q, float *scratch, float *out, int T, void *stream);

/* Data term of the m-phase objective at the offsets m [K,3] (stac_core.py:160-165), evaluated from the residuals
 *   out[0] = sum_t sum_k |y_tk - p_tk - R_tk m_k|^2
 * instead of the reference's expanded form z2 - 2 sum(m s) + T sum(m^2): the same number without the cancellation of
 * three O(z2) terms in float32 (the reference's identity-pose KAT asserts < 1e-8).  Same buffers and order as stacb_m_stats. */
int stacb_m_residual(const stacb_tree *tree, const float *kp, const float *q, const float *m, float *scratch, float *out, int T,
                     void *stream);
int stacb_m_scratch_floats(const stacb_tree *tree, int T);

/* Device epilogues of the IK pass (run on the packed device outputs before the single device-to-host copy).
 *
 * stacb_edge_crossfade -- utils.handle_edge_effects (stac_mjx/utils.py:393-461) for one packed array of `continuous` clips:
 *   in [C, F+ov, D] -> out [stacb_edge_rows(C,F,ov), D]: the last ov frames of clip c are blended with the first ov frames of clip
 *   c+1 ((1-w) a + w b in float64, rounded to float32 as numpy does), then first clip whole, middle clips [ov:], last clip [ov:-ov].
 *   w [ov] float64 (device): the reference's sigmoid 0.5 (1 + tanh(10 (x - 0.5) / 2)), x = linspace(0, 1, ov).
 * stacb_qvel -- utils.compute_velocity_from_kinematics (stac_mjx/utils.py:302-347) for C continuous clips:
 *   qpos [C,F,nq] -> qvel [C,F,nv], nv = nq-1 with a free joint (first 7 qpos), else nq; the last frame of a clip gets zero velocity.
 * Neither takes a handle: they launch on the calling thread's current CUDA device, which must be the device `stream` belongs to.
 * w may be null when ov == 0 (nothing is blended). */
long long stacb_edge_rows(int C, int F, int ov);
int stacb_edge_crossfade(const float *in, const double *w, float *out, int C, int F, int ov, int D, void *stream);
int stacb_qvel(const float *qpos, float *qvel, int C, int F, int nq, int freejoint, float dt, float max_qvel, void *stream);

/* Measurement helper for bench.py's FP32 roofline denominator: runs a dense FFMA loop
 * (blocks x threads, `iters` iterations of 16 independent FMAs per thread). out [blocks*threads]. */
int stacb_fma_peak(float *out, int blocks, int threads, int iters, void *stream);

/* Scheduling of stacb_pose_clips for this handle: -1 auto (default), 0 throughput mode (one warp per clip chain),
 * 1 latency mode (four cooperating warps per chain, speculative line search), 2 dense throughput mode (registers capped so 16
 * chains fit per SM; chosen automatically from 16 chains per SM), 3 wide latency mode (register-resident path: six warps per
 * chain speculating on three line-search candidates; general path: each of the four speculative evaluations carried out by
 * three warps sharing the bodies of a wide tree), 4 pair mode (register-resident path: two warps per chain evaluate two line-search
 * candidates at once, then the accepted point's gradient and the next extrapolation at once; chosen automatically between one
 * chain per SM and four).  On the multi-warp register-resident path (W = 2..8 warps per evaluation) 0 and 3 mean one group of W
 * warps per chain, 2 the same with registers capped for two CTAs per SM, 1 and 4 two groups of W warps per chain (pair mode; chosen
 * automatically only for W = 2 with at most one chain per SM).  Results do not depend on the mode (bit-identical).
 * The setting is a property of the handle, not of the process; do not change it while a launch of the same handle is being
 * enqueued from another thread. */
int stacb_tree_set_mode(stacb_tree *tree, int mode);

/* Which kernels serve the handle: 0 (default) = the register-resident hinge-tree solver where the model qualifies (active
 * subtree of at most 31 bodies with hinge joints plus one free joint, at most 31 keypoint sites: rodent, C. elegans), the
 * general kernels elsewhere; 1 = general kernels only.  (Jointless -- welded -- active bodies are folded into their nearest
 * jointed ancestor first, which brings the fruitfly's 49-57 active bodies down to 25 elements; hinge trees of 32-255 elements with
 * one hinge per element -- the mouse -- run the same solver on 2-8 warps per chain.)  The two paths evaluate the same mathematics in different (each
 * fixed and documented) float32 operation orders, so their results agree to rounding, not bit for bit.
 * stacb_tree_path returns 1 when the register-resident solver serves the handle, 0 otherwise. */
int stacb_tree_set_path(stacb_tree *tree, int path);
int stacb_tree_path(const stacb_tree *tree);

const char *stacb_last_error(void);
int stacb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* STACB_H_ */
