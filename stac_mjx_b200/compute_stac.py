"""Phase drivers: root / pose / offset optimisation (reference ``stac_mjx/compute_stac.py``).

Signatures and return conventions follow the reference (``compute_stac.py:17-28,107-118,170-187``).
With the CUDA-backed `StacCore` each driver is ONE fused kernel launch over all clips and frames
(``stacb_pose_clips``); with any other duck-typed solver object (the reference's own tests use a
``FakeStacCore``) the drivers fall back to the reference's per-frame call sequence through
``stac_core_obj.q_opt`` / ``m_opt`` so the seam keeps its observable behaviour.
"""

from __future__ import annotations

import time

import numpy as np
import torch

from . import stac_core, utils
from .mjcf import JNT_SLIDE


def _fused(stac_core_obj, mjx_model) -> bool:
    return isinstance(stac_core_obj, stac_core.StacCore) and isinstance(mjx_model, stac_core.StacModel)


def _root_dims(mjx_model) -> int:
    return 4 if int(mjx_model.jnt_type[0]) == JNT_SLIDE else 7  # compute_stac.py:51-54


def root_optimization(stac_core_obj, mjx_model, mjx_data, kp_data, root_kp_idx, lb, ub, site_idxs, trunk_kps, frame: int = 0):
    """Optimize root DOFs for a single frame (reference ``compute_stac.py:17-104``).

    ``kp_data`` is [F, 3K] with ``mjx_data.qpos`` [nq], or batched over clips ([C, F, 3K] / [C, nq],
    the reference's ``jax.vmap`` in ``stac.py:405-419``).
    """
    print("Root Optimization:")
    root_dims = _root_dims(mjx_model)
    print(f"Optimizing first {root_dims} qposes for root optimization")
    s = time.time()
    if not _fused(stac_core_obj, mjx_model):
        return _root_optimization_seam(stac_core_obj, mjx_model, mjx_data, kp_data, root_kp_idx, lb, ub, site_idxs, trunk_kps, frame, root_dims, s)
    eng = mjx_model.engine
    kp = eng.f32(kp_data)
    single = kp.dim() == 2
    kp3 = kp.reshape((1,) + tuple(kp.shape)) if single else kp
    qio = mjx_data.qpos.reshape(-1, eng.nq).clone().contiguous()
    out = eng.pose_clips(
        kp3[:, frame : frame + 1].contiguous(), qio, mjx_model.site_pos, lb, ub, np.zeros((0, eng.nq), bool),
        do_root=2, root_kp_idx=root_kp_idx, trunk_kps=np.asarray(trunk_kps), root_dims=root_dims,
        tol=stac_core_obj.q_solver.tol, maxiter=stac_core_obj.q_solver.maxiter, maxls=stac_core_obj.q_solver.maxls,
    )  # fmt: skip
    qnew = qio[0] if single else qio
    # FK outputs for the caller; qpos itself stays exactly what the solver wrote (already normalised once)
    new = stac_core.kinematics(mjx_model, stac_core.StacState(qpos=qnew)).replace(qpos=qnew)
    new.root_stats = out["root_stats"]
    print(f"Root optimization finished in {(time.time() - s) / 60:.2f} minutes")
    return new


def _root_optimization_seam(stac_core_obj, mjx_model, mjx_data, kp_data, root_kp_idx, lb, ub, site_idxs, trunk_kps, frame, root_dims, s):
    # the reference's call sequence, for duck-typed solver objects (compute_stac.py:57-104)
    xp = torch if isinstance(mjx_data.qpos, torch.Tensor) else np
    kp_data = xp.as_tensor(kp_data) if xp is torch else np.asarray(kp_data)
    root_xyz = kp_data[frame, 3 * root_kp_idx : 3 * root_kp_idx + 3]
    nq = mjx_data.qpos.shape[0]
    qs_to_opt = np.zeros(nq, dtype=bool)
    qs_to_opt[:root_dims] = True
    kps_to_opt = np.repeat(np.asarray(trunk_kps), 3)
    res = None
    for _ in range(2):
        q0 = mjx_data.qpos.clone() if xp is torch else np.array(mjx_data.qpos)
        q0[:3] = root_xyz
        mjx_data, res = stac_core_obj.q_opt(mjx_model, mjx_data, kp_data[frame, :], qs_to_opt, kps_to_opt, q0, lb, ub, site_idxs)
        m = xp.as_tensor(qs_to_opt) if xp is torch else qs_to_opt
        mjx_data = mjx_data.replace(qpos=utils.make_qs(q0, m, res.params))
    print(f"Root optimization finished in {(time.time() - s) / 60:.2f} minutes with an error of {res.state.error}")
    return mjx_data


def offset_optimization(stac_core_obj, mjx_model, mjx_data, kp_data, offsets, q, n_sample_frames, is_regularized, site_idxs,
                        m_reg_coef, time_indices=None, reduce_fn=None):  # fmt: skip
    """Closed-form marker offsets from sampled frames (reference ``compute_stac.py:107-167``).

    The reference samples with ``jax.random.permutation(PRNGKey(0), arange(F))[:n_sample_frames]``
    (``:136-140``); `sample_time_indices` reproduces that threefry permutation in numpy.  ``time_indices`` injects an
    explicit sample (a rank's share of it in multi-GPU fits).
    """
    F = int(kp_data.shape[0])
    if time_indices is None:
        time_indices = sample_time_indices(F, int(n_sample_frames))
    s = time.time()
    print("Begining offset optimization:")
    if _fused(stac_core_obj, mjx_model):
        eng = mjx_model.engine
        idx = torch.as_tensor(np.asarray(time_indices), device=eng.device, dtype=torch.long)
        keypoints = eng.f32(kp_data)[idx]
        qs = eng.f32(q)[idx]
        res = stac_core_obj.m_opt(mjx_model, mjx_data, keypoints, qs, offsets, is_regularized, m_reg_coef, site_idxs, reduce_fn=reduce_fn)
    else:
        res = stac_core_obj.m_opt(mjx_model, mjx_data, np.asarray(kp_data)[time_indices], np.asarray(q)[time_indices], offsets,
                                  is_regularized, m_reg_coef, site_idxs)  # fmt: skip
    offset_opt_param = res.params
    print(f"Final residual error of {float(res.error)}")
    mjx_model = utils.set_site_pos(mjx_model, offset_opt_param, site_idxs)
    if _fused(stac_core_obj, mjx_model):
        mjx_data = stac_core.kinematics(mjx_model, mjx_data)
    print(f"Offset optimization finished in {time.time() - s} seconds")
    return mjx_model, mjx_data, offset_opt_param


def sample_time_indices(n_frames: int, n_sample_frames: int, partitionable: bool = True) -> np.ndarray:
    """Frame sample of the m-phase (reference ``compute_stac.py:134-140``):
    ``jax.random.permutation(jax.random.PRNGKey(0), arange(F), independent=True)[:n_sample]``.

    Reproduced without JAX by ``jax_random`` (threefry2x32 + sort-based shuffle, integer arithmetic only), so the
    fitted offsets use the frames the reference would use; ``partitionable`` is the ``jax_threefry_partitionable``
    flag (default of the JAX versions the reference requires).
    """
    from . import jax_random

    return jax_random.permutation(0, int(n_frames), partitionable)[: int(n_sample_frames)]


def pose_optimization(stac_core_obj, mjx_model, mjx_data, kp_data, lb, ub, site_idxs, indiv_parts):
    """Pose optimisation over an entire clip (reference ``compute_stac.py:170-278``).

    Returns ``(mjx_data, qposes, xposes, xquats, marker_sites, frame_time, frame_error)``.  Unbatched
    input ([F, 3K]) gives per-frame leading dimensions; batched input ([C, F, 3K], the reference's vmap
    in ``stac.py:425-440``) gives the shapes ``np.array`` makes of the vmapped lists: qposes [C, F, nq]
    but xposes / xquats / marker_sites / frame_error with the FRAME axis first ([F, C, ...]).
    """
    s = time.time()
    print("Pose Optimization:")
    if not _fused(stac_core_obj, mjx_model):
        return _pose_optimization_seam(stac_core_obj, mjx_model, mjx_data, kp_data, lb, ub, site_idxs, indiv_parts, s)
    eng = mjx_model.engine
    kp = eng.f32(kp_data)
    single = kp.dim() == 2
    kp3 = kp.reshape((1,) + tuple(kp.shape)) if single else kp
    qio = mjx_data.qpos.reshape(-1, eng.nq).clone().contiguous()
    parts = np.asarray(indiv_parts).reshape(-1, eng.nq) if len(indiv_parts) else np.zeros((0, eng.nq), bool)
    out = eng.pose_clips(
        kp3.contiguous(), qio, mjx_model.site_pos, lb, ub, parts, do_root=0,
        tol=stac_core_obj.q_solver.tol, maxiter=stac_core_obj.q_solver.maxiter, maxls=stac_core_obj.q_solver.maxls,
    )  # fmt: skip
    F = int(kp3.shape[1])
    last = lambda t: t[0, -1] if single else t[:, -1]
    new = stac_core.StacState(qpos=qio[0] if single else qio, xpos=last(out["xpos"]), xquat=last(out["xquat"]), site_xpos=last(out["sites"]))
    new.solver_stats = {"iters": out["iters"], "ls_evals": out["ls_evals"], "status": out["status"]}
    if single:
        qposes, xposes, xquats, sites, err = out["qpos"][0], out["xpos"][0], out["xquat"][0], out["sites"][0], out["err"][0]
    else:
        qposes = out["qpos"]
        xposes, xquats = out["xpos"].transpose(0, 1), out["xquat"].transpose(0, 1)
        sites, err = out["sites"].transpose(0, 1), out["err"].transpose(0, 1)
    torch.cuda.synchronize(eng.device)
    dt = time.time() - s
    print(f"Pose Optimization finished in {dt / 60.0:.2f} minutes")
    return new, qposes, xposes, xquats, sites, [dt / max(F, 1)] * F, err


def _pose_optimization_seam(stac_core_obj, mjx_model, mjx_data, kp_data, lb, ub, site_idxs, indiv_parts, s):
    # the reference's per-frame loop, for duck-typed solver objects (compute_stac.py:206-278)
    nq = mjx_model.nq
    kps_to_opt = np.ones(kp_data.shape[1], dtype=bool)
    qs_to_opt = np.ones(nq, dtype=bool)
    qposes, xposes, xquats, marker_sites, frame_time, frame_error = [], [], [], [], [], []
    for n_frame in range(kp_data.shape[0]):
        t0 = time.time()
        q0 = mjx_data.qpos
        mjx_data, res = stac_core_obj.q_opt(mjx_model, mjx_data, kp_data[n_frame, :], qs_to_opt, kps_to_opt, q0, lb, ub, site_idxs)
        mjx_data = mjx_data.replace(qpos=res.params)
        for part in indiv_parts:
            q0 = mjx_data.qpos
            mjx_data, res = stac_core_obj.q_opt(mjx_model, mjx_data, kp_data[n_frame, :], part, kps_to_opt, q0, lb, ub, site_idxs)
            mjx_data = mjx_data.replace(qpos=utils.make_qs(q0, part, res.params))
        qposes.append(mjx_data.qpos)
        xposes.append(mjx_data.xpos)
        xquats.append(mjx_data.xquat)
        marker_sites.append(utils.get_site_xpos(mjx_data, site_idxs))
        frame_time.append(time.time() - t0)
        frame_error.append(res.state.error)
    print(f"Pose Optimization finished in {(time.time() - s) / 60.0:.2f} minutes")
    return mjx_data, np.array(qposes), xposes, xquats, marker_sites, frame_time, frame_error
