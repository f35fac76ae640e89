"""Solver seam: ``StacCore.q_opt`` / ``StacCore.m_opt`` (reference ``stac_mjx/stac_core.py``).

Same names, argument order and return conventions as the reference's duck-typed seam
(``stac_core.py:175-275``; proved a seam by ``FakeStacCore`` in the reference's
``tests/unit/test_compute_stac.py:32-51``), backed by the CUDA library instead of
MJX + jaxopt.  ``mjx_model`` is a `StacModel` (engine + current site offsets),
``mjx_data`` a `StacData_`-like state carrying ``qpos``.  Arrays may be numpy or torch;
results are torch tensors on the model's device.
"""

from __future__ import annotations

from dataclasses import dataclass, replace as _dc_replace
from types import SimpleNamespace
from typing import Any, NamedTuple

import torch

from .engine import Engine


class MOptResult(NamedTuple):
    """Result of marker offset optimization (reference ``stac_core.py:20-24``)."""

    params: Any  # [K, 3]
    error: Any  # scalar


@dataclass
class StacModel:
    """What the path reads from ``mjx.Model``: tree (inside the engine) + current site offsets."""

    engine: Engine
    site_pos: torch.Tensor  # [K, 3] offsets of the keypoint sites, device

    @property
    def nq(self) -> int:
        return self.engine.nq

    @property
    def jnt_type(self):
        return self.engine.tree.jnt_type

    def replace(self, **kw) -> "StacModel":
        if "site_pos" in kw:
            kw["site_pos"] = self.engine.f32(kw["site_pos"], (self.engine.K, 3))
        return _dc_replace(self, **kw)


@dataclass
class StacState:
    """What the path reads from ``mjx.Data``: qpos and the FK results of the last ``kinematics``."""

    qpos: torch.Tensor  # [nq] or [C, nq]
    xpos: torch.Tensor | None = None
    xquat: torch.Tensor | None = None
    site_xpos: torch.Tensor | None = None

    def replace(self, **kw) -> "StacState":
        return _dc_replace(self, **kw)


def kinematics(model: StacModel, data: StacState) -> StacState:
    """``utils.kinematics`` (reference ``utils.py:49-60``) on the GPU."""
    q = data.qpos
    single = q.dim() == 1
    qo, xp, xq, sx = model.engine.fk(q.reshape(-1, model.nq), model.site_pos)
    if single:
        qo, xp, xq, sx = qo[0], xp[0], xq[0], sx[0]
    return StacState(qpos=qo, xpos=xp, xquat=xq, site_xpos=sx)


class _QSolver:
    """Stand-in for the ``jaxopt.ProjectedGradient`` attribute the reference exposes (``.tol``, ``.maxiter``)."""

    def __init__(self, tol: float, maxiter: int):
        self.tol, self.maxiter = tol, maxiter
        self.maxls, self.decrease_factor, self.acceleration, self.stepsize = 15, 0.5, True, 0.0


def _q_opt(q_solver, mjx_model, mjx_data, marker_ref_arr, qs_to_opt, kps_to_opt, q0, lb, ub, site_idxs=None):
    """One box-constrained FISTA solve (reference ``stac_core.py:66-99``). Returns (mjx_data, res)."""
    eng = mjx_model.engine
    q0t = eng.f32(q0).reshape(1, eng.nq)
    kp = eng.f32(marker_ref_arr).reshape(1, 3 * eng.K)
    params, err, iters, ls = eng.q_opt(
        q0t, kp, qs_to_opt, kps_to_opt, mjx_model.site_pos, lb, ub, q_solver.tol, q_solver.maxiter, q_solver.maxls
    )
    state = SimpleNamespace(error=err[0], iter_num=iters[0], ls_evals=ls[0])
    return mjx_data, SimpleNamespace(params=params[0], state=state)


def _m_opt(mjx_model, mjx_data, keypoints, q, initial_offsets, is_regularized, reg_coef, site_idxs=None, reduce_fn=None):
    """Closed-form marker offsets (reference ``stac_core.py:102-172``).

    One kernel produces the sufficient statistics as ONE device buffer ``[s (3K), z2, T]``; ``reduce_fn(buf)``, when
    given, all-reduces it in place across ranks (m-phase of a multi-GPU fit: 3K+2 floats over NVLink right behind the
    kernel) before the closed form is applied redundantly on every rank.  The reported error is the objective at
    ``m*`` evaluated from the residuals (``stacb_m_residual``): the same number as the reference's expanded form
    ``z2 - 2 sum(m* s) + T sum(m*^2)`` without its float32 cancellation.
    """
    eng = mjx_model.engine
    kp, qt = eng.f32(keypoints), eng.f32(q)
    K = eng.K
    buf = eng.m_stats_buffer(kp, qt)
    T = float(kp.shape[0])
    if reduce_fn is not None:
        buf = reduce_fn(buf)
        T = float(round(float(buf[3 * K + 1])))  # total frames over all ranks
    s = buf[: 3 * K].reshape(K, 3)
    d = eng.f32(is_regularized, (K, 3))
    m0 = eng.f32(initial_offsets, (K, 3))
    reg = float(reg_coef)
    m_star = (s + reg * d * m0) / (T + reg * d)
    data_term = eng.m_residual(kp, qt, m_star)
    if reduce_fn is not None:
        data_term = reduce_fn(data_term)
    reg_term = reg * torch.sum((d * (m_star - m0)) ** 2)
    return MOptResult(params=m_star, error=data_term[0] + reg_term)


class StacCore:
    """Pose and offset optimization core (reference ``stac_core.py:175-275``)."""

    def __init__(self, tol: float = 1e-5, n_iter_q: int = 400):
        self.q_solver = _QSolver(float(tol), int(n_iter_q))

    def q_opt(self, mjx_model, mjx_data, marker_ref_arr, qs_to_opt, kps_to_opt, q0, lb, ub, site_idxs=None):
        return _q_opt(self.q_solver, mjx_model, mjx_data, marker_ref_arr, qs_to_opt, kps_to_opt, q0, lb, ub, site_idxs)

    def m_opt(self, mjx_model, mjx_data, keypoints, q, initial_offsets, is_regularized, reg_coef, site_idxs=None, reduce_fn=None):
        return _m_opt(mjx_model, mjx_data, keypoints, q, initial_offsets, is_regularized, reg_coef, site_idxs, reduce_fn)
