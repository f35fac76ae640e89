"""ctypes front-end of the CPU oracle (oracle/stac_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  Never imported by
stac_mjx_b200/.  q-phase parity is UNPINNED (see the header of stac_oracle.c
and DESIGN.md); the m-phase is pinned by the reference's own known-answer tests.

``Oracle(tree, site_bodies, dtype, mode)``
    dtype  np.float32 | np.float64
    mode   0 = MJX operation order (faithful restatement)
           1 = canonical order (bit-matched to the general CUDA kernels in float32)
           2 = what libstacb computes by default: the register-resident solver's "fast order"
               (oracle/fast_order.h) for the models it serves, mode 1 elsewhere
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIBS: dict = {}


def build() -> None:
    subprocess.run(["sh", str(HERE / "build.sh")], check=True, capture_output=True)


def _lib(dtype) -> C.CDLL:
    key = np.dtype(dtype).name
    if key not in _LIBS:
        p = HERE / "_build" / ("liboracle_f32.so" if key == "float32" else "liboracle_f64.so")
        newest = max((HERE / n).stat().st_mtime for n in ("stac_oracle.c", "fast_order.h"))
        if not p.exists() or p.stat().st_mtime < newest:
            build()
        _LIBS[key] = C.CDLL(str(p))
    return _LIBS[key]


class _OModel(C.Structure):
    _fields_ = [
        ("nbody", C.c_int32),
        ("nq", C.c_int32),
        ("njnt", C.c_int32),
        ("nsite", C.c_int32),
        ("body_parent", C.c_void_p),
        ("body_jntadr", C.c_void_p),
        ("body_jntnum", C.c_void_p),
        ("body_pos", C.c_void_p),
        ("body_quat", C.c_void_p),
        ("jnt_type", C.c_void_p),
        ("jnt_qposadr", C.c_void_p),
        ("jnt_bodyid", C.c_void_p),
        ("jnt_pos", C.c_void_p),
        ("jnt_axis", C.c_void_p),
        ("qpos0", C.c_void_p),
        ("site_body", C.c_void_p),
    ]


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, tree, site_bodies, dtype=np.float32, mode: int = 0):
        self.dtype = np.dtype(dtype)
        self.mode = int(mode)
        self.lib = _lib(dtype)
        self.tree = tree
        self.nq, self.nbody, self.K = tree.nq, tree.nbody, len(site_bodies)
        f = lambda a: np.ascontiguousarray(a, dtype=self.dtype)
        i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        # the model reaches MJX as float32 (mjx.put_model); keep that rounding in the f64 twin too
        r32 = lambda a: np.asarray(a, dtype=np.float32)
        self._keep = dict(
            body_parent=i(tree.body_parent),
            body_jntadr=i(tree.body_jntadr),
            body_jntnum=i(tree.body_jntnum),
            body_pos=f(r32(tree.body_pos)),
            body_quat=f(r32(tree.body_quat)),
            jnt_type=i(tree.jnt_type),
            jnt_qposadr=i(tree.jnt_qposadr),
            jnt_bodyid=i(tree.jnt_bodyid),
            jnt_pos=f(r32(tree.jnt_pos)),
            jnt_axis=f(r32(tree.jnt_axis)),
            qpos0=f(r32(tree.qpos0)),
            site_body=i(site_bodies),
        )
        self.m = _OModel(tree.nbody, tree.nq, tree.njnt, self.K, *[_p(self._keep[k]) for k in list(self._keep)])
        self.real = C.c_float if self.dtype == np.float32 else C.c_double

    @property
    def fast_path(self) -> bool:
        """True when mode 2 differs from mode 1 for this model (rodent, celegans: the register-resident solver)."""
        return bool(self.lib.oracle_fast_path(C.byref(self.m)))

    def _f(self, a, shape=None):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        if shape is not None:
            assert a.shape == tuple(shape), (a.shape, shape)
        return a

    @staticmethod
    def _u8(a):
        return np.ascontiguousarray(np.asarray(a).astype(np.uint8))

    # -- utils.kinematics + site_xpos -------------------------------------
    def fk(self, qpos, site_pos):
        qpos, site_pos = self._f(qpos, (self.nq,)), self._f(site_pos, (self.K, 3))
        qout = np.empty(self.nq, self.dtype)
        xpos = np.empty((self.nbody, 3), self.dtype)
        xquat = np.empty((self.nbody, 4), self.dtype)
        sx = np.empty((self.K, 3), self.dtype)
        self.lib.oracle_fk(C.byref(self.m), self.mode, _p(qpos), _p(site_pos), _p(qout), _p(xpos), _p(xquat), _p(sx))
        return qout, xpos, xquat, sx

    # -- stac_core.q_loss and its gradient --------------------------------
    def loss_grad(self, q, q0, qmask, kp, kpmask, site_pos):
        q, q0, kp, site_pos = self._f(q), self._f(q0), self._f(kp, (3 * self.K,)), self._f(site_pos, (self.K, 3))
        qmask, kpmask = self._u8(qmask), self._u8(kpmask)
        loss = self.real(0)
        grad = np.empty(self.nq, self.dtype)
        self.lib.oracle_loss_grad(
            C.byref(self.m), self.mode, _p(q), _p(q0), _p(qmask), _p(kp), _p(kpmask), _p(site_pos), C.byref(loss), _p(grad)
        )
        return self.dtype.type(loss.value), grad

    # -- stac_core._q_opt ---------------------------------------------------
    def q_opt(self, q0, lb, ub, qmask, kp, kpmask, site_pos, tol, maxiter=400, maxls=15):
        q0, lb, ub = self._f(q0), self._f(lb), self._f(ub)
        kp, site_pos = self._f(kp, (3 * self.K,)), self._f(site_pos, (self.K, 3))
        qmask, kpmask = self._u8(qmask), self._u8(kpmask)
        params = np.empty(self.nq, self.dtype)
        err = self.real(0)
        it, ls = C.c_int32(0), C.c_int32(0)
        self.lib.oracle_q_opt(
            C.byref(self.m), self.mode, _p(q0), _p(lb), _p(ub), _p(qmask), _p(kp), _p(kpmask), _p(site_pos),
            self.real(tol), int(maxiter), int(maxls), _p(params), C.byref(err), C.byref(it), C.byref(ls),
        )  # fmt: skip
        return params, self.dtype.type(err.value), it.value, ls.value

    # -- root_optimization + pose_optimization over clips -------------------
    def pose_clips(self, kp, qpos_init, site_pos, lb, ub, part_masks, *, do_root, root_kp_idx=-1, trunk_kps=None,
                   root_dims=7, tol=1e-4, maxiter=400, maxls=15, nthreads=1):  # fmt: skip
        kp = self._f(kp)
        Cn, F = kp.shape[0], kp.shape[1]
        assert kp.shape[2] == 3 * self.K
        qpos_init = self._f(np.broadcast_to(np.asarray(qpos_init, self.dtype), (Cn, self.nq)))
        site_pos, lb, ub = self._f(site_pos, (self.K, 3)), self._f(lb), self._f(ub)
        pm = self._u8(np.asarray(part_masks).reshape(-1, self.nq)) if len(part_masks) else np.zeros((0, self.nq), np.uint8)
        P = pm.shape[0]
        trunk = self._u8(trunk_kps if trunk_kps is not None else np.ones(self.K))
        out = dict(
            qpos=np.empty((Cn, F, self.nq), self.dtype),
            xpos=np.empty((Cn, F, self.nbody, 3), self.dtype),
            xquat=np.empty((Cn, F, self.nbody, 4), self.dtype),
            sites=np.empty((Cn, F, self.K, 3), self.dtype),
            err=np.empty((Cn, F), self.dtype),
            iters=np.zeros((Cn, F, 1 + P), np.int32),
            ls_evals=np.zeros((Cn, F, 1 + P), np.int32),
            root_stats=np.zeros((Cn, 4), np.int32),
        )
        self.lib.oracle_pose_clips(
            C.byref(self.m), self.mode, _p(kp), Cn, F, _p(qpos_init), _p(site_pos), _p(lb), _p(ub), _p(pm), P,
            int(bool(do_root)), int(root_kp_idx), _p(trunk), int(root_dims), self.real(tol), int(maxiter), int(maxls),
            _p(out["qpos"]), _p(out["xpos"]), _p(out["xquat"]), _p(out["sites"]), _p(out["err"]), _p(out["iters"]),
            _p(out["ls_evals"]), _p(out["root_stats"]), int(nthreads),
        )  # fmt: skip
        return out

    # -- stac_core._m_opt -----------------------------------------------------
    def m_stats(self, kp, q):
        kp, q = self._f(kp), self._f(q)
        T = kp.shape[0]
        s = np.empty((self.K, 3), self.dtype)
        z2 = self.real(0)
        self.lib.oracle_m_stats(C.byref(self.m), self.mode, _p(kp), _p(q), T, _p(s), C.byref(z2))
        return s, self.dtype.type(z2.value)

    def m_residual(self, kp, q, m):
        """sum_t sum_k |y - p - R m|^2 (the data term of the m-phase objective at offsets m, from the residuals)."""
        kp, q, m = self._f(kp), self._f(q), self._f(m, (self.K, 3))
        out = self.real(0)
        self.lib.oracle_m_residual(C.byref(self.m), self.mode, _p(kp), _p(q), _p(m), kp.shape[0], C.byref(out))
        return self.dtype.type(out.value)

    def m_opt(self, kp, q, initial_offsets, is_regularized, reg_coef):
        """Closed-form offsets (reference stac_core.py:146-172)."""
        dt = self.dtype.type
        T = dt(np.asarray(kp).shape[0])
        s, z2 = self.m_stats(kp, q)
        d = np.asarray(is_regularized, self.dtype)
        m0 = np.asarray(initial_offsets, self.dtype)
        reg = dt(reg_coef)
        denom = T + reg * d
        numer = s + reg * d * m0
        m_star = numer / denom
        data_term = z2 - dt(2.0) * np.sum(m_star * s, dtype=self.dtype) + T * np.sum(m_star**2, dtype=self.dtype)
        reg_term = reg * np.sum((d * (m_star - m0)) ** 2, dtype=self.dtype)
        return m_star.astype(self.dtype), dt(data_term + reg_term)
