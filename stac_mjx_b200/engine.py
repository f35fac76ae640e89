"""Device-side solver handle: torch supplies device memory and streams, libstacb does the work.

`Engine` owns one `stacb_tree` (the tree descriptor copied to a GPU) and exposes the C ABI of
include/stacb.h on torch CUDA tensors.  Every method enqueues on the current torch stream and
returns tensors that live on the device; nothing here synchronises.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .tree import TreeModel


def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, tree: TreeModel, site_bodies, device: int | torch.device | None = None):
        if not torch.cuda.is_available():
            raise _lib.StacbError("no CUDA device: the STAC solver path has no CPU fallback")
        L = _lib.lib()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.StacbError(f"the STAC solver path runs on CUDA devices only, not {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.tree = tree
        self.nq, self.nbody, self.K = tree.nq, tree.nbody, len(site_bodies)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        keep = [
            i32(tree.body_parent), i32(tree.body_jntadr), i32(tree.body_jntnum), f32(tree.body_pos), f32(tree.body_quat),
            i32(tree.jnt_type), i32(tree.jnt_qposadr), i32(tree.jnt_bodyid), f32(tree.jnt_pos), f32(tree.jnt_axis),
            f32(tree.qpos0), i32(site_bodies),
        ]  # fmt: skip
        desc = _lib.TreeDesc(tree.nbody, tree.nq, tree.njnt, self.K, *[a.ctypes.data_as(C.c_void_p) for a in keep])
        h = C.c_void_p()
        _lib.check(L.stacb_tree_create(C.byref(desc), self.device.index, C.byref(h)), "stacb_tree_create")
        self._h, self._L = h, L

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.stacb_tree_destroy(h)
            self._h = None

    # -- helpers ---------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def f32(self, a, shape=None) -> torch.Tensor:
        t = torch.as_tensor(np.asarray(a, dtype=np.float32) if not isinstance(a, torch.Tensor) else a)
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def u8(self, a, shape=None) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device).to(torch.uint8).contiguous()
        else:
            t = torch.as_tensor(np.asarray(a).astype(np.uint8)).to(self.device).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def empty(self, *shape, dtype=torch.float32) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def set_mode(self, mode: int) -> None:
        """Scheduling of pose_clips (this handle): -1 auto, 0 throughput, 1 latency, 2 dense throughput, 3 wide latency."""
        _lib.check(self._L.stacb_tree_set_mode(self._h, int(mode)), "stacb_tree_set_mode")

    def set_path(self, path: int) -> None:
        """0: register-resident hinge-tree solver where the model qualifies (default); 1: general kernels only."""
        _lib.check(self._L.stacb_tree_set_path(self._h, int(path)), "stacb_tree_set_path")

    @property
    def path(self) -> int:
        """1 when the register-resident solver serves this model, 0 for the general kernels."""
        return int(self._L.stacb_tree_path(self._h))

    @property
    def smem_per_chain(self) -> int:
        return int(self._L.stacb_tree_smem_per_chain(self._h))

    # -- utils.kinematics + get_site_xpos ----------------------------------
    def fk(self, qpos, site_pos):
        qpos = self.f32(qpos)
        B = qpos.shape[0]
        site_pos = self.f32(site_pos, (self.K, 3))
        qout, xpos = self.empty(B, self.nq), self.empty(B, self.nbody, 3)
        xquat, sx = self.empty(B, self.nbody, 4), self.empty(B, self.K, 3)
        rc = self._L.stacb_fk(self._h, _ptr(qpos), _ptr(site_pos), _ptr(qout), _ptr(xpos), _ptr(xquat), _ptr(sx), B, self._stream())
        _lib.check(rc, "stacb_fk")
        return qout, xpos, xquat, sx

    # -- stac_core.q_loss and its gradient ---------------------------------
    def loss_grad(self, q, q0, kp, q_mask, kp_mask, site_pos, want_grad=True):
        q, q0, kp = self.f32(q), self.f32(q0), self.f32(kp)
        B = q.shape[0]
        q_mask, kp_mask = self.u8(q_mask, (self.nq,)), self.u8(kp_mask, (3 * self.K,))
        site_pos = self.f32(site_pos, (self.K, 3))
        loss = self.empty(B)
        grad = self.empty(B, self.nq) if want_grad else None
        rc = self._L.stacb_loss_grad(
            self._h, _ptr(q), _ptr(q0), _ptr(kp), _ptr(q_mask), _ptr(kp_mask), _ptr(site_pos), _ptr(loss), _ptr(grad), B, self._stream()
        )
        _lib.check(rc, "stacb_loss_grad")
        return loss, grad

    # -- stac_core._q_opt ----------------------------------------------------
    def q_opt(self, q0, kp, q_mask, kp_mask, site_pos, lb, ub, tol, maxiter=400, maxls=15):
        q0, kp = self.f32(q0), self.f32(kp)
        B = q0.shape[0]
        q_mask, kp_mask = self.u8(q_mask, (self.nq,)), self.u8(kp_mask, (3 * self.K,))
        site_pos, lb, ub = self.f32(site_pos, (self.K, 3)), self.f32(lb, (self.nq,)), self.f32(ub, (self.nq,))
        params, err = self.empty(B, self.nq), self.empty(B)
        iters, ls = self.empty(B, dtype=torch.int32), self.empty(B, dtype=torch.int32)
        rc = self._L.stacb_q_opt(
            self._h, _ptr(q0), _ptr(kp), _ptr(q_mask), _ptr(kp_mask), _ptr(site_pos), _ptr(lb), _ptr(ub),
            float(tol), int(maxiter), int(maxls), _ptr(params), _ptr(err), _ptr(iters), _ptr(ls), B, self._stream(),
        )  # fmt: skip
        _lib.check(rc, "stacb_q_opt")
        return params, err, iters, ls

    # -- fused root_optimization + pose_optimization over clips -------------
    def pose_clips(self, kp, qpos_io, site_pos, lb, ub, part_masks, *, do_root, root_kp_idx=-1, trunk_kps=None,
                   root_dims=7, tol=1e-4, maxiter=400, maxls=15, out=None, want_stats=True, session=None):  # fmt: skip
        """kp [C,F,3K] (device), qpos_io [C,nq] (device, updated in place). Returns dict of device tensors.

        ``session=(C, F, clip_stride)``: `kp` is ONE session buffer [(C-1)*clip_stride + F, 3K] whose clips overlap in memory
        (the `continuous` look-ahead read in place by the kernel, `stacb_pose_session`)."""
        kp = self.f32(kp)
        if session is not None:
            Cn, F, stride = (int(v) for v in session)
            if kp.dim() != 2 or kp.shape[1] != 3 * self.K or (Cn > 0 and kp.shape[0] < (Cn - 1) * stride + F):
                raise ValueError("session kp must be [(C-1)*clip_stride + F, 3K]")
        else:
            if kp.dim() != 3:
                raise ValueError("kp must be [C, F, 3K]")
            Cn, F = int(kp.shape[0]), int(kp.shape[1])
            stride = F
        if kp.shape[-1] != 3 * self.K:
            raise ValueError("kp must be [C, F, 3K]")
        if not (isinstance(qpos_io, torch.Tensor) and qpos_io.is_cuda and qpos_io.dtype == torch.float32 and qpos_io.is_contiguous()):
            raise ValueError("qpos_io must be a contiguous float32 CUDA tensor [C, nq] (it is updated in place)")
        if tuple(qpos_io.shape) != (Cn, self.nq):
            raise ValueError("qpos_io must be [C, nq]")
        site_pos, lb, ub = self.f32(site_pos, (self.K, 3)), self.f32(lb, (self.nq,)), self.f32(ub, (self.nq,))
        pm = self.u8(np.asarray(part_masks).reshape(-1, self.nq)) if not isinstance(part_masks, torch.Tensor) else self.u8(part_masks)
        P = int(pm.shape[0]) if pm.numel() else 0
        trunk = self.u8(trunk_kps if trunk_kps is not None else np.ones(self.K), (self.K,))
        if out is None:
            out = {}

        def o(k, *shape, dtype=torch.float32, zero=False):
            t = out.get(k)
            if t is None:  # first use of this dict: allocate (later calls with the same C, F reuse the buffers)
                t = out[k] = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.device)
            elif not (isinstance(t, torch.Tensor) and tuple(t.shape) == shape and t.dtype == dtype and t.device == self.device and t.is_contiguous()):
                raise ValueError(f"out[{k!r}] must be a contiguous {dtype} tensor of shape {shape} on {self.device} (the kernel writes it in place)")
            return t

        qpos, xpos = o("qpos", Cn, F, self.nq), o("xpos", Cn, F, self.nbody, 3)
        xquat, sites, err = o("xquat", Cn, F, self.nbody, 4), o("sites", Cn, F, self.K, 3), o("err", Cn, F)
        if want_stats:
            iters, ls = o("iters", Cn, F, 1 + P, dtype=torch.int32), o("ls_evals", Cn, F, 1 + P, dtype=torch.int32)
            rs = o("root_stats", Cn, 4, dtype=torch.int32, zero=True)
        else:
            iters = ls = rs = None
        status = o("status", Cn, dtype=torch.int32)
        if Cn == 0 or F == 0:
            return out
        rc = self._L.stacb_pose_session(
            self._h, _ptr(kp), stride, _ptr(qpos_io), _ptr(site_pos), _ptr(lb), _ptr(ub), _ptr(pm) if P else None, P,
            int(do_root), int(root_kp_idx), _ptr(trunk), int(root_dims), float(tol), int(maxiter), int(maxls),
            _ptr(qpos), _ptr(xpos), _ptr(xquat), _ptr(sites), _ptr(err), _ptr(iters), _ptr(ls), _ptr(rs), _ptr(status),
            Cn, F, self._stream(),
        )  # fmt: skip
        _lib.check(rc, "stacb_pose_session")
        return out

    # -- stac_core._m_opt sufficient statistics ------------------------------
    def _m_scratch(self, T: int) -> torch.Tensor:
        return self.empty(int(self._L.stacb_m_scratch_floats(self._h, int(T))))

    def m_stats_buffer(self, kp, q) -> torch.Tensor:
        """One contiguous device buffer [3K+2] = { s[K,3], z2, float(T) } (all-reduced in place by multi-GPU fits)."""
        kp, q = self.f32(kp), self.f32(q)
        T = int(kp.shape[0])
        out = self.empty(3 * self.K + 2)
        rc = self._L.stacb_m_stats(self._h, _ptr(kp), _ptr(q), _ptr(self._m_scratch(T)), _ptr(out), T, self._stream())
        _lib.check(rc, "stacb_m_stats")
        return out

    def m_stats(self, kp, q):
        out = self.m_stats_buffer(kp, q)
        return out[: 3 * self.K].reshape(self.K, 3), out[3 * self.K : 3 * self.K + 1]

    def m_residual(self, kp, q, m) -> torch.Tensor:
        """sum_t sum_k |y - p - R m|^2 at the offsets m [K,3]: device tensor [1]."""
        kp, q, m = self.f32(kp), self.f32(q), self.f32(m, (self.K, 3))
        T = int(kp.shape[0])
        out = self.empty(1)
        rc = self._L.stacb_m_residual(self._h, _ptr(kp), _ptr(q), _ptr(m), _ptr(self._m_scratch(T)), _ptr(out), T, self._stream())
        _lib.check(rc, "stacb_m_residual")
        return out

    # -- device epilogues of the IK pass (utils.handle_edge_effects / compute_velocity_from_kinematics) ----------
    def edge_crossfade(self, packed: torch.Tensor, n_frames_per_clip: int, overlap: int) -> torch.Tensor:
        """`packed` [C * (F + ov), ...] (any trailing shape) -> [rows, ...] with the overlaps cross-faded and removed."""
        F, ov = int(n_frames_per_clip), int(overlap)
        x = self.f32(packed)
        tail = tuple(x.shape[1:])
        D = int(np.prod(tail)) if tail else 1
        C = x.shape[0] // (F + ov)
        if C * (F + ov) != x.shape[0]:
            raise ValueError("packed length is not a multiple of n_frames_per_clip + overlap")
        rows = int(self._L.stacb_edge_rows(C, F, ov))
        if rows < 0:
            raise ValueError("edge_crossfade: need at least one clip and n_frames_per_clip >= overlap")
        xs = np.linspace(0.0, 1.0, ov)
        w = torch.as_tensor(0.5 * (1.0 + np.tanh(10.0 * (xs - 0.5) / 2.0)), dtype=torch.float64).to(self.device)
        out = self.empty(rows, *tail)
        rc = self._L.stacb_edge_crossfade(_ptr(x), _ptr(w), _ptr(out), C, F, ov, D, self._stream())
        _lib.check(rc, "stacb_edge_crossfade")
        return out

    def qvel(self, qpos: torch.Tensor, n_frames_per_clip: int, dt: float, freejoint: bool = True, max_qvel: float = 20.0) -> torch.Tensor:
        """qpos [C * F, nq] of C continuous clips -> qvel [C * F, nv]."""
        q = self.f32(qpos)
        F = int(n_frames_per_clip)
        C, nq = q.shape[0] // F, int(q.shape[-1])
        if C * F != q.shape[0]:
            raise ValueError("qpos length is not a multiple of n_frames_per_clip")
        out = self.empty(C * F, nq - 1 if freejoint else nq)
        rc = self._L.stacb_qvel(_ptr(q), _ptr(out), C, F, nq, int(bool(freejoint)), float(dt), float(max_qvel), self._stream())
        _lib.check(rc, "stacb_qvel")
        return out

    def fma_peak_tflops(self, iters: int = 20000) -> float:
        """Measured FP32 FMA throughput (TFLOP/s) of this GPU: bench.py's roofline denominator."""
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        blocks, threads = sms * 8, 256
        out = self.empty(blocks * threads)
        s = self._stream()
        self._L.stacb_fma_peak(_ptr(out), blocks, threads, 200, s)
        torch.cuda.synchronize(self.device)
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(self._L.stacb_fma_peak(_ptr(out), blocks, threads, iters, s), "stacb_fma_peak")
            e1.record()
            torch.cuda.synchronize(self.device)
            ms = e0.elapsed_time(e1)
            best = max(best, 2.0 * 16 * iters * blocks * threads / (ms * 1e-3) / 1e12)
        return best
