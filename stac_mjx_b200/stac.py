"""`Stac`: orchestration of the fitting pipeline (reference ``stac_mjx/stac.py:91-503``).

Public surface kept from the reference: ``Stac(xml_path, cfg, kp_names)``, ``fit_offsets(kp_data)``,
``ik_only(kp_data, offsets)`` and ``_package_data`` with its output layout (including the frame-major
``marker_sites`` interleave of batched runs, ``stac.py:483-486``).  The two ``jax.vmap`` over clips
(``stac.py:405-440``) become one fused kernel launch; with ``torch.distributed`` initialised the clips
are block-partitioned across ranks.  Rendering is out of scope.
"""

from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from . import compute_stac, io, model, parallel, stac_core, utils
from .engine import Engine
from .mjcf import JNT_FREE, JNT_SLIDE
from .tree import TreeModel


class Stac:
    """Skeletal registration: model setup, pose / offset optimisation (reference ``stac.py:91``)."""

    def __init__(self, xml_path, cfg, kp_names: list[str], *, tree: TreeModel | None = None, device=None):
        self.cfg = cfg
        self._kp_names = list(kp_names)
        self._xml_path = Path(xml_path) if xml_path is not None else None
        self._marker_size = cfg.model.get("MARKER_SIZE", 0.005)
        if tree is None:
            tree = model.compile_fit_tree(self._xml_path, cfg.model)
        setup = model.make_setup(tree, cfg.model, self._kp_names)
        self._setup = setup
        self._mj_model = tree
        tree.opt = SimpleNamespace(timestep=tree.timestep)
        self._body_site_idxs = setup.site_idxs
        self._is_regularized = setup.is_regularized
        self._body_names = setup.body_names
        self._root_kp_idx = setup.root_kp_idx
        self._lb, self._ub, self._part_names = setup.lb, setup.ub, setup.part_names
        self._indiv_parts = setup.indiv_parts
        self._trunk_kps = setup.trunk_kps
        self._freejoint = bool(tree.jnt_type[0] == JNT_FREE)
        self._slidejoint = bool(tree.jnt_type[0] == JNT_SLIDE)
        self._fixed = not (self._freejoint or self._slidejoint)
        self.stac_core_obj = stac_core.StacCore(cfg.model.FTOL, cfg.model.N_ITER_Q)
        self._engine = Engine(tree, setup.site_bodies, device)
        self._offsets = None
        self.time_indices = None  # optional injection of the m-phase frame sample (see compute_stac)
        self.last_stats = None

    # ------------------------------------------------------------------
    def _get_error_stats(self, errors):
        flattened = np.asarray(errors.cpu() if isinstance(errors, torch.Tensor) else errors, dtype=np.float64).reshape(-1)
        return flattened, float(np.mean(flattened)), float(np.std(flattened))

    def _load(self, offsets):
        """``utils.mjx_load`` + ``set_site_pos`` + FK at qpos0 (reference ``stac.py:268-275,382-391``)."""
        eng = self._engine
        mjx_model = stac_core.StacModel(engine=eng, site_pos=eng.f32(offsets, (eng.K, 3)))
        mjx_data = stac_core.kinematics(mjx_model, stac_core.StacState(qpos=eng.f32(self._mj_model.qpos0)))
        return mjx_model, mjx_data

    # ------------------------------------------------------------------
    def fit_offsets(self, kp_data) -> io.StacData:
        """Alternate pose and offset optimisation (reference ``stac.py:254-354``).

        The q-phase of a fit is ONE sequential chain over all frames (``compute_stac.py:256``), so ranks of
        a multi-GPU job run it as replicas; the m-phase shards its sampled frames across ranks and
        all-reduces the 3K+2 sufficient statistics.
        """
        kp_data = np.asarray(kp_data, dtype=np.float32)
        self._offsets = np.array(self._setup.initial_offsets)
        mjx_model, mjx_data = self._load(self._offsets)
        if self._root_kp_idx == -1:
            print("ROOT_OPTIMIZATION_KEYPOINT not specified, skipping Root Optimization.")
        elif not self._fixed:
            mjx_data = compute_stac.root_optimization(
                self.stac_core_obj, mjx_model, mjx_data, kp_data, self._root_kp_idx, self._lb, self._ub,
                self._body_site_idxs, self._trunk_kps,
            )  # fmt: skip
        else:
            print("ROOT_OPTIMIZATION_KEYPOINT specified but model has fixed root, skipping Root Optimization")
        rank, ws = parallel.world()
        for n_iter in range(self.cfg.model.N_ITERS):
            print(f"Calibration iteration: {n_iter + 1}/{self.cfg.model.N_ITERS}")
            mjx_data, qposes, xposes, xquats, marker_sites, frame_time, frame_error = compute_stac.pose_optimization(
                self.stac_core_obj, mjx_model, mjx_data, kp_data, self._lb, self._ub, self._body_site_idxs, self._indiv_parts
            )
            _, mean, std = self._get_error_stats(frame_error)
            print(f"Mean: {mean}")
            print(f"Standard deviation: {std}")
            tidx = self.time_indices
            if tidx is None:
                tidx = compute_stac.sample_time_indices(kp_data.shape[0], int(self.cfg.model.N_SAMPLE_FRAMES))
            lo, hi = parallel.shard_range(len(tidx), rank, ws)
            mjx_model, mjx_data, offs = compute_stac.offset_optimization(
                self.stac_core_obj, mjx_model, mjx_data, kp_data, self._offsets, qposes, self.cfg.model.N_SAMPLE_FRAMES,
                self._is_regularized, self._body_site_idxs, self.cfg.model.M_REG_COEF,
                time_indices=np.asarray(tidx)[lo:hi], reduce_fn=parallel.allreduce_m_stats if ws > 1 else None,
            )  # fmt: skip
            self._offsets = offs.cpu().numpy()
        print("Final pose optimization", flush=True)
        mjx_data, qposes, xposes, xquats, marker_sites, frame_time, frame_error = compute_stac.pose_optimization(
            self.stac_core_obj, mjx_model, mjx_data, kp_data, self._lb, self._ub, self._body_site_idxs, self._indiv_parts
        )
        _, mean, std = self._get_error_stats(frame_error)
        print(f"Mean: {mean}")
        print(f"Standard deviation: {std}")
        self.last_stats = mjx_data.solver_stats
        return self._package_data(
            mjx_model, qposes.cpu().numpy(), xposes.cpu().numpy(), xquats.cpu().numpy(), marker_sites.cpu().numpy(), kp_data
        )

    # ------------------------------------------------------------------
    def fit_offsets_clip_split(self, kp_data, n_frames_per_clip: int | None = None) -> io.StacData:
        """OPT-IN deviation from the reference schedule: the fit's q-phase over independent clips.

        The reference runs the q-phase of ``fit_offsets`` as ONE warm-started chain over all ``n_fit_frames``
        (``stac.py:298-342``), which cannot use more than one GPU (or more than one SM).  Here the fit frames are cut into
        clips of ``n_frames_per_clip`` exactly like ``ik_only`` does (root optimisation on each clip's first frame in the
        first pass, warm start carried per clip from pass to pass), the clips are block-partitioned over ranks, and the
        m-phase statistics of every rank's sampled frames are all-reduced.  With a single clip and a single rank this is
        the reference schedule.  Results differ from `fit_offsets` only through the different warm starts.
        """
        kp_data = np.asarray(kp_data, dtype=np.float32)
        F = int(n_frames_per_clip or self.cfg.stac.n_frames_per_clip)
        clips = utils.batch_kp_data(kp_data, F, continuous=False)
        C = clips.shape[0]
        eng, q = self._engine, self.stac_core_obj.q_solver
        rank, ws = parallel.world()
        lo, hi = parallel.shard_range(C, rank, ws)
        kp_dev = eng.f32(np.ascontiguousarray(clips[lo:hi]))
        kp_flat = kp_dev.reshape(-1, kp_dev.shape[-1])
        self._offsets = np.array(self._setup.initial_offsets)
        qio = eng.f32(self._mj_model.qpos0).repeat(hi - lo, 1).contiguous()
        has_root = self._root_kp_idx != -1 and not self._fixed
        tidx_all = self.time_indices
        if tidx_all is None:
            tidx_all = compute_stac.sample_time_indices(C * F, int(self.cfg.model.N_SAMPLE_FRAMES))
        tidx_all = np.asarray(tidx_all)  # kept in the sample's own order: the statistics are summed in this order
        mine = tidx_all[(tidx_all >= lo * F) & (tidx_all < hi * F)] - lo * F  # this rank's share of the frame sample
        mjx_model = stac_core.StacModel(engine=eng, site_pos=eng.f32(self._offsets, (eng.K, 3)))
        out = None
        for n_iter in range(self.cfg.model.N_ITERS + 1):
            last = n_iter == self.cfg.model.N_ITERS
            print("Final pose optimization" if last else f"Calibration iteration: {n_iter + 1}/{self.cfg.model.N_ITERS}", flush=True)
            out = eng.pose_clips(
                kp_dev, qio, mjx_model.site_pos, self._lb, self._ub, self._indiv_parts, do_root=1 if (has_root and n_iter == 0) else 0,
                root_kp_idx=max(self._root_kp_idx, 0), trunk_kps=self._trunk_kps, root_dims=4 if self._slidejoint else 7,
                tol=q.tol, maxiter=q.maxiter, maxls=q.maxls,
            )  # fmt: skip
            if last:
                break
            idx = torch.as_tensor(mine, device=eng.device, dtype=torch.long)
            res = self.stac_core_obj.m_opt(
                mjx_model, None, kp_flat[idx], out["qpos"].reshape(-1, eng.nq)[idx], self._offsets, self._is_regularized,
                self.cfg.model.M_REG_COEF, self._body_site_idxs, reduce_fn=parallel.allreduce_m_stats if ws > 1 else None,
            )  # fmt: skip
            print(f"Final residual error of {float(res.error)}")
            mjx_model = mjx_model.replace(site_pos=res.params)
            self._offsets = res.params.cpu().numpy()
            # offset_optimization ends with utils.kinematics (compute_stac.py:163), which normalises the free / ball
            # quaternions of the carried qpos once more; keep that so a single clip reproduces fit_offsets bit for bit
            qio = eng.fk(qio, mjx_model.site_pos)[0].contiguous()
        self.last_stats = {"iters": out["iters"], "ls_evals": out["ls_evals"], "status": out["status"]}
        res = {k: out[k] for k in ("qpos", "xpos", "xquat", "sites")}
        if ws > 1:
            res = {k: parallel.allgather_blocks(v.contiguous(), C) for k, v in res.items()}
        flat = lambda t: t.reshape((C * F,) + tuple(t.shape[2:])).cpu().numpy()
        return io.StacData(
            qpos=flat(res["qpos"]), xpos=flat(res["xpos"]), xquat=flat(res["xquat"]), marker_sites=flat(res["sites"]),
            offsets=np.array(self._offsets), names_qpos=self._part_names, names_xpos=self._body_names,
            kp_data=clips.reshape(C * F, -1), kp_names=self._kp_names,
        )  # fmt: skip

    # ------------------------------------------------------------------
    def ik_only(self, kp_data, offsets, *, edge_effects: bool = False, infer_qvels: bool = False) -> io.StacData:
        """Inverse kinematics with fixed offsets over independent clips (reference ``stac.py:356-454``).

        The reference vmaps ``root_optimization`` and ``pose_optimization`` over clips; here the whole phase is ONE
        fused launch (``stacb_pose_clips`` with ``do_root=1``) on this rank's block of clips.  Outputs are produced
        directly in the reference's packed layout (``_package_data(batched=True)``, ``stac.py:483-486``): qpos / xpos /
        xquat clip-major -- which is the kernel's native layout -- and marker_sites frame-major (transposed on the GPU).

        ``edge_effects`` / ``infer_qvels`` (keyword extensions; the reference does both on the host afterwards,
        ``main.py:118-133``) run ``utils.handle_edge_effects`` and ``utils.compute_velocity_from_kinematics`` as device
        epilogues on the packed outputs before the single device-to-host copy.
        """
        _nvtx = torch.cuda.nvtx
        _nvtx.range_push("stac.ik_only/h2d")
        kp_data = np.asarray(kp_data, dtype=np.float32)
        eng = self._engine
        Fc = int(self.cfg.stac.n_frames_per_clip)
        continuous = bool(self.cfg.stac.continuous)
        ov = utils.CONTINUOUS_BATCH_OVERLAP if continuous else 0
        C, F = int(kp_data.shape[0] // Fc), Fc + ov
        if C < 1 or (continuous and kp_data.shape[0] < F):
            raise ValueError("not enough frames for one clip")
        rank, ws = parallel.world()
        lo, hi = parallel.shard_range(C, rank, ws)
        offsets = np.asarray(offsets, dtype=np.float32)
        site_pos = eng.f32(offsets, (eng.K, 3))
        # Keypoint ingest: this rank's frames go straight from the session array into ONE page-locked staging buffer -- clips are
        # windows of F frames every Fc frames of it (utils.batch_kp_data, reference utils.py:350-389), read in place by the kernel,
        # so the `continuous` look-ahead is neither materialised per clip on the host nor copied twice.  The reference wrap-pads
        # the LAST clip with its own first frames (utils.py:377-381): appended behind the session on the rank that owns it.
        n_rows = max(hi - lo - 1, 0) * Fc + F if hi > lo else 0
        stage = self._pinned("kp_in", (n_rows, kp_data.shape[1]), torch.float32)
        if hi > lo:
            src = kp_data[lo * Fc : min(lo * Fc + n_rows, C * Fc)]
            stage[: len(src)].copy_(torch.from_numpy(src))
            if len(src) < n_rows:  # only the rank holding the last clip: its look-ahead wraps onto the clip's own start
                last = np.pad(kp_data[(C - 1) * Fc : (C - 1) * Fc + F], ((0, ov), (0, 0)), mode="wrap")
                if len(last) != F:  # left-over frames behind the last clip: the reference's np.stack of unequal windows fails too
                    raise ValueError("all input arrays must have the same shape (session length is not a multiple of n_frames_per_clip)")
                stage[len(src) :].copy_(torch.from_numpy(np.ascontiguousarray(last[Fc : Fc + n_rows - len(src)])))
        kp_dev = stage.to(eng.device, non_blocking=True)
        qio = eng.f32(self._mj_model.qpos0).repeat(hi - lo, 1).contiguous()
        has_root = self._root_kp_idx != -1 and not self._fixed
        if self._root_kp_idx == -1:
            print("Missing or invalid ROOT_OPTIMIZATION_KEYPOINT, skipping root_optimization()")
        elif self._fixed:
            print("ROOT_OPTIMIZATION_KEYPOINT specified but model has fixed root, skipping root_optimization()")
        q = self.stac_core_obj.q_solver
        _nvtx.range_pop()
        _nvtx.range_push("stac.ik_only/pose_clips")
        out = eng.pose_clips(
            kp_dev, qio, site_pos, self._lb, self._ub, self._indiv_parts, do_root=1 if has_root else 0,
            root_kp_idx=max(self._root_kp_idx, 0), trunk_kps=self._trunk_kps, root_dims=4 if self._slidejoint else 7,
            tol=q.tol, maxiter=q.maxiter, maxls=q.maxls, session=(hi - lo, F, Fc),
        )  # fmt: skip
        _nvtx.range_pop()
        self.last_stats = {"iters": out["iters"], "ls_evals": out["ls_evals"], "status": out["status"], "root_stats": out["root_stats"]}
        res = {k: out[k] for k in ("qpos", "xpos", "xquat", "sites", "err")}
        if ws > 1:  # hand every rank the complete result (blocks are contiguous in the clip-major layout)
            res = {k: parallel.allgather_blocks(v.contiguous(), C) for k, v in res.items()}
        nb, K = eng.nbody, eng.K
        dev = {
            "qpos": res["qpos"].reshape(C * F, eng.nq),
            "xpos": res["xpos"].reshape(C * F, nb, 3),
            "xquat": res["xquat"].reshape(C * F, nb, 4),
            "marker_sites": res["sites"].transpose(0, 1).reshape(C * F, K, 3),  # frame-major interleave (stac.py:486)
            "err": res["err"],
        }
        # StacData.kp_data: the reference returns the batched keypoints flattened (stac.py:498-500); after the cross-fade of identical
        # look-ahead copies that is the session itself up to the blend's rounding, so it is produced on the host from the input
        kp_packed = utils.batch_kp_data(kp_data, Fc, continuous=continuous)
        kp_packed = kp_packed.reshape(-1, kp_packed.shape[-1])
        if edge_effects:  # device epilogue: sigmoid cross-fade of the look-ahead overlap, overlaps removed (utils.py:393-461)
            if not self.cfg.stac.continuous:
                raise ValueError("edge_effects needs cfg.stac.continuous clips (the overlap is what is cross-faded)")
            _nvtx.range_push("stac.ik_only/edge_crossfade")
            Fc, ov = int(self.cfg.stac.n_frames_per_clip), utils.CONTINUOUS_BATCH_OVERLAP
            for k in ("qpos", "xpos", "xquat", "marker_sites"):
                dev[k] = eng.edge_crossfade(dev[k].contiguous(), Fc, ov)
            kp_packed = utils.edge_crossfade_host(kp_packed, Fc)
            _nvtx.range_pop()
        if infer_qvels:  # device epilogue: finite-difference velocities per clip (utils.py:302-347)
            _nvtx.range_push("stac.ik_only/qvel")
            dev["qvel"] = eng.qvel(dev["qpos"].contiguous(), int(self.cfg.stac.n_frames_per_clip), float(self._mj_model.opt.timestep), self._freejoint)
            _nvtx.range_pop()
        _nvtx.range_push("stac.ik_only/d2h")
        host = {}
        for k, v in dev.items():  # the step's D2H read: ONE copy into page-locked arrays that the returned StacData owns
            buf = self._pinned_out("out_" + k, tuple(v.shape), v.dtype)
            buf.copy_(v, non_blocking=True)
            host[k] = buf
        torch.cuda.synchronize(eng.device)
        _nvtx.range_pop()
        arrays = {k: self._hand_out("out_" + k, host[k]) for k in host}
        _, mean, std = self._get_error_stats(arrays["err"])
        print(f"Mean: {mean}")
        print(f"Standard deviation: {std}")
        data = io.StacData(
            qpos=arrays["qpos"], xpos=arrays["xpos"], xquat=arrays["xquat"], marker_sites=arrays["marker_sites"], offsets=np.array(offsets),
            names_qpos=self._part_names, names_xpos=self._body_names, kp_data=kp_packed, kp_names=self._kp_names,
        )  # fmt: skip
        if infer_qvels:
            data.qvel = arrays["qvel"]
        return data

    def _pinned(self, key: str, shape, dtype) -> torch.Tensor:
        """Reusable page-locked host buffer (allocating pinned memory per call would dominate small sessions)."""
        cache = self.__dict__.setdefault("_pin_cache", {})
        buf = cache.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(tuple(shape), dtype=dtype).pin_memory() if int(np.prod(shape)) else torch.empty(tuple(shape), dtype=dtype)
            cache[key] = buf
        return buf

    def _pinned_out(self, key: str, shape, dtype) -> torch.Tensor:
        """Page-locked OUTPUT buffer.  The numpy view handed to the caller (`_hand_out`) aliases it, so it is reused by a later call
        only once every array derived from that view has been released; otherwise a fresh buffer is allocated (no copy either way)."""
        live = self.__dict__.setdefault("_pin_out_live", {})
        ref = live.get(key)
        if ref is not None and ref() is not None:  # the previous result is still referenced by the caller
            self.__dict__.setdefault("_pin_cache", {}).pop(key, None)
        return self._pinned(key, shape, dtype)

    def _hand_out(self, key: str, buf: torch.Tensor) -> np.ndarray:
        import weakref

        arr = buf.numpy()
        self.__dict__.setdefault("_pin_out_live", {})[key] = weakref.ref(arr)
        return arr

    # ------------------------------------------------------------------
    def _package_data(self, mjx_model, qposes, xposes, xquats, marker_sites, kp_data, batched: bool = False) -> io.StacData:
        """Package results (reference ``stac.py:456-503``), layout quirks included."""
        if batched:
            offsets = mjx_model.site_pos.cpu().numpy()
            qposes = qposes.reshape(-1, qposes.shape[-1])
            xposes = xposes.reshape(-1, *xposes.shape[2:], order="F")
            xquats = xquats.reshape(-1, *xquats.shape[2:], order="F")
            marker_sites = marker_sites.reshape(-1, *marker_sites.shape[2:])
        else:
            offsets = self._offsets
        offsets = np.array(offsets)
        kp_data = np.asarray(kp_data).reshape(-1, kp_data.shape[-1])
        return io.StacData(
            qpos=qposes, xpos=xposes, xquat=xquats, marker_sites=marker_sites, offsets=offsets,
            names_qpos=self._part_names, names_xpos=self._body_names, kp_data=kp_data, kp_names=self._kp_names,
        )  # fmt: skip
