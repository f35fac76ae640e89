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
    def ik_only(self, kp_data, offsets) -> io.StacData:
        """Inverse kinematics with fixed offsets over independent clips (reference ``stac.py:356-454``)."""
        kp_data = np.asarray(kp_data, dtype=np.float32)
        batched_kp_data = utils.batch_kp_data(kp_data, self.cfg.stac.n_frames_per_clip, continuous=self.cfg.stac.continuous)
        eng = self._engine
        C = batched_kp_data.shape[0]
        rank, ws = parallel.world()
        lo, hi = parallel.shard_range(C, rank, ws)
        mjx_model, data0 = self._load(np.asarray(offsets, dtype=np.float32))
        # pinned staging -> device (the step's H2D copy)
        host = torch.from_numpy(np.ascontiguousarray(batched_kp_data[lo:hi]))
        kp_dev = host.pin_memory().to(eng.device, non_blocking=True) if host.numel() else host.to(eng.device)
        mjx_data = stac_core.StacState(qpos=data0.qpos.repeat(hi - lo, 1))
        if self._root_kp_idx == -1:
            print("Missing or invalid ROOT_OPTIMIZATION_KEYPOINT, skipping root_optimization()")
        elif not self._fixed:
            mjx_data = compute_stac.root_optimization(
                self.stac_core_obj, mjx_model, mjx_data, kp_dev, self._root_kp_idx, self._lb, self._ub,
                self._body_site_idxs, self._trunk_kps,
            )  # fmt: skip
        else:
            print("ROOT_OPTIMIZATION_KEYPOINT specified but model has fixed root, skipping root_optimization()")
        mjx_data, qposes, xposes, xquats, marker_sites, frame_time, frame_error = compute_stac.pose_optimization(
            self.stac_core_obj, mjx_model, mjx_data, kp_dev, self._lb, self._ub, self._body_site_idxs, self._indiv_parts
        )
        self.last_stats = mjx_data.solver_stats
        if ws > 1:  # hand every rank the complete result, clip-major like the single-process run
            qposes = parallel.allgather_blocks(qposes.contiguous(), C)
            xposes = parallel.allgather_blocks(xposes.transpose(0, 1).contiguous(), C).transpose(0, 1)
            xquats = parallel.allgather_blocks(xquats.transpose(0, 1).contiguous(), C).transpose(0, 1)
            marker_sites = parallel.allgather_blocks(marker_sites.transpose(0, 1).contiguous(), C).transpose(0, 1)
            frame_error = parallel.allgather_blocks(frame_error.transpose(0, 1).contiguous(), C).transpose(0, 1)
        _, mean, std = self._get_error_stats(frame_error)
        print(f"Mean: {mean}")
        print(f"Standard deviation: {std}")
        return self._package_data(
            mjx_model, qposes.cpu().numpy(), xposes.cpu().numpy(), xquats.cpu().numpy(), marker_sites.cpu().numpy(),
            batched_kp_data, batched=True,
        )  # fmt: skip

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
