"""User-level API (reference ``stac_mjx/main.py``): ``load_configs`` and ``run_stac``."""

from __future__ import annotations

import time
from pathlib import Path

import numpy as np

from . import io, utils
from .config import compose_config
from .stac import Stac


def load_configs(config_dir, config_name: str = "config"):
    """Load and validate configs (reference ``main.py:18-30``)."""
    cfg = compose_config(config_dir, config_name=config_name)
    print("Config loaded and validated.")
    return cfg


def run_stac(cfg, kp_data, kp_names: list[str], base_path: Path | None = None, *, tree=None):
    """fit_offsets -> ik_only -> save, with the reference's control flow and errors (``main.py:33-139``)."""
    if base_path is None:
        base_path = Path.cwd()
    expected_cols = len(kp_names) * 3
    if kp_data.shape[1] != expected_cols:
        raise ValueError(
            f"kp_data has {kp_data.shape[1]} columns but expected {expected_cols} "
            f"({len(kp_names)} keypoints × 3). "
            f"Ensure kp_data is shaped (n_frames, n_keypoints * 3) and that "
            f"kp_names length matches the number of keypoints in kp_data."
        )
    start_time = time.time()
    fit_offsets_path = base_path / cfg.stac.fit_offsets_path
    ik_only_path = base_path / cfg.stac.ik_only_path
    xml_path = base_path / cfg.model.MJCF_PATH
    stac = Stac(xml_path, cfg, kp_names, tree=tree)
    fit_offsets_data = None
    if not cfg.stac.skip_fit_offsets:
        kps = kp_data[: cfg.stac.n_fit_frames]
        print(f"Running fit. Mocap data shape: {kps.shape}")
        fit_offsets_data = stac.fit_offsets(kps)
        print(f"saving data to {fit_offsets_path}", flush=True)
        io.save_data_to_h5(config=cfg, file_path=fit_offsets_path, **fit_offsets_data.as_dict())
    else:
        print("Skipping fit_offsets. To change this behavior, set cfg.stac.skip_fit_offsets to False.")
    if cfg.stac.skip_ik_only:
        print("Skipping IK-only phase. To change this behavior, set cfg.stac.skip_ik_only to False.")
        return fit_offsets_path, None
    elif kp_data.shape[0] % cfg.stac.n_frames_per_clip != 0:
        raise ValueError(
            f"n_frames_per_clip ({cfg.stac.n_frames_per_clip}) must divide evenly with the total number of mocap frames({kp_data.shape[0]})"
        )
    print("Running ik_only()")
    cfg, fit_offsets_data = io.load_stac_data(fit_offsets_path)
    ik_only_data = stac.ik_only(kp_data, fit_offsets_data.offsets)
    if cfg.stac.continuous:
        print("Handling edge effects...")
        ik_only_data = utils.handle_edge_effects(ik_only_data, cfg.stac.n_frames_per_clip)
    print(f"Final qpos shape: {ik_only_data.qpos.shape}")
    print(f"Saving data to {ik_only_path}. Finished in {(time.time() - start_time) / 60:.2f} minutes")
    io.save_data_to_h5(config=cfg, file_path=ik_only_path, **ik_only_data.as_dict())
    return fit_offsets_path, ik_only_path
