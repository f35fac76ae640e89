"""Pipeline entry points with the reference's call surface (``stac_mjx/main.py:18-139``).

``load_configs(config_dir, config_name)`` and ``run_stac(cfg, kp_data, kp_names, base_path)`` keep the reference's
arguments, return value ``(fit_path, ik_path | None)``, skip flags and the two ``ValueError`` conditions; the work
in between is delegated to `Stac`, whose fit / IK phases are single fused launches on the GPU.
"""

from __future__ import annotations

import time
from pathlib import Path

from . import io, utils
from .config import compose_config
from .stac import Stac


def load_configs(config_dir, config_name: str = "config"):
    cfg = compose_config(config_dir, config_name=config_name)
    print("Config loaded and validated.")
    return cfg


def _require_matching_columns(kp_data, kp_names) -> None:
    want = 3 * len(kp_names)
    if kp_data.shape[1] != want:
        raise ValueError(
            f"kp_data has {kp_data.shape[1]} columns but expected {want} ({len(kp_names)} keypoints × 3). "
            "kp_data must be shaped (n_frames, n_keypoints * 3) with one name per keypoint in kp_names."
        )


def _fit_phase(stac: Stac, cfg, kp_data, out_path: Path) -> None:
    frames = kp_data[: cfg.stac.n_fit_frames]
    print(f"Running fit. Mocap data shape: {frames.shape}")
    result = stac.fit_offsets(frames)
    print(f"saving data to {out_path}", flush=True)
    io.save_data_to_h5(config=cfg, file_path=out_path, **result.as_dict())


def _ik_phase(stac: Stac, kp_data, fit_path: Path, out_path: Path, t0: float) -> None:
    print("Running ik_only()")
    cfg, fitted = io.load_stac_data(fit_path)  # the stored config replaces the live one, as in the reference (main.py:111-113)
    # the reference cross-fades and differentiates on the host afterwards (main.py:118-133); here both are device epilogues of the IK pass
    if cfg.stac.continuous:
        print("Handling edge effects...")
    result = stac.ik_only(kp_data, fitted.offsets, edge_effects=bool(cfg.stac.continuous), infer_qvels=bool(cfg.stac.get("infer_qvels", False)))
    print(f"Final qpos shape: {result.qpos.shape}")
    print(f"Saving data to {out_path}. Finished in {(time.time() - t0) / 60:.2f} minutes")
    io.save_data_to_h5(config=cfg, file_path=out_path, **result.as_dict())


def run_stac(cfg, kp_data, kp_names: list[str], base_path: Path | None = None, *, tree=None):
    """Offset fit on the first ``n_fit_frames`` frames, then clip-parallel IK over the whole session."""
    base = Path.cwd() if base_path is None else Path(base_path)
    _require_matching_columns(kp_data, kp_names)
    t0 = time.time()
    fit_path, ik_path = base / cfg.stac.fit_offsets_path, base / cfg.stac.ik_only_path
    stac = Stac(base / cfg.model.MJCF_PATH, cfg, kp_names, tree=tree)

    if cfg.stac.skip_fit_offsets:
        print("Skipping fit_offsets. To change this behavior, set cfg.stac.skip_fit_offsets to False.")
    else:
        _fit_phase(stac, cfg, kp_data, fit_path)

    if cfg.stac.skip_ik_only:
        print("Skipping IK-only phase. To change this behavior, set cfg.stac.skip_ik_only to False.")
        return fit_path, None
    if kp_data.shape[0] % cfg.stac.n_frames_per_clip != 0:
        raise ValueError(
            f"n_frames_per_clip ({cfg.stac.n_frames_per_clip}) must divide evenly with the total number of mocap frames({kp_data.shape[0]})"
        )
    _ik_phase(stac, kp_data, fit_path, ik_path, t0)
    return fit_path, ik_path
