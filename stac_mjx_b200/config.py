"""Configuration loading (reference ``stac_mjx/config.py``).

The reference composes its YAML tree with Hydra and validates it against
dataclasses through OmegaConf (``config.py:73-88``).  When ``hydra`` and
``omegaconf`` are importable the same call is used; otherwise (this image has
neither) a small composer understands the only Hydra feature the reference's
config tree uses -- a ``defaults`` list of ``group: option`` entries -- and
returns an attribute/item-addressable mapping with the ``in`` / ``.get``
behaviour the hot path relies on (``stac.py:118,173,230``).
"""

from __future__ import annotations

from pathlib import Path
from typing import Any, Iterable

import yaml

MODEL_KEYS = (
    "MJCF_PATH FTOL ROOT_FTOL LIMB_FTOL N_ITERS N_ITER_Q KP_NAMES KEYPOINT_MODEL_PAIRS "
    "KEYPOINT_INITIAL_OFFSETS ROOT_OPTIMIZATION_KEYPOINT TRUNK_OPTIMIZATION_KEYPOINTS "
    "INDIVIDUAL_PART_OPTIMIZATION KEYPOINT_COLOR_PAIRS SCALE_FACTOR MOCAP_SCALE_FACTOR "
    "SITES_TO_REGULARIZE RENDER_FPS N_SAMPLE_FRAMES M_REG_COEF MARKER_SIZE"
).split()
STAC_KEYS = (
    "fit_offsets_path ik_only_path data_path num_clips n_fit_frames skip_fit_offsets "
    "skip_ik_only infer_qvels n_frames_per_clip mujoco continuous"
).split()


class Cfg(dict):
    """dict with attribute access; nested dicts are wrapped on the way in."""

    def __init__(self, data: dict | None = None):
        super().__init__()
        for k, v in (data or {}).items():
            self[k] = v

    @staticmethod
    def _wrap(v: Any) -> Any:
        if isinstance(v, dict) and not isinstance(v, Cfg):
            return Cfg(v)
        if isinstance(v, list):
            return [Cfg._wrap(x) for x in v]
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, Cfg._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self) -> dict:
        def un(v):
            if isinstance(v, dict):
                return {k: un(x) for k, x in v.items()}
            if isinstance(v, list):
                return [un(x) for x in v]
            return v

        return un(self)


def _apply_override(cfg: Cfg, item: str) -> None:
    key, _, val = item.partition("=")
    node = cfg
    parts = key.lstrip("+").split(".")
    for p in parts[:-1]:
        node = node.setdefault(p, Cfg())
    node[parts[-1]] = yaml.safe_load(val)


def _compose_yaml(config_dir: Path, config_name: str, overrides: Iterable[str]) -> Cfg:
    overrides = list(overrides)
    top = yaml.safe_load((config_dir / f"{config_name}.yaml").read_text()) or {}
    groups: dict[str, str] = {}
    for entry in top.pop("defaults", []):
        if isinstance(entry, dict):
            groups.update({k: v for k, v in entry.items()})
    # group selections such as "model=mouse" replace the defaults entry
    rest = []
    for ov in overrides:
        k, _, v = ov.partition("=")
        if k in groups and "." not in k:
            groups[k] = v
        elif not k.startswith("hydra"):
            rest.append(ov)
    cfg = Cfg()
    for group, option in groups.items():
        cfg[group] = yaml.safe_load((config_dir / group / f"{option}.yaml").read_text()) or {}
    for k, v in top.items():
        cfg[k] = v
    for ov in rest:
        _apply_override(cfg, ov)
    cfg.setdefault("model", Cfg()).setdefault("MARKER_SIZE", 0.005)
    return cfg


def compose_config(config_path: Path | str, config_name: str = "config", overrides: Iterable[str] | None = None):
    """Load and validate the configuration (reference ``config.py:73-88``)."""
    overrides = list(overrides or [])
    config_dir = Path(config_path).resolve()
    try:
        from hydra import compose, initialize_config_dir
        from omegaconf import OmegaConf
    except ImportError:
        return _compose_yaml(config_dir, config_name, overrides)
    overrides.extend(["hydra/job_logging=disabled", "hydra/hydra_logging=disabled"])
    with initialize_config_dir(config_dir=str(config_dir), version_base=None):
        return OmegaConf.create(OmegaConf.to_container(compose(config_name=config_name, overrides=overrides)))
