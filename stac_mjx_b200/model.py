"""Model -> fitting descriptor: sites, scaling, bounds, masks.

Restates the model-side setup of ``Stac.__init__`` (reference
``stac_mjx/stac.py:98-159``): `_build_body_spec` (``:185-207``: one site per
``KEYPOINT_MODEL_PAIRS`` entry at ``KEYPOINT_INITIAL_OFFSETS``, added *before*
``dm_scale_spec``), `_init_body_sites` (``:209-235``), bounds (``:128-133``),
part masks (``:135``) and the trunk mask (``:138-140``).
"""

from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from . import mjcf
from .config import Cfg
from .tree import TreeModel, align_joint_dims, compile_spec, part_masks

ASSET_DIR = Path(__file__).parent / "assets"


@dataclass
class FitSetup:
    """Everything `Stac` derives from (model, cfg, kp_names) before any solve."""

    tree: TreeModel
    site_idxs: np.ndarray  # int32 [K] indices into tree.site_*
    is_regularized: np.ndarray  # float32 [K, 3]
    lb: np.ndarray  # float32 [nq]
    ub: np.ndarray  # float32 [nq]
    part_names: list[str]
    indiv_parts: np.ndarray  # bool [P, nq]
    trunk_kps: np.ndarray  # bool [K]
    root_kp_idx: int
    body_names: list[str]

    @property
    def site_bodies(self) -> np.ndarray:
        return self.tree.site_bodyid[self.site_idxs]

    @property
    def initial_offsets(self) -> np.ndarray:
        return self.tree.site_pos[self.site_idxs].astype(np.float32)


def _offset(v) -> list[float]:
    if isinstance(v, str):
        return [float(p) for p in v.split(" ") if p != ""]
    return [float(p) for p in v]


def build_spec(xml_path: str | Path, model_cfg) -> mjcf.ModelSpec:
    """``Stac._build_body_spec``: parse, add keypoint sites, then scale."""
    spec = mjcf.parse_mjcf(xml_path)
    for key, body_name in model_cfg["KEYPOINT_MODEL_PAIRS"].items():
        spec.body(body_name).add_site(key, _offset(model_cfg["KEYPOINT_INITIAL_OFFSETS"][key]))
    return spec.scale(float(model_cfg["SCALE_FACTOR"]))


def compile_fit_tree(xml_path: str | Path, model_cfg) -> TreeModel:
    try:  # production: the real MuJoCo compiler, exactly as the reference does
        import mujoco

        from . import _mujoco_bridge

        return TreeModel.from_mjmodel(_mujoco_bridge.build_mjmodel(mujoco, xml_path, model_cfg))
    except ImportError:
        return compile_spec(build_spec(xml_path, model_cfg))


def make_setup(tree: TreeModel, model_cfg, kp_names: list[str]) -> FitSetup:
    pairs = model_cfg["KEYPOINT_MODEL_PAIRS"]
    site_idxs = np.array([tree.site_id(name) for name in pairs.keys()], dtype=np.int32)
    if (site_idxs < 0).any():
        raise ValueError("keypoint site missing from the compiled model")
    reg = list(model_cfg.get("SITES_TO_REGULARIZE", []) or [])
    is_regularized = np.array(
        [[1.0, 1.0, 1.0] if any(n == k for n in reg) else [0.0, 0.0, 0.0] for k in pairs.keys()],
        dtype=np.float32,
    ).reshape(-1, 3)
    lb, ub, part_names = align_joint_dims(tree.jnt_type, tree.jnt_range, tree.jnt_names)
    parts = model_cfg["INDIVIDUAL_PART_OPTIMIZATION"] if "INDIVIDUAL_PART_OPTIMIZATION" in model_cfg else None
    if "ROOT_OPTIMIZATION_KEYPOINT" in model_cfg and model_cfg["ROOT_OPTIMIZATION_KEYPOINT"] is not None:
        root_kp_idx = kp_names.index(model_cfg["ROOT_OPTIMIZATION_KEYPOINT"])
    else:
        root_kp_idx = -1
    trunk = model_cfg.get("TRUNK_OPTIMIZATION_KEYPOINTS", []) or []
    return FitSetup(
        tree=tree,
        site_idxs=site_idxs,
        is_regularized=is_regularized,
        lb=lb,
        ub=ub,
        part_names=part_names,
        indiv_parts=part_masks(part_names, parts),
        trunk_kps=np.array([n in trunk for n in kp_names], dtype=bool),
        root_kp_idx=root_kp_idx,
        body_names=list(tree.body_names),
    )


# --- committed fixtures (models compiled once from the reference's assets) ------


def save_fixture(path: Path, tree: TreeModel, model_cfg: dict, stac_cfg: dict, source: str) -> None:
    path.write_text(
        json.dumps({"source": source, "tree": tree.to_dict(), "model_cfg": model_cfg, "stac_cfg": stac_cfg})
    )


def load_fixture(name: str) -> tuple[TreeModel, Cfg]:
    """Load a pre-compiled model + config, e.g. ``load_fixture('rodent')``."""
    p = ASSET_DIR / f"{name}.json"
    if not p.exists():
        raise FileNotFoundError(f"no compiled model fixture {p}; run tools/make_fixtures.py")
    d = json.loads(p.read_text())
    cfg = Cfg({"model": d["model_cfg"], "stac": d["stac_cfg"]})
    return TreeModel.from_dict(d["tree"]), cfg
