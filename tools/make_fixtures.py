"""Compile the reference's model assets into committed descriptor fixtures.

Run in the authoring container (needs /root/reference, which the GPU box does
not have):  python tools/make_fixtures.py
Writes stac_mjx_b200/assets/<name>.json = {tree, model_cfg, stac_cfg}.
The MJCF/YAML files are read as input data; nothing is copied verbatim.
"""
import sys
from pathlib import Path

import yaml

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from stac_mjx_b200 import model  # noqa: E402

REF = Path("/root/reference")
MODELS = {
    # name: (model yaml, stac yaml)
    "rodent": ("rodent", "stac"),
    "mouse": ("mouse", "stac_mouse"),
    "celegans": ("celegans", "stac_celegans"),
    "fly_treadmill": ("fly_treadmill", "stac_fly_treadmill"),
    "fly_tethered": ("fly_tethered", "stac_fly_tethered"),
    "synth_data": ("synth_data", "stac_synth_data"),
}

for name, (myaml, syaml) in MODELS.items():
    mcfg = yaml.safe_load((REF / "configs/model" / f"{myaml}.yaml").read_text())
    scfg = yaml.safe_load((REF / "configs/stac" / f"{syaml}.yaml").read_text())
    mcfg.setdefault("MARKER_SIZE", 0.005)
    mcfg.pop("KEYPOINT_COLOR_PAIRS", None)  # rendering only
    tree = model.compile_spec(model.build_spec(REF / mcfg["MJCF_PATH"], mcfg))
    out = model.ASSET_DIR / f"{name}.json"
    model.save_fixture(out, tree, mcfg, scfg, f"{mcfg['MJCF_PATH']} + configs/model/{myaml}.yaml")
    kp_names = list(mcfg["KEYPOINT_MODEL_PAIRS"].keys())
    setup = model.make_setup(tree, mcfg, kp_names)
    act = tree.active_bodies(setup.site_bodies)
    depth = tree.body_depth()
    print(
        f"{name}: nbody={tree.nbody} njnt={tree.njnt} nq={tree.nq} nsite={tree.nsite} K={len(kp_names)} "
        f"P={setup.indiv_parts.shape[0]} active={len(act)} depth_act={depth[act].max()} depth={depth.max()} "
        f"maxjnt/body={tree.body_jntnum.max()} bytes={out.stat().st_size}"
    )
