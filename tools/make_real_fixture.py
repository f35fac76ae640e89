"""tests/golden/rodent_real250.npz: the first 250 frames of the reference's real rat23 mocap fixture
(tests/data/test_rodent_mocap_1000_frames.mat, BASELINE config 1) run through io.load_data with the rodent config,
plus the oracle's IK output on them (c32 canonical order, g32 fast order = the default CUDA path, f64 / m32 MJX order in
float64 / float32).  Needs /root/reference; run in the authoring container."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.oracle import Oracle  # noqa: E402
from stac_mjx_b200 import io, model  # noqa: E402
from stac_mjx_b200.config import Cfg  # noqa: E402

tree, cfg = model.load_fixture("rodent")
cfg = Cfg(cfg.to_dict())
cfg.stac.data_path = "tests/data/test_rodent_mocap_1000_frames.mat"
kp, names = io.load_data(cfg, base_path="/root/reference")
assert kp.shape == (1000, 69) and names == list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
kp = kp[:250]
s = model.make_setup(tree, cfg.model, names)
kw = dict(do_root=1, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=float(cfg.model.FTOL))
out = {"kp": kp}
for tag, dt, mode in (("c32", np.float32, 1), ("g32", np.float32, 2), ("f64", np.float64, 0), ("m32", np.float32, 0)):
    r = Oracle(tree, s.site_bodies, dt, mode).pose_clips(kp[None], tree.qpos0, s.initial_offsets, s.lb, s.ub, s.indiv_parts, nthreads=1, **kw)
    for k in ("qpos", "sites", "err", "iters", "ls_evals", "root_stats"):
        out[f"{tag}_{k}"] = r[k][0]
np.savez_compressed(ROOT / "tests" / "golden" / "rodent_real250.npz", **out)
res = np.linalg.norm(out["c32_sites"] - kp.reshape(250, -1, 3), axis=-1)
print("real clip: iters/frame", out["c32_iters"].sum(1).mean(), "marker residual mm mean", 1e3 * res.mean(),
      "f32-vs-f64 qpos max", np.abs(out["c32_qpos"] - out["f64_qpos"]).max())
