"""Short launch of the fused solver for ncu captures: python tools/profile_target.py [clips] [frames] [model]."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from stac_mjx_b200 import model, synth
from stac_mjx_b200.engine import Engine

C = int(sys.argv[1]) if len(sys.argv) > 1 else 72
F = int(sys.argv[2]) if len(sys.argv) > 2 else 6
name = sys.argv[3] if len(sys.argv) > 3 else "rodent"
tree, cfg = model.load_fixture(name)
kpn = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
s = model.make_setup(tree, cfg.model, kpn)
kp, _, _ = synth.synth_session(tree, s, C * F, F, seed=1)
eng = Engine(tree, s.site_bodies, 0)
import os
eng.set_mode(int(os.environ.get("STACB_MODE", "-1")))
eng.set_path(int(os.environ.get("STACB_PATH", "0")))
has_root = s.root_kp_idx >= 0 and int(tree.jnt_type[0]) in (0, 2)
for _ in range(3):
    qio = torch.tensor(np.tile(tree.qpos0.astype(np.float32), (C, 1)), device="cuda")
    out = eng.pose_clips(kp.reshape(C, F, -1), qio, s.initial_offsets, s.lb, s.ub, s.indiv_parts, do_root=1 if has_root else 0,
                         root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=float(cfg.model.FTOL))
    torch.cuda.synchronize()
print("iters", int(out["iters"].sum()), "ls", int(out["ls_evals"].sum()))
