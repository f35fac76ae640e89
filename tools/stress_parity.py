"""Randomised parity stress (one-off, GPU box): many seeded clips per model, GPU vs canonical oracle, bitwise.
    python tools/stress_parity.py [clips] [frames]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from oracle.oracle import Oracle
from stac_mjx_b200 import model, synth
from stac_mjx_b200.engine import Engine

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
F = int(sys.argv[2]) if len(sys.argv) > 2 else 6
bad = 0
for name in ("rodent", "celegans", "fly_treadmill", "fly_tethered", "synth_data"):
    tree, cfg = model.load_fixture(name)
    kpn = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
    s = model.make_setup(tree, cfg.model, kpn)
    eng, orc = Engine(tree, s.site_bodies, 0), Oracle(tree, s.site_bodies, np.float32, 2)
    has_root = s.root_kp_idx >= 0 and int(tree.jnt_type[0]) in (0, 2)
    kw = dict(do_root=1 if has_root else 0, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=float(cfg.model.FTOL))
    for seed, noise in ((101, 1e-3), (202, 5e-3), (303, 2e-2)):
        kp, _, _ = synth.synth_session(tree, s, C * F, F, seed=seed, obs_sigma=noise)
        kp = kp.reshape(C, F, -1)
        rng = np.random.default_rng(seed)
        offs = (s.initial_offsets + rng.normal(scale=2e-3, size=s.initial_offsets.shape)).astype(np.float32)
        qio = torch.tensor(np.tile(tree.qpos0.astype(np.float32), (C, 1)), device="cuda")
        out = eng.pose_clips(kp, qio, offs, s.lb, s.ub, s.indiv_parts, **kw)
        ref = orc.pose_clips(kp, tree.qpos0, offs, s.lb, s.ub, s.indiv_parts, nthreads=os.cpu_count(), **kw)
        same = all(np.array_equal(out[k].cpu().numpy(), ref[k]) for k in ("qpos", "xpos", "xquat", "sites", "err", "iters", "ls_evals"))
        dq = float(np.abs(out["qpos"].cpu().numpy() - ref["qpos"]).max())
        print(f"{name:14s} seed {seed} noise {noise:g}: {C} clips x {F} frames bitwise equal: {same} (max |dqpos| {dq:.2e}), "
              f"iters/frame {ref['iters'].sum(-1).mean():.0f}", flush=True)
        bad += 0 if same else 1
print("STRESS", "PASS" if bad == 0 else f"FAIL ({bad})")
sys.exit(1 if bad else 0)
