"""Throughput sweep (BASELINE.json configs 4 and 5): python tools/sweep.py model:frames[:clip] ...  -> JSON lines."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from stac_mjx_b200 import flops, model, synth
from stac_mjx_b200.engine import Engine

for spec in sys.argv[1:]:
    parts = spec.split(":")
    name, n_frames = parts[0], int(float(parts[1]))
    tree, cfg = model.load_fixture(name)
    F = int(parts[2]) if len(parts) > 2 else int(cfg.stac.get("n_frames_per_clip", 250))
    kpn = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
    s = model.make_setup(tree, cfg.model, kpn)
    C = max(1, n_frames // F)
    # a pool of distinct clips, tiled to the requested session length (keeps host-side generation cheap)
    pool = min(C, 128)
    kp, _, _ = synth.synth_session(tree, s, pool * F, F, seed=5)
    kp = np.tile(kp.reshape(pool, F, -1), ((C + pool - 1) // pool, 1, 1))[:C]
    eng = Engine(tree, s.site_bodies, 0)
    has_root = s.root_kp_idx >= 0 and int(tree.jnt_type[0]) in (0, 2)
    kw = dict(do_root=1 if has_root else 0, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=float(cfg.model.FTOL))
    kpd = torch.tensor(kp, device="cuda")
    q0 = torch.tensor(np.tile(tree.qpos0.astype(np.float32), (C, 1)), device="cuda")
    out = {}
    eng.pose_clips(kpd[:, :2].contiguous(), q0.clone(), s.initial_offsets, s.lb, s.ub, s.indiv_parts, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = eng.pose_clips(kpd, q0.clone(), s.initial_offsets, s.lb, s.ub, s.indiv_parts, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it, ls = int(o["iters"].sum()) + int(o["root_stats"][:, [0, 2]].sum()), int(o["ls_evals"].sum()) + int(o["root_stats"][:, [1, 3]].sum())
    pc = flops.path_cost(tree, s.site_bodies)
    fl = pc.total(it, ls, C * F, 1 + s.indiv_parts.shape[0])
    fl_exec = pc.executed(it, ls, C * F, 1 + s.indiv_parts.shape[0])
    resid = float(torch.linalg.norm(o["sites"] - kpd.reshape(C, F, -1, 3), dim=-1).mean())
    print(json.dumps({"model": name, "frames": C * F, "clips": C, "clip_frames": F, "ms": ms, "frames_per_s": C * F / ms * 1e3,
                      "iters_per_frame": it / (C * F), "ls_per_iter": ls / max(it, 1), "tflops_algorithmic": fl / ms / 1e9, "tflops_executed": fl_exec / ms / 1e9, "path": "register-resident" if eng.path else "general",
                      "mean_marker_residual_m": resid, "nonfinite": int((o["status"] != 0).sum())}), flush=True)
    del eng, kpd, out, o
    torch.cuda.empty_cache()
