"""Throughput of the scheduling modes of the register-resident path against the number of chains (GPU box).
    python tools/mode_sweep.py [model] modes... -- chains..."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from stac_mjx_b200 import model, synth
from stac_mjx_b200.engine import Engine
args = sys.argv[1:]
name = args[0]; sep = args.index("--")
modes = [int(a) for a in args[1:sep]]; chains = [int(a) for a in args[sep + 1:]]
t, cfg = model.load_fixture(name)
kpn = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys()); s = model.make_setup(t, cfg.model, kpn)
eng = Engine(t, s.site_bodies)
F = 250
has_root = s.root_kp_idx >= 0 and int(t.jnt_type[0]) in (0, 2)
kw = dict(do_root=1 if has_root else 0, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=float(cfg.model.FTOL))
base, _, _ = synth.synth_session(t, s, 64 * F, F, seed=7)
base = base.reshape(64, F, -1)
for C in chains:
    kpd = torch.tensor(np.tile(base, ((C + 63) // 64, 1, 1))[:C], device="cuda")
    for mode in modes:
        eng.set_mode(mode)
        qio = torch.tensor(np.tile(t.qpos0.astype(np.float32), (C, 1)), device="cuda")
        eng.pose_clips(kpd[:, :3].contiguous(), qio.clone(), s.initial_offsets, s.lb, s.ub, s.indiv_parts, **kw)
        torch.cuda.synchronize()
        best = 1e30
        for rep in range(2):
            q2 = qio.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o = eng.pose_clips(kpd, q2, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **kw)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{name} C={C} mode {mode}: {best:.1f} ms -> {C*F/best*1e3:.0f} frames/s", flush=True)
