"""GPU box probe of the register-resident path: bitwise diffs against oracle mode 2 in every scheduling mode, then timings.
    python tools/fast_probe.py [model] [time_clips]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from stac_mjx_b200 import model, synth
from stac_mjx_b200.engine import Engine
from oracle.oracle import Oracle

name = sys.argv[1] if len(sys.argv) > 1 else "rodent"
t, cfg = model.load_fixture(name)
kpn = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
s = model.make_setup(t, cfg.model, kpn)
K = len(kpn)
eng = Engine(t, s.site_bodies)
orc = Oracle(t, s.site_bodies, np.float32, 2)
print(name, "engine path", eng.path, "oracle fast", orc.fast_path)
off = s.initial_offsets
F = 250
kp, qtrue, _ = synth.synth_session(t, s, 4 * F, F)
tol = float(cfg.model.FTOL)
rng = np.random.default_rng(1)
qq = (qtrue[:64] + rng.normal(scale=0.02, size=qtrue[:64].shape)).astype(np.float32)
qm = np.ones(t.nq, bool); km = np.ones(3 * K, bool)
L, G = eng.loss_grad(qq, qq, kp[:64], qm, km, off)
L, G = L.cpu().numpy(), G.cpu().numpy()
dl = dg = 0
for i in range(64):
    l, g = orc.loss_grad(qq[i], qq[i], qm, kp[i], km, off)
    dl = max(dl, abs(float(l) - float(L[i])) / float(l)); dg = max(dg, float(np.abs(g - G[i]).max()))
print("loss rel maxdiff", dl, "grad maxdiff", dg, "gmax", np.abs(G).max())
has_root = s.root_kp_idx >= 0 and int(t.jnt_type[0]) in (0, 2)
nr = min(7, t.nq)
q0 = np.tile(t.qpos0.astype(np.float32), (8, 1))
if has_root: q0[:, :3] = kp[:8, 3 * s.root_kp_idx:3 * s.root_kp_idx + 3]
q0[:, -1] += 10.0  # a coordinate outside its box (passive for the rodent)
rq = np.zeros(t.nq, bool); rq[:nr] = True; tk = np.repeat(s.trunk_kps, 3) if s.trunk_kps.any() else km
p, e, it, ls = [x.cpu().numpy() for x in eng.q_opt(q0, kp[:8], rq, tk, off, s.lb, s.ub, tol)]
for i in range(8):
    po, eo, io, lo = orc.q_opt(q0[i], s.lb, s.ub, rq, kp[i], tk, off, tol)
    print(" solve", i, "iters", it[i], io, "ls", ls[i], lo, "err", e[i], eo, "dparams", np.abs(p[i] - po).max())
nclip = 4; Fs = 12
kpc = kp.reshape(nclip, F, -1)[:, :Fs].copy()
kw = dict(do_root=1 if has_root else 0, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=tol)
qinit = np.tile(t.qpos0.astype(np.float32), (nclip, 1)); qinit[:, -1] += 10.0
ref = orc.pose_clips(kpc, qinit, off, s.lb, s.ub, s.indiv_parts, nthreads=8, **kw)
for mode in (0, 1, 2, 3, 4):
    eng.set_mode(mode)
    qio = torch.tensor(qinit, device="cuda")
    out = eng.pose_clips(kpc, qio, off, s.lb, s.ub, s.indiv_parts, **kw)
    torch.cuda.synchronize()
    d = {k: float(np.abs(out[k].cpu().numpy() - ref[k]).max()) for k in ("qpos", "xpos", "xquat", "sites", "err")}
    print(" mode", mode, "clip maxdiff", d, "iters equal", np.array_equal(out["iters"].cpu().numpy(), ref["iters"]), "ls equal",
          np.array_equal(out["ls_evals"].cpu().numpy(), ref["ls_evals"]), "root", out["root_stats"].cpu().numpy()[0], ref["root_stats"][0],
          "status", out["status"].cpu().numpy(), "qio diff", float(np.abs(qio.cpu().numpy() - ref["qpos"][:, -1]).max()))
for C in [int(a) for a in sys.argv[2:]] or [72]:
    kpb, _, _ = synth.synth_session(t, s, C * F, F, seed=7)
    kpd = torch.tensor(kpb.reshape(C, F, -1), device="cuda")
    for path, mode in ((0, 1), (0, 3), (0, 4), (0, 0), (0, 2), (1, 1), (1, 0)):
        eng.set_path(path); eng.set_mode(mode)
        qio = torch.tensor(np.tile(t.qpos0.astype(np.float32), (C, 1)), device="cuda")
        o = eng.pose_clips(kpd[:, :5].contiguous(), qio.clone(), off, s.lb, s.ub, s.indiv_parts, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = eng.pose_clips(kpd, qio, off, s.lb, s.ub, s.indiv_parts, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        it = o["iters"].sum().item(); ls = o["ls_evals"].sum().item()
        print(f"path {path} mode {mode} C={C} F={F}: {ms:.1f} ms -> {C*F/ms*1e3:.0f} frames/s; iters/frame {it/(C*F):.1f} ls/iter {ls/it:.2f}; us/iter/chain {ms*1e3/(it/C):.3f}", flush=True)
    eng.set_path(0); eng.set_mode(-1)
