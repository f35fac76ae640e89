"""Exploration probe run on a GPU box: kernel-vs-oracle diffs and a first timing."""
import sys, time
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
off = s.initial_offsets
F = 250
kp, qtrue, _ = synth.synth_session(t, s, 4 * F, F)
print("smem/chain", eng.smem_per_chain)

# FK
q = qtrue[:64].astype(np.float32)
go = [x.cpu().numpy() for x in eng.fk(q, off)]
d = [0, 0, 0, 0]
for i in range(64):
    ro = orc.fk(q[i], off)
    for j in range(4):
        d[j] = max(d[j], float(np.abs(go[j][i] - ro[j]).max()))
print("fk maxdiff qpos,xpos,xquat,sites:", d)

# loss/grad
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

# single solves
q0 = np.tile(t.qpos0.astype(np.float32), (8, 1)); q0[:, :3] = kp[:8, 3 * s.root_kp_idx:3 * s.root_kp_idx + 3]
rq = np.zeros(t.nq, bool); rq[:7] = True; tk = np.repeat(s.trunk_kps, 3)
p, e, it, ls = [x.cpu().numpy() for x in eng.q_opt(q0, kp[:8], rq, tk, off, s.lb, s.ub, 1e-4)]
for i in range(8):
    po, eo, io, lo = orc.q_opt(q0[i], s.lb, s.ub, rq, kp[i], tk, off, 1e-4)
    print(" solve", i, "iters", it[i], io, "ls", ls[i], lo, "err", e[i], eo, "dparams", np.abs(p[i] - po).max())

# clips
nclip = 4; Fs = 20
kpc = kp.reshape(nclip, F, -1)[:, :Fs].copy()
qio = torch.tensor(np.tile(t.qpos0.astype(np.float32), (nclip, 1)), device="cuda")
out = eng.pose_clips(kpc, qio, off, s.lb, s.ub, s.indiv_parts, do_root=True, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=float(cfg.model.FTOL))
torch.cuda.synchronize()
ref = orc.pose_clips(kpc, t.qpos0, off, s.lb, s.ub, s.indiv_parts, do_root=True, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=float(cfg.model.FTOL), nthreads=8)
for k in ("qpos", "xpos", "xquat", "sites", "err"):
    print(" clip", k, "maxdiff", float(np.abs(out[k].cpu().numpy() - ref[k]).max()))
print(" iters equal", np.array_equal(out["iters"].cpu().numpy(), ref["iters"]), "ls equal", np.array_equal(out["ls_evals"].cpu().numpy(), ref["ls_evals"]),
      "root", out["root_stats"].cpu().numpy()[0], ref["root_stats"][0], "status", out["status"].cpu().numpy())

# timing: 72 clips x 250
import os
eng.set_mode(int(os.environ.get('STACB_MODE', '-1')))
for C in (72, 148, 592):
    kpb, _, _ = synth.synth_session(t, s, C * F, F, seed=7)
    kpd = torch.tensor(kpb.reshape(C, F, -1), device="cuda")
    qio = torch.tensor(np.tile(t.qpos0.astype(np.float32), (C, 1)), device="cuda")
    o = eng.pose_clips(kpd[:, :5].contiguous(), qio.clone(), off, s.lb, s.ub, s.indiv_parts, do_root=True, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=float(cfg.model.FTOL))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = eng.pose_clips(kpd, qio, off, s.lb, s.ub, s.indiv_parts, do_root=True, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=float(cfg.model.FTOL))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it = o["iters"].sum().item(); ls = o["ls_evals"].sum().item()
    print(f"C={C} F={F}: {ms:.1f} ms -> {C*F/ms*1e3:.0f} frames/s; iters/frame {it/(C*F):.1f} ls/iter {ls/it:.2f}; us/iter/chain {ms*1e3/(it/C):.2f}")
print("fma peak TFLOP/s", eng.fma_peak_tflops())
