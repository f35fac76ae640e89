import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from random_trees import random_tree
from stac_mjx_b200.engine import Engine
from oracle.oracle import Oracle
for seed, nb, ns in ((21, 70, 20), (22, 140, 40), (23, 250, 60)):
    t, site_idxs, lb, ub = random_tree(seed, n_bodies=nb, p_welded=0.1, max_hinges=1, n_sites=ns)
    sb, off = t.site_bodyid[site_idxs], t.site_pos[site_idxs].astype(np.float32)
    K = len(sb)
    eng, o = Engine(t, sb, 0), Oracle(t, sb, np.float32, 2)
    rng = np.random.default_rng(seed)
    C, F = 100, 20
    q = (t.qpos0 + rng.normal(scale=0.2, size=(8, t.nq))).astype(np.float32)
    base = np.stack([o.fk(q[i], off)[3].reshape(-1) for i in range(8)]).astype(np.float32)
    kp = np.stack([base[(c + np.arange(F) // 5) % 8] for c in range(C)]) + 0.003
    kpd = torch.tensor(kp, device="cuda")
    kw = dict(do_root=1, root_kp_idx=0, trunk_kps=np.ones(K, bool), tol=1e-5, maxiter=200)
    for mode in (0, 4):
        eng.set_mode(mode)
        best = 1e30
        for rep in range(3):
            qio = torch.tensor(np.tile(t.qpos0.astype(np.float32), (C, 1)), device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = eng.pose_clips(kpd, qio, off, lb, ub, np.zeros((0, t.nq), bool), **kw); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        it = int(out["iters"].sum()); ls = int(out["ls_evals"].sum())
        print(f"nbody {nb} mode {mode}: {best:.1f} ms  iters {it} ls/iter {ls/it:.2f}", flush=True)
