"""Generate tests/golden/*.npz: committed input/output vectors of the CPU oracle.

The reference (JAX + MJX + jaxopt) cannot be imported in the authoring image, so these vectors come
from the two restatements that can run here and that were validated against each other:
  - oracle/stac_oracle.c, mode 0 float64 (MJX operation order)  -> "f64_*" entries (ground truth values)
  - oracle/stac_oracle.c, mode 1 float32 (canonical order)       -> "c32_*" entries (bit-level expectation
    for the general CUDA kernels: every model with Engine.set_path(1), and the default path of the models the
    register-resident solver does not serve)
  - oracle/stac_oracle.c, mode 2 float32 (fast order)            -> "g32_*" entries, only for the models the
    register-resident solver serves (rodent, celegans, synth_data) and only for the quantities its arithmetic
    touches (loss / gradient / solves / clips): bit-level expectation for the default CUDA path
  - oracle/np_oracle.py (torch reverse-mode autodiff, float64)   -> "ad_*" entries (gradient, solver)
Run:  python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.np_oracle import TorchModel  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from stac_mjx_b200 import model, synth  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def make(name, n_eval=6, clip_frames=4, n_clips=2, seed=5):
    tree, cfg = model.load_fixture(name)
    kpn = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
    s = model.make_setup(tree, cfg.model, kpn)
    K = len(kpn)
    tol = float(cfg.model.FTOL)
    kp, qtrue, _ = synth.synth_session(tree, s, n_clips * clip_frames, clip_frames, seed=seed)
    off = s.initial_offsets
    rng = np.random.default_rng(seed)
    q = (qtrue[:n_eval] + rng.normal(scale=0.02, size=(n_eval, tree.nq))).astype(np.float32)
    q0 = (q + rng.normal(scale=0.01, size=q.shape)).astype(np.float32)
    f64, c32 = Oracle(tree, s.site_bodies, np.float64, 0), Oracle(tree, s.site_bodies, np.float32, 1)
    g32 = Oracle(tree, s.site_bodies, np.float32, 2)
    tags = (("f64", f64), ("c32", c32)) + ((("g32", g32),) if g32.fast_path else ())
    T = TorchModel(tree, s.site_bodies)
    qm_all, km_all = np.ones(tree.nq, bool), np.ones(3 * K, bool)
    qm_part = s.indiv_parts[0] if len(s.indiv_parts) else qm_all
    km_trunk = np.repeat(s.trunk_kps, 3) if s.trunk_kps.any() else km_all
    g = dict(q=q, q0=q0, kp=kp, offsets=off, qm_part=qm_part, km_trunk=km_trunk, tol=np.float32(tol))
    for tag, o in tags:
        if tag != "g32":  # FK outputs are the general cold path in every mode
            fk = [o.fk(qq, off) for qq in q]
            for i, nm in enumerate(("qpos", "xpos", "xquat", "sites")):
                g[f"{tag}_fk_{nm}"] = np.stack([r[i] for r in fk])
        lg = [o.loss_grad(q[i], q[i], qm_all, kp[i], km_all, off) for i in range(n_eval)]
        g[f"{tag}_loss"], g[f"{tag}_grad"] = np.array([r[0] for r in lg]), np.stack([r[1] for r in lg])
        lg = [o.loss_grad(q[i], q0[i], qm_part, kp[i], km_trunk, off) for i in range(n_eval)]
        g[f"{tag}_mloss"], g[f"{tag}_mgrad"] = np.array([r[0] for r in lg]), np.stack([r[1] for r in lg])
    ad = [T.loss_grad(q[i], q0[i], qm_part, kp[i], km_trunk, off) for i in range(2)]
    g["ad_mloss"], g["ad_mgrad"] = np.array([r[0] for r in ad]), np.stack([r[1] for r in ad])
    # single solves (canonical f32): root-style mask on frame 0 of each clip
    nroot = min(7, tree.nq)
    rq = np.zeros(tree.nq, bool)
    rq[:nroot] = True
    for tag, o in tags[1:]:
        sol = [o.q_opt(q0[i], s.lb, s.ub, rq, kp[i], km_trunk, off, tol, maxiter=50) for i in range(min(3, n_eval))]
        g[f"{tag}_sol_params"] = np.stack([r[0] for r in sol])
        g[f"{tag}_sol_err"] = np.array([r[1] for r in sol])
        g[f"{tag}_sol_iters"] = np.array([r[2] for r in sol])
        g[f"{tag}_sol_ls"] = np.array([r[3] for r in sol])
    g["root_mask"] = rq
    # clips
    has_root = s.root_kp_idx >= 0 and int(tree.jnt_type[0]) in (0, 2)
    kw = dict(do_root=1 if has_root else 0, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps, tol=tol)
    kpc = kp.reshape(n_clips, clip_frames, -1)
    for tag, o in tags:
        r = o.pose_clips(kpc, tree.qpos0, off, s.lb, s.ub, s.indiv_parts, nthreads=4, **kw)
        for k, v in r.items():
            g[f"{tag}_clip_{k}"] = v
    # m-phase
    for tag, o in (("f64", f64), ("c32", c32)):
        st = o.m_stats(kp[: n_clips * clip_frames], g[f"{tag}_clip_qpos"].reshape(-1, tree.nq))
        g[f"{tag}_m_s"], g[f"{tag}_m_z2"] = st
    old = np.load(OUT / f"{name}.npz") if (OUT / f"{name}.npz").exists() else {}
    drift = [k for k in old if k in g and not k.startswith("g32") and not np.array_equal(np.asarray(old[k]), np.asarray(g[k]))]
    np.savez_compressed(OUT / f"{name}.npz", **g)
    print(name, "written:", sum(v.nbytes for v in g.values()) // 1024, "KiB;",
          "clip qpos c32-vs-f64 max", np.abs(g["c32_clip_qpos"] - g["f64_clip_qpos"]).max(),
          "; g32-vs-f64", np.abs(g["g32_clip_qpos"] - g["f64_clip_qpos"]).max() if "g32_clip_qpos" in g else "-",
          "; entries that changed against the committed file:", drift or "none")


for nm in ("rodent", "celegans", "fly_treadmill", "synth_data"):
    make(nm)
make("mouse", n_eval=2, clip_frames=2, n_clips=1)
