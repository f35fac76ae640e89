"""GPU solves against the FROZEN MJX-order restatement (oracle mode 0), float32 and float64 -- evidence that does not depend on
the bit-matched oracle modes that follow the kernels.

FISTA with the reference's FTOL on these ill-conditioned IK problems amplifies rounding: once two float32 implementations of the
same formulas differ in the last bit, line-search accepts flip and the iterates of weakly determined coordinates drift apart.
What can be asserted, and is asserted here on BASELINE config 1 (the real 250-frame rat23 clip) and on all 72 clips of
config 2:
  * GPU vs float32-MJX-order: the typical (median) frame meets north_star's tolerances (1e-3 rad, 1e-4 m, 1e-3 relative loss)
    with orders of magnitude to spare; the tails (p99, max) are reported and bounded;
  * the GPU result is no farther from the float64 MJX-order solution than the float32 MJX-order run itself is:
    spread(GPU, f64) <= 1.5 x spread(f32-mjx, f64) for the median and the 99th percentile (2 x for the max, a single-sample
    statistic) of |dqpos|, per-marker distance and per-frame relative loss.
Numbers land in gpurun_out/mjx_order_parity.json when that directory exists.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

QPOS_TOL, MARKER_TOL, REL_TOL = 1e-3, 1e-4, 1e-3


def three_stats(a):
    a = np.abs(np.asarray(a, np.float64)).ravel()
    return np.array([np.median(a), np.percentile(a, 99), a.max()])


def spreads(x, y, kp3):
    """median / p99 / max of |dqpos| [rad], per-marker distance [m], per-frame relative loss difference."""
    la, lb = ((x["sites"] - kp3) ** 2).sum((-1, -2)), ((y["sites"] - kp3) ** 2).sum((-1, -2))
    return {"qpos": three_stats(x["qpos"] - y["qpos"]), "marker": three_stats(np.linalg.norm(x["sites"] - y["sites"], axis=-1)),
            "loss_rel": three_stats((la - lb) / lb)}  # fmt: skip


def check_and_report(tag, gpu, m32, f64, kp3):
    vs_m32, vs_f64, floor = spreads(gpu, m32, kp3), spreads(gpu, f64, kp3), spreads(m32, f64, kp3)
    rec = {"case": tag}
    for name, d in (("gpu_vs_f32_mjx", vs_m32), ("gpu_vs_f64_mjx", vs_f64), ("f32_mjx_vs_f64_mjx", floor)):
        rec[name] = {k: dict(zip(("median", "p99", "max"), map(float, v))) for k, v in d.items()}
    print("\n[mjx-order parity]", json.dumps(rec))
    out = ROOT / "gpurun_out"
    if out.is_dir():
        with open(out / "mjx_order_parity.json", "a") as fh:
            fh.write(json.dumps(rec) + "\n")
    # typical frame: north_star tolerances against the float32 MJX-order run
    assert vs_m32["qpos"][0] <= QPOS_TOL and vs_m32["marker"][0] <= MARKER_TOL and vs_m32["loss_rel"][0] <= REL_TOL
    # tails against the float32 MJX-order run: bounded (chaotic amplification of rounding, see the module docstring)
    assert vs_m32["qpos"][1] <= 2e-2 and vs_m32["marker"][1] <= 2e-3 and vs_m32["loss_rel"][1] <= 2e-2
    # no farther from the float64 solution than the float32 MJX-order run is
    for k in ("qpos", "marker", "loss_rel"):
        for i, factor in ((0, 1.5), (1, 1.5), (2, 2.0)):
            assert vs_f64[k][i] <= factor * floor[k][i] + 1e-9, (k, i, vs_f64[k], floor[k])


def gpu_clips(case, eng, kp):
    s = case.setup
    C = kp.shape[0]
    qio = torch.tensor(np.tile(case.tree.qpos0.astype(np.float32), (C, 1)), device=eng.device)
    out = eng.pose_clips(kp, qio, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **case.root_kw())
    assert (out["status"] == 0).all()
    return {k: out[k].cpu().numpy() for k in ("qpos", "sites", "iters")}


def test_real_clip_against_mjx_order(rodent, engine_of):
    """BASELINE config 1: the reference's real rat23 recording, 250 frames; MJX-order runs are committed vectors."""
    g = np.load(ROOT / "tests" / "golden" / "rodent_real250.npz")
    gpu = gpu_clips(rodent, engine_of(rodent), g["kp"][None])
    m32 = {"qpos": g["m32_qpos"][None], "sites": g["m32_sites"][None]}
    f64 = {"qpos": g["f64_qpos"][None], "sites": g["f64_sites"][None]}
    check_and_report("config1 real rat23 clip, 250 frames", gpu, m32, f64, g["kp"].reshape(1, 250, -1, 3))


def test_all_72_clips_against_mjx_order(rodent, engine_of):
    """BASELINE config 2: every clip of the 18 000-frame synthetic session, MJX-order oracle run live on the host cores."""
    s = rodent.setup
    C, F = 72, 250
    kp, _, _ = rodent.session(C * F, F, seed=20260101)
    kp = kp.reshape(C, F, -1)
    gpu = gpu_clips(rodent, engine_of(rodent), kp)
    n = os.cpu_count() or 1
    runs = {}
    for tag, dt in (("m32", np.float32), ("f64", np.float64)):
        runs[tag] = rodent.oracle(dt, 0).pose_clips(kp, rodent.tree.qpos0, s.initial_offsets, s.lb, s.ub, s.indiv_parts, nthreads=n, **rodent.root_kw())
    check_and_report("config2 synthetic session, 72 clips x 250 frames", gpu, runs["m32"], runs["f64"], kp.reshape(C, F, -1, 3))
    # and bit-for-bit against the fast-order oracle on every clip (the kernel's own arithmetic)
    g32 = rodent.oracle(np.float32, 2).pose_clips(kp, rodent.tree.qpos0, s.initial_offsets, s.lb, s.ub, s.indiv_parts, nthreads=n, **rodent.root_kw())
    assert np.array_equal(gpu["qpos"], g32["qpos"]) and np.array_equal(gpu["sites"], g32["sites"]) and np.array_equal(gpu["iters"], g32["iters"])
