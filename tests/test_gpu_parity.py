"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.

Tolerances are the ones BASELINE.json's north_star states: final qpos within 1e-3 rad, per-marker
position within 1e-4 m, per-frame loss / solver error within 1e-3 relative.  The kernels are written to
reproduce the oracle's canonical float32 arithmetic, so in practice the differences are zero; the tests
assert the stated tolerances and additionally report whether the match was bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import ROOT, get_case

pytestmark = pytest.mark.gpu

QPOS_TOL, MARKER_TOL, REL_TOL = 1e-3, 1e-4, 1e-3
MODELS = ["rodent", "celegans", "fly_treadmill", "synth_data", "mouse"]


def golden(name):
    return np.load(ROOT / "tests" / "golden" / f"{name}.npz")


def npy(t):
    return t.detach().cpu().numpy()


def E(g, key, path=0):
    """Bit-level expectation for the kernels that serve the model: "g32_*" (oracle mode 2, fast order) where the
    register-resident solver applies and the default path is used, "c32_*" (mode 1, canonical order) otherwise."""
    return g[f"g32_{key}"] if path == 0 and f"g32_{key}" in g.files else g[f"c32_{key}"]


@pytest.mark.parametrize("name", MODELS)
def test_fk_matches_golden(name, engine_of):
    c, g = get_case(name), golden(name)
    eng = engine_of(c)
    qo, xp, xq, sx = [npy(t) for t in eng.fk(g["q"], g["offsets"])]
    np.testing.assert_allclose(xp, g["c32_fk_xpos"], atol=MARKER_TOL, rtol=0)
    np.testing.assert_allclose(sx, g["c32_fk_sites"], atol=MARKER_TOL, rtol=0)
    np.testing.assert_allclose(xq, g["c32_fk_xquat"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(qo, g["c32_fk_qpos"], atol=1e-6, rtol=0)
    # and against the float64 MJX-order values (the mathematics, not the rounding)
    np.testing.assert_allclose(xp, g["f64_fk_xpos"], atol=5e-6 * max(1.0, np.abs(g["f64_fk_xpos"]).max()))
    assert np.array_equal(xp, g["c32_fk_xpos"]) and np.array_equal(xq, g["c32_fk_xquat"]), "FK no longer bit-identical to the canonical oracle"


@pytest.mark.parametrize("name", MODELS)
def test_loss_and_gradient_match_golden(name, engine_of):
    c, g = get_case(name), golden(name)
    eng = engine_of(c)
    qm, km = np.ones(c.tree.nq, bool), np.ones(3 * c.K, bool)
    L, G = [npy(t) for t in eng.loss_grad(g["q"], g["q"], g["kp"][: len(g["q"])], qm, km, g["offsets"])]
    np.testing.assert_allclose(L, E(g, "loss"), rtol=REL_TOL)
    np.testing.assert_allclose(G, E(g, "grad"), atol=1e-5 * max(1.0, np.abs(E(g, "grad")).max()))
    np.testing.assert_allclose(G, g["f64_grad"], atol=1e-4 * max(1.0, np.abs(g["f64_grad"]).max()))
    L2, G2 = [npy(t) for t in eng.loss_grad(g["q"], g["q0"], g["kp"][: len(g["q"])], g["qm_part"], g["km_trunk"], g["offsets"])]
    np.testing.assert_allclose(L2, E(g, "mloss"), rtol=REL_TOL)
    np.testing.assert_allclose(G2, E(g, "mgrad"), atol=1e-5 * max(1.0, np.abs(E(g, "mgrad")).max()))
    assert (G2[:, ~g["qm_part"].astype(bool)] == 0).all()  # masked-out coordinates have zero gradient
    assert np.array_equal(G, E(g, "grad")) and np.array_equal(L2, E(g, "mloss")), "loss/grad no longer bit-identical"
    Lonly, none = eng.loss_grad(g["q"], g["q"], g["kp"][: len(g["q"])], qm, km, g["offsets"], want_grad=False)
    assert none is None and np.array_equal(npy(Lonly), L)


@pytest.mark.parametrize("name", MODELS)
def test_single_solves_match_golden(name, engine_of):
    c, g = get_case(name), golden(name)
    eng = engine_of(c)
    n = len(g["c32_sol_iters"])
    p, e, it, ls = [npy(t) for t in eng.q_opt(g["q0"][:n], g["kp"][:n], g["root_mask"], g["km_trunk"], g["offsets"], c.setup.lb, c.setup.ub,
                                               float(g["tol"]), maxiter=50)]  # fmt: skip
    np.testing.assert_array_equal(it, E(g, "sol_iters"))
    np.testing.assert_array_equal(ls, E(g, "sol_ls"))
    np.testing.assert_allclose(p, E(g, "sol_params"), atol=QPOS_TOL, rtol=0)
    np.testing.assert_allclose(e, E(g, "sol_err"), rtol=REL_TOL)
    assert np.array_equal(p, E(g, "sol_params"))


@pytest.mark.parametrize("name", MODELS)
def test_clips_match_golden(name, engine_of):
    """root_optimization + pose_optimization over clips, committed oracle output (2 clips x 4 frames)."""
    c, g = get_case(name), golden(name)
    eng = engine_of(c)
    C, F = g["c32_clip_qpos"].shape[:2]
    qio = torch.tensor(np.tile(c.tree.qpos0.astype(np.float32), (C, 1)), device=eng.device)
    out = eng.pose_clips(g["kp"].reshape(C, F, -1), qio, g["offsets"], c.setup.lb, c.setup.ub, c.setup.indiv_parts, **c.root_kw())
    o = {k: npy(v) for k, v in out.items()}
    assert (o["status"] == 0).all()
    np.testing.assert_array_equal(o["iters"], E(g, "clip_iters"))
    np.testing.assert_array_equal(o["ls_evals"], E(g, "clip_ls_evals"))
    np.testing.assert_allclose(o["qpos"], E(g, "clip_qpos"), atol=QPOS_TOL, rtol=0)
    np.testing.assert_allclose(o["sites"], E(g, "clip_sites"), atol=MARKER_TOL, rtol=0)
    np.testing.assert_allclose(o["xpos"], E(g, "clip_xpos"), atol=MARKER_TOL, rtol=0)
    np.testing.assert_allclose(o["err"], E(g, "clip_err"), rtol=REL_TOL)
    np.testing.assert_array_equal(npy(qio), E(g, "clip_qpos")[:, -1])  # qpos_io carries the last frame's pose
    for k in ("qpos", "xpos", "xquat", "sites", "err"):
        assert np.array_equal(o[k], E(g, f"clip_{k}")), f"{k} no longer bit-identical to the canonical oracle"


@pytest.mark.parametrize("name", ["rodent", "celegans", "synth_data", "fly_treadmill", "mouse"])
def test_general_kernels_still_serve_the_fast_models(name, engine_of):
    """Engine.set_path(1): the general kernels (canonical order, oracle mode 1) on the models the register-resident solver
    serves by default -- loss / gradient, single solves and clips against the committed c32 vectors, bit for bit."""
    c, g = get_case(name), golden(name)
    eng = engine_of(c)
    assert eng.path == 1
    try:
        eng.set_path(1)
        assert eng.path == 0
        qm, km = np.ones(c.tree.nq, bool), np.ones(3 * c.K, bool)
        L, G = [npy(t) for t in eng.loss_grad(g["q"], g["q"], g["kp"][: len(g["q"])], qm, km, g["offsets"])]
        assert np.array_equal(L, g["c32_loss"]) and np.array_equal(G, g["c32_grad"])
        n = len(g["c32_sol_iters"])
        p, e, it, ls = [npy(t) for t in eng.q_opt(g["q0"][:n], g["kp"][:n], g["root_mask"], g["km_trunk"], g["offsets"], c.setup.lb, c.setup.ub,
                                                   float(g["tol"]), maxiter=50)]  # fmt: skip
        assert np.array_equal(p, g["c32_sol_params"]) and np.array_equal(it, g["c32_sol_iters"])
        C, F = g["c32_clip_qpos"].shape[:2]
        for mode in (0, 1):
            eng.set_mode(mode)
            qio = torch.tensor(np.tile(c.tree.qpos0.astype(np.float32), (C, 1)), device=eng.device)
            out = eng.pose_clips(g["kp"].reshape(C, F, -1), qio, g["offsets"], c.setup.lb, c.setup.ub, c.setup.indiv_parts, **c.root_kw())
            for k in ("qpos", "xpos", "xquat", "sites", "err", "iters", "ls_evals"):
                assert np.array_equal(npy(out[k]), g[f"c32_clip_{k}"]), k
    finally:
        eng.set_path(0)
        eng.set_mode(-1)
    # the two paths agree to rounding on a single evaluation
    np.testing.assert_allclose(g["g32_loss"], g["c32_loss"], rtol=1e-5)
    np.testing.assert_allclose(g["g32_grad"], g["c32_grad"], atol=1e-5 * max(1.0, np.abs(g["c32_grad"]).max()))


@pytest.mark.parametrize("name", ["rodent", "celegans", "fly_treadmill", "mouse"])
def test_latency_and_throughput_modes_are_bit_identical(name, engine_of):
    """Throughput (0), latency (1: four cooperating warps, speculative line search), dense throughput (2) and grouped
    latency (3: three member warps per evaluation for wide trees; falls back to 1 for one-body-per-lane models) modes:
    same bits, same counters."""
    c = get_case(name)
    eng = engine_of(c)
    F = 3 if name == "mouse" else 6
    kp, _, _ = c.session(5 * F, F, seed=41)
    kp = kp.reshape(5, F, -1)
    outs = []
    try:
        for mode in (0, 1, 2, 3) + ((4,) if eng.path == 1 else ()):  # 4 = pair mode of the register-resident path
            eng.set_mode(mode)
            qio = torch.tensor(np.tile(c.tree.qpos0.astype(np.float32), (5, 1)), device=eng.device)
            o = eng.pose_clips(kp, qio, c.setup.initial_offsets, c.setup.lb, c.setup.ub, c.setup.indiv_parts, **c.root_kw())
            outs.append({k: npy(v) for k, v in o.items()} | {"qio": npy(qio)})
    finally:
        eng.set_mode(-1)
    for k in outs[0]:
        assert all(np.array_equal(outs[0][k], o[k]) for o in outs[1:]), k
    ref = c.oracle(np.float32, 2).pose_clips(kp, c.tree.qpos0, c.setup.initial_offsets, c.setup.lb, c.setup.ub, c.setup.indiv_parts,
                                              nthreads=4, **c.root_kw())  # fmt: skip
    np.testing.assert_allclose(outs[1]["qpos"], ref["qpos"], atol=QPOS_TOL, rtol=0)
    np.testing.assert_array_equal(outs[1]["iters"], ref["iters"])
    np.testing.assert_array_equal(outs[1]["ls_evals"], ref["ls_evals"])


def test_rodent_clip_against_live_oracle(rodent, engine_of):
    """Fresh seeded inputs (not the committed ones): 3 clips x 12 frames of the rodent, all 6 solves per frame."""
    eng = engine_of(rodent)
    s = rodent.setup
    kp, _, _ = rodent.session(36, 12, seed=123)
    kp = kp.reshape(3, 12, -1)
    qio = torch.tensor(np.tile(rodent.tree.qpos0.astype(np.float32), (3, 1)), device=eng.device)
    out = eng.pose_clips(kp, qio, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **rodent.root_kw())
    ref = rodent.oracle(np.float32, 2).pose_clips(kp, rodent.tree.qpos0, s.initial_offsets, s.lb, s.ub, s.indiv_parts, nthreads=4, **rodent.root_kw())
    np.testing.assert_array_equal(npy(out["iters"]), ref["iters"])
    np.testing.assert_array_equal(npy(out["root_stats"]), ref["root_stats"])
    np.testing.assert_allclose(npy(out["qpos"]), ref["qpos"], atol=QPOS_TOL, rtol=0)
    np.testing.assert_allclose(npy(out["sites"]), ref["sites"], atol=MARKER_TOL, rtol=0)
    # per-frame loss at the solution, recomputed by the oracle from the GPU's qpos
    res_gpu = ((npy(out["sites"]) - kp.reshape(3, 12, -1, 3)) ** 2).sum((-1, -2))
    res_ref = ((ref["sites"] - kp.reshape(3, 12, -1, 3)) ** 2).sum((-1, -2))
    np.testing.assert_allclose(res_gpu, res_ref, rtol=REL_TOL)


def test_pose_without_root_and_warm_start_chain(rodent, engine_of):
    """fit_offsets-style use: do_root=0, second pass warm-started from qpos_io of the first."""
    eng = engine_of(rodent)
    s, o = rodent.setup, rodent.oracle(np.float32, 2)
    kp, _, _ = rodent.session(5, 5, seed=9)
    kp = kp.reshape(1, 5, -1)
    qio = torch.tensor(rodent.tree.qpos0.astype(np.float32)[None], device=eng.device)
    kw = dict(do_root=0, tol=rodent.tol)
    a = eng.pose_clips(kp, qio, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **kw)
    a_q = npy(a["qpos"]).copy()
    b = eng.pose_clips(kp, qio, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **kw)  # continues from pass 1
    r1 = o.pose_clips(kp, rodent.tree.qpos0, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **kw)
    r2 = o.pose_clips(kp, r1["qpos"][:, -1], s.initial_offsets, s.lb, s.ub, s.indiv_parts, **kw)
    np.testing.assert_allclose(a_q, r1["qpos"], atol=QPOS_TOL, rtol=0)
    np.testing.assert_allclose(npy(b["qpos"]), r2["qpos"], atol=QPOS_TOL, rtol=0)
    np.testing.assert_array_equal(npy(b["iters"]), r2["iters"])


def test_passive_coordinates_outside_their_box(rodent, engine_of):
    """Hinges outside the active subtree have zero gradient: the first whole-body solve projects them into the box and
    their move enters that solve's first line search (register-resident path: kept outside the solver slots)."""
    eng = engine_of(rodent)
    s = rodent.setup
    kp, _, _ = rodent.session(8, 4, seed=77)
    kp = kp.reshape(2, 4, -1)
    q0 = np.tile(rodent.tree.qpos0.astype(np.float32), (2, 1))
    q0[:, -1] += 10.0
    q0[1, -5] -= 7.0
    ref = rodent.oracle(np.float32, 2).pose_clips(kp, q0, s.initial_offsets, s.lb, s.ub, s.indiv_parts, nthreads=2, **rodent.root_kw())
    try:
        for mode in (0, 1, 3, 4):
            eng.set_mode(mode)
            qio = torch.tensor(q0, device=eng.device)
            out = eng.pose_clips(kp, qio, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **rodent.root_kw())
            for k in ("qpos", "sites", "err", "iters", "ls_evals", "root_stats"):
                assert np.array_equal(npy(out[k]), ref[k]), (mode, k)
            assert (npy(out["qpos"])[..., -1] == s.ub[-1]).all()
    finally:
        eng.set_mode(-1)
    p, e, it, ls = [npy(t) for t in eng.q_opt(q0, kp[:, 0], np.ones(rodent.tree.nq, bool), np.ones(3 * rodent.K, bool), s.initial_offsets,
                                               s.lb, s.ub, 1e-4, maxiter=7)]  # fmt: skip
    for i in range(2):
        po, eo, ito, lso = rodent.oracle(np.float32, 2).q_opt(q0[i], s.lb, s.ub, np.ones(rodent.tree.nq, bool), kp[i, 0], np.ones(3 * rodent.K, bool),
                                                               s.initial_offsets, 1e-4, maxiter=7)  # fmt: skip
        assert np.array_equal(p[i], po) and (it[i], ls[i]) == (ito, lso)


def test_edge_cases(rodent, engine_of):
    eng = engine_of(rodent)
    s = rodent.setup
    kp, _, _ = rodent.session(4, 1, seed=2)
    # F = 1 clips, P = 0 (no part solves), root only (do_root=2)
    qio = torch.tensor(np.tile(rodent.tree.qpos0.astype(np.float32), (4, 1)), device=eng.device)
    out = eng.pose_clips(kp.reshape(4, 1, -1), qio, s.initial_offsets, s.lb, s.ub, np.zeros((0, rodent.tree.nq), bool), **rodent.root_kw())
    ref = rodent.oracle(np.float32, 2).pose_clips(kp.reshape(4, 1, -1), rodent.tree.qpos0, s.initial_offsets, s.lb, s.ub, [], **rodent.root_kw())
    assert out["iters"].shape == (4, 1, 1)
    np.testing.assert_allclose(npy(out["qpos"]), ref["qpos"], atol=QPOS_TOL, rtol=0)
    q2 = torch.tensor(np.tile(rodent.tree.qpos0.astype(np.float32), (4, 1)), device=eng.device)
    eng.pose_clips(kp.reshape(4, 1, -1), q2, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **{**rodent.root_kw(), "do_root": 2})
    assert not torch.equal(q2[:, :7], torch.tensor(rodent.tree.qpos0[:7].astype(np.float32), device=eng.device).expand(4, 7))
    assert torch.equal(q2[:, 7:], torch.zeros_like(q2[:, 7:]))  # only the root DOFs moved
    # C = 0 is a no-op; bad shapes raise
    empty = eng.pose_clips(np.zeros((0, 3, 3 * rodent.K), np.float32), torch.zeros(0, rodent.tree.nq, device=eng.device), s.initial_offsets,
                           s.lb, s.ub, s.indiv_parts, do_root=0)  # fmt: skip
    assert empty["qpos"].shape[0] == 0
    with pytest.raises(ValueError):
        eng.pose_clips(np.zeros((1, 3, 5), np.float32), qio[:1], s.initial_offsets, s.lb, s.ub, s.indiv_parts, do_root=0)
    # a reused output dict is written in place: buffers of another launch shape are refused, matching ones are reused
    with pytest.raises(ValueError, match="out\\['qpos'\\]"):
        eng.pose_clips(kp.reshape(4, 1, -1)[:2], qio[:2].clone(), s.initial_offsets, s.lb, s.ub, np.zeros((0, rodent.tree.nq), bool), out=out, **rodent.root_kw())
    ptr = out["qpos"].data_ptr()
    again = eng.pose_clips(kp.reshape(4, 1, -1), torch.tensor(np.tile(rodent.tree.qpos0.astype(np.float32), (4, 1)), device=eng.device),
                           s.initial_offsets, s.lb, s.ub, np.zeros((0, rodent.tree.nq), bool), out=out, **rodent.root_kw())  # fmt: skip
    assert again["qpos"].data_ptr() == ptr
    np.testing.assert_allclose(npy(again["qpos"]), ref["qpos"], atol=QPOS_TOL, rtol=0)
    # maxiter = 1: exactly one FISTA iteration per solve
    one = eng.pose_clips(kp.reshape(4, 1, -1), qio.clone(), s.initial_offsets, s.lb, s.ub, s.indiv_parts, do_root=0, maxiter=1)
    assert (npy(one["iters"]) == 1).all()
    # non-finite keypoints poison the loss (reference has no NaN guard) and are flagged
    bad = kp.reshape(4, 1, -1).copy()
    bad[2, 0, 5] = np.nan
    st = eng.pose_clips(bad, qio.clone(), s.initial_offsets, s.lb, s.ub, s.indiv_parts, do_root=0, maxiter=3)["status"]
    assert npy(st).tolist() == [0, 0, 1, 0]


def test_kat_m_opt_on_gpu():
    """The known-answer tests of reference tests/unit/test_m_opt.py:72-225 through StacCore.m_opt on the GPU."""
    from test_oracle_cpu import GT_A, GT_B, chain_tree
    from stac_mjx_b200 import stac_core
    from stac_mjx_b200.engine import Engine

    t = chain_tree()
    eng = Engine(t, t.site_bodyid, 0)
    core = stac_core.StacCore()
    Z, ONE = np.zeros((3, 3), np.float32), np.ones((3, 3), np.float32)

    def gen(q, off):
        return eng.fk(q, off)[3].reshape(len(q), -1)

    def m_opt(kp, q, m0, reg_mask, coef):
        mdl = stac_core.StacModel(engine=eng, site_pos=eng.f32(m0))
        r = core.m_opt(mdl, None, kp, q, m0, reg_mask, coef, None)
        return npy(r.params), float(r.error)

    q = np.zeros((5, 3), np.float32)
    p, err = m_opt(gen(q, GT_A), q, Z, Z, 0.0)
    np.testing.assert_allclose(p, GT_A, atol=1e-5)
    assert err < 1e-8  # reference tests/unit/test_m_opt.py:89; the objective is evaluated from the residuals at m*
    q = np.random.RandomState(42).randn(10, 3).astype(np.float32) * 0.5
    np.testing.assert_allclose(m_opt(gen(q, GT_A), q, Z, Z, 0.0)[0], GT_A, atol=1e-5)
    q = np.zeros((8, 3), np.float32)
    q[:, 0] = np.linspace(0.0, np.pi / 4, 8)
    np.testing.assert_allclose(m_opt(gen(q, GT_B), q, GT_B, Z, 0.0)[0], GT_B, atol=1e-5)
    q = np.random.RandomState(99).randn(15, 3).astype(np.float32) * 1.5
    np.testing.assert_allclose(m_opt(gen(q, GT_B), q, Z, Z, 0.0)[0], GT_B, atol=1e-4)
    q = np.random.RandomState(42).randn(10, 3).astype(np.float32) * 0.3
    kp = gen(q, GT_A)
    np.testing.assert_allclose(m_opt(kp, q, ONE * 99.0, ONE, 0.0)[0], GT_A, atol=1e-5)
    np.testing.assert_allclose(m_opt(kp, q, Z, ONE, 1e6)[0], Z, atol=1e-3)
    q = np.zeros((10, 3), np.float32)
    gt = np.full((3, 3), 0.5, np.float32)
    is_reg = Z.copy()
    is_reg[0] = 1.0
    strong, noreg = m_opt(gen(q, gt), q, Z, is_reg, 1e4)[0], m_opt(gen(q, gt), q, Z, is_reg, 0.0)[0]
    assert np.linalg.norm(strong[0]) < np.linalg.norm(noreg[0])
    np.testing.assert_allclose(strong[1:], gt[1:], atol=1e-5)


@pytest.mark.parametrize("name", ["rodent", "mouse"])
def test_m_stats_match_oracle(name, engine_of):
    c, g = get_case(name), golden(name)
    eng = engine_of(c)
    q = g["c32_clip_qpos"].reshape(-1, c.tree.nq)
    s, z2 = eng.m_stats(g["kp"][: len(q)], q)
    np.testing.assert_allclose(npy(s), g["c32_m_s"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(npy(z2)[0], g["c32_m_z2"], rtol=1e-5)
    np.testing.assert_allclose(npy(s), g["f64_m_s"], rtol=0, atol=5e-3 * np.abs(g["f64_m_s"]).max())  # f64 run has its own qpos
    assert np.array_equal(npy(s), g["c32_m_s"])


def test_full_size_properties(rodent, engine_of):
    """BASELINE config 2 shape (72 clips x 250 frames): size-independent properties + determinism."""
    eng = engine_of(rodent)
    s = rodent.setup
    C, F = 72, 250
    kp, _, _ = rodent.session(C * F, F, seed=20260101)
    kpd = torch.tensor(kp.reshape(C, F, -1), device=eng.device)
    q0 = torch.tensor(np.tile(rodent.tree.qpos0.astype(np.float32), (C, 1)), device=eng.device)
    runs = []
    for _ in range(2):
        qio = q0.clone()
        out = eng.pose_clips(kpd, qio, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **rodent.root_kw())
        runs.append({k: npy(v) for k, v in out.items()})
    a, b = runs
    for k in a:
        assert np.array_equal(a[k], b[k]), f"{k} differs between two identical launches"
    assert (a["status"] == 0).all() and np.isfinite(a["qpos"]).all()
    np.testing.assert_allclose(np.linalg.norm(a["qpos"][..., 3:7], axis=-1), 1.0, atol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(a["xquat"], axis=-1), 1.0, atol=1e-4)
    lb, ub = s.lb[7:], s.ub[7:]
    assert (a["qpos"][..., 7:] >= lb).all() and (a["qpos"][..., 7:] <= ub).all()  # box constraints hold exactly
    assert (a["iters"] >= 1).all() and (a["iters"] <= 400).all() and (a["ls_evals"] >= a["iters"]).all()
    resid = np.linalg.norm(a["sites"] - kp.reshape(C, F, -1, 3), axis=-1)
    assert resid.mean() < 4e-3  # 1 mm observation noise + 2 mm offset perturbation in the synthetic session
    conv = a["iters"][..., -1] < 400
    assert (a["err"][conv] <= rodent.tol).all()
    # oracle spot check on one full-length clip at full size
    ci = 17
    ref = rodent.oracle(np.float32, 2).pose_clips(kp.reshape(C, F, -1)[ci : ci + 1], rodent.tree.qpos0, s.initial_offsets, s.lb, s.ub,
                                                  s.indiv_parts, **rodent.root_kw())  # fmt: skip
    np.testing.assert_allclose(a["qpos"][ci], ref["qpos"][0], atol=QPOS_TOL, rtol=0)
    np.testing.assert_allclose(a["sites"][ci], ref["sites"][0], atol=MARKER_TOL, rtol=0)
    np.testing.assert_array_equal(a["iters"][ci], ref["iters"][0])


def test_real_mocap_clip_matches_golden(rodent, engine_of):
    """BASELINE config 1: 250 frames of the reference's real rat23 recording, root optimisation + 6 solves per frame."""
    g = golden("rodent_real250")
    eng = engine_of(rodent)
    s = rodent.setup
    qio = torch.tensor(rodent.tree.qpos0.astype(np.float32)[None], device=eng.device)
    out = eng.pose_clips(g["kp"][None], qio, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **rodent.root_kw())
    np.testing.assert_array_equal(npy(out["iters"])[0], E(g, "iters"))
    np.testing.assert_array_equal(npy(out["root_stats"])[0], E(g, "root_stats"))
    np.testing.assert_allclose(npy(out["qpos"])[0], E(g, "qpos"), atol=QPOS_TOL, rtol=0)
    np.testing.assert_allclose(npy(out["sites"])[0], E(g, "sites"), atol=MARKER_TOL, rtol=0)
    np.testing.assert_allclose(npy(out["err"])[0], E(g, "err"), rtol=REL_TOL)
    assert np.array_equal(npy(out["qpos"])[0], E(g, "qpos"))


def test_all_joint_types_on_gpu():
    """The rare-joint paths of the kernels (ball, slide, a second free joint, 3 joints on one body) against the oracle."""
    from mixed_model import mixed_tree, random_qpos
    from oracle.oracle import Oracle
    from stac_mjx_b200.engine import Engine

    t, site_idxs, lb, ub = mixed_tree()
    sb = t.site_bodyid[site_idxs]
    off = t.site_pos[site_idxs].astype(np.float32)
    K = len(sb)
    eng = Engine(t, sb, 0)
    o = Oracle(t, sb, np.float32, 2)
    assert eng.path == 0 and not o.fast_path  # ball / slide joints: general kernels
    rng = np.random.default_rng(1)
    q = random_qpos(t, rng, 6).astype(np.float32)
    got = [npy(x) for x in eng.fk(q, off)]
    kp = np.stack([o.fk(q[(i + 1) % 6], off)[3].reshape(-1) for i in range(6)]).astype(np.float32) + 0.005
    for i in range(6):
        ref = o.fk(q[i], off)
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(a[i], b)
    qm, km = np.ones(t.nq, bool), np.ones(3 * K, bool)
    L, G = [npy(x) for x in eng.loss_grad(q, q, kp, qm, km, off)]
    for i in range(6):
        l, g = o.loss_grad(q[i], q[i], qm, kp[i], km, off)
        assert float(l) == float(L[i])
        np.testing.assert_array_equal(G[i], g)
    # masked solve (only the arm / wrist DOFs) and a full clip with root optimisation, all three scheduling modes
    part = np.zeros(t.nq, bool)
    part[7:16] = True
    p, e, it, ls = [npy(x) for x in eng.q_opt(q[:3], kp[:3], part, km, off, lb, ub, 1e-5, maxiter=80)]
    for i in range(3):
        po, eo, ito, lso = o.q_opt(q[i], lb, ub, part, kp[i], km, off, 1e-5, maxiter=80)
        assert (it[i], ls[i]) == (ito, lso)
        np.testing.assert_array_equal(p[i], po)
    kpc = kp.reshape(2, 3, -1)
    kw = dict(do_root=1, root_kp_idx=0, trunk_kps=np.ones(K, bool), tol=1e-5, maxiter=60)
    ref = o.pose_clips(kpc, t.qpos0, off, lb, ub, part[None], **kw)
    try:
        for mode in (0, 1, 2):
            eng.set_mode(mode)
            qio = torch.tensor(np.tile(t.qpos0.astype(np.float32), (2, 1)), device=eng.device)
            out = eng.pose_clips(kpc, qio, off, lb, ub, part[None], **kw)
            np.testing.assert_array_equal(npy(out["iters"]), ref["iters"])
            for k in ("qpos", "xpos", "xquat", "sites", "err"):
                np.testing.assert_array_equal(npy(out[k]), ref[k])
    finally:
        eng.set_mode(-1)
    np.testing.assert_allclose(np.linalg.norm(ref["qpos"][..., 3:7], axis=-1), 1.0, atol=1e-6)


def test_slide_root_model_uses_four_root_dofs():
    """A model whose first joint is a slide (the reference's root_dims = 4 branch, compute_stac.py:51-54)."""
    from oracle.oracle import Oracle
    from stac_mjx_b200 import mjcf, tree
    from stac_mjx_b200.engine import Engine

    xml = """<mujoco><compiler angle="radian"/><worldbody>
      <body name="cart" pos="0 0 0.1">
        <joint name="sx" type="slide" axis="1 0 0"/><joint name="sy" type="slide" axis="0 1 0"/><joint name="yaw" type="hinge" axis="0 0 1"/>
        <site name="k0" pos="0.02 0 0.01"/>
        <body name="pole" pos="0 0 0.05"><joint name="tilt" type="hinge" axis="0 1 0" range="-1 1"/><site name="k1" pos="0 0.01 0.2"/>
          <body name="tip" pos="0 0 0.3"><joint name="bend" type="hinge" axis="1 0 0" range="-1 1"/><site name="k2" pos="0.01 0 0.1"/></body>
        </body>
      </body></worldbody></mujoco>"""
    t = tree.compile_spec(mjcf.parse_mjcf(xml, from_string=True))
    sidx = np.array([t.site_id(n) for n in ("k0", "k1", "k2")])
    sb, off = t.site_bodyid[sidx], t.site_pos[sidx].astype(np.float32)
    lb, ub, _ = tree.align_joint_dims(t.jnt_type, t.jnt_range, t.jnt_names)
    assert int(t.jnt_type[0]) == 2 and t.nq == 5
    eng, o = Engine(t, sb, 0), Oracle(t, sb, np.float32, 2)
    rng = np.random.default_rng(4)
    qt = rng.normal(scale=0.3, size=(2, 4, t.nq)).astype(np.float32)
    kp = np.stack([[o.fk(qt[c, f], off)[3].reshape(-1) for f in range(4)] for c in range(2)]).astype(np.float32)
    kw = dict(do_root=1, root_kp_idx=0, trunk_kps=np.array([1, 1, 0], bool), root_dims=4, tol=1e-6, maxiter=100)
    qio = torch.zeros(2, t.nq, device=eng.device)
    out = eng.pose_clips(kp, qio, off, lb, ub, np.zeros((0, t.nq), bool), **kw)
    ref = o.pose_clips(kp, np.zeros(t.nq), off, lb, ub, [], **kw)
    np.testing.assert_array_equal(npy(out["root_stats"]), ref["root_stats"])
    for k in ("qpos", "sites", "err", "iters"):
        np.testing.assert_array_equal(npy(out[k]), ref[k])
    assert np.isfinite(npy(out["qpos"])).all()


def test_c_abi_without_torch(rodent):
    """The C ABI driven with cuda-python device memory and ctypes only (no torch types cross the boundary)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("abi_demo_no_torch", ROOT / "tools" / "abi_demo_no_torch.py")
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    r = demo.run(n_clips=2, n_frames=3, seed=3)
    s = rodent.setup
    ref = rodent.oracle(np.float32, 2).pose_clips(r["kp"], rodent.tree.qpos0, s.initial_offsets, s.lb, s.ub, s.indiv_parts, nthreads=2,
                                                  **rodent.root_kw())  # fmt: skip
    np.testing.assert_array_equal(r["iters"], ref["iters"])
    np.testing.assert_allclose(r["qpos"], ref["qpos"], atol=QPOS_TOL, rtol=0)
    np.testing.assert_allclose(r["sites"], ref["sites"], atol=MARKER_TOL, rtol=0)


def test_fly_tethered_config_against_live_oracle(engine_of):
    """BASELINE config 4's second fruitfly configuration (30 keypoints, no root optimisation key): 2 clips x 5 frames."""
    c = get_case("fly_tethered")
    eng = engine_of(c)
    kp, _, _ = c.session(10, 5, seed=21)
    kp = kp.reshape(2, 5, -1)
    qio = torch.tensor(np.tile(c.tree.qpos0.astype(np.float32), (2, 1)), device=eng.device)
    out = eng.pose_clips(kp, qio, c.setup.initial_offsets, c.setup.lb, c.setup.ub, c.setup.indiv_parts, **c.root_kw())
    ref = c.oracle(np.float32, 2).pose_clips(kp, c.tree.qpos0, c.setup.initial_offsets, c.setup.lb, c.setup.ub, c.setup.indiv_parts,
                                              nthreads=2, **c.root_kw())  # fmt: skip
    np.testing.assert_array_equal(npy(out["iters"]), ref["iters"])
    np.testing.assert_allclose(npy(out["qpos"]), ref["qpos"], atol=QPOS_TOL, rtol=0)
    np.testing.assert_allclose(npy(out["sites"]), ref["sites"], atol=MARKER_TOL, rtol=0)
    assert np.array_equal(npy(out["qpos"]), ref["qpos"])


def test_concurrent_launches_on_two_streams(rodent, engine_of):
    """Entry points are re-entrant: two launches of the same tree handle overlapping on different streams."""
    eng = engine_of(rodent)
    s = rodent.setup
    kp, _, _ = rodent.session(2 * 20 * 6, 6, seed=55)
    kp = kp.reshape(2, 20, 6, -1)
    base = []
    for i in range(2):
        qio = torch.tensor(np.tile(rodent.tree.qpos0.astype(np.float32), (20, 1)), device=eng.device)
        base.append(npy(eng.pose_clips(kp[i], qio, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **rodent.root_kw())["qpos"]).copy())
    streams = [torch.cuda.Stream(device=eng.device) for _ in range(2)]
    kpd = [torch.tensor(kp[i], device=eng.device) for i in range(2)]
    qios = [torch.tensor(np.tile(rodent.tree.qpos0.astype(np.float32), (20, 1)), device=eng.device) for _ in range(2)]
    torch.cuda.synchronize()
    outs = []
    for i in range(2):
        with torch.cuda.stream(streams[i]):
            outs.append(eng.pose_clips(kpd[i], qios[i], s.initial_offsets, s.lb, s.ub, s.indiv_parts, **rodent.root_kw()))
    torch.cuda.synchronize()
    for i in range(2):
        np.testing.assert_array_equal(npy(outs[i]["qpos"]), base[i])


def test_large_batches_use_the_multi_warp_cta_path(rodent, engine_of):
    """B > 2 x SM count: four independent items per CTA in the batch kernel (fk / loss_grad / q_opt / m_stats)."""
    eng = engine_of(rodent)
    s, o = rodent.setup, rodent.oracle(np.float32, 2)
    B = 700
    kp, qtrue, _ = rodent.session(B, B, seed=66)
    q = (qtrue + np.random.default_rng(0).normal(scale=0.02, size=qtrue.shape)).astype(np.float32)
    qo, xp, xq, sx = [npy(t) for t in eng.fk(q, s.initial_offsets)]
    qm, km = np.ones(rodent.tree.nq, bool), np.ones(3 * rodent.K, bool)
    L, G = [npy(t) for t in eng.loss_grad(q, q, kp, qm, km, s.initial_offsets)]
    st, z2 = eng.m_stats(kp, q)
    for i in (0, 1, 255, 256, 511, 699):
        r = o.fk(q[i], s.initial_offsets)
        np.testing.assert_array_equal(xp[i], r[1])
        np.testing.assert_array_equal(sx[i], r[3])
        l, g = o.loss_grad(q[i], q[i], qm, kp[i], km, s.initial_offsets)
        assert float(l) == float(L[i])
        np.testing.assert_array_equal(G[i], g)
    so, z2o = o.m_stats(kp, q)
    np.testing.assert_array_equal(npy(st), so)
    assert float(npy(z2)[0]) == float(z2o)
    p, e, it, ls = [npy(t) for t in eng.q_opt(q[:600], kp[:600], qm, km, s.initial_offsets, s.lb, s.ub, 1e-4, maxiter=5)]
    for i in (0, 300, 599):
        po, eo, ito, lso = o.q_opt(q[i], s.lb, s.ub, qm, kp[i], km, s.initial_offsets, 1e-4, maxiter=5)
        np.testing.assert_array_equal(p[i], po)
        assert (it[i], ls[i]) == (ito, lso)


@pytest.mark.parametrize("seed", [0, 2, 3, 5, 7, 8, 10, 11])
def test_random_trees_with_welded_bodies_on_the_register_resident_path(seed):
    """Random hinge trees with welded bodies (tests/random_trees.py): the library and the oracle must agree on WHETHER the
    register-resident path serves the model (same folding, same variant choice) and then bit for bit on loss / gradient,
    single solves (masked, with frozen coordinates outside their box) and clips in every scheduling mode."""
    from oracle.oracle import Oracle
    from random_trees import random_tree
    from stac_mjx_b200.engine import Engine

    t, site_idxs, lb, ub = random_tree(seed, n_bodies=40 + 3 * seed, p_welded=0.3 + 0.03 * seed, n_sites=min(31, 8 + 2 * seed))
    sb, off = t.site_bodyid[site_idxs], t.site_pos[site_idxs].astype(np.float32)
    K = len(sb)
    eng, o = Engine(t, sb, 0), Oracle(t, sb, np.float32, 2)
    assert eng.path == 1 and o.fast_path
    rng = np.random.default_rng(seed)
    q = (t.qpos0 + rng.normal(scale=0.2, size=(6, t.nq))).astype(np.float32)
    q0 = (q + rng.normal(scale=0.05, size=q.shape)).astype(np.float32)
    kp = np.stack([o.fk(q[(i + 1) % 6], off)[3].reshape(-1) for i in range(6)]).astype(np.float32) + 0.004
    qm, km = np.ones(t.nq, bool), np.ones(3 * K, bool)
    part = rng.random(t.nq) < 0.5
    for mask in (qm, part):
        L, G = [npy(x) for x in eng.loss_grad(q, q0, kp, mask, km, off)]
        for i in range(6):
            l, g = o.loss_grad(q[i], q0[i], mask, kp[i], km, off)
            assert float(l) == float(L[i])
            np.testing.assert_array_equal(G[i], g)
    qs = q[:3].copy()
    qs[:, 8:12] += 9.0  # frozen (masked-out) and optimised coordinates outside their boxes
    p, e, it, ls = [npy(x) for x in eng.q_opt(qs, kp[:3], part, km, off, lb, ub, 1e-5, maxiter=40)]
    for i in range(3):
        po, eo, ito, lso = o.q_opt(qs[i], lb, ub, part, kp[i], km, off, 1e-5, maxiter=40)
        assert (it[i], ls[i]) == (ito, lso)
        np.testing.assert_array_equal(p[i], po)
    kpc = kp.reshape(2, 3, -1)
    kw = dict(do_root=1, root_kp_idx=0, trunk_kps=rng.random(K) < 0.7, tol=1e-5, maxiter=60)
    kw["trunk_kps"][0] = True
    ref = o.pose_clips(kpc, t.qpos0, off, lb, ub, part[None], **kw)
    for mode in (0, 1, 2, 3, 4):
        eng.set_mode(mode)
        qio = torch.tensor(np.tile(t.qpos0.astype(np.float32), (2, 1)), device=eng.device)
        out = eng.pose_clips(kpc, qio, off, lb, ub, part[None], **kw)
        np.testing.assert_array_equal(npy(out["iters"]), ref["iters"])
        for k in ("qpos", "xpos", "xquat", "sites", "err"):
            np.testing.assert_array_equal(npy(out[k]), ref[k], err_msg=f"mode {mode} {k}")
    # the general kernels on the same model agree to rounding
    eng.set_mode(-1)
    eng.set_path(1)
    L1, G1 = [npy(x) for x in eng.loss_grad(q, q0, kp, qm, km, off)]
    L0, G0 = [npy(x) for x in Engine(t, sb, 0).loss_grad(q, q0, kp, qm, km, off)]
    np.testing.assert_allclose(L1, L0, rtol=2e-5)
    np.testing.assert_allclose(G1, G0, atol=2e-5 * max(1.0, np.abs(G0).max()))


@pytest.mark.parametrize("seed,n_bodies,n_sites,p_welded", [(21, 70, 20, 0.1), (22, 140, 40, 0.1), (23, 250, 60, 0.1), (25, 250, 140, 0.02), (26, 254, 160, 0.0)])
def test_random_wide_trees_on_the_multi_warp_path(seed, n_bodies, n_sites, p_welded):
    """Random single-hinge trees of 34 .. 200+ jointed elements: the multi-warp register-resident kernels (2, 4, 6, 8 warps per
    chain) against oracle mode 2, bit for bit -- loss / gradient, masked solves, a clip with root optimisation, both register caps and
    the pair mode (two groups of W warps per chain)."""
    from oracle.oracle import Oracle
    from random_trees import random_tree
    from stac_mjx_b200.engine import Engine

    t, site_idxs, lb, ub = random_tree(seed, n_bodies=n_bodies, p_welded=p_welded, max_hinges=1, n_sites=n_sites)
    sb, off = t.site_bodyid[site_idxs], t.site_pos[site_idxs].astype(np.float32)
    K = len(sb)
    eng, o = Engine(t, sb, 0), Oracle(t, sb, np.float32, 2)
    assert eng.path == 1 and o.fast_path
    rng = np.random.default_rng(seed)
    q = (t.qpos0 + rng.normal(scale=0.2, size=(4, t.nq))).astype(np.float32)
    q0 = (q + rng.normal(scale=0.05, size=q.shape)).astype(np.float32)
    kp = np.stack([o.fk(q[(i + 1) % 4], off)[3].reshape(-1) for i in range(4)]).astype(np.float32) + 0.004
    qm, km = np.ones(t.nq, bool), np.ones(3 * K, bool)
    part = rng.random(t.nq) < 0.5
    for mask in (qm, part):
        L, G = [npy(x) for x in eng.loss_grad(q, q0, kp, mask, km, off)]
        for i in range(4):
            l, g = o.loss_grad(q[i], q0[i], mask, kp[i], km, off)
            assert float(l) == float(L[i])
            np.testing.assert_array_equal(G[i], g)
    qs = q[:2].copy()
    qs[:, 8:12] += 9.0
    p, e, it, ls = [npy(x) for x in eng.q_opt(qs, kp[:2], part, km, off, lb, ub, 1e-5, maxiter=25)]
    for i in range(2):
        po, eo, ito, lso = o.q_opt(qs[i], lb, ub, part, kp[i], km, off, 1e-5, maxiter=25)
        assert (it[i], ls[i]) == (ito, lso)
        np.testing.assert_array_equal(p[i], po)
    kpc = kp.reshape(2, 2, -1)
    kw = dict(do_root=1, root_kp_idx=0, trunk_kps=np.ones(K, bool), tol=1e-5, maxiter=30)
    ref = o.pose_clips(kpc, t.qpos0, off, lb, ub, part[None], **kw)
    for mode in (0, 2, 4, -1):  # one group of W warps, the same with capped registers, pair mode (two groups), auto (= pair for few chains)
        eng.set_mode(mode)
        qio = torch.tensor(np.tile(t.qpos0.astype(np.float32), (2, 1)), device=eng.device)
        out = eng.pose_clips(kpc, qio, off, lb, ub, part[None], **kw)
        np.testing.assert_array_equal(npy(out["iters"]), ref["iters"])
        for k in ("qpos", "xpos", "xquat", "sites", "err"):
            np.testing.assert_array_equal(npy(out[k]), ref[k], err_msg=f"mode {mode} {k}")
