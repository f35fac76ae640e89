"""The CPU oracle: pinned against the reference's own known-answer tests where they exist, against an
independent reverse-mode-autodiff restatement, and against the committed golden vectors.

q-phase parity with the real jaxopt/MJX stack is UNPINNED (no golden qpos exists in the reference,
and neither dependency can be imported here); see DESIGN.md section 3.
"""
import numpy as np
import pytest

from oracle.np_oracle import TorchModel
from oracle.oracle import Oracle
from stac_mjx_b200 import mjcf, tree

from conftest import ROOT, get_case

# The 3-body hinge chain of reference tests/unit/test_m_opt.py:17-37 (axes z/x/y, offsets 1 0 0 / 0 1 0 / 0 0 1)
CHAIN_XML = """<mujoco><worldbody>
 <body name="b1" pos="1 0 0"><joint name="j1" type="hinge" axis="0 0 1"/><site name="s1" pos="0.1 0.2 0.3"/>
  <body name="b2" pos="0 1 0"><joint name="j2" type="hinge" axis="1 0 0"/><site name="s2" pos="0.4 0.5 0.6"/>
   <body name="b3" pos="0 0 1"><joint name="j3" type="hinge" axis="0 1 0"/><site name="s3" pos="0.15 0.25 0.35"/></body>
  </body></body></worldbody></mujoco>"""
GT_A = np.array([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6], [0.15, 0.25, 0.35]])
GT_B = np.array([[0.2, -0.1, 0.4], [0.3, 0.4, -0.2], [-0.1, 0.3, 0.1]])


def chain_tree():
    return tree.compile_spec(mjcf.parse_mjcf(CHAIN_XML, from_string=True))


def gen_keypoints(o, q, offsets):
    return np.stack([o.fk(qt, offsets)[3].reshape(-1) for qt in q])


@pytest.fixture(params=[(np.float32, 0), (np.float32, 1), (np.float64, 0)], ids=["f32-mjx", "f32-canon", "f64-mjx"])
def chain(request):
    t = chain_tree()
    return Oracle(t, t.site_bodyid, *request.param)


Z = np.zeros((3, 3))


# ---- the six known-answer tests of reference tests/unit/test_m_opt.py:72-225 ----------------------
def test_kat_identity_pose(chain):
    q = np.zeros((5, 3))
    p, err = chain.m_opt(gen_keypoints(chain, q, GT_A), q, Z, Z, 0.0)
    np.testing.assert_allclose(p, GT_A, atol=1e-5)
    # the reference asserts error < 1e-8; exact cancellation holds in MJX summation order, the
    # canonical (lane-butterfly) order leaves one ulp of z2 ~ 7 (documented in DESIGN.md)
    assert float(err) < (1e-8 if chain.mode == 0 else 2e-6)


def test_kat_varied_random_poses(chain):
    q = np.random.RandomState(42).randn(10, 3).astype(np.float32) * 0.5
    p, _ = chain.m_opt(gen_keypoints(chain, q, GT_A), q, Z, Z, 0.0)
    np.testing.assert_allclose(p, GT_A, atol=1e-5)


def test_kat_sweeping_single_joint(chain):
    q = np.zeros((8, 3))
    q[:, 0] = np.linspace(0.0, np.pi / 4, 8)
    p, _ = chain.m_opt(gen_keypoints(chain, q, GT_B), q, GT_B, Z, 0.0)
    np.testing.assert_allclose(p, GT_B, atol=1e-5)


def test_kat_large_rotations(chain):
    q = np.random.RandomState(99).randn(15, 3).astype(np.float32) * 1.5
    p, _ = chain.m_opt(gen_keypoints(chain, q, GT_B), q, Z, Z, 0.0)
    np.testing.assert_allclose(p, GT_B, atol=1e-4)


def test_kat_reg_zero_vs_strong(chain):
    q = np.random.RandomState(42).randn(10, 3).astype(np.float32) * 0.3
    kp = gen_keypoints(chain, q, GT_A)
    p, _ = chain.m_opt(kp, q, np.ones((3, 3)) * 99.0, np.ones((3, 3)), 0.0)
    np.testing.assert_allclose(p, GT_A, atol=1e-5)
    p, _ = chain.m_opt(kp, q, Z, np.ones((3, 3)), 1e6)
    np.testing.assert_allclose(p, Z, atol=1e-3)


def test_kat_partial_regularization(chain):
    q = np.zeros((10, 3))
    gt = np.array([[0.5, 0.5, 0.5]] * 3)
    kp = gen_keypoints(chain, q, gt)
    is_reg = np.zeros((3, 3))
    is_reg[0] = 1.0
    strong, _ = chain.m_opt(kp, q, Z, is_reg, 1e4)
    noreg, _ = chain.m_opt(kp, q, Z, is_reg, 0.0)
    assert np.linalg.norm(strong[0]) < np.linalg.norm(noreg[0])
    np.testing.assert_allclose(strong[1:], gt[1:], atol=1e-5)


# ---- independent restatement (torch reverse-mode AD) ------------------------------------------------
@pytest.mark.parametrize("name", ["rodent", "celegans", "fly_treadmill"])
def test_gradient_matches_reverse_mode_autodiff(name):
    c = get_case(name)
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    T = TorchModel(c.tree, c.setup.site_bodies)
    for mode in (0, 1):
        o = c.oracle(np.float64, mode)
        for i in range(2):
            L, G = T.loss_grad(g["q"][i], g["q0"][i], g["qm_part"], g["kp"][i], g["km_trunk"], g["offsets"])
            l, gr = o.loss_grad(g["q"][i], g["q0"][i], g["qm_part"], g["kp"][i], g["km_trunk"], g["offsets"])
            assert abs(float(l) - L) <= 1e-12 * max(1.0, abs(L))
            # analytic Jacobian transpose vs AD: differences only from the float32-rounded (slightly
            # non-unit) joint axes, which AD differentiates through
            np.testing.assert_allclose(gr, G, atol=1e-9 * max(1.0, np.abs(G).max()))


def test_gradient_matches_central_differences(rodent):
    g = np.load(ROOT / "tests" / "golden" / "rodent.npz")
    o = rodent.oracle(np.float64, 0)
    qm, km = np.ones(rodent.tree.nq, bool), np.ones(3 * rodent.K, bool)
    q, kp, off = g["q"][0].astype(np.float64), g["kp"][0], g["offsets"]
    _, gr = o.loss_grad(q, q, qm, kp, km, off)
    h = 1e-6
    for i in [0, 2, 3, 6, 7, 12, 20, 33, 45, 60, 73]:
        e = np.zeros_like(q)
        e[i] = h
        num = (float(o.loss_grad(q + e, q, qm, kp, km, off)[0]) - float(o.loss_grad(q - e, q, qm, kp, km, off)[0])) / (2 * h)
        assert abs(num - gr[i]) <= 1e-7 * max(1.0, abs(num)), (i, num, gr[i])


def test_solver_matches_python_restatement_of_jaxopt():
    # jaxopt 0.8.5 ProjectedGradient restated in Python over torch autodiff vs the C solver, float64:
    # same accepted step sizes, same iteration and line-search counts, same parameters.
    t = chain_tree()
    o = Oracle(t, t.site_bodyid, np.float64, 0)
    T = TorchModel(t, t.site_bodyid)
    qt = np.array([0.4, -0.3, 0.5])
    kp = o.fk(qt, GT_A)[3].reshape(-1)
    lb, ub = np.full(3, -0.45), np.full(3, 2.0)  # the box is active for joint 3
    qm, km = np.array([1, 1, 1], bool), np.ones(9, bool)
    x, e, it, ls = T.projected_gradient(np.zeros(3), lb, ub, qm, kp, km, GT_A, 1e-6, maxiter=80)
    xo, eo, ito, lso = o.q_opt(np.zeros(3), lb, ub, qm, kp, km, GT_A, 1e-6, maxiter=80)
    assert (it, ls) == (ito, lso) and it > 3
    np.testing.assert_allclose(xo, x, atol=1e-12)
    assert abs(e - float(eo)) < 1e-12


def test_solver_rodent_root_solve_matches_python_restatement(rodent):
    o = rodent.oracle(np.float64, 0)
    T = TorchModel(rodent.tree, rodent.setup.site_bodies)
    g = np.load(ROOT / "tests" / "golden" / "rodent.npz")
    s = rodent.setup
    q0 = rodent.tree.qpos0.copy()
    q0[:3] = g["kp"][0, 3 * s.root_kp_idx : 3 * s.root_kp_idx + 3]
    x, e, it, ls = T.projected_gradient(q0, s.lb, s.ub, g["root_mask"], g["kp"][0], g["km_trunk"], g["offsets"], 1e-4, maxiter=5)
    xo, eo, ito, lso = o.q_opt(q0, s.lb, s.ub, g["root_mask"], g["kp"][0], g["km_trunk"], g["offsets"], 1e-4, maxiter=5)
    assert (it, ls) == (ito, lso)
    np.testing.assert_allclose(xo, x, atol=1e-12)


# ---- golden vectors ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["rodent", "celegans", "fly_treadmill", "synth_data", "mouse"])
def test_oracle_reproduces_golden_vectors(name):
    c = get_case(name)
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    qm, km = np.ones(c.tree.nq, bool), np.ones(3 * c.K, bool)
    for tag, o, tol in (("f64", c.oracle(np.float64, 0), 1e-12), ("c32", c.oracle(np.float32, 1), 0.0)):
        n = g[f"{tag}_loss"].shape[0]
        for i in range(n):
            r = o.fk(g["q"][i], g["offsets"])
            for j, nm in enumerate(("qpos", "xpos", "xquat", "sites")):
                np.testing.assert_allclose(r[j], g[f"{tag}_fk_{nm}"][i], atol=tol, rtol=0)
            l, gr = o.loss_grad(g["q"][i], g["q"][i], qm, g["kp"][i], km, g["offsets"])
            np.testing.assert_allclose(l, g[f"{tag}_loss"][i], atol=tol, rtol=tol)
            np.testing.assert_allclose(gr, g[f"{tag}_grad"][i], atol=tol, rtol=0)
    kw = c.root_kw()
    C, F = g["c32_clip_qpos"].shape[:2]
    fast = c.oracle(np.float32, 2).fast_path
    assert fast == ("g32_clip_qpos" in g.files)  # every bundled model is served by the register-resident path (mouse: 6 warps)
    for tag, mode in (("c32", 1),) + ((("g32", 2),) if fast else ()):
        o = c.oracle(np.float32, mode)
        r = o.pose_clips(g["kp"].reshape(C, F, -1), c.tree.qpos0, g["offsets"], c.setup.lb, c.setup.ub, c.setup.indiv_parts, nthreads=4, **kw)
        for k in ("qpos", "xpos", "xquat", "sites", "err", "iters", "ls_evals"):
            np.testing.assert_array_equal(r[k], g[f"{tag}_clip_{k}"])
        for i in range(g[f"{tag}_loss"].shape[0]):
            l, gr = o.loss_grad(g["q"][i], g["q"][i], qm, g["kp"][i], km, g["offsets"])
            assert float(l) == float(g[f"{tag}_loss"][i])
            np.testing.assert_array_equal(gr, g[f"{tag}_grad"][i])
            l, gr = o.loss_grad(g["q"][i], g["q0"][i], g["qm_part"], g["kp"][i], g["km_trunk"], g["offsets"])
            assert float(l) == float(g[f"{tag}_mloss"][i])
            np.testing.assert_array_equal(gr, g[f"{tag}_mgrad"][i])
    assert fast


# Mode 0 (the faithful restatement of MJX + jaxopt) is FROZEN: the digests below were taken when the goldens were first
# committed and must never change when modes 1 / 2 follow a kernel change (tools/make_golden.py also reports drift).
MODE0_DIGESTS = {
    "rodent": "c162ee256d089aa2a7bb972c4300a37b",
    "celegans": "2ebe18ab3f738af786ce5cae1321d194",
    "fly_treadmill": "e20fe3b09e66f438d1d0f51b496b623a",
    "synth_data": "0e7b57ad7d4e4ddd7b98ae5bc1de1e16",
    "mouse": "b11191388fd6423c076acb0ad79fb183",
    "rodent_real250": "d105185d2d6723bde419bf046fed2061",
}


@pytest.mark.parametrize("name", sorted(MODE0_DIGESTS))
def test_mode0_goldens_are_frozen(name):
    import hashlib

    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    h = hashlib.sha256()
    for k in sorted(g.files):
        if k.startswith("f64_") or k.startswith("m32_"):
            h.update(k.encode())
            h.update(np.ascontiguousarray(g[k]).tobytes())
    assert h.hexdigest()[:32] == MODE0_DIGESTS[name], "the MJX-order (mode 0) golden vectors changed"


def test_mode0_solver_output_is_frozen(rodent):
    """The mode-0 code itself (not only its stored output): float64 and float32 MJX-order runs of the first frames of the
    real clip reproduce the committed values (float64 to libm rounding, float32 exactly on this libm)."""
    g = np.load(ROOT / "tests" / "golden" / "rodent_real250.npz")
    s, n = rodent.setup, 6
    for tag, dt, tol in (("f64", np.float64, 1e-9), ("m32", np.float32, 2e-3)):
        r = rodent.oracle(dt, 0).pose_clips(g["kp"][None, :n], rodent.tree.qpos0, s.initial_offsets, s.lb, s.ub, s.indiv_parts, **rodent.root_kw())
        np.testing.assert_array_equal(r["iters"][0][:, 1:], g[f"{tag}_iters"][:n, 1:])
        np.testing.assert_allclose(r["qpos"][0], g[f"{tag}_qpos"][:n], atol=tol, rtol=0)


@pytest.mark.parametrize("name", ["rodent", "celegans", "synth_data", "fly_treadmill", "mouse"])
def test_fast_order_agrees_with_mjx_order(name):
    """mode 2 (register-resident kernel arithmetic) vs mode 0: the same mathematics -- to 1e-12 in float64, to float32
    rounding in float32 -- including the analytic gradient against reverse-mode autodiff and a solve that starts with
    passive coordinates outside their box."""
    c = get_case(name)
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    np.testing.assert_allclose(g["g32_loss"], g["f64_loss"], rtol=2e-4)
    gs = np.abs(g["f64_grad"]).max()
    np.testing.assert_allclose(g["g32_grad"], g["f64_grad"], atol=3e-5 * gs)
    np.testing.assert_allclose(g["g32_mgrad"], g["f64_mgrad"], atol=3e-5 * max(np.abs(g["f64_mgrad"]).max(), 1e-9))
    o0, o2 = c.oracle(np.float64, 0), c.oracle(np.float64, 2)
    assert o2.fast_path
    # float64 agreement: 1e-12 where the model's body quaternions are exactly unit in float32; the fruitfly's are unit only to
    # float32 rounding, and MJX's rotate() scales by |q|^2 where the fast order's rotq() does not (a 1e-7 relative difference in
    # the MODEL, far below every tolerance of the path)
    ftol = 1e-12 if name != "fly_treadmill" else 1e-6
    T = TorchModel(c.tree, c.setup.site_bodies)
    for i in range(2):
        a = o0.loss_grad(g["q"][i], g["q0"][i], g["qm_part"], g["kp"][i], g["km_trunk"], g["offsets"])
        b = o2.loss_grad(g["q"][i], g["q0"][i], g["qm_part"], g["kp"][i], g["km_trunk"], g["offsets"])
        L, G = T.loss_grad(g["q"][i], g["q0"][i], g["qm_part"], g["kp"][i], g["km_trunk"], g["offsets"])
        assert abs(float(a[0]) - float(b[0])) < ftol * max(1.0, float(a[0])) and abs(float(b[0]) - L) < ftol * max(1.0, L)
        np.testing.assert_allclose(b[1], a[1], atol=10 * ftol * max(1.0, np.abs(a[1]).max()))
        np.testing.assert_allclose(b[1], G, atol=max(2e-8, ftol) * max(1.0, np.abs(G).max()))
    q0 = g["q0"][0].astype(np.float64).copy()
    q0[-1] = c.setup.ub[-1] + 0.5 if np.isfinite(c.setup.ub[-1]) else q0[-1]  # outside the box (a passive hinge for the rodent)
    qm, km = np.ones(c.tree.nq, bool), np.ones(3 * c.K, bool)
    r0 = o0.q_opt(q0, c.setup.lb, c.setup.ub, qm, g["kp"][0], km, g["offsets"], 1e-9, maxiter=25)
    r2 = o2.q_opt(q0, c.setup.lb, c.setup.ub, qm, g["kp"][0], km, g["offsets"], 1e-9, maxiter=25)
    assert (r0[2], r0[3]) == (r2[2], r2[3])
    np.testing.assert_allclose(r2[0], r0[0], atol=1e-9 if name != "fly_treadmill" else 1e-5)


@pytest.mark.parametrize("name", ["rodent", "celegans", "fly_treadmill", "mouse"])
def test_canonical_order_agrees_with_mjx_order(name):
    """mode 1 (kernel arithmetic order) vs mode 0 (MJX order): same mathematics to float32 rounding."""
    c = get_case(name)
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    scale = np.abs(g["f64_fk_xpos"]).max()
    np.testing.assert_allclose(g["c32_fk_xpos"], g["f64_fk_xpos"], atol=2e-6 * max(scale, 1.0))
    np.testing.assert_allclose(g["c32_fk_sites"], g["f64_fk_sites"], atol=2e-6 * max(scale, 1.0))
    np.testing.assert_allclose(g["c32_loss"], g["f64_loss"], rtol=2e-4)
    gs = np.abs(g["f64_grad"]).max()
    np.testing.assert_allclose(g["c32_grad"], g["f64_grad"], atol=3e-5 * gs)
    np.testing.assert_allclose(g["c32_mgrad"], g["f64_mgrad"], atol=3e-5 * max(np.abs(g["f64_mgrad"]).max(), 1e-9))
    o0, o1 = c.oracle(np.float64, 0), c.oracle(np.float64, 1)
    for i in range(2):
        a, b = o0.fk(g["q"][i], g["offsets"]), o1.fk(g["q"][i], g["offsets"])
        for x, y in zip(a, b):
            np.testing.assert_allclose(x, y, atol=1e-13)


def test_solver_recovers_noise_free_poses(rodent):
    """keypoints generated by FK from in-limit poses are fitted back.

    FISTA on this ill-conditioned problem is slow: with the reference's iteration cap (400 per solve) the
    per-marker misfit of a cold-started clip is a few mm and shrinks frame over frame; with a 10x larger
    cap it drops below 0.3 mm.  (The 1e-4 m figure in SURVEY 8(c) is not reachable at the reference's cap.)
    """
    from stac_mjx_b200 import synth

    t, s = rodent.tree, rodent.setup
    rng = np.random.default_rng(3)
    q = synth.synth_trajectory(t, s.lb, s.ub, 6, rng)
    off = s.initial_offsets.astype(np.float64)
    kp = synth.site_positions(t, s.site_bodies, off, q).reshape(1, 6, -1).astype(np.float32)
    o = rodent.oracle(np.float32, 1)
    worst = {}
    for maxiter in (400, 4000):
        r = o.pose_clips(kp, t.qpos0, off, s.lb, s.ub, s.indiv_parts, nthreads=1, **{**rodent.root_kw(), "tol": 1e-6, "maxiter": maxiter})
        d = np.linalg.norm(r["sites"][0] - kp[0].reshape(6, -1, 3), axis=-1).max(axis=1)
        worst[maxiter] = d
        np.testing.assert_allclose(np.linalg.norm(r["qpos"][0][:, 3:7], axis=1), 1.0, atol=1e-6)
        assert (r["qpos"][0][:, 7:] >= s.lb[7:] - 1e-7).all() and (r["qpos"][0][:, 7:] <= s.ub[7:] + 1e-7).all()
    assert worst[400].max() < 3e-3 and worst[400][-1] < worst[400][0]
    assert worst[4000].max() < 3e-4


def test_noise_floor_is_reported(rodent, capsys):
    """Not an assertion of closeness: documents how far float32 and float64 runs of the SAME algorithm drift."""
    g = np.load(ROOT / "tests" / "golden" / "rodent.npz")
    dq = np.abs(g["c32_clip_qpos"] - g["f64_clip_qpos"])
    ds = np.linalg.norm(g["c32_clip_sites"] - g["f64_clip_sites"], axis=-1)
    with capsys.disabled():
        print(f"\n[noise floor] rodent f32(canonical) vs f64(mjx order): qpos max {dq.max():.2e} median {np.median(dq):.2e} rad;"
              f" marker max {ds.max():.2e} m")  # fmt: skip
    assert ds.max() < 5e-3


def test_oracle_reproduces_real_mocap_clip(rodent):
    """BASELINE config 1: the first frames of the reference's real rat23 recording (committed, tools/make_real_fixture.py)."""
    g = np.load(ROOT / "tests" / "golden" / "rodent_real250.npz")
    s = rodent.setup
    n = 25
    for tag, mode in (("c32", 1), ("g32", 2)):
        r = rodent.oracle(np.float32, mode).pose_clips(g["kp"][None, :n], rodent.tree.qpos0, s.initial_offsets, s.lb, s.ub, s.indiv_parts,
                                                       **rodent.root_kw())  # fmt: skip
        np.testing.assert_array_equal(r["qpos"][0], g[f"{tag}_qpos"][:n])
        np.testing.assert_array_equal(r["iters"][0], g[f"{tag}_iters"][:n])


def test_all_joint_types_gradient_and_orders():
    """free + extra free + hinge + slide + ball joints, 2-3 joints per body: analytic gradient vs reverse-mode autodiff,
    canonical vs MJX order, un-normalised quaternions included."""
    from mixed_model import mixed_tree, random_qpos

    t, site_idxs, lb, ub = mixed_tree()
    sb = t.site_bodyid[site_idxs]
    off = t.site_pos[site_idxs]
    rng = np.random.default_rng(0)
    q = random_qpos(t, rng, 4)
    T = TorchModel(t, sb)
    o0, o1 = Oracle(t, sb, np.float64, 0), Oracle(t, sb, np.float64, 1)
    qm, km = np.ones(t.nq, bool), np.ones(3 * len(sb), bool)
    for i in range(4):
        kp = o0.fk(q[(i + 1) % 4], off)[3].reshape(-1) + 0.01
        L, G = T.loss_grad(q[i], q[i], qm, kp, km, off)
        for o in (o0, o1):
            l, g = o.loss_grad(q[i], q[i], qm, kp, km, off)
            assert abs(float(l) - L) < 1e-12 * max(1.0, L)
            np.testing.assert_allclose(g, G, atol=2e-8 * max(1.0, np.abs(G).max()))
        a, b = o0.fk(q[i], off), o1.fk(q[i], off)
        for x, y in zip(a, b):
            np.testing.assert_allclose(x, y, atol=1e-13)
        np.testing.assert_allclose(np.linalg.norm(a[0][3:7]), 1.0, atol=1e-12)  # quaternions normalised in the returned qpos
    f32 = Oracle(t, sb, np.float32, 1)
    p, e, it, ls = f32.q_opt(q[0], lb, ub, qm, kp, km, off, 1e-5, maxiter=60)
    assert it > 3 and np.isfinite(p).all()


@pytest.mark.parametrize("seed", [0, 3, 6, 9, 11])
def test_fast_order_on_random_trees_with_welded_bodies(seed):
    """Random hinge trees (free root, 1-3 hinges per jointed body, 30-60 % welded bodies, sites on jointed and welded bodies):
    the element folding of the register-resident path (oracle mode 2) computes the same loss and gradient as the MJX order
    and as reverse-mode autodiff.  Body quaternions are unit to float32 rounding only, so float64 agreement is ~1e-7
    relative (MJX's rotate() scales by |q|^2)."""
    from random_trees import n_active, random_tree

    t, site_idxs, lb, ub = random_tree(seed, n_bodies=40 + 3 * seed, p_welded=0.3 + 0.03 * seed, n_sites=min(31, 8 + 2 * seed))
    sb, off = t.site_bodyid[site_idxs], t.site_pos[site_idxs]
    na, nj = n_active(t, sb)
    o0, o2 = Oracle(t, sb, np.float64, 0), Oracle(t, sb, np.float64, 2)
    assert o2.fast_path and nj <= 31 and (seed < 3) == (na <= 31)  # seeds >= 3 need the folding
    rng = np.random.default_rng(seed)
    T = TorchModel(t, sb)
    qm, km = np.ones(t.nq, bool), np.ones(3 * len(sb), bool)
    part = rng.random(t.nq) < 0.5
    for i in range(2):
        q = t.qpos0 + rng.normal(scale=0.2, size=t.nq)
        q0 = q + rng.normal(scale=0.05, size=t.nq)
        kp = o0.fk(t.qpos0 + rng.normal(scale=0.2, size=t.nq), off)[3].reshape(-1) + 0.003
        for mask in (qm, part):
            a, b = o0.loss_grad(q, q0, mask, kp, km, off), o2.loss_grad(q, q0, mask, kp, km, off)
            L, G = T.loss_grad(q, q0, mask, kp, km, off)
            assert abs(float(a[0]) - float(b[0])) < 2e-6 * max(1e-3, float(a[0]))
            np.testing.assert_allclose(b[1], a[1], atol=2e-6 * max(1.0, np.abs(a[1]).max()))
            np.testing.assert_allclose(b[1], G, atol=2e-6 * max(1.0, np.abs(G).max()))
            assert (b[1][~mask] == 0).all()
    r0 = o0.q_opt(q, lb, ub, qm, kp, km, off, 1e-9, maxiter=15)
    r2 = o2.q_opt(q, lb, ub, qm, kp, km, off, 1e-9, maxiter=15)
    np.testing.assert_allclose(r2[0], r0[0], atol=2e-4)


@pytest.mark.parametrize("seed,n_bodies,n_sites,p_welded", [(21, 70, 20, 0.1), (22, 140, 40, 0.1), (23, 250, 60, 0.1), (25, 250, 140, 0.02), (26, 254, 160, 0.0)])
def test_wide_order_on_random_trees(seed, n_bodies, n_sites, p_welded):
    """Random single-hinge trees of 34 .. 200+ jointed elements (2, 4, 6, 8 warps per evaluation, more sites than one warp holds,
    welded bodies folded): oracle mode 2 in its multi-warp order against the MJX order and reverse-mode autodiff."""
    from random_trees import n_active, random_tree

    t, site_idxs, lb, ub = random_tree(seed, n_bodies=n_bodies, p_welded=p_welded, max_hinges=1, n_sites=n_sites)
    sb, off = t.site_bodyid[site_idxs], t.site_pos[site_idxs]
    na, nj = n_active(t, sb)
    o0, o2 = Oracle(t, sb, np.float64, 0), Oracle(t, sb, np.float64, 2)
    assert o2.fast_path and nj > 31
    rng = np.random.default_rng(seed)
    T = TorchModel(t, sb)
    qm, km = np.ones(t.nq, bool), np.ones(3 * len(sb), bool)
    part = rng.random(t.nq) < 0.5
    q = t.qpos0 + rng.normal(scale=0.2, size=t.nq)
    q0 = q + rng.normal(scale=0.05, size=t.nq)
    kp = o0.fk(t.qpos0 + rng.normal(scale=0.2, size=t.nq), off)[3].reshape(-1) + 0.003
    for mask in (qm, part):
        a, b = o0.loss_grad(q, q0, mask, kp, km, off), o2.loss_grad(q, q0, mask, kp, km, off)
        L, G = T.loss_grad(q, q0, mask, kp, km, off)
        assert abs(float(a[0]) - float(b[0])) < 2e-6 * max(1e-3, float(a[0]))
        np.testing.assert_allclose(b[1], a[1], atol=2e-6 * max(1.0, np.abs(a[1]).max()))
        np.testing.assert_allclose(b[1], G, atol=2e-6 * max(1.0, np.abs(G).max()))
    r0 = o0.q_opt(q, lb, ub, qm, kp, km, off, 1e-9, maxiter=10)
    r2 = o2.q_opt(q, lb, ub, qm, kp, km, off, 1e-9, maxiter=10)
    np.testing.assert_allclose(r2[0], r0[0], atol=2e-4)
