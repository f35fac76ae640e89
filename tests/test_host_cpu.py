"""Host-side logic that needs no GPU: clip batching, the duck-typed solver seam, packaging layout, pipeline errors.

Mirrors reference tests/unit/test_utils_math.py:41-50, tests/unit/test_compute_stac.py:54-182,
tests/unit/test_stac_package.py and tests/unit/test_main_run_stac.py.
"""
import types

import numpy as np
import pytest

from stac_mjx_b200 import compute_stac, config, io, main, parallel, utils


def test_batch_kp_data_shapes():
    assert utils.batch_kp_data(np.zeros((10, 6), np.float32), 4).shape == (2, 4, 6)
    assert utils.batch_kp_data(np.zeros((30, 6), np.float32), 10, continuous=True).shape == (3, 20, 6)


def test_batch_kp_data_continuous_overlap_and_wrap():
    kp = np.arange(30, dtype=np.float32).reshape(30, 1).repeat(3, 1)
    b = utils.batch_kp_data(kp, 10, continuous=True)
    np.testing.assert_array_equal(b[0, :, 0], np.arange(20))
    np.testing.assert_array_equal(b[1, :, 0], np.arange(10, 30))
    np.testing.assert_array_equal(b[2, :, 0], np.r_[np.arange(20, 30), np.arange(20, 30)])  # mode="wrap" pad of the last clip


class FakeData:
    def __init__(self, qpos, site_xpos=None, xpos=None, xquat=None):
        self.qpos = qpos
        self.site_xpos = site_xpos if site_xpos is not None else np.zeros((2, 3))
        self.xpos = xpos if xpos is not None else np.zeros((2, 3))
        self.xquat = xquat if xquat is not None else np.zeros((2, 4))

    def replace(self, **kw):
        return FakeData(kw.get("qpos", self.qpos), kw.get("site_xpos", self.site_xpos), kw.get("xpos", self.xpos), kw.get("xquat", self.xquat))


class FakeModel:
    def __init__(self, nq, jnt_type, site_pos):
        self.nq, self.jnt_type, self.site_pos = nq, jnt_type, site_pos

    def replace(self, **kw):
        m = FakeModel(self.nq, self.jnt_type, self.site_pos)
        m.__dict__.update(kw)
        return m


class FakeStacCore:
    def __init__(self):
        self.q_calls, self.m_calls, self.q0_args = 0, 0, []

    def q_opt(self, *args, **kwargs):
        self.q_calls += 1
        q0 = args[5]
        self.q0_args.append(np.array(q0))
        return args[1], types.SimpleNamespace(params=q0, state=types.SimpleNamespace(error=0.0))

    def m_opt(self, mjx_model, mjx_data, keypoints, q, initial_offsets, *args, **kwargs):
        self.m_calls += 1
        return types.SimpleNamespace(params=np.asarray(initial_offsets).reshape(-1, 3), error=0.0)


def test_root_optimization_calls_q_opt_twice_and_seeds_translation():
    core = FakeStacCore()
    mdl = FakeModel(7, np.array([0]), np.zeros((2, 3)))
    data = FakeData(np.array([91.0, 92.0, 93.0, 4.0, 5.0, 6.0, 7.0]))
    kp = np.array([[11.0, 12.0, 13.0, 21.0, 22.0, 23.0]])
    out = compute_stac.root_optimization(core, mdl, data, kp, 1, np.zeros(7), np.ones(7), np.array([0, 1]), np.array([True, True]))
    assert isinstance(out, FakeData) and core.q_calls == 2
    expected = np.array([21.0, 22.0, 23.0, 4.0, 5.0, 6.0, 7.0])
    np.testing.assert_allclose(core.q0_args[0], expected)
    np.testing.assert_allclose(core.q0_args[1], expected)


def test_offset_optimization_calls_m_opt_once_and_writes_site_pos():
    core = FakeStacCore()
    mdl = FakeModel(7, np.array([0]), np.zeros((2, 3)))
    offsets = np.zeros((2, 3))
    mdl2, data2, off = compute_stac.offset_optimization(
        core, mdl, FakeData(np.zeros(7)), np.zeros((4, 6)), offsets, np.zeros((4, 7)), 2, np.zeros((2, 3)), np.array([0, 1]), 0.0
    )
    assert core.m_calls == 1
    np.testing.assert_allclose(off, offsets)
    np.testing.assert_allclose(mdl2.site_pos, offsets)


def test_pose_optimization_runs_all_frames_through_the_seam():
    core = FakeStacCore()
    mdl = FakeModel(7, np.array([0]), np.zeros((2, 3)))
    res = compute_stac.pose_optimization(core, mdl, FakeData(np.zeros(7)), np.zeros((2, 6)), np.zeros(7), np.ones(7), np.array([0, 1]), [])
    _, qposes, _, _, marker_sites, _, frame_error = res
    assert qposes.shape == (2, 7) and len(marker_sites) == 2 and len(frame_error) == 2
    assert core.q_calls == 2
    core = FakeStacCore()
    parts = [np.array([1, 1, 1, 0, 0, 0, 0], bool), np.array([0, 0, 0, 1, 1, 1, 1], bool)]
    compute_stac.pose_optimization(core, mdl, FakeData(np.zeros(7)), np.zeros((3, 6)), np.zeros(7), np.ones(7), np.array([0, 1]), parts)
    assert core.q_calls == 3 * (1 + 2)  # 1 + P solves per frame (compute_stac.py:219-250)


def test_sample_time_indices_covers_all_frames_when_sample_exceeds_clip():
    idx = compute_stac.sample_time_indices(10, 100)
    assert sorted(idx.tolist()) == list(range(10))
    idx = compute_stac.sample_time_indices(1000, 100)
    assert len(idx) == 100 and len(set(idx.tolist())) == 100


def test_package_layout_quirk_of_batched_runs():
    """reference stac.py:483-486: xpos/xquat flattened clip-major (order='F' on [F,C,...]) but marker_sites
    flattened with a C-order reshape of [F,C,K,3], i.e. frame-major interleave."""
    from stac_mjx_b200.stac import Stac

    C, F, nb, K, nq = 3, 4, 2, 2, 5
    tag = lambda c, f: 100 * c + f
    qposes = np.array([[[tag(c, f)] * nq for f in range(F)] for c in range(C)], np.float32)
    xposes = np.array([[[[tag(c, f)] * 3] * nb for c in range(C)] for f in range(F)], np.float32)  # [F,C,nb,3]
    xquats = np.array([[[[tag(c, f)] * 4] * nb for c in range(C)] for f in range(F)], np.float32)
    sites = np.array([[[[tag(c, f)] * 3] * K for c in range(C)] for f in range(F)], np.float32)
    kp = np.zeros((C, F, 3 * K), np.float32)
    fake = types.SimpleNamespace(_part_names=["q"] * nq, _body_names=["b"] * nb, _kp_names=["k"] * K, _offsets=None)
    mdl = types.SimpleNamespace(site_pos=types.SimpleNamespace(cpu=lambda: types.SimpleNamespace(numpy=lambda: np.zeros((K, 3)))))
    d = Stac._package_data(fake, mdl, qposes, xposes, xquats, sites, kp, batched=True)
    clip_major = [tag(c, f) for c in range(C) for f in range(F)]
    frame_major = [tag(c, f) for f in range(F) for c in range(C)]
    assert d.qpos[:, 0].tolist() == clip_major
    assert d.xpos[:, 0, 0].tolist() == clip_major and d.xquat[:, 0, 0].tolist() == clip_major
    assert d.marker_sites[:, 0, 0].tolist() == frame_major
    assert d.kp_data.shape == (C * F, 3 * K)


def test_run_stac_rejects_bad_column_count_and_indivisible_clips(monkeypatch):
    cfg = config.Cfg({"model": {"MJCF_PATH": "x.xml"}, "stac": {"fit_offsets_path": "a.h5", "ik_only_path": "b.h5", "skip_fit_offsets": True,
                                                                  "skip_ik_only": False, "n_frames_per_clip": 4, "n_fit_frames": 2}})  # fmt: skip
    with pytest.raises(ValueError, match="columns but expected"):
        main.run_stac(cfg, np.zeros((8, 5), np.float32), ["a", "b"])

    class DummyStac:
        def __init__(self, *a, **k):
            pass

    monkeypatch.setattr(main, "Stac", DummyStac)
    with pytest.raises(ValueError, match="must divide evenly"):
        main.run_stac(cfg, np.zeros((10, 6), np.float32), ["a", "b"])
    cfg.stac.skip_ik_only = True
    fit_path, ik_path = main.run_stac(cfg, np.zeros((10, 6), np.float32), ["a", "b"])
    assert ik_path is None and str(fit_path).endswith("a.h5")


def test_config_composer_reads_a_hydra_style_tree(tmp_path):
    (tmp_path / "model").mkdir()
    (tmp_path / "stac").mkdir()
    (tmp_path / "config.yaml").write_text("defaults:\n  - stac: demo\n  - model: toy\n  - _self_\n")
    (tmp_path / "model" / "toy.yaml").write_text("MJCF_PATH: m.xml\nFTOL: 1.0e-4\nN_ITER_Q: 400\nKEYPOINT_MODEL_PAIRS: {a: b}\n")
    (tmp_path / "model" / "other.yaml").write_text("MJCF_PATH: o.xml\n")
    (tmp_path / "stac" / "demo.yaml").write_text("n_frames_per_clip: 250\ncontinuous: False\nmujoco: {solver: newton}\n")
    cfg = config._compose_yaml(tmp_path, "config", [])
    assert cfg.model.MJCF_PATH == "m.xml" and cfg.stac.n_frames_per_clip == 250 and cfg.stac.mujoco.solver == "newton"
    assert "ROOT_OPTIMIZATION_KEYPOINT" not in cfg.model and cfg.model.get("SITES_TO_REGULARIZE", []) == []
    cfg = config._compose_yaml(tmp_path, "config", ["model=other", "stac.n_frames_per_clip=10"])
    assert cfg.model.MJCF_PATH == "o.xml" and cfg.stac.n_frames_per_clip == 10


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 7, 72, 73):
        for ws in (1, 2, 3, 8):
            blocks = [parallel.shard_range(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_edge_effect_crossfade_shapes():
    n_clip, F, ov = 3, 20, utils.CONTINUOUS_BATCH_OVERLAP
    mk = lambda *tail: np.random.default_rng(0).normal(size=(n_clip * (F + ov),) + tail)
    d = io.StacData(qpos=mk(5), xpos=mk(2, 3), xquat=mk(2, 4), marker_sites=mk(2, 3), offsets=np.zeros((2, 3)), kp_data=mk(6),
                    names_qpos=[], names_xpos=[], kp_names=[])  # fmt: skip
    out = utils.handle_edge_effects(d, F)
    assert out.qpos.shape == (n_clip * F, 5) and out.xpos.shape == (n_clip * F, 2, 3)


def test_velocity_from_kinematics():
    # reference tests/unit/test_utils_math.py: constant joint ramp without a free joint, and a pure z rotation with one
    dt = 0.1
    q = np.stack([np.linspace(0, 1, 5), np.linspace(0, -2, 5)], axis=1)
    v = utils.compute_velocity_from_kinematics(q, dt, freejoint=False)
    np.testing.assert_allclose(v[:-1], np.tile([2.5, -5.0], (4, 1)), rtol=1e-5)
    np.testing.assert_allclose(v[-1], 0.0)
    ang = np.linspace(0, 0.4, 5)
    qf = np.zeros((5, 9))
    qf[:, 0] = np.linspace(0, 1, 5)
    qf[:, 3], qf[:, 6] = np.cos(ang / 2), np.sin(ang / 2)
    qf[:, 8] = 100.0 * np.arange(5)
    v = utils.compute_velocity_from_kinematics(qf, dt)
    assert v.shape == (5, 8)
    np.testing.assert_allclose(v[:-1, 0], 2.5, rtol=1e-5)
    np.testing.assert_allclose(v[:-1, 5], 1.0, rtol=1e-4)  # 0.1 rad per 0.1 s about z
    np.testing.assert_allclose(v[:-1, 3:5], 0.0, atol=1e-5)
    assert (v[:-1, 7] == 20.0).all()  # joint velocities clipped, root velocities not


def test_real_mocap_fixture_is_what_load_data_produces(tmp_path):
    """io.load_data on a DANNCE-style .mat: order by KEYPOINT_MODEL_PAIRS, scale, flatten keypoint-major (io.py:39-98)."""
    import scipy.io as spio

    from conftest import ROOT, get_case
    from stac_mjx_b200.config import Cfg

    c = get_case("rodent")
    g = np.load(ROOT / "tests" / "golden" / "rodent_real250.npz")
    # rebuild a [frames, xyz, keypoints] millimetre array in the mocap's own KP_NAMES order and round-trip it
    names_model = c.kp_names
    names_mocap = list(c.cfg.model.KP_NAMES)
    kp = g["kp"][:20].reshape(20, len(names_model), 3)
    pred = np.zeros((20, 3, len(names_mocap)))
    for k, n in enumerate(names_model):
        pred[:, :, names_mocap.index(n)] = kp[:, k, :] / c.cfg.model.MOCAP_SCALE_FACTOR
    spio.savemat(tmp_path / "m.mat", {"pred": pred})
    cfg = Cfg(c.cfg.to_dict())
    cfg.stac.data_path = "m.mat"
    out, names = io.load_data(cfg, base_path=tmp_path)
    assert names == names_model and out.shape == (20, 69) and out.dtype == np.float32
    np.testing.assert_allclose(out, g["kp"][:20], rtol=1e-6, atol=1e-9)


def test_load_data_keypoint_name_errors_and_label3d_names(tmp_path):
    """reference tests/unit/test_utils.py:48-80: names from a label3d file, no names at all, fewer names than keypoints."""
    import scipy.io as spio

    from conftest import get_case
    from stac_mjx_b200.config import Cfg

    c = get_case("rodent")
    names_mocap = list(c.cfg.model.KP_NAMES)
    K = len(names_mocap)
    rng = np.random.default_rng(3)
    pred = rng.normal(size=(7, 3, K))
    spio.savemat(tmp_path / "m.mat", {"pred": pred})
    spio.savemat(tmp_path / "names.mat", {"joint_names": np.array([[n] for n in names_mocap], dtype=object)})

    def cfg_with(**model):
        d = c.cfg.to_dict()
        d["stac"]["data_path"] = "m.mat"
        for k, v in model.items():
            if v is None:
                d["model"].pop(k, None)
            else:
                d["model"][k] = v
        return Cfg(d)

    base, names0 = io.load_data(cfg_with(), base_path=tmp_path)
    assert base.shape == (7, 3 * K) and len(names0) == K
    # keypoint names read from the label3d file instead of the config
    out, names = io.load_data(cfg_with(KP_NAMES=None, KP_NAMES_LABEL3D_PATH=str(tmp_path / "names.mat")), base_path=tmp_path)
    assert names == names0
    np.testing.assert_array_equal(out, base)
    with pytest.raises(ValueError, match="Keypoint names not provided"):
        io.load_data(cfg_with(KP_NAMES=None), base_path=tmp_path)
    with pytest.raises(ValueError, match="is not the same as the number of keypoints"):
        io.load_data(cfg_with(KP_NAMES=names_mocap[:-2]), base_path=tmp_path)
    bad = cfg_with()
    bad.stac.data_path = "m.csv"
    with pytest.raises(ValueError, match="Unsupported file extension"):
        io.load_data(bad, base_path=tmp_path)


def test_package_data_unbatched():
    # reference tests/unit/test_stac_package.py:11-38
    from stac_mjx_b200.stac import Stac

    dummy = types.SimpleNamespace(_offsets=np.zeros((2, 3)), _part_names=["root"], _body_names=["body"], _kp_names=["kp1", "kp2"],
                                  _body_site_idxs=np.array([0, 1]))  # fmt: skip
    out = Stac._package_data(dummy, None, np.zeros((2, 1)), np.zeros((2, 1, 3)), np.zeros((2, 1, 4)), np.zeros((2, 2, 3)),
                             np.zeros((2, 6)), batched=False)  # fmt: skip
    assert out.offsets.shape == (2, 3) and out.kp_names == ["kp1", "kp2"] and out.qpos.shape == (2, 1)


def test_velocity_no_freejoint_reference_case():
    # reference tests/unit/test_utils_math.py: three frames of a linear ramp, dt = 1
    q = np.array([[0.0, 0.0, 0.0], [1.0, 2.0, 3.0], [2.0, 4.0, 6.0]])
    v = utils.compute_velocity_from_kinematics(q, dt=1.0, freejoint=False, max_qvel=100.0)
    assert v.shape == (3, 3)
    np.testing.assert_allclose(v[0], [1, 2, 3])
    np.testing.assert_allclose(v[1], [1, 2, 3])
    np.testing.assert_allclose(v[2], [0, 0, 0])


def test_threefry_known_answers_and_jax_permutation():
    """jax.random.permutation(PRNGKey(0), arange(F)) restated in numpy (reference compute_stac.py:136-140).

    Pins: the three Random123 known-answer vectors for threefry2x32_20 (also JAX's own prng test) and the published
    value of jax.random.split(PRNGKey(0)) under the original key derivation."""
    from stac_mjx_b200 import jax_random as jr

    kat = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]  # fmt: skip
    for key, ctr, want in kat:
        a, b = jr.threefry2x32(key, [ctr[0]], [ctr[1]])
        assert (int(a[0]), int(b[0])) == want
    assert jr.prng_key(0) == (0, 0) and jr.prng_key((5 << 32) | 7) == (5, 7)
    assert jr.split((0, 0), 2, partitionable=False) == [(4146024105, 967050713), (2718843009, 1272950319)]
    assert jr.split((0, 0), 2, partitionable=True) == [(1797259609, 2579123966), (928981903, 3453687069)]
    for n in (1, 2, 10, 1000, 2000):  # 2000 needs two sort rounds (ceil(3 ln n / ln(2^32-1)))
        for part in (True, False):
            p = jr.permutation(0, n, part)
            assert sorted(p.tolist()) == list(range(n))
    # the config-3 sample (1000 fit frames, N_SAMPLE_FRAMES=100): fixed forever, integer arithmetic only
    idx = compute_stac.sample_time_indices(1000, 100)
    assert idx[:10].tolist() == [166, 872, 474, 336, 210, 769, 0, 475, 36, 835]
    assert len(set(idx.tolist())) == 100


def test_make_qs_accepts_torch_and_numpy_in_any_mix():
    import torch

    q0, q = np.arange(5, dtype=np.float32), -np.ones(5, dtype=np.float32)
    m = np.array([1, 0, 1, 0, 0], bool)
    want = np.where(m, q, q0)
    assert np.array_equal(utils.make_qs(q0, m, q), want)
    assert np.array_equal(utils.make_qs(q0, m.astype(np.float32), q), want)
    for mm in (m, torch.from_numpy(m), torch.from_numpy(m.astype(np.float32))):
        for a, b in ((torch.from_numpy(q0), torch.from_numpy(q)), (q0, torch.from_numpy(q)), (torch.from_numpy(q0), q)):
            out = utils.make_qs(a, mm, b)
            assert isinstance(out, torch.Tensor) and np.array_equal(out.numpy(), want)


class _FakeH5Dataset:
    def __init__(self, data, compression):
        self.data, self.compression = np.asarray(data), compression

    def __getitem__(self, key):
        return self.data[key]

    def __iter__(self):
        return iter(self.data)


class _FakeH5File(dict):
    """The slice of the h5py.File API the writer / reader use, backed by a dict that outlives the `with` block."""

    stores: dict = {}

    def __init__(self, path, mode):
        super().__init__()
        self.path, self.mode = str(path), mode
        if mode == "r":
            self.update(self.stores[self.path])

    def create_dataset(self, name, data=None, compression=None):
        assert self.mode == "w" and name not in self
        self[name] = _FakeH5Dataset(data, compression)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.mode == "w":
            self.stores[self.path] = dict(self)
        return False


def test_h5_writer_and_reader_run_for_real_against_the_reference_layout(monkeypatch, tmp_path):
    """`io.save_data_to_h5` / `io.load_stac_data` (reference io.py:194-278) executed end to end.  h5py is absent from this
    image, so an in-memory stand-in with the h5py.File surface the two functions touch is injected: the test pins the dataset
    names and order, the string / bytes encodings, which datasets are gzip-compressed, and the round trip -- with the real
    h5py the same body runs below when it is installed."""
    import sys
    import types

    from stac_mjx_b200 import io
    from stac_mjx_b200.model import load_fixture

    _, cfg = load_fixture("rodent")
    rng = np.random.default_rng(0)
    d = io.StacData(
        qpos=rng.normal(size=(6, 74)).astype(np.float32), xpos=rng.normal(size=(6, 67, 3)).astype(np.float32),
        xquat=rng.normal(size=(6, 67, 4)).astype(np.float32), marker_sites=rng.normal(size=(6, 23, 3)).astype(np.float32),
        offsets=rng.normal(size=(23, 3)).astype(np.float32), kp_data=rng.normal(size=(6, 69)).astype(np.float32),
        names_qpos=[f"q{i}" for i in range(74)], names_xpos=[f"b{i}" for i in range(67)], kp_names=[f"k{i}" for i in range(23)],
        qvel=rng.normal(size=(6, 73)).astype(np.float32),
    )  # fmt: skip

    def roundtrip(path):
        io.save_data_to_h5(config=cfg, file_path=path, **d.as_dict())
        cfg2, d2 = io.load_stac_data(path)
        assert cfg2.to_dict() == cfg.to_dict()
        for k in ("qpos", "xpos", "xquat", "marker_sites", "offsets", "kp_data", "qvel"):
            np.testing.assert_array_equal(getattr(d2, k), getattr(d, k))
        assert d2.names_qpos == d.names_qpos and d2.names_xpos == d.names_xpos and d2.kp_names == d.kp_names

    fake = types.ModuleType("h5py")
    fake.File = _FakeH5File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    path = tmp_path / "fit.h5"
    roundtrip(path)
    store = _FakeH5File.stores[str(path)]
    # the reference's datasets, in its order (io.py:225-237)
    assert list(store) == ["config", "kp_names", "names_qpos", "names_xpos", "kp_data", "marker_sites", "offsets", "qpos", "qvel", "xpos", "xquat"]
    assert store["config"].data.dtype.kind == "S" and store["config"].data.shape == () and store["config"].compression is None
    for k in ("kp_names", "names_qpos", "names_xpos"):
        assert store[k].data.dtype.kind == "S" and store[k].compression is None
    for k in ("kp_data", "marker_sites", "offsets", "qpos", "qvel", "xpos", "xquat"):
        assert store[k].compression == "gzip" and store[k].data.dtype == np.float32
    import yaml

    assert yaml.safe_load(store["config"].data[()].decode("utf-8"))["model"]["N_ITER_Q"] == cfg.model.N_ITER_Q
    monkeypatch.delitem(sys.modules, "h5py")
    try:
        import h5py  # noqa: F401
    except ImportError:
        return
    roundtrip(tmp_path / "real.h5")
