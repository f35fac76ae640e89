import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Case:
    """A model fixture with everything Stac derives from it."""

    def __init__(self, name):
        from stac_mjx_b200 import model

        self.name = name
        self.tree, self.cfg = model.load_fixture(name)
        self.kp_names = list(self.cfg.model.KEYPOINT_MODEL_PAIRS.keys())
        self.setup = model.make_setup(self.tree, self.cfg.model, self.kp_names)
        self.K = len(self.kp_names)
        self.tol = float(self.cfg.model.FTOL)

    def oracle(self, dtype=np.float32, mode=2):
        from oracle.oracle import Oracle

        return Oracle(self.tree, self.setup.site_bodies, dtype, mode)

    def session(self, n_frames, clip, seed=11):
        from stac_mjx_b200 import synth

        return synth.synth_session(self.tree, self.setup, n_frames, clip, seed=seed)

    def root_kw(self):
        s = self.setup
        has_root = s.root_kp_idx >= 0 and int(self.tree.jnt_type[0]) in (0, 2)
        return dict(do_root=1 if has_root else 0, root_kp_idx=s.root_kp_idx, trunk_kps=s.trunk_kps,
                    root_dims=4 if int(self.tree.jnt_type[0]) == 2 else 7, tol=self.tol)  # fmt: skip


_CASES = {}


def get_case(name):
    if name not in _CASES:
        _CASES[name] = Case(name)
    return _CASES[name]


@pytest.fixture(scope="session")
def rodent():
    return get_case("rodent")


@pytest.fixture(scope="session")
def engine_of():
    engines = {}

    def make(case):
        from stac_mjx_b200.engine import Engine

        if case.name not in engines:
            engines[case.name] = Engine(case.tree, case.setup.site_bodies, 0)
        return engines[case.name]

    return make
