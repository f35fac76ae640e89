"""The C-ABI library loads without a GPU and exports every symbol include/stacb.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (ROOT / "include" / "stacb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stacb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    from stac_mjx_b200 import _lib

    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from stac_mjx_b200 import _lib

    if not _lib.LIB_PATH.exists():
        pytest.fail("stac_mjx_b200/libstacb.so is not built; run `python __graft_entry__.py`")
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    for sym in header_symbols():
        assert hasattr(L, sym), f"{sym} missing from libstacb.so"
    L.stacb_version.restype = ctypes.c_int
    m = re.search(r"#define STACB_VERSION (\d+)", (ROOT / "include" / "stacb.h").read_text())
    assert L.stacb_version() == int(m.group(1))


def test_bad_arguments_are_rejected_without_a_gpu():
    from stac_mjx_b200 import _lib

    L = _lib.lib()
    assert L.stacb_fk(None, None, None, None, None, None, None, 1, None) == -1
    assert b"stacb_fk" in L.stacb_last_error()
    assert L.stacb_tree_create(None, 0, None) == -1


def test_product_never_imports_the_oracle():
    pkg = ROOT / "stac_mjx_b200"
    for p in pkg.rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), flags=re.M), p
        assert "oracle/" not in p.read_text(), p
    for p in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.sh")) + [ROOT / "include" / "stacb.h"]:
        assert not re.search(r"#include.*oracle|oracle/", p.read_text()), p


def test_engine_fails_loudly_without_cuda():
    import torch

    from conftest import get_case
    from stac_mjx_b200 import _lib
    from stac_mjx_b200.engine import Engine

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    c = get_case("synth_data")
    with pytest.raises(_lib.StacbError):
        Engine(c.tree, c.setup.site_bodies)


def test_jax_ffi_front_end_is_guarded():
    """No jax in this image: the FFI front-end reports unavailable and refuses loudly instead of falling back."""
    from stac_mjx_b200 import jax_ffi

    try:
        import jax.ffi  # noqa: F401

        pytest.skip("jax is installed here")
    except ImportError:
        pass
    assert jax_ffi.available() is False
    with pytest.raises(RuntimeError, match="JAX FFI path unavailable"):
        jax_ffi.register()


def test_tree_create_rejects_body_orders_that_are_not_depth_first():
    """The gradient's subtree = contiguous site range logic needs DFS pre-order body ids (what MuJoCo produces): a merely
    topological order (every parent before its children, but a subtree interleaved with a sibling) is refused -- before any
    CUDA call, so this runs without a GPU."""
    import ctypes as C

    import numpy as np

    from stac_mjx_b200 import _lib

    L = _lib.lib()
    nb = 5
    parent = np.array([0, 0, 0, 1, 2], np.int32)  # body 3 hangs off body 1, but body 2 (a sibling subtree) sits in between
    jadr = np.array([-1, 0, 1, 2, 3], np.int32)
    jnum = np.array([0, 1, 1, 1, 1], np.int32)
    pos, quat = np.zeros((nb, 3), np.float32), np.tile(np.array([1, 0, 0, 0], np.float32), (nb, 1))
    jt, jq, jb = np.full(4, 3, np.int32), np.arange(4, dtype=np.int32), np.arange(1, 5, dtype=np.int32)
    jp, ja = np.zeros((4, 3), np.float32), np.tile(np.array([0, 0, 1], np.float32), (4, 1))
    q0, sb = np.zeros(4, np.float32), np.array([3, 4], np.int32)
    keep = [parent, jadr, jnum, pos, quat, jt, jq, jb, jp, ja, q0, sb]
    d = _lib.TreeDesc(nb, 4, 4, 2, *[a.ctypes.data_as(C.c_void_p) for a in keep])
    h = C.c_void_p()
    assert L.stacb_tree_create(C.byref(d), 0, C.byref(h)) == -1
    assert b"depth-first pre-order" in L.stacb_last_error()


def test_tree_create_rejects_malformed_descriptions():
    """Joint ranges, joint types and qpos addresses are validated before anything is dereferenced on their strength (no GPU)."""
    import ctypes as C

    import numpy as np

    from stac_mjx_b200 import _lib

    L = _lib.lib()

    def create(nq=2, jadr=(-1, 0, 1), jnum=(0, 1, 1), jtype=(3, 3), jq=(0, 1), null_field=None):
        nb = 3
        parent = np.array([0, 0, 1], np.int32)
        pos, quat = np.zeros((nb, 3), np.float32), np.tile(np.array([1, 0, 0, 0], np.float32), (nb, 1))
        jp, ja = np.zeros((2, 3), np.float32), np.tile(np.array([0, 0, 1], np.float32), (2, 1))
        keep = [parent, np.array(jadr, np.int32), np.array(jnum, np.int32), pos, quat, np.array(jtype, np.int32), np.array(jq, np.int32),
                np.array([1, 2], np.int32), jp, ja, np.zeros(max(nq, 1), np.float32), np.array([2], np.int32)]  # fmt: skip
        ptrs = [a.ctypes.data_as(C.c_void_p) for a in keep]
        if null_field is not None:
            ptrs[null_field] = C.c_void_p()
        d = _lib.TreeDesc(nb, nq, 2, 1, *ptrs)
        h = C.c_void_p()
        rc = L.stacb_tree_create(C.byref(d), 0, C.byref(h))
        return rc, L.stacb_last_error()

    rc, msg = create(null_field=3)
    assert rc == -1 and b"null array" in msg
    rc, msg = create(jadr=(-1, 0, 2))  # joint range past the joint arrays
    assert rc == -1 and b"body_jntadr" in msg
    rc, msg = create(jnum=(0, 1, 0))  # a joint nobody owns
    assert rc == -1 and b"exactly one" in msg
    rc, msg = create(jtype=(3, 7))
    assert rc == -1 and b"unknown joint type" in msg
    rc, msg = create(jq=(0, 2))  # coordinate address outside qpos
    assert rc == -1 and b"jnt_qposadr" in msg
    rc, msg = create(jtype=(3, 1), jq=(0, 1))  # a ball joint needs four coordinates
    assert rc == -1 and b"jnt_qposadr" in msg


def test_xla_ffi_handlers_type_check_against_a_mock_of_the_ffi_api():
    """csrc/stacb_xla_ffi.cc cannot be built here (no jaxlib headers).  It is compiled (-fsyntax-only) against a mock with the shape
    of xla/ffi/api/ffi.h: every handler must be invocable with exactly the argument list its binding declares, and every C ABI
    call must match include/stacb.h -- so a change of the C ABI cannot silently rot the FFI layer."""
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I", str(ROOT / "tests" / "xla_ffi_mock"), "-I", str(ROOT / "include"),
           "-I", "/usr/local/cuda/include", str(ROOT / "stac_mjx_b200" / "csrc" / "stacb_xla_ffi.cc")]  # fmt: skip
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
