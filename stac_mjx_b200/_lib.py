"""ctypes binding of libstacb.so (include/stacb.h). No fallback: a missing library is an error."""

from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libstacb.so"

# every symbol include/stacb.h declares
SYMBOLS = (
    "stacb_tree_create",
    "stacb_tree_destroy",
    "stacb_tree_smem_per_chain",
    "stacb_fk",
    "stacb_loss_grad",
    "stacb_q_opt",
    "stacb_pose_clips",
    "stacb_pose_session",
    "stacb_m_stats",
    "stacb_m_residual",
    "stacb_m_scratch_floats",
    "stacb_edge_rows",
    "stacb_edge_crossfade",
    "stacb_qvel",
    "stacb_fma_peak",
    "stacb_tree_set_mode",
    "stacb_tree_set_path",
    "stacb_tree_path",
    "stacb_last_error",
    "stacb_version",
)


class StacbError(RuntimeError):
    pass


class TreeDesc(C.Structure):
    _fields_ = [
        ("nbody", C.c_int32),
        ("nq", C.c_int32),
        ("njnt", C.c_int32),
        ("nsite", C.c_int32),
        ("body_parent", C.c_void_p),
        ("body_jntadr", C.c_void_p),
        ("body_jntnum", C.c_void_p),
        ("body_pos", C.c_void_p),
        ("body_quat", C.c_void_p),
        ("jnt_type", C.c_void_p),
        ("jnt_qposadr", C.c_void_p),
        ("jnt_bodyid", C.c_void_p),
        ("jnt_pos", C.c_void_p),
        ("jnt_axis", C.c_void_p),
        ("qpos0", C.c_void_p),
        ("site_body", C.c_void_p),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load libstacb.so once; raise loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise StacbError(
                f"{LIB_PATH} is missing: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` or stac_mjx_b200/csrc/build.sh. "
                "There is no CPU fallback."
            )
        L = C.CDLL(str(LIB_PATH))
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        L.stacb_tree_create.argtypes = [C.POINTER(TreeDesc), i32, C.POINTER(vp)]
        L.stacb_tree_destroy.argtypes = [vp]
        L.stacb_tree_destroy.restype = None
        L.stacb_tree_smem_per_chain.argtypes = [vp]
        L.stacb_fk.argtypes = [vp] * 7 + [i32, vp]
        L.stacb_loss_grad.argtypes = [vp] * 9 + [i32, vp]
        L.stacb_q_opt.argtypes = [vp] * 8 + [f32, i32, i32] + [vp] * 4 + [i32, vp]
        L.stacb_pose_clips.argtypes = (
            [vp] * 7 + [i32, i32, i32, vp, i32, f32, i32, i32] + [vp] * 9 + [i32, i32, vp]
        )
        L.stacb_pose_session.argtypes = (
            [vp, vp, i32] + [vp] * 5 + [i32, i32, i32, vp, i32, f32, i32, i32] + [vp] * 9 + [i32, i32, vp]
        )
        L.stacb_m_stats.argtypes = [vp] * 5 + [i32, vp]
        L.stacb_m_residual.argtypes = [vp] * 6 + [i32, vp]
        L.stacb_m_scratch_floats.argtypes = [vp, i32]
        L.stacb_edge_rows.argtypes = [i32, i32, i32]
        L.stacb_edge_rows.restype = C.c_longlong
        L.stacb_edge_crossfade.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
        L.stacb_qvel.argtypes = [vp, vp, i32, i32, i32, i32, f32, f32, vp]
        L.stacb_fma_peak.argtypes = [vp, i32, i32, i32, vp]
        L.stacb_tree_set_mode.argtypes = [vp, i32]
        L.stacb_tree_set_path.argtypes = [vp, i32]
        L.stacb_tree_path.argtypes = [vp]
        L.stacb_last_error.restype = C.c_char_p
        L.stacb_version.restype = i32
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().stacb_last_error().decode(errors="replace")
        raise StacbError(f"{what} failed (code {rc}): {msg}")
