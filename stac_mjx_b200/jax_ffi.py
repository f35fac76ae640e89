"""JAX front-end of libstacb: the C ABI registered as XLA FFI custom calls (``north_star``).

Only usable where ``jax`` (>= 0.4.31, ``jax.ffi``) is installed and ``csrc/build.sh`` found jaxlib's FFI headers and built
``libstacb_xla_ffi.so``.  Neither is true in the authoring image, so this module is import-guarded and has NOT been executed
there; it contains no arithmetic -- it only declares output shapes and forwards arrays to the handlers of
``csrc/stacb_xla_ffi.cc``.  The torch/ctypes path (`engine.Engine`) is what the tests and the benchmark exercise.
"""

from __future__ import annotations

import ctypes
from pathlib import Path

import numpy as np

FFI_LIB_PATH = Path(__file__).resolve().parent / "libstacb_xla_ffi.so"
_registered = False


def available() -> bool:
    try:
        import jax.ffi  # noqa: F401
    except Exception:
        return False
    return FFI_LIB_PATH.exists()


def register() -> None:
    """``jax.ffi.register_ffi_target`` for every handler; idempotent."""
    global _registered
    if _registered:
        return
    if not available():
        raise RuntimeError("JAX FFI path unavailable: needs jax.ffi and libstacb_xla_ffi.so (see csrc/build.sh)")
    import jax

    lib = ctypes.CDLL(str(FFI_LIB_PATH))
    for name, sym in (("stacb_pose_clips", "StacbPoseClips"), ("stacb_q_opt", "StacbQOpt"), ("stacb_fk", "StacbFk"), ("stacb_m_stats", "StacbMStats")):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, sym)), platform="CUDA")
    _registered = True


def pose_clips(tree_handle: int, nbody: int, kp, qpos_init, site_pos, lb, ub, part_masks, trunk_kps, *, do_root: int, root_kp_idx: int,
               root_dims: int = 7, tol: float = 1e-4, maxiter: int = 400, maxls: int = 15):  # fmt: skip
    """Fused root + pose optimisation over clips on JAX arrays (jit-compatible). `tree_handle` = int(stacb_tree*)."""
    import jax
    import jax.numpy as jnp

    register()
    C, F, k3 = kp.shape
    nq, K, P = qpos_init.shape[-1], k3 // 3, part_masks.shape[0]
    f32, i32 = jnp.float32, jnp.int32
    out_types = [
        jax.ShapeDtypeStruct((C, F, nq), f32), jax.ShapeDtypeStruct((C, F, nbody, 3), f32), jax.ShapeDtypeStruct((C, F, nbody, 4), f32),
        jax.ShapeDtypeStruct((C, F, K, 3), f32), jax.ShapeDtypeStruct((C, F), f32), jax.ShapeDtypeStruct((C, nq), f32),
        jax.ShapeDtypeStruct((C, F, 1 + P), i32), jax.ShapeDtypeStruct((C, F, 1 + P), i32), jax.ShapeDtypeStruct((C, 4), i32),
        jax.ShapeDtypeStruct((C,), i32),
    ]  # fmt: skip
    call = jax.ffi.ffi_call("stacb_pose_clips", out_types)
    qpos, xpos, xquat, sites, err, qpos_last, iters, ls_evals, root_stats, status = call(
        kp.astype(f32), qpos_init.astype(f32), site_pos.astype(f32), lb.astype(f32), ub.astype(f32), part_masks.astype(jnp.uint8),
        trunk_kps.astype(jnp.uint8), tree=np.int64(tree_handle), do_root=np.int32(do_root), root_kp_idx=np.int32(root_kp_idx),
        root_dims=np.int32(root_dims), tol=np.float32(tol), maxiter=np.int32(maxiter), maxls=np.int32(maxls),
    )  # fmt: skip
    return dict(qpos=qpos, xpos=xpos, xquat=xquat, sites=sites, err=err, qpos_last=qpos_last, iters=iters, ls_evals=ls_evals,
                root_stats=root_stats, status=status)  # fmt: skip


def q_opt(tree_handle: int, q0, kp, q_mask, kp_mask, site_pos, lb, ub, *, tol: float, maxiter: int = 400, maxls: int = 15):
    """Batch of independent ``StacCore.q_opt`` solves on JAX arrays."""
    import jax
    import jax.numpy as jnp

    register()
    B, nq = q0.shape
    out_types = [jax.ShapeDtypeStruct((B, nq), jnp.float32), jax.ShapeDtypeStruct((B,), jnp.float32),
                 jax.ShapeDtypeStruct((B,), jnp.int32), jax.ShapeDtypeStruct((B,), jnp.int32)]  # fmt: skip
    return jax.ffi.ffi_call("stacb_q_opt", out_types)(
        q0, kp, q_mask.astype(jnp.uint8), kp_mask.astype(jnp.uint8), site_pos, lb, ub, tree=np.int64(tree_handle), tol=np.float32(tol),
        maxiter=np.int32(maxiter), maxls=np.int32(maxls),
    )  # fmt: skip
