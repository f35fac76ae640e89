"""libstacb through its C ABI with NO torch: device memory from cuda-python (cudaMalloc / cudaMemcpy), entry points via ctypes.

Shows that the boundary of include/stacb.h is plain pointers and sizes.  Runs one small rodent clip (root optimisation + pose
optimisation) and prints a checksum;  tests/test_gpu_parity.py::test_c_abi_without_torch compares it with the torch-backed path.
    python tools/abi_demo_no_torch.py
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cuda.bindings import runtime as rt  # noqa: E402

from stac_mjx_b200 import _lib, model, synth  # noqa: E402  (no torch import in these modules)


def ck(ret):
    err, *rest = ret if isinstance(ret, tuple) else (ret,)
    if int(err) != 0:
        raise RuntimeError(f"CUDA error {err}")
    return rest[0] if len(rest) == 1 else rest


def to_dev(a: np.ndarray) -> int:
    a = np.ascontiguousarray(a)
    p = ck(rt.cudaMalloc(max(a.nbytes, 4)))
    if a.nbytes:
        ck(rt.cudaMemcpy(p, a.ctypes.data, a.nbytes, rt.cudaMemcpyKind.cudaMemcpyHostToDevice))
    return int(p)


def from_dev(p: int, shape, dtype) -> np.ndarray:
    out = np.empty(shape, dtype)
    ck(rt.cudaMemcpy(out.ctypes.data, p, out.nbytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost))
    return out


def run(n_clips=2, n_frames=3, seed=3):
    tree, cfg = model.load_fixture("rodent")
    kp_names = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
    s = model.make_setup(tree, cfg.model, kp_names)
    kp, _, _ = synth.synth_session(tree, s, n_clips * n_frames, n_frames, seed=seed)
    L = _lib.lib()
    ck(rt.cudaSetDevice(0))
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    keep = [i32(tree.body_parent), i32(tree.body_jntadr), i32(tree.body_jntnum), f32(tree.body_pos), f32(tree.body_quat), i32(tree.jnt_type),
            i32(tree.jnt_qposadr), i32(tree.jnt_bodyid), f32(tree.jnt_pos), f32(tree.jnt_axis), f32(tree.qpos0), i32(s.site_bodies)]  # fmt: skip
    desc = _lib.TreeDesc(tree.nbody, tree.nq, tree.njnt, len(kp_names), *[a.ctypes.data_as(C.c_void_p) for a in keep])
    h = C.c_void_p()
    _lib.check(L.stacb_tree_create(C.byref(desc), 0, C.byref(h)), "stacb_tree_create")
    nq, nb, K, P = tree.nq, tree.nbody, len(kp_names), s.indiv_parts.shape[0]
    Cn, F = n_clips, n_frames
    d_kp, d_qio = to_dev(kp.reshape(Cn, F, -1)), to_dev(np.tile(f32(tree.qpos0), (Cn, 1)))
    d_off, d_lb, d_ub = to_dev(f32(s.initial_offsets)), to_dev(s.lb), to_dev(s.ub)
    d_pm, d_trunk = to_dev(s.indiv_parts.astype(np.uint8)), to_dev(s.trunk_kps.astype(np.uint8))
    outs = {"qpos": (Cn, F, nq), "xpos": (Cn, F, nb, 3), "xquat": (Cn, F, nb, 4), "sites": (Cn, F, K, 3), "err": (Cn, F)}
    d_out = {k: to_dev(np.zeros(v, np.float32)) for k, v in outs.items()}
    d_it, d_ls = to_dev(np.zeros((Cn, F, 1 + P), np.int32)), to_dev(np.zeros((Cn, F, 1 + P), np.int32))
    d_rs, d_st = to_dev(np.zeros((Cn, 4), np.int32)), to_dev(np.zeros(Cn, np.int32))
    vp = C.c_void_p
    rc = L.stacb_pose_clips(h, vp(d_kp), vp(d_qio), vp(d_off), vp(d_lb), vp(d_ub), vp(d_pm), P, 1, int(s.root_kp_idx), vp(d_trunk), 7,
                            float(cfg.model.FTOL), 400, 15, vp(d_out["qpos"]), vp(d_out["xpos"]), vp(d_out["xquat"]), vp(d_out["sites"]),
                            vp(d_out["err"]), vp(d_it), vp(d_ls), vp(d_rs), vp(d_st), Cn, F, None)  # fmt: skip
    _lib.check(rc, "stacb_pose_clips")
    ck(rt.cudaDeviceSynchronize())
    res = {k: from_dev(d_out[k], v, np.float32) for k, v in outs.items()}
    res["iters"] = from_dev(d_it, (Cn, F, 1 + P), np.int32)
    res["kp"] = kp.reshape(Cn, F, -1)
    for p in [d_kp, d_qio, d_off, d_lb, d_ub, d_pm, d_trunk, d_it, d_ls, d_rs, d_st, *d_out.values()]:
        ck(rt.cudaFree(p))
    L.stacb_tree_destroy(h)
    return res


if __name__ == "__main__":
    assert "torch" not in sys.modules
    r = run()
    assert "torch" not in sys.modules, "this demo must not pull torch in"
    print("iters per frame:", r["iters"].sum(-1).reshape(-1).tolist())
    print("qpos checksum: %.9f" % float(np.abs(r["qpos"]).sum()))
