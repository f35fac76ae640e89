"""Algorithmic work of the STAC hot path: the single source of the FLOP formulas.

Counted on the ACTIVE subtree in the reference's operation sequence (FMA = 2, sin/cos/sqrt/div = 1; primitive
costs rotate 40, quat_mul 28, axis_angle_to_quat 7, normalize4 13); speculative or skipped work is not counted.

    F_fwd = 71 (nb_act - 1) + 161 nh_act + 13 n_free + 13 nb_act + 43 K + 12 K      FK + sites + masked SSE
    F_bwd = 21 K + 6 (nb_act - 1) + 17 nh_act + 70 n_free                            analytic reverse sweep
    VG    = F_fwd + F_bwd
    flops = sum_solves [ iters (2 VG + 20 nq) + ls_evals F_fwd ] + frames (1 + P) F_fwd          ("reference sequence")

The kernels execute less than that sequence: the gradient at the accepted line-search point reuses the forward pass of the
accepted candidate (the reference runs a second full value_and_grad there), and the 13 nb_act term (a per-body quaternion
normalisation the survey charged to MJX's kinematics) is performed by neither the kernels nor the oracle:
    F_fwd' = F_fwd - 13 nb_act
    executed = sum_solves [ iters (F_fwd' + 2 F_bwd + 20 nq) + ls_evals F_fwd' ] + frames F_fk_full                ("executed")
Both are reported by bench.py; speculative evaluations of the latency modes are counted in neither.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .mjcf import JNT_FREE, JNT_HINGE, JNT_SLIDE
from .tree import TreeModel


@dataclass
class PathCost:
    nb_act: int
    nh_act: int
    n_free: int
    K: int
    nq: int
    f_fwd: int
    f_bwd: int
    f_fk_full: int = 0

    @property
    def vg(self) -> int:
        return self.f_fwd + self.f_bwd

    def per_iteration(self, n_ls: float = 1.0) -> float:
        return 2 * self.vg + 20 * self.nq + n_ls * self.f_fwd

    def total(self, iters: int, ls_evals: int, n_frames: int, n_stages: int) -> float:
        """The reference's operation sequence."""
        return iters * (2 * self.vg + 20 * self.nq) + ls_evals * self.f_fwd + n_frames * n_stages * self.f_fwd

    def executed(self, iters: int, ls_evals: int, n_frames: int, n_stages: int) -> float:
        """What the sequential semantics of the kernels need: FK of the accepted candidate reused for its gradient, no
        per-body quaternion normalisation, one full-model FK per frame for the outputs."""
        f_fwd = self.f_fwd - 13 * self.nb_act
        return iters * (f_fwd + 2 * self.f_bwd + 20 * self.nq) + ls_evals * f_fwd + n_frames * self.f_fk_full


def path_cost(tree: TreeModel, site_bodies) -> PathCost:
    act = tree.active_bodies(site_bodies)
    on_act = np.isin(tree.jnt_bodyid, act)
    nh = int(np.sum(on_act & ((tree.jnt_type == JNT_HINGE) | (tree.jnt_type == JNT_SLIDE))))
    nf = int(np.sum(on_act & (tree.jnt_type == JNT_FREE)))
    nb, K = len(act), len(site_bodies)
    f_fwd = 71 * (nb - 1) + 161 * nh + 13 * nf + 13 * nb + 43 * K + 12 * K
    f_bwd = 21 * K + 6 * (nb - 1) + 17 * nh + 70 * nf
    all_h = int(np.sum((tree.jnt_type == JNT_HINGE) | (tree.jnt_type == JNT_SLIDE)))
    f_fk_full = 71 * (tree.nbody - 2) + 161 * all_h + 13 * int(np.sum(tree.jnt_type == JNT_FREE)) + 43 * K
    return PathCost(nb, nh, nf, K, tree.nq, f_fwd, f_bwd, f_fk_full)


def hbm_bytes_per_frame(tree: TreeModel, K: int, n_stages: int) -> int:
    """Algorithmic HBM traffic of one frame: keypoints in; qpos, xpos, xquat, sites, err and solver counters out."""
    return 4 * (3 * K) + 4 * (tree.nq + 3 * tree.nbody + 4 * tree.nbody + 3 * K + 1) + 8 * n_stages
