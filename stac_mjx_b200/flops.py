"""Algorithmic work of the STAC hot path: the single source of the FLOP formulas.

Counted on the ACTIVE subtree in the reference's operation sequence (FMA = 2, sin/cos/sqrt/div = 1; primitive
costs rotate 40, quat_mul 28, axis_angle_to_quat 7, normalize4 13); speculative or skipped work is not counted.

    F_fwd = 71 (nb_act - 1) + 161 nh_act + 13 n_free + 13 nb_act + 43 K + 12 K      FK + sites + masked SSE
    F_bwd = 21 K + 6 (nb_act - 1) + 17 nh_act + 70 n_free                            analytic reverse sweep
    VG    = F_fwd + F_bwd
    flops = sum_solves [ iters (2 VG + 20 nq) + ls_evals F_fwd ] + frames (1 + P) F_fwd
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

    @property
    def vg(self) -> int:
        return self.f_fwd + self.f_bwd

    def per_iteration(self, n_ls: float = 1.0) -> float:
        return 2 * self.vg + 20 * self.nq + n_ls * self.f_fwd

    def total(self, iters: int, ls_evals: int, n_frames: int, n_stages: int) -> float:
        return iters * (2 * self.vg + 20 * self.nq) + ls_evals * self.f_fwd + n_frames * n_stages * self.f_fwd


def path_cost(tree: TreeModel, site_bodies) -> PathCost:
    act = tree.active_bodies(site_bodies)
    on_act = np.isin(tree.jnt_bodyid, act)
    nh = int(np.sum(on_act & ((tree.jnt_type == JNT_HINGE) | (tree.jnt_type == JNT_SLIDE))))
    nf = int(np.sum(on_act & (tree.jnt_type == JNT_FREE)))
    nb, K = len(act), len(site_bodies)
    f_fwd = 71 * (nb - 1) + 161 * nh + 13 * nf + 13 * nb + 43 * K + 12 * K
    f_bwd = 21 * K + 6 * (nb - 1) + 17 * nh + 70 * nf
    return PathCost(nb, nh, nf, K, tree.nq, f_fwd, f_bwd)


def hbm_bytes_per_frame(tree: TreeModel, K: int, n_stages: int) -> int:
    """Algorithmic HBM traffic of one frame: keypoints in; qpos, xpos, xquat, sites, err and solver counters out."""
    return 4 * (3 * K) + 4 * (tree.nq + 3 * tree.nbody + 4 * tree.nbody + 3 * K + 1) + 8 * n_stages
