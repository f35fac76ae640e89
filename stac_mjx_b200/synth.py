"""Deterministic synthetic mocap sessions.

Used by bench.py and the tests to make keypoint clips of a model's shape
without any dataset: a smooth random qpos trajectory per clip is pushed through
a batched NumPy forward kinematics (float64, vectorised over frames) with
perturbed marker offsets, and observation noise is added.  This is input
generation only; it is not on the measured path.
"""

from __future__ import annotations

import numpy as np

from .mjcf import JNT_BALL, JNT_FREE, JNT_HINGE
from .tree import TreeModel


def _rotate(v, q):
    s, u = q[..., :1], q[..., 1:]
    return 2 * np.sum(u * v, -1, keepdims=True) * u + (s * s - np.sum(u * u, -1, keepdims=True)) * v + 2 * s * np.cross(u, v)


def _qmul(a, b):
    return np.stack(
        [
            a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3],
            a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2],
            a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1],
            a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0],
        ],
        -1,
    )


def batched_fk(tree: TreeModel, qpos: np.ndarray):
    """World pose of every body for qpos[N, nq] (float64). Returns xpos[N,nbody,3], xquat[N,nbody,4]."""
    N = qpos.shape[0]
    xpos = np.zeros((N, tree.nbody, 3))
    xquat = np.zeros((N, tree.nbody, 4))
    xquat[:, 0, 0] = 1.0
    for b in range(1, tree.nbody):
        p = int(tree.body_parent[b])
        pos = xpos[:, p] + _rotate(np.broadcast_to(tree.body_pos[b], (N, 3)), xquat[:, p])
        quat = _qmul(xquat[:, p], np.broadcast_to(tree.body_quat[b], (N, 4)))
        for jj in range(int(tree.body_jntnum[b])):
            j = int(tree.body_jntadr[b]) + jj
            adr, typ = int(tree.jnt_qposadr[j]), int(tree.jnt_type[j])
            jpos = np.broadcast_to(tree.jnt_pos[j], (N, 3))
            if typ == JNT_FREE:
                pos = qpos[:, adr : adr + 3]
                quat = qpos[:, adr + 3 : adr + 7] / np.linalg.norm(qpos[:, adr + 3 : adr + 7], axis=-1, keepdims=True)
                continue
            anchor = _rotate(jpos, quat) + pos
            if typ == JNT_HINGE:
                half = 0.5 * (qpos[:, adr] - tree.qpos0[adr])
                ql = np.concatenate([np.cos(half)[:, None], tree.jnt_axis[j][None] * np.sin(half)[:, None]], -1)
            elif typ == JNT_BALL:
                ql = qpos[:, adr : adr + 4] / np.linalg.norm(qpos[:, adr : adr + 4], axis=-1, keepdims=True)
            else:
                pos = pos + _rotate(np.broadcast_to(tree.jnt_axis[j], (N, 3)), quat) * (qpos[:, adr] - tree.qpos0[adr])[:, None]
                continue
            quat = _qmul(quat, ql)
            pos = anchor - _rotate(jpos, quat)
        xpos[:, b], xquat[:, b] = pos, quat
    return xpos, xquat


def site_positions(tree: TreeModel, site_bodies, offsets, qpos):
    xpos, xquat = batched_fk(tree, qpos)
    sb = np.asarray(site_bodies)
    N = qpos.shape[0]
    return xpos[:, sb] + _rotate(np.broadcast_to(offsets, (N,) + offsets.shape), xquat[:, sb])


def _random_unit_quat_walk(rng, n, sigma_rad):
    q = np.zeros((n, 4))
    q0 = rng.normal(size=4)
    q0 = np.array([1.0, 0, 0, 0]) + 0.15 * q0
    q[0] = q0 / np.linalg.norm(q0)
    for i in range(1, n):
        w = rng.normal(scale=sigma_rad, size=3)
        dq = np.concatenate([[1.0], 0.5 * w])
        q[i] = _qmul(q[i - 1], dq / np.linalg.norm(dq))
        q[i] /= np.linalg.norm(q[i])
    return q


def synth_trajectory(tree: TreeModel, lb, ub, n_frames: int, rng, center=(0.34, 0.04, 0.04)):
    """Smooth qpos[n_frames, nq] inside the joint limits (root random walk + band-limited hinges)."""
    q = np.tile(tree.qpos0, (n_frames, 1))
    tt = np.arange(n_frames)
    for j in range(tree.njnt):
        adr, typ = int(tree.jnt_qposadr[j]), int(tree.jnt_type[j])
        if typ == JNT_FREE:
            q[:, adr : adr + 3] = np.asarray(center) + np.cumsum(rng.normal(scale=0.15e-3, size=(n_frames, 3)), 0)
            q[:, adr + 3 : adr + 7] = _random_unit_quat_walk(rng, n_frames, np.deg2rad(0.5))
        elif typ == JNT_BALL:
            q[:, adr : adr + 4] = _random_unit_quat_walk(rng, n_frames, np.deg2rad(0.5))
        else:
            lo, hi = float(lb[adr]), float(ub[adr])
            if not np.isfinite(lo) or not np.isfinite(hi):
                lo, hi = -0.05, 0.05
            lo, hi = lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo)
            sig = np.zeros(n_frames)
            for _ in range(3):
                period = rng.uniform(25, 250)
                sig += rng.uniform(0.2, 1.0) * np.sin(2 * np.pi * tt / period + rng.uniform(0, 2 * np.pi))
            sig = sig / 3.0  # in [-1, 1]
            q[:, adr] = 0.5 * (lo + hi) + 0.5 * (hi - lo) * 0.6 * sig
    return q


def synth_session(tree: TreeModel, setup, n_frames: int, n_frames_per_clip: int, seed: int = 20260101,
                  offset_sigma: float = 2e-3, obs_sigma: float = 1e-3):  # fmt: skip
    """Keypoints [n_frames, 3K] float32 (+ ground truth) for a session cut into independent clips."""
    rng = np.random.default_rng(seed)
    offsets_true = setup.initial_offsets.astype(np.float64) + rng.normal(scale=offset_sigma, size=(len(setup.site_idxs), 3))
    n_clips = max(1, n_frames // n_frames_per_clip)
    qs = []
    for _ in range(n_clips):
        qs.append(synth_trajectory(tree, setup.lb, setup.ub, n_frames_per_clip, rng))
    q = np.concatenate(qs, 0)[:n_frames]
    if q.shape[0] < n_frames:
        q = np.concatenate([q, synth_trajectory(tree, setup.lb, setup.ub, n_frames - q.shape[0], rng)], 0)
    sites = site_positions(tree, setup.site_bodies, offsets_true, q)
    kp = sites + rng.normal(scale=obs_sigma, size=sites.shape)
    return kp.reshape(n_frames, -1).astype(np.float32), q, offsets_true
