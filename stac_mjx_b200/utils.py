"""Host-side helpers of the STAC path (reference ``stac_mjx/utils.py``).

Only what the hot path and its callers need: clip batching (``utils.py:350-389``),
``make_qs`` (``:129-144``) and the site-offset accessors (``:94-126``).  FK itself
(``kinematics`` / ``replace_qs``) runs on the GPU through `engine.Engine.fk`.
"""

from __future__ import annotations

import numpy as np

CONTINUOUS_BATCH_OVERLAP = 10


def batch_kp_data(kp_data: np.ndarray, n_frames_per_clip: int, continuous: bool = False) -> np.ndarray:
    """Reshape [N, 3K] keypoints into clips [C, F(+10), 3K] (reference ``utils.py:350-389``)."""
    kp_data = np.asarray(kp_data)
    n_frames = n_frames_per_clip
    total_frames = kp_data.shape[0]
    n_batches = int(total_frames // n_frames)
    if continuous:
        window = n_frames + CONTINUOUS_BATCH_OVERLAP
        if total_frames < window:
            return kp_data.reshape((n_batches, window) + kp_data.shape[1:])
        starts = np.arange(0, n_batches * n_frames, n_frames)
        batches = [kp_data[s : s + window] for s in starts]
        batches[-1] = np.pad(batches[-1], ((0, CONTINUOUS_BATCH_OVERLAP), (0, 0)), mode="wrap")
        return np.stack(batches, axis=0)
    out = kp_data[: n_batches * n_frames]
    return out.reshape((n_batches, n_frames) + kp_data.shape[1:])


def make_qs(q0, qs_to_opt, q):
    """``(1 - mask) * q0 + mask * q`` (reference ``utils.py:129-144``) as a select, for numpy or torch operands
    in any mix (bool or 0/1 mask); identical to the arithmetic form for finite values."""
    import torch

    if isinstance(q0, torch.Tensor) or isinstance(q, torch.Tensor):
        ref = q if isinstance(q, torch.Tensor) else q0
        as_t = lambda a, dt: a.to(device=ref.device, dtype=dt) if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a), device=ref.device).to(dt)
        return torch.where(as_t(qs_to_opt, torch.bool), as_t(q, ref.dtype), as_t(q0, ref.dtype))
    mask = np.asarray(qs_to_opt.cpu() if isinstance(qs_to_opt, torch.Tensor) else qs_to_opt).astype(bool)
    return np.where(mask, np.asarray(q), np.asarray(q0))


def get_site_xpos(data, site_idxs=None):
    """World positions of the keypoint sites (``utils.py:77-91``); `data.site_xpos` is already [K, 3]."""
    return data.site_xpos


def get_site_pos(model, site_idxs=None):
    return model.site_pos


def set_site_pos(model, offsets, site_idxs=None):
    return model.replace(site_pos=offsets)


def edge_crossfade_host(data, n_frames_per_clip: int):
    """One packed array through the reference's cross-fade (``utils.py:436-453``), on the host."""
    ov = CONTINUOUS_BATCH_OVERLAP

    def crossfade(a, b, center=0.5, steepness=10.0):
        n = a.shape[0]
        x = np.linspace(0.0, 1.0, n)
        m = 0.5 * (1.0 + np.tanh(steepness * (x - center) / 2.0))
        m = m.reshape((n,) + (1,) * (a.ndim - 1))
        return (1.0 - m) * a + m * b

    data = np.array(data)
    b = data.reshape((-1, n_frames_per_clip + ov) + data.shape[1:])
    for i in range(b.shape[0] - 1):
        b[i, -ov:] = crossfade(b[i, -ov:], b[i + 1, :ov])
    first, middle, last = b[0], b[1:-1, ov:], b[-1, ov:-ov]
    return np.concatenate([first, middle.reshape((-1,) + middle.shape[2:]), last], axis=0)


def handle_edge_effects(ik_only_data, n_frames_per_clip: int):
    """Sigmoid cross-fade of overlapping clip boundaries (reference ``utils.py:393-461``), host version: same signature and
    semantics as the reference; ``Stac.ik_only(..., edge_effects=True)`` does the same on the GPU before the results leave it."""
    for k in ("qpos", "kp_data", "xpos", "xquat", "marker_sites"):
        setattr(ik_only_data, k, edge_crossfade_host(getattr(ik_only_data, k), n_frames_per_clip))
    return ik_only_data


def _quat_mul(a, b):
    return np.stack(
        [
            a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3],
            a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2],
            a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1],
            a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0],
        ],
        axis=-1,
    )


def compute_velocity_from_kinematics(qpos_trajectory, dt: float, freejoint: bool = True, max_qvel: float = 20.0):
    """Finite-difference qvel of one continuous clip (reference ``utils.py:302-347``), vectorised over frames.

    The last frame is repeated (zero velocity there); with a free joint the angular velocity is the axis-angle of
    ``conj(q_t) * q_{t+1}`` divided by dt and only the joint velocities are clipped to ``max_qvel``.
    """
    q = np.asarray(qpos_trajectory, dtype=np.float32)
    q = np.concatenate([q, q[-1:]], axis=0)
    if not freejoint:
        return np.clip((q[1:] - q[:-1]) / dt, -max_qvel, max_qvel)
    joints = (q[1:, 7:] - q[:-1, 7:]) / dt
    trans = (q[1:, :3] - q[:-1, :3]) / dt
    conj = q[:-1, 3:7] * np.array([1.0, -1.0, -1.0, -1.0], dtype=np.float32)
    diff = _quat_mul(conj, q[1:, 3:7])
    diff = diff / np.linalg.norm(diff, axis=-1, keepdims=True)
    angle = 2.0 * np.arccos(np.clip(diff[:, 0], -1.0, 1.0))
    half_sin = np.sin(angle / 2.0)
    wrapped = (angle + np.pi) % (2.0 * np.pi) - np.pi
    with np.errstate(divide="ignore", invalid="ignore"):
        gyro = np.where((angle < 1e-10)[:, None], 0.0, diff[:, 1:4] / half_sin[:, None] * wrapped[:, None]) / dt
    return np.concatenate([trans, gyro, np.clip(joints, -max_qvel, max_qvel)], axis=1).astype(np.float32)
