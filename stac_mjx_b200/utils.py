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
    """``(1 - mask) * q0 + mask * q`` (reference ``utils.py:129-144``); numpy or torch."""
    return (1 - qs_to_opt) * q0 + qs_to_opt * q


def get_site_xpos(data, site_idxs=None):
    """World positions of the keypoint sites (``utils.py:77-91``); `data.site_xpos` is already [K, 3]."""
    return data.site_xpos


def get_site_pos(model, site_idxs=None):
    return model.site_pos


def set_site_pos(model, offsets, site_idxs=None):
    return model.replace(site_pos=offsets)


def handle_edge_effects(ik_only_data, n_frames_per_clip: int):
    """Sigmoid cross-fade of overlapping clip boundaries (reference ``utils.py:393-461``)."""
    ov = CONTINUOUS_BATCH_OVERLAP

    def crossfade(a, b, center=0.5, steepness=10.0):
        n = a.shape[0]
        x = np.linspace(0.0, 1.0, n)
        m = 0.5 * (1.0 + np.tanh(steepness * (x - center) / 2.0))
        m = m.reshape((n,) + (1,) * (a.ndim - 1))
        return (1.0 - m) * a + m * b

    def f(data):
        data = np.array(data)
        b = data.reshape((-1, n_frames_per_clip + ov) + data.shape[1:])
        for i in range(b.shape[0] - 1):
            b[i, -ov:] = crossfade(b[i, -ov:], b[i + 1, :ov])
        first, middle, last = b[0], b[1:-1, ov:], b[-1, ov:-ov]
        return np.concatenate([first, middle.reshape((-1,) + middle.shape[2:]), last], axis=0)

    for k in ("qpos", "kp_data", "xpos", "xquat", "marker_sites"):
        setattr(ik_only_data, k, f(getattr(ik_only_data, k)))
    return ik_only_data
