"""STAC output container and HDF5 layout (reference ``stac_mjx/io.py:17-36,194-278``).

The dataset names, dtypes and gzip compression of the reference's ``.h5`` files are kept so files
written here are readable by the reference and vice versa.  ``h5py`` / ``omegaconf`` are imported
lazily (neither is installed in the authoring image).
"""

from __future__ import annotations

from dataclasses import asdict, dataclass, field
from pathlib import Path

import numpy as np
import yaml


@dataclass
class StacData:
    """Data structure for STAC output (reference ``io.py:17-36``)."""

    qpos: np.ndarray
    xpos: np.ndarray
    xquat: np.ndarray
    marker_sites: np.ndarray
    offsets: np.ndarray
    kp_data: np.ndarray
    names_qpos: list
    names_xpos: list
    kp_names: list
    qvel: np.ndarray = field(default_factory=lambda: np.array([]))

    def as_dict(self) -> dict:
        return asdict(self)


def _config_yaml(config) -> str:
    plain = config.to_dict() if hasattr(config, "to_dict") else config
    try:
        from omegaconf import DictConfig, OmegaConf
    except ImportError:
        return yaml.safe_dump(dict(plain), sort_keys=False)
    # this package's Cfg is a dict subclass, which OmegaConf does not accept as a container: go through a plain dict
    return OmegaConf.to_yaml(plain if isinstance(plain, DictConfig) else OmegaConf.create(dict(plain)))


def save_data_to_h5(config, kp_names, names_qpos, names_xpos, kp_data, marker_sites, offsets, qpos, xpos, xquat, qvel, file_path):
    """Write config + STAC data with the reference's dataset layout (``io.py:194-237``)."""
    import h5py

    with h5py.File(file_path, "w") as f:
        f.create_dataset("config", data=np.bytes_(_config_yaml(config)))
        f.create_dataset("kp_names", data=np.array(kp_names, dtype="S"))
        f.create_dataset("names_qpos", data=np.array(names_qpos, dtype="S"))
        f.create_dataset("names_xpos", data=np.array(names_xpos, dtype="S"))
        # the arrays may be views of page-locked result buffers (Stac.ik_only hands them out without a second host copy):
        # h5py reads them in place
        for name, arr in (("kp_data", kp_data), ("marker_sites", marker_sites), ("offsets", offsets), ("qpos", qpos),
                          ("qvel", qvel), ("xpos", xpos), ("xquat", xquat)):  # fmt: skip
            f.create_dataset(name, data=arr, compression="gzip")


def load_stac_data(file_path: str | Path):
    """Read a STAC ``.h5`` back into (config, StacData) (``io.py:240-278``)."""
    import h5py

    from .config import Cfg

    with h5py.File(file_path, "r") as f:
        config = Cfg(yaml.safe_load(f["config"][()].decode("utf-8")))
        names = lambda k: [n.decode("utf-8") for n in f[k]]
        data = StacData(
            kp_names=names("kp_names"), names_qpos=names("names_qpos"), names_xpos=names("names_xpos"),
            kp_data=f["kp_data"][()], marker_sites=f["marker_sites"][()], offsets=f["offsets"][()],
            qpos=f["qpos"][()], qvel=f["qvel"][()], xpos=f["xpos"][()], xquat=f["xquat"][()],
        )  # fmt: skip
    return config, data


def load_data(cfg, base_path: Path | None = None):
    """Load mocap, order keypoints like ``KEYPOINT_MODEL_PAIRS``, scale, flatten (reference ``io.py:39-98``).

    Returns ``(kp_data [n_frames, 3K] float32, sorted keypoint names)``; the flattened layout is keypoint-major,
    xyz-minor -- the wire layout of every kernel entry point.
    """
    base_path = Path.cwd() if base_path is None else Path(base_path)
    file_path = base_path / cfg.stac.data_path
    if file_path.suffix == ".mat":
        data, kp_names = load_dannce(str(file_path), names_filename=cfg.model.get("KP_NAMES_LABEL3D_PATH", None))
    elif file_path.suffix == ".nwb":
        data, kp_names = load_nwb(file_path)
    elif file_path.suffix == ".h5":
        data, kp_names = load_h5(file_path)
    else:
        raise ValueError("Unsupported file extension. Please provide a .mat, .nwb, or .h5 file.")
    kp_names = kp_names or cfg.model.get("KP_NAMES", None)
    if kp_names is None:
        raise ValueError(
            "Keypoint names not provided. Please provide an ordered list of keypoint names corresponding to the keypoint data order."
        )
    if len(kp_names) != data.shape[2]:
        raise ValueError(
            f"Number of keypoint names ({len(kp_names)}) is not the same as the number of keypoints in data ({data.shape[2]})"
        )
    model_inds = [list(kp_names).index(src) for src in cfg.model.KEYPOINT_MODEL_PAIRS.keys()]
    sorted_names = [kp_names[i] for i in model_inds]
    data = (np.asarray(data) * cfg.model.MOCAP_SCALE_FACTOR)[:, :, model_inds]
    data = np.transpose(data, (0, 2, 1)).reshape(data.shape[0], -1)
    return np.ascontiguousarray(data, dtype=np.float32), sorted_names


def load_dannce(filename, names_filename=None):
    """DANNCE ``.mat``: ``pred`` [frames, xyz, keypoints] in millimetres (reference ``io.py:101-124``)."""
    import scipy.io as spio

    names = None
    if names_filename is not None:
        mat = spio.loadmat(names_filename)
        names = [item[0] for sub in mat["joint_names"] for item in sub]
    return np.asarray(spio.loadmat(filename)["pred"]), names


def load_nwb(filename):
    """NWB pose estimation: [frames, xyz, keypoints] + node names (reference ``io.py:127-148``)."""
    from pynwb import NWBHDF5IO

    with NWBHDF5IO(filename, mode="r", load_namespaces=True) as f:
        pose = f.read().processing["behavior"]["PoseEstimation"]
        names = pose.nodes[:].tolist()
        data = np.stack([pose[n].data[:] for n in names], axis=-1)
    return data, names


def load_h5(filename):
    """``tracks`` dataset [frames, 1, keypoints, xyz] -> [frames, xyz, keypoints] (reference ``io.py:151-171``)."""
    import h5py

    with h5py.File(filename, "r") as f:
        data = np.array(f["tracks"][()])
    return np.transpose(np.squeeze(data, axis=1), (0, 2, 1)), None
