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
    try:
        from omegaconf import OmegaConf

        return OmegaConf.to_yaml(config)
    except ImportError:
        return yaml.safe_dump(config.to_dict() if hasattr(config, "to_dict") else dict(config), sort_keys=False)


def save_data_to_h5(config, kp_names, names_qpos, names_xpos, kp_data, marker_sites, offsets, qpos, xpos, xquat, qvel, file_path):
    """Write config + STAC data with the reference's dataset layout (``io.py:194-237``)."""
    import h5py

    with h5py.File(file_path, "w") as f:
        f.create_dataset("config", data=np.bytes_(_config_yaml(config)))
        f.create_dataset("kp_names", data=np.array(kp_names, dtype="S"))
        f.create_dataset("names_qpos", data=np.array(names_qpos, dtype="S"))
        f.create_dataset("names_xpos", data=np.array(names_xpos, dtype="S"))
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
