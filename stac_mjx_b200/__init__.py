"""stac_mjx_b200: B200-native STAC fitting hot path (drop-in for talmolab/stac-mjx's solver path)."""

__version__ = "0.1.0"
