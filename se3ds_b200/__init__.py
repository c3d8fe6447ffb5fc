"""se3ds_b200 -- B200-native geometric guidance path of SE3DS.

Drop-in for the reference's `utils/pano_utils.py`, `utils/point_cloud_utils.py` and
`inference/perturbation_utils.py` hot-path functions (same names, argument order and error
behaviour), backed by hand-written sm_100a CUDA behind a C ABI (include/se3ds_geom.h).
Nothing here falls back to the CPU.
"""
from . import constants  # noqa: F401

__version__ = '0.1.0'
