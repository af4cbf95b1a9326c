"""B200-native tile-based immersive-video streaming simulator (MANSY hot path).

Only what the hot path needs lives here: the CUDA kernels + C ABI (``csrc/``), the ctypes binding,
the host-side mirrors of the reference's env / vector-env interface, the table packer and the
synthetic-data generator.  There is no CPU fallback: every compute entry point goes through
``csrc/libmansy_b200.so`` and raises if it is missing.
"""
from .config import (MANSY_OBS_FLOATS, MANSY_OBS_SEGMENTS, MANSY_OBS_STRIDE, OBS_MODE_MANSY, OBS_MODE_NONE,  # noqa: F401
                     OBS_MODE_SIMPLE, REWARD_QOE, REWARD_QOE_NORM, SIMPLE_OBS_FLOATS, SIMPLE_OBS_SEGMENTS,
                     SIMPLE_OBS_STRIDE, SimConfig)
from .tables import SimTables  # noqa: F401

__version__ = "0.1.0"
