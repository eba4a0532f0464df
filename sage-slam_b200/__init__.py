"""sage-slam_b200: B200-native dense bundle-adjustment backend for SAGE-SLAM's factor hot path.

Layout: csrc/ (sm_100a CUDA kernels + the C ABI of include/sage_ba.h), capi.py (ctypes binding),
ops.py (host mirror of the reference df::*_calculate operator API), local_ba.py (batched LM +
multi-GPU plumbing), mapper.py (MappingStep / UpdateMap adapter), frames.py / synthetic.py (data contract and synthetic inputs).
Importing the package never touches CUDA; creating a Context does, and fails loudly without a GPU.
"""
from . import capi, factors, frames, synthetic  # noqa: F401
from .frames import Keyframe  # noqa: F401
from .local_ba import LocalBA  # noqa: F401
from .mapper import BatchedMapper, Map, MapperOptions  # noqa: F401
from .ops import Context, DeviceKeyframe, SageError  # noqa: F401

__all__ = ["capi", "frames", "synthetic", "Keyframe", "LocalBA", "BatchedMapper", "Map", "MapperOptions", "Context", "DeviceKeyframe", "SageError"]
