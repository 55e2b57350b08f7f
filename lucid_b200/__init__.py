"""lucid_b200 -- B200-native exact-OIT rasteriser behind the LucidRenderer interface.

The package holds only what the hot path needs: csrc/ (sm_100a kernels + the C ABI), host/
(C++ input preparation and the C++ LucidRenderer facade), a ctypes binding (api) and the synthetic
scene generators used by tests and bench.py.
"""
from .api import (LucidConfig, LucidError, LucidRenderer, build_instances, decode_stats, load_library,  # noqa: F401
                  make_camera, make_config, prepare_frame, split_info, verify_info)
