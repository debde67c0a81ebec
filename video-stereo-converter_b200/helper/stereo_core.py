"""`helper.stereo_core` — same import path and `__all__` as the reference module
(/root/reference/helper/stereo_core.py:22-29), implemented by vsc_b200 (CUDA, sm_100a)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from vsc_b200.stereo_core import (StereoGenerator, StereoParams, apply_depth_gamma, forward_warp_stereo,  # noqa: E402,F401
                                  load_image_pair, normalize_depth)

__all__ = [
    'load_image_pair',
    'normalize_depth',
    'apply_depth_gamma',
    'forward_warp_stereo',
    'StereoParams',
    'StereoGenerator'
]
