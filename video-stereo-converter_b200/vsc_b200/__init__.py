"""vsc_b200 — Python host side of the B200-native SBS hot path.

`vsc_b200.stereo_core` mirrors the reference's `helper/stereo_core.py` call surface
(StereoParams, StereoGenerator.process_frame, ...) on top of the C ABI in
include/vsc_b200.h (ctypes, see `_lib.py`).  There is no CPU fallback: importing
`_lib` without the built CUDA library raises.
"""
from .stereo_core import (StereoGenerator, StereoParams, apply_depth_gamma, forward_warp_stereo,  # noqa: F401
                          load_image_pair, normalize_depth)

__all__ = ['load_image_pair', 'normalize_depth', 'apply_depth_gamma', 'forward_warp_stereo',
           'StereoParams', 'StereoGenerator']
