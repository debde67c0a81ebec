"""CPU oracle for the SBS hot path — Python composition over libvsc_oracle.so.

TEST INFRASTRUCTURE ONLY (see vsc_oracle.c header).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

`process_frame(rgb, depth, params, taps=None)` restates
StereoGenerator.process_frame (/root/reference/helper/stereo_core.py:225-311) step by step; the
numbered comments are the reference's line numbers.  Every stage can be tapped so the CUDA path
and the unmodified reference (oracle/ref_runner.py) can be compared stage-wise.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libvsc_oracle.so')
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ('vsc_oracle.c', 'march_model.c')):
        subprocess.check_call(['make', '-C', _HERE, '-s', '-B'])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_powf.restype = C.c_float
        _lib.orc_powf.argtypes = [C.c_float, C.c_float]
        _lib.orc_bilateral_tables.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


@dataclass
class Params:
    """Mirror of StereoParams (stereo_core.py:193-202)."""
    max_disparity: float = 50.0
    convergence: float = -10.0
    super_sampling: float = 3.0
    edge_softness: float = 20.0
    artifact_smoothing: float = 1.0
    depth_gamma: float = 0.2
    sharpen: float = 14.0


# --------------------------------------------------------------------------- geometry (a3)
def geometry(h: int, w: int, p) -> Dict[str, int]:
    """Host scalar geometry, same Python float64 expressions as stereo_core.py:249-251,275-289."""
    total_buffer = 2.0 * p.max_disparity + abs(p.convergence)
    stretch_factor = 1.0 + (total_buffer / w)
    stretched_w = int(w * stretch_factor)
    ss = p.super_sampling > 1.0
    hs = int(h * p.super_sampling) if ss else h
    ws = int(stretched_w * p.super_sampling) if ss else stretched_w
    base = (stretched_w - w) // 2
    cs = int(round(p.convergence))
    lo, ro = base + cs, base - cs
    if ss:
        ratio = ws / stretched_w
        lc, rc, cw = int(lo * ratio), int(ro * ratio), int(w * ratio)
    else:
        lc, rc, cw = lo, ro, w
    return dict(stretched_w=stretched_w, hs=hs, ws=ws, left_crop=lc, right_crop=rc, crop_w=cw, ss=int(ss))


# --------------------------------------------------------------------------- stage wrappers
def lanczos4_h(src: np.ndarray, dw: int) -> np.ndarray:
    """cv2.resize(src, (dw, H), INTER_LANCZOS4) (stereo_core.py:253-254)."""
    src = np.ascontiguousarray(src)
    h, w = src.shape[:2]
    if src.dtype == np.uint8:
        c = 1 if src.ndim == 2 else src.shape[2]
        out = np.empty((h, dw) if src.ndim == 2 else (h, dw, c), np.uint8)
        lib().orc_lanczos4_h_u8(_p(src), h, w, c, dw, _p(out))
    elif src.dtype == np.uint16:
        assert src.ndim == 2
        out = np.empty((h, dw), np.uint16)
        lib().orc_lanczos4_h_u16(_p(src), h, w, dw, _p(out))
    elif src.dtype == np.float32:
        assert src.ndim == 2
        out = np.empty((h, dw), np.float32)
        lib().orc_lanczos4_h_f32(_p(src), h, w, dw, _p(out))
    else:
        raise TypeError(f'unsupported dtype {src.dtype}')
    return out


def normalize_depth(d: np.ndarray) -> np.ndarray:
    d = np.ascontiguousarray(d, np.float32)
    out = np.empty_like(d)
    lib().orc_normalize(_p(d), C.c_size_t(d.size), _p(out), None, None)
    return out


def bilinear_up(x: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """x [C,H,W] or [H,W] f32 -> same rank at (oh, ow)."""
    x = np.ascontiguousarray(x, np.float32)
    sq = x.ndim == 2
    if sq:
        x = x[None]
    c, h, w = x.shape
    out = np.empty((c, oh, ow), np.float32)
    lib().orc_bilinear_up(_p(x), c, h, w, oh, ow, _p(out))
    return out[0] if sq else out


def _torch_sum_f32(x: np.ndarray) -> np.float32:
    """torch.sum of a short contiguous f32 vector as ATen's CPU SumKernel evaluates it
    (identified empirically, 300/300 random vectors per length): n < 8 -> four interleaved
    scalar accumulators, remainder into acc0, combined in order; n >= 8 -> 8-lane vector
    accumulation (4 interleaved vector accumulators), then the scalar tail summed from 0,
    then the 8 lanes added in order."""
    f32 = np.float32
    n = len(x)
    if n < 8:
        acc = [f32(0)] * 4
        nb = n // 4
        for b in range(nb):
            for k in range(4):
                acc[k] = f32(acc[k] + x[b * 4 + k])
        for v in x[nb * 4:]:
            acc[0] = f32(acc[0] + v)
        for k in range(1, 4):
            acc[0] = f32(acc[0] + acc[k])
        return acc[0]
    vs = n // 8
    acc = [np.zeros(8, f32) for _ in range(4)]
    nb = vs // 4
    for b in range(nb):
        for k in range(4):
            acc[k] = (acc[k] + x[(b * 4 + k) * 8:(b * 4 + k + 1) * 8]).astype(f32)
    for i in range(nb * 4, vs):
        acc[0] = (acc[0] + x[i * 8:(i + 1) * 8]).astype(f32)
    for k in range(1, 4):
        acc[0] = (acc[0] + acc[k]).astype(f32)
    s = f32(0)
    for v in x[vs * 8:]:
        s = f32(s + v)
    for k in range(8):
        s = f32(s + acc[0][k])
    return s


def gauss_taps(k: int, sigma: float) -> np.ndarray:
    """kornia get_gaussian_kernel1d in f32 (A.4): exp(-(n-k//2)^2 / (2 sigma^2)) / sum, k odd."""
    n = np.arange(k, dtype=np.float32) - np.float32(k // 2)
    arg = (-(n * n) / np.float32(2.0 * float(sigma) ** 2)).astype(np.float32)
    g = np.exp(arg.astype(np.float64)).astype(np.float32)
    return (g / _torch_sum_f32(g)).astype(np.float32)


def gauss_blur(x: np.ndarray, k: int, sigma: float, taps: Optional[np.ndarray] = None) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    sq = x.ndim == 2
    if sq:
        x = x[None]
    c, h, w = x.shape
    g = np.ascontiguousarray(gauss_taps(k, sigma) if taps is None else taps, np.float32)
    out = np.empty_like(x)
    lib().orc_gauss_blur(_p(x), c, h, w, k, _p(g), _p(out))
    return out[0] if sq else out


def soft_kernel_size(edge_softness: float) -> int:
    return max(5, min(int(edge_softness * 6) | 1, 31))   # stereo_core.py:384


def apply_gamma(d: np.ndarray, gamma: float) -> np.ndarray:
    d = np.ascontiguousarray(d, np.float32)
    out = np.empty_like(d)
    lib().orc_gamma(_p(d), C.c_size_t(d.size), C.c_float(gamma), _p(out))
    return out


def warp(image: np.ndarray, depth: np.ndarray, max_disparity: float, sign: int) -> Tuple[np.ndarray, np.ndarray]:
    """image [C,H,W] f32, depth [H,W] f32 -> (warped [C,H,W] f32, mask [H,W] u8)."""
    image = np.ascontiguousarray(image, np.float32)
    depth = np.ascontiguousarray(depth, np.float32)
    c, h, w = image.shape
    out = np.empty_like(image)
    mask = np.empty((h, w), np.uint8)
    lib().orc_warp(_p(image), c, _p(depth), h, w, C.c_float(max_disparity), sign, _p(out), _p(mask))
    return out, mask


def bilateral_params(artifact_smoothing: float) -> Tuple[int, float, float]:
    d = max(5, min(int(artifact_smoothing * 4), 15))      # stereo_core.py:409
    return d, 30.0, artifact_smoothing * 25               # :410


def bilateral(img: np.ndarray, d: int, sigma_color: float, sigma_space: float, use_fma: bool = True) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w, _ = img.shape
    out = np.empty_like(img)
    lib().orc_bilateral_u8c3(_p(img), h, w, d, C.c_double(sigma_color), C.c_double(sigma_space), int(use_fma), _p(out))
    return out


def dilate3(mask: np.ndarray) -> np.ndarray:
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w = mask.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib().orc_dilate3(_p(mask), h, w, _p(out))
    return out


def telea(img: np.ndarray, mask: np.ndarray, radius: int = 3, return_t: bool = False):
    img = np.ascontiguousarray(img, np.uint8).copy()
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w, _ = img.shape
    t = np.empty((h + 2, w + 2), np.float32) if return_t else None
    lib().orc_telea_u8c3(_p(img), _p(mask), h, w, radius, _p(t) if return_t else None)
    return (img, t) if return_t else img


def telea_two_pass(img: np.ndarray, mask: np.ndarray, radius: int = 3, return_order: bool = False):
    """Same result as telea(), computed as 'arrival times + order from the mask alone, then colours in that order'
    (orc_telea_u8c3_two_pass: the decomposition the GPU march is heading for)."""
    img = np.ascontiguousarray(img, np.uint8).copy()
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w, _ = img.shape
    order = np.empty((h + 2, w + 2), np.int32) if return_order else None
    lib().orc_telea_u8c3_two_pass(_p(img), _p(mask), h, w, radius, _p(order) if return_order else None)
    return (img, order[1:-1, 1:-1]) if return_order else img


def resize_linear_f32(src: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.float32)
    out = np.empty((out_h, out_w), np.float32)
    lib().orc_resize_linear_f32(_p(src), src.shape[0], src.shape[1], out_h, out_w, _p(out))
    return out


def depth_post(depth: np.ndarray, size, bits: int = 16):
    """depth_map_generator.py:217-236: bilinear resize to size = (width, height), min/max normalise, quantise.
    Returns the u8 / u16 map, or None when the resized map is flat (the reference writes no file then)."""
    depth = np.ascontiguousarray(depth, np.float32)
    w, h = int(size[0]), int(size[1])
    out = np.empty((h, w), np.uint16 if bits == 16 else np.uint8)
    lib().orc_depth_post.restype = C.c_int
    ok = lib().orc_depth_post(_p(depth), depth.shape[0], depth.shape[1], h, w, int(bits), _p(out))
    return out if ok else None


MARCH_STATS = ('generations', 'tasks', 'largest_bucket', 'sweeps', 'max_sweeps', 'sorted_buckets', 'max_distinct_t',
               'evaluations')


def march_model(mask: np.ndarray, radius: int = 3):
    """The GPU march's bulk-synchronous schedule, run sequentially (march_model.c): arrival times [h+2,w+2] (as
    telea(return_t=True)), computation order [h,w] (as telea_two_pass(return_order=True)) and the shape of the work
    (generations, bucket sizes, sweeps) of the single march over ring and holes."""
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w = mask.shape
    t = np.empty((h + 2, w + 2), np.float32)
    order = np.empty((h + 2, w + 2), np.int32)
    stats = np.zeros(len(MARCH_STATS), np.int64)
    lib().orc_march_model(_p(mask), h, w, radius, _p(t), _p(order), _p(stats))
    return t, order[1:-1, 1:-1], dict(zip(MARCH_STATS, map(int, stats)))


def sharpen(x: np.ndarray, strength: float) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    c, h, w = x.shape
    g = gauss_taps(5, 1.0)
    out = np.empty_like(x)
    lib().orc_sharpen(_p(x), c, h, w, _p(g), C.c_float(strength), _p(out))
    return out


def area_pool(x: np.ndarray, oh: int, ow: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    c, h, w = x.shape
    out = np.empty((c, oh, ow), np.float32)
    lib().orc_area_pool(_p(x), c, h, w, oh, ow, _p(out))
    return out


def to_u8_trunc(x: np.ndarray) -> np.ndarray:
    """_to_numpy_uint8 (stereo_core.py:346): clamp(0,255) then astype(uint8) = truncation."""
    return np.clip(x, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------- composition
def postprocess_view(warped: np.ndarray, valid: np.ndarray, artifact_smoothing: float,
                     taps: Optional[dict], side: str) -> np.ndarray:
    """_postprocess_view (stereo_core.py:459-485). warped [3,H,W] f32 -> u8 [H,W,3]."""
    inpaint_mask = ((1 - valid.astype(np.float32)) * 255).astype(np.uint8)            # :477
    hwc = np.ascontiguousarray(warped.transpose(1, 2, 0))
    if artifact_smoothing > 0:                                                        # :479
        if hwc.max() > 1.0:                                                           # :404
            img = hwc.astype(np.uint8)                                                # :405 (truncation)
        else:
            img = (hwc * 255).astype(np.uint8)                                        # :407
        d, sc, ss = bilateral_params(artifact_smoothing)
        img = bilateral(img, d, sc, ss)                                               # :410
    else:
        img = to_u8_trunc(hwc)                                                        # :482
    if taps is not None:
        taps['smooth_' + side] = img
        taps['inpaint_mask_' + side] = inpaint_mask
    if inpaint_mask.any():                                                            # :452
        m = dilate3(inpaint_mask)                                                     # :455-456
        img = telea(img, m, 3)                                                        # :457
    if taps is not None:
        taps['inpaint_' + side] = img
    return img


def process_frame(rgb: np.ndarray, depth: np.ndarray, params=None, taps: Optional[dict] = None) -> np.ndarray:
    """StereoGenerator.process_frame (stereo_core.py:225-311)."""
    p = params or Params()
    h, w = rgb.shape[:2]
    g = geometry(h, w, p)
    sw = g['stretched_w']
    if g['left_crop'] < 0 or g['right_crop'] < 0:
        # the reference's slices come back short/empty and _sharpen_image's reflect pad raises
        raise RuntimeError('convergence crop window starts left of the stretched view')
    if depth.dtype not in (np.uint8, np.uint16, np.float32):
        raise TypeError('depth dtype must be uint8, uint16 or float32')
    rgb_st = lanczos4_h(rgb, sw)                                                      # :253
    depth_st = lanczos4_h(depth, sw)                                                  # :254
    depth_f = depth_st.astype(np.float32)                                             # :328
    depth_norm = normalize_depth(depth_f)                                             # :258
    rgb_t = np.ascontiguousarray(rgb_st.astype(np.float32).transpose(2, 0, 1))        # :330
    if taps is not None:
        taps['rgb_stretched'] = rgb_st
        taps['depth_stretched'] = depth_f
        taps['depth_norm'] = depth_norm
    if p.super_sampling > 1.0:                                                        # :260-262
        depth_norm = bilinear_up(depth_norm, g['hs'], g['ws'])
        rgb_t = bilinear_up(rgb_t, g['hs'], g['ws'])
        if taps is not None:
            taps['depth_up'] = depth_norm
    if p.edge_softness > 0:                                                           # :264-265
        depth_norm = gauss_blur(depth_norm, soft_kernel_size(p.edge_softness), p.edge_softness)
        if taps is not None:
            taps['depth_soft'] = depth_norm
    if p.depth_gamma != 1.0:                                                          # :267-268
        depth_norm = apply_gamma(depth_norm, p.depth_gamma)
    if taps is not None:
        taps['depth_ss'] = depth_norm
        taps['rgb_ss'] = rgb_t
    lw, lm = warp(rgb_t, depth_norm, p.max_disparity, +1)                             # :270, :187
    rw, rm = warp(rgb_t, depth_norm, p.max_disparity, -1)                             # :188
    if taps is not None:
        taps.update(warp_left=lw, warp_right=rw, mask_left=lm, mask_right=rm)
    left = postprocess_view(lw, lm, p.artifact_smoothing, taps, 'left')               # :272
    right = postprocess_view(rw, rm, p.artifact_smoothing, taps, 'right')             # :273
    left = np.ascontiguousarray(left.astype(np.float32).transpose(2, 0, 1))           # :485
    right = np.ascontiguousarray(right.astype(np.float32).transpose(2, 0, 1))
    lc, rc, cw = g['left_crop'], g['right_crop'], g['crop_w']
    left = np.ascontiguousarray(left[:, :, lc:lc + cw])                               # :291 / :301
    right = np.ascontiguousarray(right[:, :, rc:rc + cw])                             # :292 / :302
    if left.shape[2] < 3 or right.shape[2] < 3:
        raise RuntimeError('crop window too narrow for the reflect pad of _sharpen_image')
    if p.sharpen > 0:                                                                 # :294-296 / :304-306
        left = sharpen(left, p.sharpen)
        right = sharpen(right, p.sharpen)
        if taps is not None:
            taps['sharp_left'], taps['sharp_right'] = left, right
    if p.super_sampling > 1.0:                                                        # :298-299
        left = area_pool(left, h, w)
        right = area_pool(right, h, w)
    left_np = to_u8_trunc(left.transpose(1, 2, 0))                                    # :308
    right_np = to_u8_trunc(right.transpose(1, 2, 0))                                  # :309
    return np.hstack([left_np, right_np])                                             # :311
