"""Run the UNMODIFIED reference `helper/stereo_core.py` on CPU with stage taps.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference is mounted (the build container);
nothing under tests/ (-m gpu), bench.py or smoke() may import this at run time — they use the
golden vectors this script family commits under tests/golden/ instead.

The reference module is imported as-is; the only shim is `kornia.filters.gaussian_blur2d`
(oracle/refshim/kornia, because kornia is not installed).  `run_reference(...)` wraps the
reference's own functions (never replaces their bodies) to record every intermediate of
StereoGenerator.process_frame (stereo_core.py:225-311) so each stage of the restatement and of
the CUDA path can be compared in isolation.
"""
from __future__ import annotations

import importlib
import os
import sys
from typing import Any, Dict

import numpy as np

REFERENCE_ROOT = os.environ.get('VSC_REFERENCE_ROOT', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'refshim')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'helper', 'stereo_core.py'))


def import_reference():
    """Import the reference's helper.stereo_core (unmodified) and return the module."""
    if not reference_available():
        raise RuntimeError('reference tree not mounted at ' + REFERENCE_ROOT)
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # a product-side drop-in package is also called `helper`; make sure we get the reference's
    for name in list(sys.modules):
        if name == 'helper' or name.startswith('helper.'):
            mod = sys.modules[name]
            f = getattr(mod, '__file__', '') or ''
            if not f.startswith(REFERENCE_ROOT):
                del sys.modules[name]
    return importlib.import_module('helper.stereo_core')


def run_reference(rgb: np.ndarray, depth: np.ndarray, params: Dict[str, float] | None = None,
                  taps: bool = True) -> Dict[str, Any]:
    """process_frame on CPU; returns {'sbs': u8[H,2W,3], <stage taps>...}."""
    import cv2
    import torch
    sc = import_reference()
    p = sc.StereoParams(**(params or {}))
    gen = sc.StereoGenerator('cpu')
    rec: Dict[str, Any] = {}
    if not taps:
        rec['sbs'] = gen.process_frame(rgb, depth, p)
        return rec

    saved = {}

    def wrap_mod(name, fn):
        saved[name] = getattr(sc, name)
        setattr(sc, name, fn)

    o_norm, o_gamma, o_warp = sc.normalize_depth, sc.apply_depth_gamma, sc.forward_warp_stereo

    def t_norm(d):
        rec['depth_stretched'] = d.squeeze().numpy().copy()
        out = o_norm(d)
        rec['depth_norm'] = out.squeeze().numpy().copy()
        return out

    def t_gamma(d, g):
        out = o_gamma(d, g)
        rec['depth_gamma'] = out.squeeze().numpy().copy()
        return out

    def t_warp(img, d, md):
        rec['rgb_ss'] = img.squeeze(0).numpy().copy()          # [3,Hs,Ws] f32
        rec['depth_ss'] = d.squeeze().numpy().copy()            # [Hs,Ws] f32 (input of the warp)
        lw, lm, rw, rm = o_warp(img, d, md)
        rec['warp_left'] = lw.squeeze(0).numpy().copy()
        rec['warp_right'] = rw.squeeze(0).numpy().copy()
        rec['mask_left'] = lm.squeeze().numpy().astype(np.uint8)
        rec['mask_right'] = rm.squeeze().numpy().astype(np.uint8)
        return lw, lm, rw, rm

    wrap_mod('normalize_depth', t_norm)
    wrap_mod('apply_depth_gamma', t_gamma)
    wrap_mod('forward_warp_stereo', t_warp)

    cls = sc.StereoGenerator
    m_saved = {}

    def wrap_m(name, fn):
        m_saved[name] = getattr(cls, name)
        setattr(cls, name, fn)

    o_up, o_soft = cls._depth_upsampling, cls._soft_depth_edges
    o_smooth, o_inp, o_sharp = cls._smooth_warping_artifacts, cls._inpaint_missing_regions, cls._sharpen_image
    cnt = {'smooth': 0, 'inp': 0, 'sharp': 0}
    side = ['left', 'right']

    def t_up(self, d, s):
        out = o_up(self, d, s)
        rec['depth_up'] = out.squeeze().numpy().copy()
        return out

    def t_soft(self, d, e):
        out = o_soft(self, d, e)
        rec['depth_soft'] = out.squeeze().numpy().copy()
        return out

    def t_smooth(self, img, s):
        out = o_smooth(self, img, s)
        rec['smooth_' + side[cnt['smooth'] % 2]] = out.squeeze(0).permute(1, 2, 0).numpy().astype(np.uint8)
        cnt['smooth'] += 1
        return out

    def t_inp(self, img, m):
        k = side[cnt['inp'] % 2]
        rec['inpaint_in_' + k] = img.copy()
        rec['inpaint_mask_' + k] = m.copy()
        out = o_inp(self, img, m)
        rec['inpaint_' + k] = out.copy()
        cnt['inp'] += 1
        return out

    def t_sharp(self, img, s):
        out = o_sharp(self, img, s)
        rec['sharp_' + side[cnt['sharp'] % 2]] = out.squeeze(0).numpy().copy()
        cnt['sharp'] += 1
        return out

    wrap_m('_depth_upsampling', t_up)
    wrap_m('_soft_depth_edges', t_soft)
    wrap_m('_smooth_warping_artifacts', t_smooth)
    wrap_m('_inpaint_missing_regions', t_inp)
    wrap_m('_sharpen_image', t_sharp)

    o_resize = cv2.resize
    rs = []

    def t_resize(*a, **k):
        out = o_resize(*a, **k)
        rs.append(out.copy())
        return out

    cv2.resize = t_resize
    try:
        with torch.no_grad():
            rec['sbs'] = gen.process_frame(rgb, depth, p)
    finally:
        cv2.resize = o_resize
        for k, v in saved.items():
            setattr(sc, k, v)
        for k, v in m_saved.items():
            setattr(cls, k, v)
    if len(rs) >= 2:
        rec['rgb_stretched'], rec['depth_stretched_raw'] = rs[0], rs[1]
    return rec
