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
import importlib.util
import os
import sys
from typing import Any, Dict

import numpy as np

REFERENCE_ROOT = os.environ.get('VSC_REFERENCE_ROOT', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, 'refshim')
STAGED_ROOT = os.path.join(_HERE, '_ref')          # git-ignored copy made by oracle/make_ref.py (travels to the GPU box)
_module = None


def _roots():
    return [r for r in (REFERENCE_ROOT, STAGED_ROOT) if os.path.isfile(os.path.join(r, 'helper', 'stereo_core.py'))]


def reference_available() -> bool:
    return bool(_roots())


def reference_origin() -> str:
    """'/root/reference', the staged copy, or '' - where import_reference() loads the module from."""
    r = _roots()
    return r[0] if r else ''


def import_reference():
    """Load the reference's helper/stereo_core.py (unmodified) BY FILE PATH and return the module.

    Loading by path (not `import helper.stereo_core`) matters: the product ships a drop-in package that is also
    called `helper`, and a regular package always wins over the reference's namespace package, whatever the order of
    sys.path.  The only name injected is `kornia` (the shim), because kornia is not installed."""
    global _module
    if _module is not None:
        return _module
    roots = _roots()
    if not roots:
        raise RuntimeError(f'reference not found at {REFERENCE_ROOT} nor staged at {STAGED_ROOT} (python oracle/make_ref.py)')
    path = os.path.join(roots[0], 'helper', 'stereo_core.py')
    if 'kornia' not in sys.modules:
        shim_dir = _SHIM if os.path.isdir(os.path.join(_SHIM, 'kornia')) else roots[0]
        spec = importlib.util.spec_from_file_location('kornia', os.path.join(shim_dir, 'kornia', '__init__.py'),
                                                      submodule_search_locations=[os.path.join(shim_dir, 'kornia')])
        kornia = importlib.util.module_from_spec(spec)
        sys.modules['kornia'] = kornia
        spec.loader.exec_module(kornia)
    spec = importlib.util.spec_from_file_location('vsc_reference_stereo_core', path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules['vsc_reference_stereo_core'] = mod       # dataclasses looks the module up by name
    spec.loader.exec_module(mod)
    assert os.path.abspath(mod.__file__).startswith(os.path.abspath(roots[0])), mod.__file__
    _module = mod
    return mod


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
