"""C-ABI surface: the shared library loads, exports every symbol include/vsc_b200.h declares, the host-only
entry points work without a GPU, and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle as O
from vsc_b200 import _lib
from conftest import HAS_GPU, ROOT


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'vsc_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vsc_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= set(names), sorted(set(_lib.EXPORTS) - set(names))
    assert lib.vsc_abi_version() == 1


def test_default_params_match_stereoparams():
    from vsc_b200 import StereoParams
    p = _lib.VscParams()
    _lib.load().vsc_default_params(C.byref(p))
    d = StereoParams()
    for k in ('max_disparity', 'convergence', 'super_sampling', 'edge_softness', 'artifact_smoothing', 'depth_gamma', 'sharpen'):
        assert getattr(p, k) == getattr(d, k) == getattr(O.Params(), k)
    assert (d.max_disparity, d.convergence, d.super_sampling, d.edge_softness, d.artifact_smoothing, d.depth_gamma, d.sharpen) == \
           (50.0, -10.0, 3.0, 20.0, 1.0, 0.2, 14.0)


def _sweep():
    rng = np.random.default_rng(0)
    for _ in range(400):
        yield (int(rng.integers(64, 4400)), int(rng.integers(64, 7700)),
               O.Params(max_disparity=float(rng.integers(10, 201)) / 2, convergence=float(rng.integers(-50, 51)),
                        super_sampling=float(rng.integers(10, 41)) / 10, edge_softness=float(rng.integers(0, 61)) / 2,
                        artifact_smoothing=float(rng.integers(0, 51)) / 10, depth_gamma=float(rng.integers(2, 41)) / 20,
                        sharpen=float(rng.integers(0, 33)) / 2))


def test_geometry_matches_python_float_expressions():
    """vsc_geometry must reproduce stereo_core.py:249-251,275-289 (Python float64, int(), round()) exactly."""
    n_valid = 0
    for h, w, p in _sweep():
        g = O.geometry(h, w, p)
        valid = g['left_crop'] >= 0 and g['right_crop'] >= 0
        if not valid:
            with pytest.raises(_lib.VscError) as e:
                _lib.geometry(h, w, p)
            assert e.value.code == _lib.VSC_E_PARAMS
            continue
        k = O.soft_kernel_size(p.edge_softness) if p.edge_softness > 0 else 0
        if k // 2 >= min(g['hs'], g['ws']):
            continue
        c = _lib.geometry(h, w, p)
        n_valid += 1
        assert (c.stretched_w, c.ss_h, c.ss_w, c.left_crop, c.right_crop, c.crop_w) == \
               (g['stretched_w'], g['hs'], g['ws'], g['left_crop'], g['right_crop'], g['crop_w']), (h, w, p)
        assert c.blur_k == k
        assert c.bilateral_d == (O.bilateral_params(p.artifact_smoothing)[0] if p.artifact_smoothing > 0 else 0)
    assert n_valid > 200


def test_geometry_known_answers_and_banker_rounding():
    from vsc_b200 import StereoParams
    assert _lib.geometry(1080, 1920, StereoParams()).stretched_w == 2030
    assert _lib.geometry(2160, 3840, StereoParams()).stretched_w == 3949
    assert _lib.geometry(3840, 7680, StereoParams(max_disparity=100.0, convergence=-50.0)).stretched_w == 7929
    # int(round(2.5)) == 2 and int(round(3.5)) == 4 in Python
    a = _lib.geometry(200, 400, StereoParams(convergence=2.5, super_sampling=1.0))
    b = _lib.geometry(200, 400, StereoParams(convergence=3.5, super_sampling=1.0))
    base_a, base_b = (a.stretched_w - 400) // 2, (b.stretched_w - 400) // 2
    assert a.left_crop - base_a == 2 and b.left_crop - base_b == 4


def test_invalid_arguments_return_codes():
    lib = _lib.load()
    g = _lib.VscGeom()
    p = _lib.make_params(O.Params())
    assert lib.vsc_geometry(4, 4, C.byref(p), C.byref(g)) == _lib.VSC_E_INVALID
    assert b'8x8' in lib.vsc_last_error()
    assert lib.vsc_geometry(100, 100, None, C.byref(g)) == _lib.VSC_E_INVALID
    bad = _lib.make_params(O.Params(super_sampling=0.0))
    assert lib.vsc_geometry(100, 100, C.byref(bad), C.byref(g)) == _lib.VSC_E_INVALID
    h = C.c_void_p()
    assert lib.vsc_create(0, 0, C.byref(h)) == _lib.VSC_E_INVALID


@pytest.mark.skipif(HAS_GPU, reason='checks the no-GPU behaviour')
def test_no_gpu_means_loud_failure_not_a_fallback():
    from vsc_b200 import StereoGenerator
    with pytest.raises(_lib.VscCudaError) as e:
        _lib.Context(0, 1)
    assert 'no CPU fallback' in str(e.value)
    with pytest.raises(RuntimeError):
        StereoGenerator('cuda')


def test_cpu_device_is_refused_everywhere():
    from vsc_b200 import StereoGenerator
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        StereoGenerator('cpu')


def test_missing_library_raises_importerror(monkeypatch, tmp_path):
    import importlib
    monkeypatch.setenv('VSC_B200_LIB', str(tmp_path / 'nope.so'))
    mod = importlib.reload(_lib)
    try:
        with pytest.raises(ImportError, match='no CPU fallback'):
            mod.load()
    finally:
        monkeypatch.delenv('VSC_B200_LIB')
        importlib.reload(_lib)


def test_product_code_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use oracle/."""
    pkg = os.path.join(ROOT, 'video-stereo-converter_b200')
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.sh')):
                t = open(os.path.join(d, f), errors='ignore').read()
                if re.search(r'import\s+oracle|from\s+oracle|libvsc_oracle|oracle/', t.replace('see oracle.py', '')):
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders


def test_helper_dropin_exports_the_reference_names():
    import importlib.util
    spec = importlib.util.spec_from_file_location('dropin_stereo_core', os.path.join(ROOT, 'video-stereo-converter_b200', 'helper', 'stereo_core.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.__all__ == ['load_image_pair', 'normalize_depth', 'apply_depth_gamma', 'forward_warp_stereo', 'StereoParams', 'StereoGenerator']
    for n in m.__all__:
        assert hasattr(m, n)
