"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py [fixture names]

Each fixture holds the synthetic inputs, the stereo parameters, the reference's SBS output, its hole
masks (bit-packed) and SHA-256 digests of the larger intermediates, produced by importing
/root/reference/helper/stereo_core.py as-is (kornia shimmed, see oracle/ref_runner.py) on CPU with
torch {torch}, OpenCV {cv2}, numpy {numpy}.  TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)      # NOT the product directory: its drop-in `helper` package must never be importable here

import importlib.util  # noqa: E402

import ref_runner  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    'vsc_synthetic', os.path.join(ROOT, 'video-stereo-converter_b200', 'vsc_b200', 'synthetic.py'))
_syn = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_syn)
make_pair = _syn.make_pair

CASES = {
    'default_u8': ((120, 160), np.uint8, 7, {}),
    'default_u16': ((135, 240), np.uint16, 8, {}),
    'ss1_sharp_edges': ((100, 180), np.uint8, 9, dict(super_sampling=1.0, edge_softness=0.0, depth_gamma=1.0, max_disparity=30.0,
                                                       convergence=5.0, artifact_smoothing=5.0)),
    'ss2p5_nosmooth_nosharpen': ((90, 150), np.uint16, 10, dict(super_sampling=2.5, edge_softness=3.0, depth_gamma=0.5, max_disparity=20.0,
                                                                convergence=-7.0, artifact_smoothing=0.0, sharpen=0.0)),
    'aggressive_band_in_crop': ((96, 200), np.uint16, 11, dict(super_sampling=2.0, edge_softness=0.0, depth_gamma=1.0, max_disparity=100.0,
                                                               convergence=-50.0, artifact_smoothing=5.0)),
    'ss4_max_sliders': ((64, 256), np.uint8, 12, dict(super_sampling=4.0, edge_softness=30.0, depth_gamma=2.0, max_disparity=100.0,
                                                      convergence=50.0, artifact_smoothing=2.5, sharpen=16.0)),
    'min_sliders': ((80, 120), np.uint16, 13, dict(super_sampling=1.3, edge_softness=0.5, depth_gamma=0.1, max_disparity=5.0,
                                                   convergence=0.0, artifact_smoothing=0.1, sharpen=0.5)),
    'float_depth': ((72, 128), np.float32, 14, dict(super_sampling=2.0)),
    'flat_depth': ((64, 96), 'flat', 15, {}),
    'near_black': ((64, 96), 'ones', 16, {}),
}


# Full-size frames: the reference's SBS output is not stored (12 MB of noise per 1080p frame).  The fixture holds its
# SHA-256 plus the sparse list of values in which it differs from the ORACLE's output at generation time, so a test
# can rebuild the reference's frame exactly (oracle output + patches), prove it with the digest, and then apply the
# tolerance to the oracle and to the CUDA path.  Inputs are regenerated from the seed.
BIG_CASES = {
    'full_1080p_u8': ((1080, 1920), np.uint8, 0, {}),
    'band_4k_u16': ((256, 3840), np.uint16, 1, {}),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import cv2
    import torch
    out_dir = os.environ.get('VSC_GOLDEN_OUT', os.path.join(ROOT, 'tests', 'golden'))
    os.makedirs(out_dir, exist_ok=True)
    only = set(sys.argv[1:])          # optional: fixture names to (re)generate
    for name, (shape, dt, seed, kw) in CASES.items():
        if only and name not in only:
            continue
        if dt == 'flat':
            rgb, _ = make_pair(*shape, seed=seed)
            depth = np.full(shape, 77, np.uint8)
        elif dt == 'ones':
            _, depth = make_pair(*shape, seed=seed)
            rgb = np.ones(shape + (3,), np.uint8)
        else:
            rgb, depth = make_pair(*shape, seed=seed, depth_dtype=dt)
        cv2.ipp.setUseIPP(True)
        r = ref_runner.run_reference(rgb, depth, kw)
        cv2.ipp.setUseIPP(False)       # OpenCV's own (documented) code paths instead of the closed-source IPP ones
        sbs_noipp = ref_runner.run_reference(rgb, depth, kw, taps=False)['sbs']
        cv2.ipp.setUseIPP(True)
        fx = dict(rgb=rgb, depth=depth, params=np.array(json.dumps(kw)),
                  sbs=r['sbs'], sbs_noipp=sbs_noipp,
                  mask_left=np.packbits(r['mask_left']), mask_right=np.packbits(r['mask_right']), mask_shape=np.array(r['mask_left'].shape),
                  sha_rgb_stretched=sha(r['rgb_stretched']), sha_depth_stretched=sha(r['depth_stretched']),
                  sha_depth_norm=sha(r['depth_norm']), sha_rgb_ss=sha(r['rgb_ss']),
                  depth_ss_f16=r['depth_ss'].astype(np.float16), sha_depth_ss=sha(r['depth_ss']),
                  sha_warp_left=sha(r['warp_left']), sha_warp_right=sha(r['warp_right']),
                  versions=np.array([torch.__version__, cv2.__version__, np.__version__]))
        if 'depth_soft' in r:
            fx['sha_depth_soft'] = sha(r['depth_soft'])
        np.savez_compressed(os.path.join(out_dir, name + '.npz'), **fx)
        print(name, shape, r['sbs'].shape, 'holes L/R', int((r['mask_left'] == 0).sum()), int((r['mask_right'] == 0).sum()),
              'IPP on/off differ in', int((r['sbs'] != sbs_noipp).sum()), 'values')
    import oracle as O
    for name, (shape, dt, seed, kw) in BIG_CASES.items():
        if only and name not in only:
            continue
        rgb, depth = make_pair(*shape, seed=seed, depth_dtype=dt)
        cv2.ipp.setUseIPP(True)
        r = ref_runner.run_reference(rgb, depth, kw)
        cv2.ipp.setUseIPP(False)
        sbs_noipp = ref_runner.run_reference(rgb, depth, kw, taps=False)['sbs']
        cv2.ipp.setUseIPP(True)
        mine = O.process_frame(rgb, depth, O.Params(**kw))
        fx = dict(shape=np.array(shape), seed=np.array(seed), depth_dtype=np.array(np.dtype(dt).name), params=np.array(json.dumps(kw)),
                  sha_sbs=sha(r['sbs']), sha_sbs_noipp=sha(sbs_noipp),
                  sha_mask_left=sha(np.packbits(r['mask_left'])), sha_mask_right=sha(np.packbits(r['mask_right'])),
                  versions=np.array([torch.__version__, cv2.__version__, np.__version__]))
        for tag, ref in (('', r['sbs']), ('_noipp', sbs_noipp)):
            idx = np.flatnonzero(ref.ravel() != mine.ravel())
            fx['patch_idx' + tag] = idx.astype(np.int64)
            fx['patch_val' + tag] = ref.ravel()[idx]
        np.savez_compressed(os.path.join(out_dir, name + '.big.npz'), **fx)
        print(name, shape, 'oracle differs from the reference in', len(fx['patch_idx']), 'values (IPP) /', len(fx['patch_idx_noipp']), '(IPP off)')


if __name__ == '__main__':
    main()
