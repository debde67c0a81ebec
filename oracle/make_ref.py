"""Stage the UNMODIFIED reference module for the bench's reference arm.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_ref.py          (run where /root/reference is mounted: the build container)

The reference is a set of Python scripts without a setup.py / pyproject.toml, so it cannot be pip-installed
(DESIGN.md section 6), and /root/reference does not exist on the GPU box.  This recipe places a byte-for-byte copy
of the one file the hot path lives in, /root/reference/helper/stereo_core.py, under oracle/_ref/helper/ next to the
kornia shim (oracle/refshim, kornia is not installed anywhere).  oracle/_ref/ is git-ignored - reference sources
never enter the history - but it is NOT gpurun-ignored, so it travels to the GPU box like a built .so, and
`bench.py --impl reference` can time the real StereoGenerator('cpu').process_frame there (cpu_baseline.kind
"reference").  A manifest records the SHA-256 of the copied file so a stale copy is detectable.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get('VSC_REFERENCE_ROOT', '/root/reference')
DEST = os.path.join(HERE, '_ref')


def main() -> int:
    src = os.path.join(REFERENCE_ROOT, 'helper', 'stereo_core.py')
    if not os.path.isfile(src):
        print(f'make_ref: {src} not found; nothing staged (the bench falls back to the oracle port)')
        return 0
    os.makedirs(os.path.join(DEST, 'helper'), exist_ok=True)
    dst = os.path.join(DEST, 'helper', 'stereo_core.py')
    shutil.copyfile(src, dst)
    shim_dst = os.path.join(DEST, 'kornia')
    if os.path.isdir(shim_dst):
        shutil.rmtree(shim_dst)
    shutil.copytree(os.path.join(HERE, 'refshim', 'kornia'), shim_dst, ignore=shutil.ignore_patterns('__pycache__'))
    digest = hashlib.sha256(open(dst, 'rb').read()).hexdigest()
    with open(os.path.join(DEST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'sha256': digest, 'note': 'unmodified copy; git-ignored; see oracle/make_ref.py'}, f, indent=1)
    print(f'make_ref: staged {src} -> {dst} (sha256 {digest[:16]})')
    return 0


if __name__ == '__main__':
    sys.exit(main())
