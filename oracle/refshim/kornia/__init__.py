"""Minimal stand-in for the `kornia` package (absent from this image, unpinned in the
reference's requirements.txt:4).  TEST INFRASTRUCTURE ONLY: it exists so that the UNMODIFIED
reference module /root/reference/helper/stereo_core.py (which does
`from kornia.filters import gaussian_blur2d`, stereo_core.py:19) can be imported in the build
container to generate golden vectors.  Never imported by the product path."""
