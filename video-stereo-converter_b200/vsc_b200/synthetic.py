"""Synthetic frame / depth generators for parity tests and benchmarks.

There is no dataset access, so every test and bench input is produced here from a seed
(SURVEY.md §8(d)): a smooth colour field with Gaussian noise, and a depth map made of a
vertical ramp with max-composited discs (sharp depth edges => real disocclusions).
"""
from __future__ import annotations

import numpy as np

__all__ = ['make_rgb', 'make_depth', 'make_pair']


def make_rgb(h: int, w: int, seed: int = 0) -> np.ndarray:
    """uint8 [h, w, 3] RGB: 128 + 100*sin(x/97+c)*cos(y/61-c) per channel c, + N(0,12)."""
    rng = np.random.default_rng(seed)
    y = np.arange(h, dtype=np.float64)[:, None]
    x = np.arange(w, dtype=np.float64)[None, :]
    out = np.empty((h, w, 3), np.float64)
    for c in range(3):
        out[:, :, c] = 128.0 + 100.0 * np.sin(x / 97.0 + c) * np.cos(y / 61.0 - c)
    out += rng.normal(0.0, 12.0, size=out.shape)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def make_depth(h: int, w: int, seed: int = 0, dtype=np.uint8, discs: int = 12) -> np.ndarray:
    """[h, w] depth: ramp 0.2+0.3*y/h with `discs` random discs, quantised to `dtype`."""
    rng = np.random.default_rng(seed + 1_000_003)
    y = np.arange(h, dtype=np.float64)[:, None]
    x = np.arange(w, dtype=np.float64)[None, :]
    d = np.broadcast_to(0.2 + 0.3 * y / h, (h, w)).copy()
    for _ in range(discs):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        r = rng.uniform(h / 20.0, h / 5.0)
        lvl = rng.uniform(0.5, 1.0)
        inside = (y - cy) ** 2 + (x - cx) ** 2 <= r * r
        d = np.where(inside, np.maximum(d, lvl), d)
    dt = np.dtype(dtype)
    if dt == np.uint8:
        return np.rint(d * 255.0).astype(np.uint8)
    if dt == np.uint16:
        return np.rint(d * 65535.0).astype(np.uint16)
    return d.astype(dt)


def make_pair(h: int, w: int, seed: int = 0, depth_dtype=np.uint8):
    return make_rgb(h, w, seed), make_depth(h, w, seed, depth_dtype)
