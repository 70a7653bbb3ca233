"""Seeded synthetic MRC volumes (there is no network for real data).

Header: 256 int32 words, [0..2] = nx, ny, nz, [3] = 2 (MRC mode 2 = float32), rest 0.
Distributions:  G  normal(0,1), seed 1234   -- planes 0-2 incompressible at b=0 (worst case)
                P  poisson(2.0), seed 4321  -- counting-detector proxy, planes 0-1 all zero
                S  sin(x)cos(y)sin(z/2) on a 4*pi grid + normal(0,0.25), seed 7
"""
from __future__ import annotations

import numpy as np

SEEDS = {"G": 1234, "P": 4321, "S": 7}
MRC_HEADER_WORDS = 256


def mrc_header(nx: int, ny: int, nz: int) -> np.ndarray:
    h = np.zeros(MRC_HEADER_WORDS, dtype=np.int32)
    h[0:3] = (nx, ny, nz)
    h[3] = 2
    return h.view(np.uint32)


def volume_data(kind: str, shape, seed: int | None = None) -> np.ndarray:
    """float32 data of `shape` (nz, ny, nx) for one of the three named distributions."""
    rng = np.random.default_rng(SEEDS[kind] if seed is None else seed)
    n = int(np.prod(shape))
    if kind == "G":
        return rng.standard_normal(n, dtype=np.float32)
    if kind == "P":
        return rng.poisson(2.0, n).astype(np.float32)
    if kind == "S":
        nz, ny, nx = shape
        z = np.linspace(0, 4 * np.pi, nz, dtype=np.float32)[:, None, None]
        y = np.linspace(0, 4 * np.pi, ny, dtype=np.float32)[None, :, None]
        x = np.linspace(0, 4 * np.pi, nx, dtype=np.float32)[None, None, :]
        v = (np.sin(x) * np.cos(y) * np.sin(z / 2)).astype(np.float32).reshape(-1)
        v += rng.normal(0, 0.25, n).astype(np.float32)
        return v
    raise ValueError(f"unknown distribution {kind!r}")


def mrc_volume(kind: str, shape, seed: int | None = None) -> np.ndarray:
    """uint32 words of a whole synthetic MRC file: 256 header words + prod(shape) float32."""
    nz, ny, nx = shape
    data = volume_data(kind, shape, seed)
    return np.concatenate([mrc_header(nx, ny, nz), data.view(np.uint32)])


def mrc_words(kind: str, n_data_words: int, seed: int | None = None) -> np.ndarray:
    """Same, for an arbitrary (ragged) number of data words."""
    data = volume_data(kind, (1, 1, n_data_words) if kind != "S" else (1, 1, n_data_words), seed)
    return np.concatenate([mrc_header(n_data_words, 1, 1), data.view(np.uint32)])
