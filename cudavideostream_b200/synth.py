"""Synthetic camera.

numpy twin of the device generator in csrc/cvs_filter_kernels.cuh (k_synth_base / k_synth_next): a
counter-based splitmix64, so host and device produce identical bytes for the same (seed, frame index).
"""
from __future__ import annotations

import numpy as np

from .api import _check, load_library

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
DEFAULT_SEED = 0xC0DA5EED


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _splitmix64_scalar(x: int) -> int:
    return int(splitmix64(np.array([x & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0])


def base_frame(width: int, height: int, seed: int = DEFAULT_SEED) -> np.ndarray:
    """Diagonal gradient + noise (non-degenerate gray histogram).  Returns N = 3*W*H bytes."""
    n = 3 * width * height
    i = np.arange(n, dtype=np.uint64)
    px = i // np.uint64(3)
    ch = (i - px * np.uint64(3)).astype(np.int64)
    x = (px % np.uint64(width)).astype(np.int64)
    y = (px // np.uint64(width)).astype(np.int64)
    den = width + height - 2
    grad = ((x + y) * 255) // den if den else np.zeros(n, dtype=np.int64)
    with np.errstate(over="ignore"):
        h = splitmix64((np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + i) & _M64)
    v = grad + (h & np.uint64(63)).astype(np.int64) - 32 + 3 * ch
    return np.clip(v, 0, 255).astype(np.uint8)


def frame_key(seed: int, frame_index: int) -> int:
    return _splitmix64_scalar((seed ^ (((frame_index + 1) * 0xD6E8FEB86659FD93) & 0xFFFFFFFFFFFFFFFF))
                              & 0xFFFFFFFFFFFFFFFF)


def next_frame(prev: np.ndarray, seed: int, frame_index: int, density_ppm: int) -> np.ndarray:
    """Frame t from frame t-1: each byte changes by +-U[21,80] with probability density_ppm/1e6, otherwise
    drifts by U[-3,3] (exercises the negative-feedback accumulation)."""
    prev = np.ascontiguousarray(prev, dtype=np.uint8).reshape(-1)
    n = prev.size
    key = np.uint64(frame_key(seed, frame_index))
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = splitmix64((key + i) & _M64)
    u = h & np.uint64(0xFFFFF)
    big = ((u * np.uint64(1000000)) >> np.uint64(20)) < np.uint64(density_ppm)
    p = prev.astype(np.int64)
    delta = 21 + ((h >> np.uint64(20)) % np.uint64(60)).astype(np.int64)
    up = ((h >> np.uint64(40)) & np.uint64(1)).astype(bool)
    v_big = np.where(up, p + delta, p - delta)
    v_big = np.where(v_big > 255, p - delta, v_big)
    v_big = np.where(v_big < 0, p + delta, v_big)
    v_small = np.clip(p + ((h >> np.uint64(24)) % np.uint64(7)).astype(np.int64) - 3, 0, 255)
    return np.where(big, v_big, v_small).astype(np.uint8)


def sequence(width: int, height: int, nframes: int, density_ppm: int, seed: int = DEFAULT_SEED):
    """Returns (base[N], frames[nframes, N])."""
    base = base_frame(width, height, seed)
    out = np.empty((nframes, base.size), dtype=np.uint8)
    prev = base
    for t in range(nframes):
        prev = next_frame(prev, seed, t, density_ppm)
        out[t] = prev
    return base, out


# device generator (same bytes) ----------------------------------------------------------------------
def base_frame_device(d_out: int, width: int, height: int, seed: int = DEFAULT_SEED, stream: int = 0) -> None:
    _check(load_library().cvs_synth_base_device(d_out, width, height, seed, stream or None))


def next_frame_device(d_prev: int, d_out: int, width: int, height: int, seed: int, frame_index: int,
                      density_ppm: int, stream: int = 0) -> None:
    _check(load_library().cvs_synth_next_device(d_prev, d_out, width, height, seed, frame_index, density_ppm,
                                                stream or None))
