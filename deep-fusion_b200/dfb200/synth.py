"""Deterministic synthetic tensors for the deep-fusion hot path.

Counter-based generator: value(i) = splitmix64(seed * 2^32 + i), so any slice of any tensor can
be regenerated independently on any rank without state.  Used by tests/, bench.py and smoke().
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _raw(seed: int, count: int, start: int = 0) -> np.ndarray:
    idx = np.arange(start, start + count, dtype=np.uint64)
    return splitmix64(idx + (np.uint64(seed) << np.uint64(32)))


def uniform_int(seed: int, shape, lo: int, hi: int, dtype, start: int = 0) -> np.ndarray:
    """Uniform integers in [lo, hi] (inclusive)."""
    count = int(np.prod(shape))
    span = np.uint64(hi - lo + 1)
    v = (_raw(seed, count, start) >> np.uint64(16)) % span
    return (v.astype(np.int64) + lo).astype(dtype).reshape(shape)


def src_u8(seed, shape, lo=0, hi=255, start=0):
    return uniform_int(seed, shape, lo, hi, np.uint8, start)


def wei_s8(seed, shape, lo=-127, hi=127):
    return uniform_int(seed, shape, lo, hi, np.int8)


def bias(seed, n, dt: str, mag=4096):
    if dt == "s32":
        return uniform_int(seed, (n,), -mag, mag, np.int32)
    if dt == "s8":
        return uniform_int(seed, (n,), -128, 127, np.int8)
    if dt == "u8":
        return uniform_int(seed, (n,), 0, 255, np.uint8)
    if dt == "f32":
        return (uniform_int(seed, (n,), -mag * 8, mag * 8, np.int32).astype(np.float32) / np.float32(8))
    raise ValueError(dt)


def channel_scales(n: int, k: int) -> np.ndarray:
    """scale[o] = 2^-k * (1 + (o mod 13)/32)  (SURVEY.md §8d)."""
    o = np.arange(n, dtype=np.float32)
    return (np.float32(2.0) ** np.float32(-k) * (np.float32(1) + np.mod(o, 13) / np.float32(32))).astype(np.float32)


def pick_scale_exponent(sample_acc: np.ndarray, target_nonzero=0.5, max_sat=0.05) -> int:
    """Smallest k such that <= max_sat of relu(acc)*2^-k saturates u8."""
    a = np.maximum(sample_acc.astype(np.float64), 0.0) * (1 + 6 / 32)
    for k in range(0, 31):
        if np.mean(a * 2.0 ** (-k) > 255.0) <= max_sat:
            return k
    return 30
