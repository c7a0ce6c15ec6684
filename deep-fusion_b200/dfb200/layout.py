"""Weight / activation layout tools for the deep-fusion path.

The reference consumes s8 weights in ``OIhw4i16o4i`` -- [O/16][I/16][kh][kw][4i][16o][4i]
(reference src/jit_conv_kernel.cc:333-338, :161-175) -- and ships no converter; these helpers
are the host-side tooling callers need.  Pure numpy index arithmetic, no dependency on oracle/.
"""
from __future__ import annotations

import numpy as np


def oihw_to_blocked(w: np.ndarray) -> np.ndarray:
    """(O, I, kh, kw) -> flat OIhw4i16o4i buffer (same dtype)."""
    O, I, KH, KW = w.shape
    if O % 16 or I % 16:
        raise ValueError("OIhw4i16o4i needs O and I to be multiples of 16")
    # o = ob*16 + o16 ; i = ib*16 + i4o*4 + i4i
    v = w.reshape(O // 16, 16, I // 16, 4, 4, KH, KW)       # ob o16 ib i4o i4i kh kw
    v = v.transpose(0, 2, 5, 6, 3, 1, 4)                     # ob ib kh kw i4o o16 i4i
    return np.ascontiguousarray(v).reshape(-1)


def blocked_to_oihw(b: np.ndarray, O: int, I: int, KH: int, KW: int) -> np.ndarray:
    v = b.reshape(O // 16, I // 16, KH, KW, 4, 16, 4)        # ob ib kh kw i4o o16 i4i
    v = v.transpose(0, 5, 1, 4, 6, 2, 3)                     # ob o16 ib i4o i4i kh kw
    return np.ascontiguousarray(v).reshape(O, I, KH, KW)


def nchw_to_nhwc(x: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(x.transpose(0, 2, 3, 1))


def nhwc_to_nchw(x: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))
