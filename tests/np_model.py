"""Independent numpy model of the deep-fusion arithmetic contract.

Written from the contract, not from oracle/df_oracle.c: integer accumulation in int64 through an
explicit im2col + matmul, the f32 epilogue as separately rounded numpy float32 operations.  Used
only by tests to pin the oracle from a second direction.
"""
import numpy as np

F32, S32, S8, U8 = 1, 2, 3, 4
INT_MIN = np.int32(-(2 ** 31))


def cvt_f32_s32(t, down):
    t = np.asarray(t, dtype=np.float32)
    r = np.floor(t) if down else np.rint(t)
    bad = np.isnan(t) | ~(r < np.float32(2147483648.0)) | (r < np.float32(-2147483648.0))
    out = np.where(bad, 0, r).astype(np.int64).astype(np.int32)
    return np.where(bad, INT_MIN, out).astype(np.int32)


def relu_x86(t):
    # vmaxps(zero, t): second source when NaN or both zero
    return np.where(np.float32(0) > t, np.float32(0), t).astype(np.float32)


def usat8(v):
    u = v.astype(np.int64) & 0xFFFFFFFF
    return np.where(u > 255, 255, u).astype(np.uint8)


def ssat8(v):
    return np.clip(v.astype(np.int64), -128, 127).astype(np.int8)


def bias_f32(b):
    return None if b is None else np.asarray(b).astype(np.float32)


def epilogue(acc, bias, scale):
    t = acc.astype(np.float32)                    # RN int->float
    if bias is not None:
        t = (t + bias_f32(bias)[None, :]).astype(np.float32)
    s = np.asarray(scale, dtype=np.float32)
    s = np.broadcast_to(s if s.size > 1 else s.reshape(1), (acc.shape[1],)) if s.size > 1 else np.full(acc.shape[1], s.reshape(-1)[0], np.float32)
    return (t * s[None, :]).astype(np.float32)


def im2col(src, kh, kw, sh, sw, ph, pw):
    n, h, w, c = src.shape
    oh = (h + 2 * ph - kh) // sh + 1
    ow = (w + 2 * pw - kw) // sw + 1
    p = np.zeros((n, h + 2 * ph, w + 2 * pw, c), dtype=np.int64)
    p[:, ph:ph + h, pw:pw + w, :] = src
    cols = np.empty((n, oh, ow, kh, kw, c), dtype=np.int64)
    for a in range(kh):
        for b in range(kw):
            cols[:, :, :, a, b, :] = p[:, a:a + sh * oh:sh, b:b + sw * ow:sw, :]
    return cols.reshape(n * oh * ow, kh * kw * c), oh, ow


def finish(t, dst_dt, relu, down):
    if relu or dst_dt == U8:
        t = relu_x86(t)
    if dst_dt == F32:
        return t
    q = cvt_f32_s32(t, down)
    if dst_dt == S32:
        return q
    return ssat8(q) if dst_dt == S8 else usat8(q)


def conv_fused(src, w_oihw, bias0, scale0, w1_oi, bias1, scale1, dst_dt, stride=(1, 1), pad=(1, 1),
               relu0=False, relu1=False, down0=False, down1=False):
    """src NHWC u8; w_oihw (O,I,kh,kw) s8; w1_oi (O1,O) s8 or None (conv0 only)."""
    O, I, KH, KW = w_oihw.shape
    cols, oh, ow = im2col(src, KH, KW, stride[0], stride[1], pad[0], pad[1])
    wm = w_oihw.transpose(2, 3, 1, 0).reshape(KH * KW * I, O).astype(np.int64)
    acc0 = cols @ wm
    assert np.abs(acc0).max() < 2 ** 31
    t0 = epilogue(acc0.astype(np.int32), bias0, scale0)
    n = src.shape[0]
    if w1_oi is None:
        return finish(t0, dst_dt, relu0, down0).reshape(n, oh, ow, O)
    mid = usat8(cvt_f32_s32(relu_x86(t0), down0))
    acc1 = mid.astype(np.int64) @ w1_oi.astype(np.int64).T
    t1 = epilogue(acc1.astype(np.int32), bias1, scale1)
    return finish(t1, dst_dt, relu1, down1).reshape(n, oh, ow, w1_oi.shape[0])


def concat(srcs, dt, relu):
    out = np.concatenate(srcs, axis=-1)
    if not relu:
        return out
    if dt == F32:
        return relu_x86(out)
    if dt == S32:      # vpmaxsw: per 16-bit half
        h = out.view(np.int16)
        return np.where(h < 0, 0, h).astype(np.int16).view(np.int32)
    b = out.view(np.int8)   # s8 and u8: vpmaxsb
    return np.where(b < 0, 0, b).astype(np.int8).view(out.dtype)
