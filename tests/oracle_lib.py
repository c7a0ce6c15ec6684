"""ctypes binding of oracle/build/libdforacle.so -- the CPU oracle (TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "build", "libdforacle.so")

UNDEF, F32, S32, S8, U8 = 0, 1, 2, 3, 4
NP_OF = {F32: np.float32, S32: np.int32, S8: np.int8, U8: np.uint8}
DT_OF = {"f32": F32, "s32": S32, "s8": S8, "u8": U8, None: UNDEF}


class ConvDesc(C.Structure):
    _fields_ = [(k, C.c_int) for k in (
        "n", "ih", "iw", "ic", "oc", "oc1", "kh", "kw", "sh", "sw", "ph", "pw", "dst_dt", "bia0_dt",
        "bia1_dt", "relu0", "relu1", "round0", "round1", "nscale0", "nscale1",
        "literal_f32_intermediate")]


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.dfo_wei_off.restype = C.c_size_t
        _lib.dfo_epilogue_f32.restype = C.c_float
        _lib.dfo_relu_f32.restype = C.c_float
        _lib.dfo_relu_f32.argtypes = [C.c_float]
        _lib.dfo_cvt_f32_s32.argtypes = [C.c_float, C.c_int]
        _lib.dfo_usat8.restype = C.c_uint8
        _lib.dfo_ssat8.restype = C.c_int8
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_desc(n, ih, iw, ic, oc, oc1, dst_dt, bia0_dt=UNDEF, bia1_dt=UNDEF, k=3, stride=1, pad=1,
              relu0=0, relu1=0, round0=0, round1=0, nscale0=1, nscale1=1, literal=0):
    kh, kw = (k, k) if isinstance(k, int) else k
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    ph, pw = (pad, pad) if isinstance(pad, int) else pad
    return ConvDesc(n, ih, iw, ic, oc, oc1, kh, kw, sh, sw, ph, pw, dst_dt, bia0_dt, bia1_dt,
                    relu0, relu1, round0, round1, nscale0, nscale1, literal)


def out_hw(d):
    l = lib()
    return (l.dfo_conv_output_size(d.ih, d.kh, d.sh, d.ph), l.dfo_conv_output_size(d.iw, d.kw, d.sw, d.pw))


def _conv(fn, d, src, wei, bia0, scale0, wei1, bia1, scale1, out=None):
    oh, ow = out_hw(d)
    oc_out = d.oc1 if d.oc1 else d.oc
    # `out`: a caller-owned destination (timing loops must not pay for a fresh allocation + page faults per call)
    dst = out if out is not None else np.zeros((d.n, oh, ow, oc_out), dtype=NP_OF[d.dst_dt])
    assert dst.shape == (d.n, oh, ow, oc_out) and dst.dtype == NP_OF[d.dst_dt] and dst.flags.c_contiguous
    scale0 = np.ascontiguousarray(scale0, dtype=np.float32)
    scale1 = None if scale1 is None else np.ascontiguousarray(scale1, dtype=np.float32)
    rc = fn(C.byref(d), _p(src), _p(wei), _p(bia0), _p(scale0), _p(wei1), _p(bia1), _p(scale1), _p(dst))
    if rc:
        raise RuntimeError(f"oracle conv rejected: {rc}")
    return dst


def conv(d, src, wei, bia0, scale0, wei1=None, bia1=None, scale1=None, out=None):
    return _conv(lib().dfo_conv, d, src, wei, bia0, scale0, wei1, bia1, scale1, out)


def replay_conv(d, src, wei, bia0, scale0, wei1=None, bia1=None, scale1=None, out=None):
    return _conv(lib().dfr_conv, d, src, wei, bia0, scale0, wei1, bia1, scale1, out)


def conv_sum(d, src, wei, bia0, scale0, residual, wei1=None, bia1=None, scale1=None):
    """dfo_conv_sum: the operator with an eltwise sum of `residual` (destination type / layout) before the ReLU."""
    oh, ow = out_hw(d)
    oc_out = d.oc1 if d.oc1 else d.oc
    dst = np.zeros((d.n, oh, ow, oc_out), dtype=NP_OF[d.dst_dt])
    residual = np.ascontiguousarray(residual)
    assert residual.shape == dst.shape and residual.dtype == dst.dtype
    scale0 = np.ascontiguousarray(scale0, dtype=np.float32)
    scale1 = None if scale1 is None else np.ascontiguousarray(scale1, dtype=np.float32)
    rc = lib().dfo_conv_sum(C.byref(d), _p(src), _p(wei), _p(bia0), _p(scale0), _p(wei1), _p(bia1), _p(scale1), _p(residual), _p(dst))
    if rc:
        raise RuntimeError(f"oracle conv_sum rejected: {rc}")
    return dst


POOL_MAX, POOL_AVG_INCLUDE, POOL_AVG_EXCLUDE = 0, 1, 2


def pool(src, kind, k, stride, pad, out_hw_=None, round_mode=0):
    """dfo_pool over an NHWC numpy array (u8 / s8 / s32 / f32)."""
    src = np.ascontiguousarray(src)
    dt = {np.dtype(np.uint8): U8, np.dtype(np.int8): S8, np.dtype(np.int32): S32, np.dtype(np.float32): F32}[src.dtype]
    n, h, w, c = src.shape
    kh, kw = (k, k) if isinstance(k, int) else k
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    ph, pw = (pad, pad) if isinstance(pad, int) else pad
    oh, ow = out_hw_ if out_hw_ else ((h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1)
    dst = np.zeros((n, oh, ow, c), dtype=src.dtype)
    rc = lib().dfo_pool(dt, kind, _p(src), _p(dst), n, h, w, c, kh, kw, sh, sw, ph, pw, oh, ow, round_mode)
    if rc:
        raise RuntimeError(f"oracle pool rejected: {rc}")
    return dst


def conv_intermediate(d, src, wei, bia0, scale0):
    oh, ow = out_hw(d)
    mid = np.zeros((d.n, oh, ow, d.oc), dtype=np.uint8)
    scale0 = np.ascontiguousarray(scale0, dtype=np.float32)
    rc = lib().dfo_conv_intermediate(C.byref(d), _p(src), _p(wei), _p(bia0), _p(scale0), _p(mid))
    if rc:
        raise RuntimeError(f"oracle rejected: {rc}")
    return mid


def _concat(fn, dt, relu, srcs, out=None):
    n = len(srcs)
    ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in srcs])
    ic = (C.c_int * n)(*[s.shape[-1] for s in srcs])
    npix = int(np.prod(srcs[0].shape[:-1]))
    dst = out if out is not None else np.zeros(srcs[0].shape[:-1] + (sum(s.shape[-1] for s in srcs),), dtype=srcs[0].dtype)
    rc = fn(dt, int(relu), n, ptrs, ic, _p(dst), C.c_long(npix))
    if rc:
        raise RuntimeError(f"oracle concat rejected: {rc}")
    return dst


def concat(dt, relu, srcs):
    return _concat(lib().dfo_concat, dt, relu, srcs)


def replay_concat(dt, relu, srcs, out=None):
    return _concat(lib().dfr_concat, dt, relu, srcs, out)


def replay_supported():
    return bool(lib().dfr_supported())


# ---- oracle/_ref/libdfref.so: the reference's own kernel generators (unmodified sources) executed through
#      the recording Xbyak stand-in (oracle/ref_driver.cc).  Built only where /root/reference exists; the
#      built .so travels to the GPU box with the snapshot.
_REF_SO = os.path.join(ROOT, "oracle", "_ref", "libdfref.so")
_ref = None


def ref_lib():
    """The reference-executing library, or None when it was never built / the host lacks AVX-512 VNNI."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO) and os.path.exists("/root/reference/src/jit_conv_kernel.cc"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref"])
        if not os.path.exists(_REF_SO):
            return None
        l = C.CDLL(_REF_SO)
        l.dfref_last_kernel_instructions.restype = C.c_long
        if not l.dfref_supported():
            return None
        _ref = l
    return _ref


def ref_conv(d, src, wei, bia0, scale0, wei1=None, bia1=None, scale1=None):
    return _conv(ref_lib().dfref_conv, d, src, wei, bia0, scale0, wei1, bia1, scale1)


def ref_concat(dt, relu, srcs):
    return _concat(ref_lib().dfref_concat, dt, relu, srcs)


def ref_blocking():
    out = (C.c_int * 5)()
    ref_lib().dfref_last_blocking(out)
    return list(out)
