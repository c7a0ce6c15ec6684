"""ctypes view of the C++ host layer (lib/libdeepfusion.so through include/deepfusion_c.h).

`Memory`, `concat`, `conv` and `Op.submit()` are the reference's public API (include/deepfusion.h)
one-to-one; tests written with them read like the reference's own test/test_concat.cc.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import LIB_DIR

HOST_LIB_PATH = os.path.join(LIB_DIR, "libdeepfusion.so")

# deepfusion::memory::format / ::dtype / round_mode
FMT = {"x": 1, "nchw": 2, "oihw": 2, "nhwc": 3, "OIhw4i16o4i": 4, "gOIhw4i16o4i": 5}
DT = {"f32": 1, "s32": 2, "s8": 3, "u8": 4}
NP = {"f32": np.float32, "s32": np.int32, "s8": np.int8, "u8": np.uint8}
NEAREST, DOWN = 0, 1

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise ImportError(f"{HOST_LIB_PATH} is missing: run `make -C deep-fusion_b200`")
        l = C.CDLL(HOST_LIB_PATH)
        l.dfh_memory_create_nchw.restype = C.c_void_p
        l.dfh_memory_create_nchw.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
        l.dfh_memory_create.restype = C.c_void_p
        l.dfh_memory_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int]
        l.dfh_memory_data.restype = C.c_void_p
        l.dfh_memory_data.argtypes = [C.c_void_p]
        l.dfh_memory_bytes.restype = C.c_size_t
        l.dfh_memory_bytes.argtypes = [C.c_void_p]
        l.dfh_memory_device.restype = C.c_void_p
        l.dfh_memory_device.argtypes = [C.c_void_p]
        for f in ("dfh_memory_pin", "dfh_memory_to_device", "dfh_memory_to_host", "dfh_memory_destroy",
                  "dfh_op_submit", "dfh_op_destroy"):
            getattr(l, f).argtypes = [C.c_void_p]
            getattr(l, f).restype = None
        l.dfh_concat_create.restype = C.c_void_p
        l.dfh_concat_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int]
        l.dfh_conv_create.restype = C.c_void_p
        l.dfh_conv_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int,
                                      C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int]
        l.dfh_conv_sharded_create.restype = C.c_void_p
        l.dfh_conv_sharded_create.argtypes = [C.POINTER(C.c_int), C.c_int] + l.dfh_conv_create.argtypes
        l.dfh_concat_conv_create.restype = C.c_void_p
        l.dfh_concat_conv_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int] + l.dfh_conv_create.argtypes[1:]
        l.dfh_concat_conv_is_fused.argtypes = [C.c_void_p]
        l.dfh_conv_pool_create.restype = C.c_void_p
        l.dfh_conv_pool_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p,
                                           C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int]
        l.dfh_pool_create.restype = C.c_void_p
        l.dfh_pool_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        l.dfh_conv_sum_create.restype = C.c_void_p
        l.dfh_conv_sum_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int,
                                          C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int]
        for f in ("dfh_sharded_upload", "dfh_sharded_sync", "dfh_sharded_download"):
            getattr(l, f).argtypes = [C.c_void_p]
            getattr(l, f).restype = None
        l.dfh_op_submit_device.argtypes = [C.c_void_p, C.c_void_p]
        l.dfh_op_submit_device.restype = None
        l.dfh_op_launches.argtypes = [C.c_void_p]
        l.dfh_sync.argtypes = [C.c_void_p]
        l.dfh_sync.restype = None
        _lib = l
    return _lib


class Memory:
    """deepfusion::memory.  `dims` of length-4 tuple with nchw=True are logical N,C,H,W / O,I,H,W."""

    def __init__(self, dims, fmt: str, dt: str, nchw: bool = True, alignment: int = 4096):
        self.dt, self.fmt = dt, fmt
        arr = (C.c_int * len(dims))(*dims)
        if nchw:
            assert len(dims) == 4
            self.h = lib().dfh_memory_create_nchw(arr, FMT[fmt], DT[dt], alignment)
            n, c, hh, w = dims
            self.shape = (n, hh, w, c) if fmt == "nhwc" else tuple(dims)
        else:
            self.h = lib().dfh_memory_create(arr, len(dims), FMT[fmt], DT[dt], alignment)
            self.shape = tuple(dims)
        self.nbytes = lib().dfh_memory_bytes(self.h)

    def array(self) -> np.ndarray:
        """numpy view of the HOST buffer (memory::data())."""
        buf = (C.c_char * self.nbytes).from_address(lib().dfh_memory_data(self.h))
        return np.frombuffer(buf, dtype=NP[self.dt]).reshape(self.shape)

    def set(self, a: np.ndarray):
        self.array()[...] = np.asarray(a, dtype=NP[self.dt]).reshape(self.shape)

    def pin(self):
        lib().dfh_memory_pin(self.h)

    def to_device(self):
        lib().dfh_memory_to_device(self.h)

    def to_host(self):
        lib().dfh_memory_to_host(self.h)

    def device_ptr(self) -> int:
        return lib().dfh_memory_device(self.h)

    def __del__(self):
        try:
            lib().dfh_memory_destroy(self.h)
        except Exception:
            pass


class Op:
    def __init__(self, h, keep):
        self.h, self._keep = h, keep

    def submit(self):
        lib().dfh_op_submit(self.h)

    def submit_device(self, stream=None):
        lib().dfh_op_submit_device(self.h, stream)

    def launches(self) -> int:
        return lib().dfh_op_launches(self.h)

    def __del__(self):
        try:
            lib().dfh_op_destroy(self.h)
        except Exception:
            pass


def sync(stream=None):
    lib().dfh_sync(stream)


def concat(srcs, dst: Memory, post_relu: bool = False) -> Op:
    arr = (C.c_void_p * len(srcs))(*[s.h for s in srcs])
    return Op(lib().dfh_concat_create(arr, len(srcs), dst.h, int(post_relu)), (list(srcs), dst))


def conv(src, wei, bia, stride, padding, dst, wei1x1=None, bia1x1=None, conv0_relu=False, conv0_scales=(1.0,),
         conv0_round_mode=NEAREST, conv1_relu=False, conv1_scales=(1.0,), conv1_round_mode=NEAREST) -> Op:
    s0 = np.ascontiguousarray(conv0_scales, dtype=np.float32)
    s1 = np.ascontiguousarray(conv1_scales, dtype=np.float32)
    st = (C.c_int * 2)(*stride)
    pd = (C.c_int * 2)(*padding)
    fp = C.POINTER(C.c_float)
    h = lib().dfh_conv_create(src.h, wei.h, bia.h if bia else None, st, pd, wei1x1.h if wei1x1 else None,
                              bia1x1.h if bia1x1 else None, dst.h, int(conv0_relu), s0.ctypes.data_as(fp), s0.size,
                              conv0_round_mode, int(conv1_relu), s1.ctypes.data_as(fp), s1.size, conv1_round_mode)
    return Op(h, (src, wei, bia, wei1x1, bia1x1, dst))


def concat_conv(srcs, concat_relu, wei, bia, stride, padding, dst, wei1x1=None, bia1x1=None, conv0_relu=False,
                conv0_scales=(1.0,), conv0_round_mode=NEAREST, conv1_relu=False, conv1_scales=(1.0,),
                conv1_round_mode=NEAREST) -> Op:
    """deepfusion::ext::concat_conv (include/deepfusion_ext.h): concat(+ReLU) fused into the conv's input load."""
    s0 = np.ascontiguousarray(conv0_scales, dtype=np.float32)
    s1 = np.ascontiguousarray(conv1_scales, dtype=np.float32)
    st = (C.c_int * 2)(*stride)
    pd = (C.c_int * 2)(*padding)
    fp = C.POINTER(C.c_float)
    arr = (C.c_void_p * len(srcs))(*[m.h for m in srcs])
    h = lib().dfh_concat_conv_create(arr, len(srcs), int(concat_relu), wei.h, bia.h if bia else None, st, pd,
                                     wei1x1.h if wei1x1 else None, bia1x1.h if bia1x1 else None, dst.h, int(conv0_relu),
                                     s0.ctypes.data_as(fp), s0.size, conv0_round_mode, int(conv1_relu),
                                     s1.ctypes.data_as(fp), s1.size, conv1_round_mode)
    return Op(h, (list(srcs), wei, bia, wei1x1, bia1x1, dst))


def conv_pool(src, wei, bia, stride, padding, conv_dst, pool_dst, kind, pool_kernel, pool_stride, pool_padding, conv_relu=True,
              conv_scales=(1.0,), conv_round_mode=NEAREST, pool_round_mode=NEAREST) -> Op:
    """deepfusion::ext::conv_pool: conv (+ReLU) + pooling (kind 0 max, 1 avg incl. padding, 2 avg excl. padding)."""
    s0 = np.ascontiguousarray(conv_scales, dtype=np.float32)
    i2 = lambda v: (C.c_int * 2)(*v)
    h = lib().dfh_conv_pool_create(src.h, wei.h, bia.h if bia else None, i2(stride), i2(padding), conv_dst.h, pool_dst.h, kind,
                                   i2(pool_kernel), i2(pool_stride), i2(pool_padding), int(conv_relu),
                                   s0.ctypes.data_as(C.POINTER(C.c_float)), s0.size, conv_round_mode, pool_round_mode)
    return Op(h, (src, wei, bia, conv_dst, pool_dst))


def pool(src, dst, kind, pool_kernel, pool_stride, pool_padding, pool_round_mode=NEAREST) -> Op:
    """deepfusion::ext::pool: the pooling stage on its own."""
    i2 = lambda v: (C.c_int * 2)(*v)
    return Op(lib().dfh_pool_create(src.h, dst.h, kind, i2(pool_kernel), i2(pool_stride), i2(pool_padding), pool_round_mode), (src, dst))


def conv_sum(src, wei, bia, stride, padding, residual, dst, wei1x1=None, bia1x1=None, conv0_relu=True, conv0_scales=(1.0,),
             conv0_round_mode=NEAREST, conv1_relu=True, conv1_scales=(1.0,), conv1_round_mode=NEAREST) -> Op:
    """deepfusion::ext::conv_sum: conv / fused conv + eltwise sum of `residual` + ReLU."""
    s0 = np.ascontiguousarray(conv0_scales, dtype=np.float32)
    s1 = np.ascontiguousarray(conv1_scales, dtype=np.float32)
    i2 = lambda v: (C.c_int * 2)(*v)
    fp = C.POINTER(C.c_float)
    h = lib().dfh_conv_sum_create(src.h, wei.h, bia.h if bia else None, i2(stride), i2(padding), wei1x1.h if wei1x1 else None,
                                  bia1x1.h if bia1x1 else None, residual.h, dst.h, int(conv0_relu), s0.ctypes.data_as(fp), s0.size,
                                  conv0_round_mode, int(conv1_relu), s1.ctypes.data_as(fp), s1.size, conv1_round_mode)
    return Op(h, (src, wei, bia, wei1x1, bia1x1, residual, dst))


def concat_conv_is_fused(op: Op) -> bool:
    return bool(lib().dfh_concat_conv_is_fused(op.h))


class ShardedOp(Op):
    """ext::conv_sharded: device-resident helpers for timing."""

    def upload(self):
        lib().dfh_sharded_upload(self.h)

    def sync(self):
        lib().dfh_sharded_sync(self.h)

    def download(self):
        lib().dfh_sharded_download(self.h)


def conv_sharded(devices, src, wei, bia, stride, padding, dst, wei1x1=None, bia1x1=None, conv0_relu=False, conv0_scales=(1.0,),
                 conv0_round_mode=NEAREST, conv1_relu=False, conv1_scales=(1.0,), conv1_round_mode=NEAREST) -> ShardedOp:
    """deepfusion::ext::conv_sharded (include/deepfusion_ext.h): the batch split over `devices` in contiguous slabs."""
    s0 = np.ascontiguousarray(conv0_scales, dtype=np.float32)
    s1 = np.ascontiguousarray(conv1_scales, dtype=np.float32)
    st = (C.c_int * 2)(*stride)
    pd = (C.c_int * 2)(*padding)
    dv = (C.c_int * len(devices))(*devices)
    fp = C.POINTER(C.c_float)
    h = lib().dfh_conv_sharded_create(dv, len(devices), src.h, wei.h, bia.h if bia else None, st, pd, wei1x1.h if wei1x1 else None,
                                      bia1x1.h if bia1x1 else None, dst.h, int(conv0_relu), s0.ctypes.data_as(fp), s0.size,
                                      conv0_round_mode, int(conv1_relu), s1.ctypes.data_as(fp), s1.size, conv1_round_mode)
    return ShardedOp(h, (src, wei, bia, wei1x1, bia1x1, dst))
