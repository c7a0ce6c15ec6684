"""dfb200 -- thin ctypes binding of the deep-fusion B200 C-ABI (include/dfcuda.h).

This is plumbing for tests/, bench.py and smoke(): it loads ``lib/libdfcuda.so`` (the in-tree CUDA
build) and fails loudly if it is missing -- there is no CPU fallback and nothing here touches
``oracle/``.  Device buffers are plain C-ABI allocations (``df_malloc``); numpy arrays are the
host side.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("DFB200_LIB_DIR") or os.path.join(os.path.dirname(_HERE), "lib")  # (override: development builds, e.g. lib_trace)
LIB_PATH = os.path.join(LIB_DIR, "libdfcuda.so")

UNDEF, F32, S32, S8, U8 = 0, 1, 2, 3, 4
NEAREST, DOWN = 0, 1
DT_OF = {"f32": F32, "s32": S32, "s8": S8, "u8": U8, None: UNDEF}
NP_OF = {F32: np.float32, S32: np.int32, S8: np.int8, U8: np.uint8}

# every symbol include/dfcuda.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "df_last_error", "df_version", "df_device_count", "df_set_device", "df_get_device", "df_device_sm_count", "df_malloc",
    "df_free", "df_memset", "df_host_register", "df_host_unregister", "df_h2d", "df_d2h", "df_stream_create",
    "df_stream_sync", "df_stream_destroy", "df_event_create", "df_event_record", "df_event_record_node", "df_stream_wait_event", "df_event_elapsed_ms",
    "df_event_destroy", "df_concat_check", "df_concat_run", "df_conv_create", "df_conv_run", "df_conv_query", "df_conv_create_concat", "df_conv_run_concat", "df_conv_create_sum", "df_conv_run_sum", "df_pool_check", "df_pool_run",
    "df_conv_destroy", "df_conv_debug_trace", "df_graph_begin", "df_graph_end", "df_graph_launch", "df_graph_destroy",
    "df_wei_blocked_offset", "df_repack_oihw_to_blocked", "df_repack_blocked_to_oihw", "df_repack_goihw_to_blocked",
    "df_repack_blocked_to_goihw", "df_nchw_to_nhwc", "df_nhwc_to_nchw",
]


class ConvDesc(C.Structure):
    _fields_ = [(k, C.c_int) for k in (
        "n", "ih", "iw", "ic", "oc", "oc1", "kh", "kw", "sh", "sw", "ph", "pw", "dst_dt", "bia0_dt", "bia1_dt",
        "relu0", "relu1", "round0", "round1", "nscale0", "nscale1")]


class PoolDesc(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("dtype", "kind", "n", "h", "w", "c", "kh", "kw", "sh", "sw", "ph", "pw", "oh", "ow", "round_mode")]


POOL_MAX, POOL_AVG_INCLUDE, POOL_AVG_EXCLUDE = 0, 1, 2


class ConvInfo(C.Structure):
    _fields_ = [("tiles_per_launch", C.c_int), ("grid", C.c_int), ("block", C.c_int), ("smem_bytes", C.c_int),
                ("w0_resident", C.c_int), ("w1_resident", C.c_int), ("a_stages", C.c_int), ("b_stages", C.c_int),
                ("padded_w", C.c_int), ("padded_h", C.c_int), ("macs_per_image", C.c_double),
                ("mma_efficiency", C.c_double)]


class DfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dfcuda error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """The loaded C-ABI library; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C deep-fusion_b200` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        l.df_last_error.restype = C.c_char_p
        l.df_version.restype = C.c_char_p
        l.df_malloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        l.df_free.argtypes = [C.c_void_p]
        l.df_memset.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
        l.df_host_register.argtypes = [C.c_void_p, C.c_size_t]
        l.df_host_unregister.argtypes = [C.c_void_p]
        l.df_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.df_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.df_stream_create.argtypes = [C.POINTER(C.c_void_p)]
        l.df_stream_sync.argtypes = [C.c_void_p]
        l.df_stream_destroy.argtypes = [C.c_void_p]
        l.df_event_create.argtypes = [C.POINTER(C.c_void_p)]
        l.df_event_record.argtypes = [C.c_void_p, C.c_void_p]
        l.df_event_record_node.argtypes = [C.c_void_p, C.c_void_p]
        l.df_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        l.df_event_destroy.argtypes = [C.c_void_p]
        l.df_stream_wait_event.argtypes = [C.c_void_p, C.c_void_p]
        l.df_graph_begin.argtypes = [C.c_void_p]
        l.df_graph_end.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        l.df_graph_launch.argtypes = [C.c_void_p, C.c_void_p]
        l.df_graph_destroy.argtypes = [C.c_void_p]
        l.df_concat_check.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
        l.df_concat_run.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_void_p,
                                    C.c_long, C.c_void_p]
        l.df_conv_create.argtypes = [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.POINTER(C.c_void_p)]
        l.df_conv_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        l.df_conv_create_concat.argtypes = [C.POINTER(ConvDesc), C.c_int, C.POINTER(C.c_int), C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_void_p)]
        l.df_conv_run_concat.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_void_p]
        l.df_conv_create_sum.argtypes = l.df_conv_create.argtypes
        l.df_conv_run_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        l.df_pool_check.argtypes = [C.POINTER(PoolDesc)]
        l.df_pool_run.argtypes = [C.POINTER(PoolDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        l.df_conv_query.argtypes = [C.c_void_p, C.POINTER(ConvInfo)]
        l.df_conv_destroy.argtypes = [C.c_void_p]
        l.df_conv_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise DfError(rc, lib().df_last_error().decode())


def device_count() -> int:
    n = C.c_int(0)
    check(lib().df_device_count(C.byref(n)))
    return n.value


def set_device(i: int):
    check(lib().df_set_device(i))


def sm_count() -> int:
    n = C.c_int(0)
    check(lib().df_device_sm_count(C.byref(n)))
    return n.value


def sync(stream=None):
    check(lib().df_stream_sync(stream))


class DeviceBuffer:
    """A df_malloc allocation."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib().df_malloc(self.nbytes, C.byref(p)))
        self.ptr = p.value

    @classmethod
    def from_numpy(cls, a: np.ndarray, stream=None):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        b.upload(a, stream)
        return b

    def upload(self, a: np.ndarray, stream=None):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        check(lib().df_h2d(self.ptr, a.ctypes.data, a.nbytes, stream))
        sync(stream)

    def download(self, shape, dtype, stream=None) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(lib().df_d2h(out.ctypes.data, self.ptr, out.nbytes, stream))
        sync(stream)
        return out

    def fill(self, byte: int, stream=None):
        check(lib().df_memset(self.ptr, byte, self.nbytes, stream))

    def free(self):
        if self.ptr:
            lib().df_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Event:
    def __init__(self):
        p = C.c_void_p()
        check(lib().df_event_create(C.byref(p)))
        self.ptr = p.value

    def record(self, stream=None):
        check(lib().df_event_record(self.ptr, stream))

    def record_node(self, stream):
        """Inside a Graph capture: the event becomes a node of the graph, time-stamped by every replay."""
        check(lib().df_event_record_node(self.ptr, stream))

    def elapsed_ms(self, later: "Event") -> float:
        ms = C.c_float(0)
        check(lib().df_event_elapsed_ms(self.ptr, later.ptr, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            lib().df_event_destroy(self.ptr)
        except Exception:
            pass


class Stream:
    def __init__(self):
        p = C.c_void_p()
        check(lib().df_stream_create(C.byref(p)))
        self.ptr = p.value

    def sync(self):
        sync(self.ptr)

    def __del__(self):
        try:
            lib().df_stream_destroy(self.ptr)
        except Exception:
            pass


class Graph:
    """Captured sequence of launches on a Stream (df_graph_*): `with Graph(stream) as g: ...` then g.launch()."""

    def __init__(self, stream: Stream):
        self.stream, self.exec = stream, None

    def __enter__(self):
        check(lib().df_graph_begin(self.stream.ptr))
        return self

    def __exit__(self, et, ev, tb):
        p = C.c_void_p()
        rc = lib().df_graph_end(self.stream.ptr, C.byref(p))
        if et is None:
            check(rc)
        self.exec = p.value
        return False

    def launch(self):
        check(lib().df_graph_launch(self.exec, self.stream.ptr))

    def __del__(self):
        try:
            if self.exec:
                lib().df_graph_destroy(self.exec)
        except Exception:
            pass


def _ptr(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data


class Conv:
    """Fused conv3x3+ReLU+conv1x1+ReLU handle (df_conv_create / df_conv_run)."""

    def __init__(self, n, ih, iw, ic, oc, oc1, dst_dt, wei_blocked, wei1_blocked, bias0=None, bias1=None,
                 scale0=(1.0,), scale1=(1.0,), bia0_dt=UNDEF, bia1_dt=UNDEF, relu0=False, relu1=False,
                 round0=NEAREST, round1=NEAREST, k=3, stride=1, pad=1, with_sum=False):
        scale0 = np.ascontiguousarray(scale0, dtype=np.float32)
        scale1 = np.ascontiguousarray(scale1, dtype=np.float32)
        kh, kw = (k, k) if isinstance(k, int) else k
        sh, sw = (stride, stride) if isinstance(stride, int) else stride
        ph, pw = (pad, pad) if isinstance(pad, int) else pad
        self.desc = ConvDesc(n, ih, iw, ic, oc, oc1, kh, kw, sh, sw, ph, pw, dst_dt, bia0_dt, bia1_dt,
                             int(relu0), int(relu1), round0, round1, scale0.size, scale1.size)
        h = C.c_void_p()
        keep = [np.ascontiguousarray(x) if x is not None else None for x in (wei_blocked, wei1_blocked, bias0, bias1)]
        create = lib().df_conv_create_sum if with_sum else lib().df_conv_create
        check(create(C.byref(self.desc), _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), _ptr(keep[3]),
                     scale0.ctypes.data, scale1.ctypes.data, C.byref(h)))
        self.handle = h.value
        self.oh = (ih + 2 * ph - kh) // sh + 1
        self.ow = (iw + 2 * pw - kw) // sw + 1
        self.out_c = oc1 if oc1 else oc
        self.dst_dt = dst_dt

    def info(self) -> ConvInfo:
        i = ConvInfo()
        check(lib().df_conv_query(self.handle, C.byref(i)))
        return i

    def run(self, src_dev, dst_dev, n=None, stream=None):
        """Asynchronous launch on device pointers (ints or DeviceBuffer)."""
        s = src_dev.ptr if isinstance(src_dev, DeviceBuffer) else src_dev
        d = dst_dev.ptr if isinstance(dst_dev, DeviceBuffer) else dst_dev
        check(lib().df_conv_run(self.handle, s, d, self.desc.n if n is None else n, stream))

    def run_sum(self, src_dev, res_dev, dst_dev, n=None, stream=None):
        """df_conv_run_sum: the operator with an eltwise sum of `res_dev` (destination type / layout) before the ReLU."""
        p = [b.ptr if isinstance(b, DeviceBuffer) else b for b in (src_dev, res_dev, dst_dev)]
        check(lib().df_conv_run_sum(self.handle, p[0], p[1], p[2], self.desc.n if n is None else n, stream))

    def __call__(self, src: np.ndarray, residual: np.ndarray = None) -> np.ndarray:
        """Host convenience: H2D, run, D2H."""
        n = src.shape[0]
        sbuf = DeviceBuffer.from_numpy(src)
        shape = (n, self.oh, self.ow, self.out_c)
        nbytes = int(np.prod(shape)) * np.dtype(NP_OF[self.dst_dt]).itemsize
        dbuf = DeviceBuffer(nbytes)
        dbuf.fill(0xCD)
        if residual is None:
            self.run(sbuf, dbuf, n)
        else:
            rbuf = DeviceBuffer.from_numpy(residual)
            self.run_sum(sbuf, rbuf, dbuf, n)
        sync()
        return dbuf.download(shape, NP_OF[self.dst_dt])

    def close(self):
        if self.handle:
            lib().df_conv_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ConcatConv(Conv):
    """concat(+ReLU) fused into the conv's A-operand load (df_conv_create_concat / df_conv_run_concat): the conv's
    source is the channel concatenation of several NHWC u8 tensors that is never written to memory."""

    def __init__(self, n, ih, iw, src_ics, concat_relu, oc, oc1, dst_dt, wei_blocked, wei1_blocked, bias0=None, bias1=None,
                 scale0=(1.0,), scale1=(1.0,), bia0_dt=UNDEF, bia1_dt=UNDEF, relu0=False, relu1=False,
                 round0=NEAREST, round1=NEAREST):
        scale0 = np.ascontiguousarray(scale0, dtype=np.float32)
        scale1 = np.ascontiguousarray(scale1, dtype=np.float32)
        self.src_ics = list(src_ics)
        ic = sum(self.src_ics)
        self.desc = ConvDesc(n, ih, iw, ic, oc, oc1, 3, 3, 1, 1, 1, 1, dst_dt, bia0_dt, bia1_dt,
                             int(relu0), int(relu1), round0, round1, scale0.size, scale1.size)
        h = C.c_void_p()
        keep = [np.ascontiguousarray(x) if x is not None else None for x in (wei_blocked, wei1_blocked, bias0, bias1)]
        ics = (C.c_int * len(self.src_ics))(*self.src_ics)
        check(lib().df_conv_create_concat(C.byref(self.desc), len(self.src_ics), ics, int(concat_relu), _ptr(keep[0]),
                                          _ptr(keep[1]), _ptr(keep[2]), _ptr(keep[3]), scale0.ctypes.data,
                                          scale1.ctypes.data, C.byref(h)))
        self.handle = h.value
        self.oh, self.ow = ih, iw
        self.out_c = oc1 if oc1 else oc
        self.dst_dt = dst_dt

    def run(self, src_devs, dst_dev, n=None, stream=None):
        ptrs = (C.c_void_p * len(src_devs))(*[b.ptr if isinstance(b, DeviceBuffer) else b for b in src_devs])
        d = dst_dev.ptr if isinstance(dst_dev, DeviceBuffer) else dst_dev
        check(lib().df_conv_run_concat(self.handle, ptrs, d, self.desc.n if n is None else n, stream))

    def __call__(self, srcs) -> np.ndarray:
        n = srcs[0].shape[0]
        bufs = [DeviceBuffer.from_numpy(s) for s in srcs]
        shape = (n, self.oh, self.ow, self.out_c)
        dbuf = DeviceBuffer(int(np.prod(shape)) * np.dtype(NP_OF[self.dst_dt]).itemsize)
        dbuf.fill(0xCD)
        self.run(bufs, dbuf, n)
        sync()
        return dbuf.download(shape, NP_OF[self.dst_dt])


def pool(src: np.ndarray, dtype, kind, k, stride, pad, out_hw=None, round_mode=NEAREST) -> np.ndarray:
    """Host convenience around df_pool_run: NHWC numpy in, NHWC numpy out."""
    n, h, w, c = src.shape
    kh, kw = (k, k) if isinstance(k, int) else k
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    ph, pw = (pad, pad) if isinstance(pad, int) else pad
    oh, ow = out_hw if out_hw else ((h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1)
    d = PoolDesc(dtype, kind, n, h, w, c, kh, kw, sh, sw, ph, pw, oh, ow, round_mode)
    sbuf = DeviceBuffer.from_numpy(src)
    out = DeviceBuffer(max(16, n * oh * ow * c * src.dtype.itemsize))
    out.fill(0xCD)
    check(lib().df_pool_run(C.byref(d), sbuf.ptr, out.ptr, n, None))
    sync()
    return out.download((n, oh, ow, c), src.dtype)


class ConcatCall:
    """A prepared df_concat_run call (argument marshalling done once; used by benchmarks)."""

    def __init__(self, dtype, relu, src_ptrs, ics, dst_ptr, n_pixels, stream=None):
        n = len(src_ptrs)
        self._args = (dtype, int(relu), n, (C.c_void_p * n)(*src_ptrs), (C.c_int * n)(*ics), C.c_void_p(dst_ptr),
                      C.c_long(n_pixels), stream)
        self._fn = lib().df_concat_run

    def __call__(self):
        rc = self._fn(*self._args)
        if rc:
            check(rc)


def concat_run(dtype, relu, src_ptrs, ics, dst_ptr, n_pixels, stream=None):
    n = len(src_ptrs)
    ptrs = (C.c_void_p * n)(*src_ptrs)
    ic = (C.c_int * n)(*ics)
    check(lib().df_concat_run(dtype, int(relu), n, ptrs, ic, dst_ptr, n_pixels, stream))


def concat(srcs, dtype, relu=False) -> np.ndarray:
    """Host convenience: concat NHWC numpy arrays along channels on the GPU."""
    bufs = [DeviceBuffer.from_numpy(s) for s in srcs]
    oc = sum(s.shape[-1] for s in srcs)
    shape = tuple(srcs[0].shape[:-1]) + (oc,)
    npix = int(np.prod(srcs[0].shape[:-1]))
    out = DeviceBuffer(max(16, npix * oc * srcs[0].dtype.itemsize))
    out.fill(0xCD)
    concat_run(dtype, relu, [b.ptr for b in bufs], [s.shape[-1] for s in srcs], out.ptr, npix)
    sync()
    return out.download(shape, srcs[0].dtype)
