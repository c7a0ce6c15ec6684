"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports exactly
what include/dfcuda.h declares; argument validation that needs no device behaves like the
reference's init_conf rules."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dfb200 as df
from dfb200 import hostapi


def _declared(header, prefix):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", text)))


def test_dfcuda_exports_every_declared_symbol(root):
    names = _declared(os.path.join(root, "include", "dfcuda.h"), "df_")
    assert len(names) >= 20
    lib = df.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dfcuda.h but not exported"
    assert sorted(df.ABI_SYMBOLS) == names


def test_host_library_exports_c_view(root):
    names = _declared(os.path.join(root, "include", "deepfusion_c.h"), "dfh_")
    lib = hostapi.lib()
    for n in names:
        assert hasattr(lib, n), n


def test_host_library_exports_reference_cpp_api(root):
    import subprocess
    out = subprocess.check_output(["nm", "-DC", os.path.join(root, "deep-fusion_b200", "lib", "libdeepfusion.so")]).decode()
    for sym in ["deepfusion::memory::memory(std::array<int, 4ul> const&", "deepfusion::memory::memory(std::vector<int",
                "deepfusion::memory::size()", "deepfusion::memory::buffer_size()", "deepfusion::op::submit()",
                "deepfusion::concat(", "deepfusion::conv("]:
        assert sym in out, sym
    assert out.count("deepfusion::conv(") >= 2  # both overloads


def test_version_and_error_strings():
    assert b"sm_100a" in df.lib().df_version()
    assert isinstance(df.lib().df_last_error(), bytes)


def test_concat_check_matches_reference_rules():
    lib = df.lib()

    def chk(dt, ics):
        return lib.df_concat_check(dt, len(ics), (C.c_int * len(ics))(*ics))

    assert chk(df.U8, [64, 128, 32, 32]) == 0
    assert chk(df.S8, [16]) == 0
    assert chk(df.U8, [24, 16]) != 0            # 1-byte types need multiples of 16
    assert chk(df.F32, [4, 8]) == 0 and chk(df.S32, [6]) != 0
    assert chk(0, [16]) != 0 and chk(df.U8, []) != 0
    assert b"multiple" in lib.df_last_error() or b"inputs" in lib.df_last_error()


def _create(**kw):
    base = dict(n=1, ih=8, iw=8, ic=64, oc=64, oc1=128, kh=3, kw=3, sh=1, sw=1, ph=1, pw=1, dst_dt=df.U8,
                bia0_dt=0, bia1_dt=0, relu0=0, relu1=0, round0=0, round1=0, nscale0=1, nscale1=1)
    base.update(kw)
    d = df.ConvDesc(**base)
    w = np.zeros(base["oc"] * base["ic"] * base["kh"] * base["kw"] + 64, np.int8)
    w1 = np.zeros(max(1, base["oc1"]) * base["oc"] + 64, np.int8)
    s = np.ones(2048, np.float32)
    h = C.c_void_p()
    rc = df.lib().df_conv_create(C.byref(d), w.ctypes.data, w1.ctypes.data, None, None, s.ctypes.data, s.ctypes.data,
                                 C.byref(h))
    if rc == 0:
        df.lib().df_conv_destroy(h)
    return rc


def test_conv_create_rejects_like_reference_without_touching_the_gpu():
    INVALID, UNSUPPORTED = -1, -2
    assert _create(ic=60) == INVALID           # jit_conv_kernel.cc:590
    assert _create(oc=40) == INVALID
    assert _create(oc1=100) == INVALID         # :616
    assert _create(nscale0=3) == INVALID       # :665
    assert _create(nscale1=5) == INVALID       # :668
    assert _create(dst_dt=7) == INVALID
    assert _create(round1=3) == INVALID
    assert _create(ph=6, pw=6, ih=8, iw=8) == INVALID      # l_pad > ur_w (:657-661)
    # accepted by the reference, outside the B200 path: documented as unsupported, never a CPU fallback
    assert _create(kh=3, kw=3, ph=2, pw=2, ih=8, iw=8) == UNSUPPORTED   # output larger than the input
    # the conv0-only operator (oc1 = 0) and other stride-1 windows are on the B200 path: they get as far as the device
    assert _create(oc1=0) not in (INVALID, UNSUPPORTED)
    assert _create(kh=1, kw=1, ph=0, pw=0) not in (INVALID, UNSUPPORTED)
    assert _create(kh=5, kw=5, ph=2, pw=2) not in (INVALID, UNSUPPORTED)
    assert _create(sh=2, sw=2) not in (INVALID, UNSUPPORTED)       # strided windows
    assert _create(oc=512) not in (INVALID, UNSUPPORTED)           # more channels than one accumulator: composite operator
    assert _create(iw=300) not in (INVALID, UNSUPPORTED)           # rows wider than one TMA box


def test_compute_entry_points_fail_loudly_without_a_device():
    n = C.c_int(0)
    if df.lib().df_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    rc = _create()
    assert rc > 0, "df_conv_create must surface the CUDA error when no device exists (no CPU fallback)"
    assert df.lib().df_last_error() != b""


def test_format_tooling_matches_the_reference_layout_formula():
    """df_repack_* / df_nchw_to_nhwc (include/dfcuda.h): OIhw4i16o4i per jit_conv_kernel.cc:333-338,
    gOIhw4i16o4i = one such block per group; checked against the oracle's offset function, the numpy tools and a
    round trip."""
    import ctypes as C
    import numpy as np
    import dfb200 as df
    from dfb200 import layout
    import oracle_lib as O
    L = df.lib()
    L.df_wei_blocked_offset.restype = C.c_size_t
    rng = np.random.default_rng(7)
    for (g, oc, ic, kh, kw) in [(1, 32, 48, 3, 3), (1, 64, 16, 1, 1), (3, 16, 32, 3, 3)]:
        w = rng.integers(-128, 128, size=(g, oc, ic, kh, kw), dtype=np.int8)
        blocked = np.zeros(w.size, np.int8)
        assert L.df_repack_goihw_to_blocked(w.ctypes.data_as(C.c_void_p), blocked.ctypes.data_as(C.c_void_p), g, oc, ic, kh, kw) == 0
        for gi in range(g):
            want = layout.oihw_to_blocked(w[gi])
            assert np.array_equal(blocked.reshape(g, -1)[gi], want)
        for (o, i, h, x) in [(0, 0, 0, 0), (oc - 1, ic - 1, kh - 1, kw - 1), (17 % oc, 5, 0, kw - 1)]:
            assert L.df_wei_blocked_offset(o, i, h, x, ic, kh, kw) == O.lib().dfo_wei_off(o, i, h, x, ic, kh, kw)
        back = np.zeros_like(w)
        assert L.df_repack_blocked_to_goihw(blocked.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p), g, oc, ic, kh, kw) == 0
        assert np.array_equal(back, w)
    assert L.df_repack_oihw_to_blocked(None, None, 16, 16, 1, 1) != 0
    w = np.zeros((20, 16, 1, 1), np.int8)
    assert L.df_repack_oihw_to_blocked(w.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), 20, 16, 1, 1) != 0  # oc % 16
    for dt in (np.uint8, np.float32):
        x = rng.integers(0, 200, size=(2, 5, 3, 4)).astype(dt)
        y = np.zeros((2, 3, 4, 5), dt)
        assert L.df_nchw_to_nhwc(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), 2, 5, 3, 4, x.itemsize) == 0
        assert np.array_equal(y, layout.nchw_to_nhwc(x))
        z = np.zeros_like(x)
        assert L.df_nhwc_to_nchw(y.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), 2, 5, 3, 4, x.itemsize) == 0
        assert np.array_equal(z, x)
