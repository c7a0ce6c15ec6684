"""Pins the oracle to the REFERENCE ITSELF: oracle/_ref/libdfref.so is the reference's own kernel generator
sources (src/jit_conv_kernel.cc, src/jit_concat_kernel.cc, src/op_concat.cc, src/deepfusion.cc ...), compiled
unmodified against a recording stand-in for the un-vendored Xbyak and executed instruction by instruction with
the host's AVX-512 units (oracle/xbyak_shim/xbyak/xbyak.h, oracle/ref_driver.cc).  These tests assert that the
C oracle (oracle/df_oracle.c) -- the checker every GPU parity test uses -- is bit-identical to it.

CPU only.  Skipped when libdfref.so is absent (it is built where /root/reference exists) or the host has no
AVX-512 VNNI."""
import numpy as np
import pytest

import cases
import oracle_lib as O

pytestmark = pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built or host lacks AVX-512 VNNI")


def _desc(c, literal):
    return O.make_desc(c.n, c.h, c.w, c.ic, c.oc, c.oc1, O.DT_OF[c.dst], O.DT_OF[c.b0], O.DT_OF[c.b1], relu0=c.relu0,
                       relu1=c.relu1, round0=c.r0, round1=c.r1, nscale0=c.oc if c.per_channel else 1,
                       nscale1=c.oc1 if c.per_channel else 1, literal=literal)


@pytest.mark.parametrize("c", cases.SMALL_CONV, ids=lambda c: c.name)
def test_fused_conv_oracle_equals_reference(c):
    """The reference generator has defect D3 for f32 destinations (the u8 intermediate is made by saturating
    float BIT PATTERNS, jit_conv_kernel.cc:267,275-277); the oracle reproduces it with literal_f32_intermediate
    and is compared in that mode there.  Every other destination type has no such switch."""
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    literal = 1 if c.dst == "f32" else 0
    want = O.ref_conv(_desc(c, 0), src, wb, b0, s0, w1b, b1, s1)
    got = O.conv(_desc(c, literal), src, wb, b0, s0, w1b, b1, s1)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), f"{c.name}: oracle differs from the reference"


@pytest.mark.parametrize("c", [c for c in cases.SMALL_CONV if c.dst != "f32"][:6], ids=lambda c: c.name)
def test_fused_conv_replay_equals_reference(c):
    """The AVX-512 intrinsics replay (the timed CPU baseline) against the reference as well."""
    if not O.replay_supported():
        pytest.skip("host lacks AVX-512 VNNI")
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    want = O.ref_conv(_desc(c, 0), src, wb, b0, s0, w1b, b1, s1)
    got = O.replay_conv(_desc(c, 0), src, wb, b0, s0, w1b, b1, s1)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("dst", ["u8", "s8", "s32", "f32"])
@pytest.mark.parametrize("relu0,r0", [(0, 0), (1, 1)])
def test_conv0_only_oracle_equals_reference(dst, relu0, r0):
    """The 9-argument conv() (include/deepfusion.h:121-129): jit_conv_kernel with fuse_conv1x1 = false."""
    c = cases.ConvCase("c0", 2, 7, 9, 32, 48, 0, dst, "s32", None, r0=r0, relu0=relu0, k0=11)
    src = cases.synth.src_u8(1, (c.n, c.h, c.w, c.ic))
    w0 = cases.synth.wei_s8(2, (c.oc, c.ic, 3, 3))
    b0 = cases.synth.bias(4, c.oc, "s32")
    s0 = cases.synth.channel_scales(c.oc, c.k0)
    wb = cases.layout.oihw_to_blocked(w0)
    d = O.make_desc(c.n, c.h, c.w, c.ic, c.oc, 0, O.DT_OF[dst], O.S32, O.UNDEF, relu0=relu0, round0=r0, nscale0=c.oc)
    want = O.ref_conv(d, src, wb, b0, s0)
    got = O.conv(d, src, wb, b0, s0)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def test_reference_blocking_matches_the_oracle_helper():
    """nb_ic_blocking / nb_oc_blocking / ur_w / ur_w_tail as jit_conv_kernel::init_conf (:643-655) picked them."""
    import ctypes as C
    for (ic, oc, oc1, h, w) in [(64, 64, 256, 6, 56), (128, 128, 512, 4, 28), (256, 256, 1024, 3, 14), (96, 80, 144, 3, 9)]:
        c = cases.ConvCase("b", 1, h, w, ic, oc, oc1, "u8", None, None)
        src, w0, w1, b0, b1, s0, s1 = c.tensors()
        wb, w1b = c.blocked(w0, w1)
        O.ref_conv(_desc(c, 0), src, wb, b0, s0, w1b, b1, s1)
        got = O.ref_blocking()
        out = (C.c_int * 4)()
        O.lib().dfo_conv_blocking(ic, oc, w, 3, 3, out)
        assert list(out) == got[:4], (ic, oc, w, list(out), got)
        assert got[4] == 1  # the exact (VNNI) accumulation path is the one parity is defined on


CONCAT_ALL = [(dt, srcs) for dt in ("u8", "s8", "s32", "f32")
              for srcs, _ in cases.CONCAT_BASIC + (cases.CONCAT_32BIT_EXTRA if dt in ("s32", "f32") else [])]


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("data", ["reference-range", "full"])
@pytest.mark.parametrize("dt,src_dims", CONCAT_ALL, ids=lambda v: v if isinstance(v, str) else "x".join(str(d[1]) for d in v))
def test_concat_oracle_equals_reference(dt, src_dims, data, relu):
    """The complete reference path: deepfusion::concat() -> op_concat<T>::infer -> jit_concat_kernel, on the
    reference's own shape list (test/test_concat.cc:122-153), its data range and the full dtype range."""
    ins = cases.concat_inputs(dt, src_dims, data)
    want = O.ref_concat(O.DT_OF[dt], relu, ins)
    got = O.concat(O.DT_OF[dt], relu, ins)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


GENERAL = [  # (kh, kw), stride, pad -- windows jit_conv_kernel::init_conf accepts besides 3x3 s1 p1
    ((1, 1), 1, 0), ((3, 3), 1, 0), ((3, 3), 1, (1, 0)), ((5, 5), 1, 2), ((5, 5), 1, (1, 2)), ((7, 7), 1, 3), ((1, 3), 1, (0, 1)),
    ((3, 1), 1, (1, 0)), ((2, 2), 1, 0), ((3, 3), 2, 1), ((1, 1), 2, 0), ((5, 5), 2, 2), ((3, 3), (2, 1), 1), ((7, 7), 2, 3),
    ((3, 3), 1, 2), ((2, 2), 2, 0),
]


@pytest.mark.parametrize("k,stride,pad", GENERAL, ids=lambda v: str(v).replace(" ", ""))
def test_general_window_oracle_equals_reference(k, stride, pad):
    """conv-only operator with other windows, strides and paddings: the reference's generator vs the oracle."""
    n, h, w, ic, oc = 2, 11, 13, 32, 48
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = cases.layout.oihw_to_blocked(cases.synth.wei_s8(2, (oc, ic) + tuple(k)))
    b0 = cases.synth.bias(4, oc, "s32")
    s0 = cases.synth.channel_scales(oc, 10 + (k[0] * k[1] > 8))
    for dst in ("u8", "s32"):
        d = O.make_desc(n, h, w, ic, oc, 0, O.DT_OF[dst], O.S32, O.UNDEF, k=k, stride=stride, pad=pad, relu0=1, nscale0=oc)
        want = O.ref_conv(d, src, wb, b0, s0)
        got = O.conv(d, src, wb, b0, s0)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("k,pad", [(1, 0), (5, 2), ((1, 3), (0, 1)), (7, 3)], ids=str)
def test_fused_general_window_oracle_equals_reference(k, pad):
    n, h, w, ic, oc, oc1 = 2, 12, 10, 64, 64, 144
    kk = (k, k) if isinstance(k, int) else k
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = cases.layout.oihw_to_blocked(cases.synth.wei_s8(2, (oc, ic) + tuple(kk)))
    w1b = cases.layout.oihw_to_blocked(cases.synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1))
    b0, b1 = cases.synth.bias(4, oc, "s32"), cases.synth.bias(5, oc1, "s32")
    s0, s1 = cases.synth.channel_scales(oc, 12), cases.synth.channel_scales(oc1, 12)
    d = O.make_desc(n, h, w, ic, oc, oc1, O.U8, O.S32, O.S32, k=k, pad=pad, nscale0=oc, nscale1=oc1)
    assert np.array_equal(O.conv(d, src, wb, b0, s0, w1b, b1, s1), O.ref_conv(d, src, wb, b0, s0, w1b, b1, s1))
