"""CPU tests of the oracle (test infrastructure) -- no GPU.

The oracle is pinned from every direction available without the (unbuildable) reference binary:
 * known answers for the x86 instruction semantics it restates (SURVEY.md §8a C2-C6 probes),
 * the reference's own helper KATs (test/test_misc.cc:25-36),
 * an independent numpy int64 model and the AVX-512-VNNI replay of the emitted instructions,
 * the committed golden fixtures,
 * on the reference's own concat test list and data range (test/test_concat.cc:122-153,
   test/test_utils.h:49-63) plain concat + true ReLU.
"""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import np_model as M
import oracle_lib as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def f32(x):
    return C.c_float(x)


# ------------------------------------------------------------------ element semantics (KATs)
def test_cvt_round_nearest_even_ties():
    L = O.lib()
    for t, want in [(0.5, 0), (1.5, 2), (2.5, 2), (254.5, 254), (255.5, 256), (-0.5, 0), (-1.5, -2), (3.49, 3)]:
        assert L.dfo_cvt_f32_s32(f32(t), 0) == want  # mode 0 = nearest


def test_cvt_round_down_and_indefinite():
    L = O.lib()
    assert L.dfo_cvt_f32_s32(f32(2.9), 1) == 2
    assert L.dfo_cvt_f32_s32(f32(-2.1), 1) == -3
    for bad in (float("nan"), float("inf"), 2147483648.0, -4e9, 1e30):
        assert L.dfo_cvt_f32_s32(f32(bad), 0) == -(2 ** 31)
    assert L.dfo_cvt_f32_s32(f32(-2147483648.0), 0) == -(2 ** 31)
    assert L.dfo_cvt_f32_s32(f32(2147483520.0), 0) == 2147483520


def test_saturations():
    L = O.lib()
    assert L.dfo_usat8(256) == 255 and L.dfo_usat8(255) == 255 and L.dfo_usat8(0) == 0
    assert L.dfo_usat8(-(2 ** 31)) == 255  # vpmovusdb treats the dword as unsigned
    assert L.dfo_ssat8(-(2 ** 31)) == -128 and L.dfo_ssat8(300) == 127 and L.dfo_ssat8(-5) == -5


def test_relu_is_x86_maxps():
    L = O.lib()
    assert L.dfo_relu_f32(-3.0) == 0.0 and L.dfo_relu_f32(2.5) == 2.5
    assert np.isnan(L.dfo_relu_f32(float("nan")))
    assert np.signbit(np.float32(L.dfo_relu_f32(-0.0)))  # -0.0 stays -0.0


def test_epilogue_is_two_roundings_not_fma():
    L = O.lib()
    acc, bias, scale = 16777217, np.array([3], np.int32), np.float32(1.0 / 3.0)
    want = np.float32(np.float32(np.float32(acc) + np.float32(3)) * scale)
    got = L.dfo_epilogue_f32(acc, O.S32, bias.ctypes.data_as(C.c_void_p), 0, f32(scale))
    assert np.float32(got) == want
    assert np.float32(acc) == np.float32(16777216)  # s32 -> f32 is inexact above 2^24 (C5)


# ----------------------------------------------------------------------- reference helper KATs
def test_misc_kats_from_reference():
    L = O.lib()

    def dividable_of(val, *ds):
        arr = (C.c_int * len(ds))(*ds)
        return L.dfo_dividable_of(val, arr, len(ds))

    assert dividable_of(12, 5, 4, 3, 2) == 4   # test/test_misc.cc:25-27
    assert dividable_of(12, 3, 4, 5, 2) == 3
    assert dividable_of(5, 3, 2) == 1
    for (v, d), want in {(14, 8): 7, (12, 5): 4, (12, 3): 3, (5, 4): 1, (5, 5): 5, (5, 8): 5}.items():
        assert L.dfo_find_dividable(v, d) == want  # test/test_misc.cc:29-34


def test_balance211_partitions_everything():
    L = O.lib()
    for n, team in [(4, 3), (56, 8), (64 * 28, 28), (1, 8), (0, 4), (100, 1)]:
        covered, sizes = 0, []
        for tid in range(team):
            s, e = C.c_long(), C.c_long()
            L.dfo_balance211(C.c_long(n), team, tid, C.byref(s), C.byref(e))
            assert s.value == covered
            covered = e.value
            sizes.append(e.value - s.value)
        assert covered == n and max(sizes) - min(sizes) <= 1


def test_blocking_matches_survey_table():
    L = O.lib()
    out = (C.c_int * 4)()
    for (ic, oc, ow), want in {(64, 64, 56): (4, 4, 5, 1), (128, 128, 28): (8, 4, 5, 3), (256, 256, 14): (8, 4, 5, 4)}.items():
        L.dfo_conv_blocking(ic, oc, ow, 3, 3, out)
        assert tuple(out) == want


def test_weight_offset_formula():
    L = O.lib()
    L.dfo_wei_off.argtypes = [C.c_int] * 7
    # [O/16][I/16][kh][kw][4i][16o][4i]
    assert L.dfo_wei_off(0, 0, 0, 0, 64, 3, 3) == 0
    assert L.dfo_wei_off(1, 0, 0, 0, 64, 3, 3) == 4
    assert L.dfo_wei_off(0, 1, 0, 0, 64, 3, 3) == 1
    assert L.dfo_wei_off(0, 4, 0, 0, 64, 3, 3) == 64
    assert L.dfo_wei_off(0, 0, 0, 1, 64, 3, 3) == 256
    assert L.dfo_wei_off(0, 16, 0, 0, 64, 3, 3) == 9 * 256
    assert L.dfo_wei_off(16, 0, 0, 0, 64, 3, 3) == 4 * 9 * 256
    from dfb200 import layout
    w = np.arange(32 * 48 * 9, dtype=np.int64).astype(np.int8).reshape(32, 48, 3, 3)
    blocked = layout.oihw_to_blocked(w)
    for o, i, h, ww in [(0, 0, 0, 0), (17, 5, 2, 1), (31, 47, 1, 2), (16, 16, 0, 0)]:
        assert blocked[L.dfo_wei_off(o, i, h, ww, 48, 3, 3)] == w[o, i, h, ww]
    assert np.array_equal(layout.blocked_to_oihw(blocked, 32, 48, 3, 3), w)


# ------------------------------------------------------------------------------- acceptance
def test_conv_check_rules():
    L = O.lib()
    ok = O.make_desc(1, 56, 56, 64, 64, 256, O.U8, O.S32, O.S32, nscale0=64, nscale1=256)
    assert L.dfo_conv_check(C.byref(ok)) == 0
    for field, val in [("ic", 60), ("oc", 40), ("oc1", 100), ("nscale0", 3), ("nscale1", 7), ("dst_dt", 9), ("round0", 2)]:
        d = O.make_desc(1, 56, 56, 64, 64, 256, O.U8, O.S32, O.S32, nscale0=64, nscale1=256)
        setattr(d, field, val)
        assert L.dfo_conv_check(C.byref(d)) != 0, field
    d = O.make_desc(1, 56, 56, 64, 64, 256, O.U8, nscale0=64, nscale1=256, pad=6)  # l_pad > ur_w (=5)
    assert L.dfo_conv_check(C.byref(d)) != 0


# ------------------------------------------------------------ three implementations agree
def _oracle_conv(c, fn=O.conv):
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    d = O.make_desc(c.n, c.h, c.w, c.ic, c.oc, c.oc1, cases.DT[c.dst], cases.DT[c.b0], cases.DT[c.b1], relu0=c.relu0,
                    relu1=c.relu1, round0=c.r0, round1=c.r1, nscale0=s0.size, nscale1=s1.size)
    return fn(d, src, wb, b0, s0, w1b, b1, s1)


@pytest.mark.parametrize("c", cases.SMALL_CONV, ids=lambda c: c.name)
def test_conv_oracle_vs_numpy_model(c):
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    want = M.conv_fused(src, w0, b0, s0, w1, b1, s1, cases.DT[c.dst], relu0=c.relu0, relu1=c.relu1, down0=c.r0,
                        down1=c.r1)
    got = _oracle_conv(c)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("c", cases.SMALL_CONV, ids=lambda c: c.name)
def test_conv_oracle_vs_x86_replay(c):
    if not O.replay_supported():
        pytest.skip("host CPU lacks AVX-512 VNNI")
    assert np.array_equal(_oracle_conv(c).view(np.uint8), _oracle_conv(c, O.replay_conv).view(np.uint8))


def test_conv_oracle_vs_x86_replay_cfg1_full():
    if not O.replay_supported():
        pytest.skip("host CPU lacks AVX-512 VNNI")
    c = cases.FULL_CONV[0]
    assert np.array_equal(_oracle_conv(c), _oracle_conv(c, O.replay_conv))


def test_conv0_only_mode_all_dst():
    """conv without the 1x1 stage (reference 9-argument conv(), 'next' row of SURVEY §8f)."""
    c = cases.ConvCase("c0", 2, 7, 6, 32, 48, 0, b0="s32")
    src = c.tensors()[0]
    from dfb200 import layout, synth
    w0 = synth.wei_s8(2, (48, 32, 3, 3))
    b0 = synth.bias(4, 48, "s32")
    s0 = synth.channel_scales(48, 10)
    for dst in ("u8", "s8", "s32", "f32"):
        for relu0 in (0, 1):
            d = O.make_desc(2, 7, 6, 32, 48, 0, cases.DT[dst], O.S32, 0, relu0=relu0, nscale0=48)
            got = O.conv(d, src, layout.oihw_to_blocked(w0), b0, s0)
            want = M.conv_fused(src, w0, b0, s0, None, None, None, cases.DT[dst], relu0=relu0)
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (dst, relu0)
            if O.replay_supported():
                rep = O.replay_conv(d, src, layout.oihw_to_blocked(w0), b0, s0)
                assert np.array_equal(got.view(np.uint8), rep.view(np.uint8)), (dst, relu0)


def test_literal_f32_intermediate_switch_documents_defect_d3():
    c = cases.ConvCase("d3", 1, 4, 4, 16, 16, 16, "f32", None, None, k0=8)
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    d = O.make_desc(1, 4, 4, 16, 16, 16, O.F32, nscale0=16, nscale1=16)
    intended = O.conv_intermediate(d, src, wb, None, s0)
    d.literal_f32_intermediate = 1
    literal = O.conv_intermediate(d, src, wb, None, s0)
    # literally, vpmovusdb saturates the float BIT PATTERN: every positive value becomes 255
    assert set(np.unique(literal)) <= {0, 255}
    assert np.array_equal(literal == 0, intended == 0) or (literal[intended > 0] == 255).all()
    assert not np.array_equal(literal, intended)


# ---------------------------------------------------------------------------------- golden
def test_conv_golden_fixtures():
    g = np.load(os.path.join(GOLD, "conv_small.npz"))
    by_name = {c.name: c for c in cases.SMALL_CONV}
    assert len(g.files) >= 10
    for key in g.files:
        got = _oracle_conv(by_name[key[len("conv_"):]])
        assert np.array_equal(got.view(np.uint8), g[key].view(np.uint8)), key


def test_concat_golden_fixtures():
    g = np.load(os.path.join(GOLD, "concat_relu.npz"))
    for key in g.files:
        _, dt, ci, data = key.split("_")
        ins = cases.concat_inputs(dt, cases.CONCAT_BASIC[int(ci)][0], data)
        got = O.concat(cases.DT[dt], 1, ins)
        assert np.array_equal(got.view(np.uint8), g[key].view(np.uint8)), key


# ---------------------------------------------------------------------------------- concat
@pytest.mark.parametrize("dt", ["u8", "s8", "s32", "f32"])
@pytest.mark.parametrize("relu", [False, True])
def test_concat_reference_test_list(dt, relu):
    """On the reference's own shapes and data range the op is plain concat + true ReLU
    (what its MKL-DNN comparison checks, test/test_concat.cc:31-87)."""
    shapes = cases.CONCAT_BASIC + (cases.CONCAT_32BIT_EXTRA if dt in ("f32", "s32") else [])
    for srcs, dst in shapes:
        ins = cases.concat_inputs(dt, srcs, "reference-range")
        got = O.concat(cases.DT[dt], relu, ins)
        want = np.concatenate(ins, axis=-1)
        if relu:
            want = np.maximum(want, 0).astype(want.dtype)
        n, c, h, w = dst
        assert got.shape == (n, h, w, c)
        assert np.array_equal(got, want)
        if O.replay_supported():
            assert np.array_equal(O.replay_concat(cases.DT[dt], relu, ins).view(np.uint8), got.view(np.uint8))


def test_concat_literal_relu_quirks():
    """Outside the reference's data range the emitted vpmaxsb / vpmaxsw are NOT a true ReLU (C6)."""
    u = np.array([0, 127, 128, 200, 255] + [1] * 11, np.uint8).reshape(1, 1, 1, 16)
    assert O.concat(O.U8, True, [u]).reshape(-1)[:5].tolist() == [0, 127, 0, 0, 0]
    s = np.array([40000, -65531, 100000, -1, 7, 32767, 32768, -32768], np.int32).reshape(1, 1, 1, 8)
    assert O.concat(O.S32, True, [s]).reshape(-1).tolist() == [0, 5, 65536, 0, 7, 32767, 0, 0]
    f = np.array([-0.0, np.nan, -1.5, 2.0], np.float32).reshape(1, 1, 1, 4)
    r = O.concat(O.F32, True, [f]).reshape(-1)
    assert np.signbit(r[0]) and np.isnan(r[1]) and r[2] == 0 and r[3] == 2.0


def test_concat_block_rule():
    L = O.lib()

    def block(dt, ics):
        return L.dfo_concat_block(dt, len(ics), (C.c_int * len(ics))(*ics))

    assert block(O.U8, [64, 128, 32, 32]) == 32   # BASELINE configs[1] -> ymm
    assert block(O.U8, [64, 128]) == 64
    assert block(O.S8, [16, 48]) == 16
    assert block(O.U8, [24, 16]) == 0
    assert block(O.F32, [4, 8]) == 4 and block(O.S32, [16, 8]) == 8 and block(O.F32, [6]) == 0
