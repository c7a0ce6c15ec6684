"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the oracle.

Bar: bit-exact for u8 / s8 / s32 destinations; f32 destinations are expected bit-exact too (the
same two separately rounded f32 operations) and are asserted to <= 1e-6 relative with the maximum
ULP distance reported (the reference's own criterion is 1e-4 relative, test/test_utils.h:77-80).
Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest

import cases
import oracle_lib as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
F32_RTOL = 1e-6


@pytest.fixture(scope="module")
def df():
    import dfb200
    assert dfb200.device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    dfb200.set_device(0)
    return dfb200


def _gpu_conv(df, c, tensors=None):
    src, w0, w1, b0, b1, s0, s1 = tensors or c.tensors()
    wb, w1b = c.blocked(w0, w1)
    op = df.Conv(c.n, c.h, c.w, c.ic, c.oc, c.oc1, cases.DT[c.dst], wb, w1b, b0, b1, s0, s1, cases.DT[c.b0],
                 cases.DT[c.b1], relu0=bool(c.relu0), relu1=bool(c.relu1), round0=c.r0, round1=c.r1)
    out = op(src)
    op.close()
    return out


def _oracle_conv(c, tensors=None, fast=True):
    src, w0, w1, b0, b1, s0, s1 = tensors or c.tensors()
    wb, w1b = c.blocked(w0, w1)
    d = O.make_desc(c.n, c.h, c.w, c.ic, c.oc, c.oc1, cases.DT[c.dst], cases.DT[c.b0], cases.DT[c.b1], relu0=c.relu0,
                    relu1=c.relu1, round0=c.r0, round1=c.r1, nscale0=s0.size, nscale1=s1.size)
    fn = O.replay_conv if (fast and O.replay_supported()) else O.conv
    return fn(d, src, wb, b0, s0, w1b, b1, s1)


def _assert_same(got, want, dst):
    if dst == "f32":
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        denom = np.maximum(np.abs(want[fin]), 1e-30)
        rel = np.abs(got[fin] - want[fin]) / denom
        assert rel.max(initial=0.0) <= F32_RTOL, f"max rel err {rel.max()}"
        ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
        assert ulp.max() == 0, f"f32 output not bit-exact: max ULP distance {ulp.max()}"
    else:
        bad = np.argwhere(got != want)
        assert bad.size == 0, f"{len(bad)} mismatches, first at {bad[0].tolist()}: got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}"


@pytest.mark.parametrize("c", cases.SMALL_CONV, ids=lambda c: c.name)
def test_conv_small_vs_scalar_oracle(df, c):
    _assert_same(_gpu_conv(df, c), _oracle_conv(c, fast=False), c.dst)


def test_conv_golden_fixtures(df):
    g = np.load(os.path.join(GOLD, "conv_small.npz"))
    by_name = {c.name: c for c in cases.SMALL_CONV}
    for key in g.files:
        c = by_name[key[len("conv_"):]]
        _assert_same(_gpu_conv(df, c), g[key], c.dst)


@pytest.mark.parametrize("c", cases.FULL_CONV, ids=lambda c: c.name)
def test_conv_baseline_configs_full_size(df, c):
    """BASELINE.json configs at full size.  Checker: the AVX-512 replay (all images) when the host
    has VNNI, otherwise the scalar oracle on a sample of images."""
    t = c.tensors()
    got = _gpu_conv(df, c, t)
    if O.replay_supported():
        _assert_same(got, _oracle_conv(c, t), c.dst)
    else:
        idx = [0, c.n // 2, c.n - 1]
        sub = cases.ConvCase(c.name, len(idx), c.h, c.w, c.ic, c.oc, c.oc1, c.dst, c.b0, c.b1, c.r0, c.r1, c.relu0, c.relu1)
        ts = (np.ascontiguousarray(t[0][idx]),) + t[1:]
        _assert_same(got[idx], _oracle_conv(sub, ts, fast=False), c.dst)


def test_conv_batch_properties_at_full_size(df):
    """Size-independent properties on cfg3 (N=64): images are independent, so (a) a permuted batch
    gives the permuted output, (b) any sub-batch run through the same handle (n < created n) equals
    the corresponding slice, (c) the result does not depend on what else is in the batch."""
    c = cases.FULL_CONV[1]
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    op = df.Conv(c.n, c.h, c.w, c.ic, c.oc, c.oc1, cases.DT[c.dst], wb, w1b, b0, b1, s0, s1, cases.DT[c.b0], cases.DT[c.b1])
    full = op(src)
    perm = np.random.RandomState(0).permutation(c.n)
    assert np.array_equal(op(np.ascontiguousarray(src[perm])), full[perm])
    for n in (1, 3, 37):
        assert np.array_equal(op(np.ascontiguousarray(src[:n])), full[:n])
    other = src.copy()
    other[1:] = 255 - other[1:]
    assert np.array_equal(op(other)[0], full[0])
    # checksum of checksums, recorded in the test log for cross-run comparison
    print("cfg3 checksum", int(full.astype(np.uint64).sum()), int((full.astype(np.uint64) * (np.arange(full.size, dtype=np.uint64).reshape(full.shape) % 251)).sum()))
    op.close()


# conv() without the 1x1 stage (reference 9-argument overload, include/deepfusion.h:121-129)
CONV0_ONLY = [  # (n, h, w, ic, oc, bias dtype, scale exponent)
    (2, 7, 6, 32, 48, "s32", 10),     # ragged, one block of 32 columns + one of 16
    (3, 9, 11, 64, 64, "s8", 12),
    (2, 14, 14, 128, 128, None, 13),  # two accumulators, one chunk
    (2, 5, 9, 64, 256, "f32", 12),    # OC = 256: one accumulator, two chunks of 128 columns
    (1, 28, 28, 256, 144, "u8", 14),  # weights streamed, partial second chunk
]


@pytest.mark.parametrize("shape", CONV0_ONLY, ids=lambda s: f"{s[3]}to{s[4]}_{s[1]}x{s[2]}")
@pytest.mark.parametrize("dst", ["u8", "s8", "s32", "f32"])
def test_conv0_only_operator(df, shape, dst):
    n, h, w, ic, oc, bdt, k0 = shape
    from dfb200 import layout, synth
    src = synth.src_u8(1, (n, h, w, ic))
    w0 = synth.wei_s8(2, (oc, ic, 3, 3))
    b0 = synth.bias(4, oc, bdt) if bdt else None
    s0 = synth.channel_scales(oc, k0)
    wb = layout.oihw_to_blocked(w0)
    for relu0, r0 in ((0, 0), (1, 1)):
        op = df.Conv(n, h, w, ic, oc, 0, cases.DT[dst], wb, None, b0, None, s0, (1.0,), cases.DT[bdt], 0, relu0=bool(relu0), round0=r0)
        got = op(src)
        op.close()
        d = O.make_desc(n, h, w, ic, oc, 0, cases.DT[dst], cases.DT[bdt], 0, relu0=relu0, round0=r0, nscale0=oc)
        want = O.conv(d, src, wb, b0, s0)
        assert got.shape == want.shape == (n, h, w, oc)
        _assert_same(got, want, dst)


def test_conv_zero_input_gives_bias_only(df):
    c = cases.ConvCase("zero", 2, 10, 10, 64, 64, 128, "s32", "s32", "s32")
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    src[...] = 0
    got = _gpu_conv(df, c, (src, w0, w1, b0, b1, s0, s1))
    _assert_same(got, _oracle_conv(c, (src, w0, w1, b0, b1, s0, s1)), "s32")
    assert (got == got[0, 0, 0]).all()  # every pixel sees the same (bias-only) pipeline


def test_conv_nonfinite_scale_takes_the_nan_safe_kernel(df):
    c = cases.ConvCase("nan", 1, 6, 6, 32, 32, 48, "u8", "f32", "f32")
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    s0 = s0.copy(); s1 = s1.copy(); b0 = b0.copy()
    s0[3] = np.inf; s0[5] = np.nan; b0[7] = -np.inf; s1[2] = np.nan; s1[11] = -np.inf
    t = (src, w0, w1, b0, b1, s0, s1)
    _assert_same(_gpu_conv(df, c, t), _oracle_conv(c, t, fast=False), "u8")
    for dst in ("s8", "s32", "f32"):
        c2 = cases.ConvCase("nan", 1, 6, 6, 32, 32, 48, dst, "f32", "f32")
        got, want = _gpu_conv(df, c2, t), _oracle_conv(c2, t, fast=False)
        if dst == "f32":
            # NaN sign/payload is the one thing not reproduced (x86 yields 0xFFC00000 "real
            # indefinite", the GPU 0x7FFFFFFF); NaN positions and every non-NaN bit must match
            assert np.array_equal(np.isnan(got), np.isnan(want))
            ok = ~np.isnan(want)
            assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32))
        else:
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), dst


def test_conv_empty_batch_is_a_no_op(df):
    c = cases.SMALL_CONV[0]
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    op = df.Conv(c.n, c.h, c.w, c.ic, c.oc, c.oc1, cases.DT[c.dst], wb, w1b, b0, b1, s0, s1)
    buf = df.DeviceBuffer(64)
    op.run(buf, buf, n=0)
    with pytest.raises(df.DfError):
        op.run(buf, buf, n=c.n + 1)
    op.close()


# ---------------------------------------------------------------------------------- concat
@pytest.mark.parametrize("dt", ["u8", "s8", "s32", "f32"])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("data", ["reference-range", "full"])
def test_concat_reference_test_list(df, dt, relu, data):
    shapes = cases.CONCAT_BASIC + (cases.CONCAT_32BIT_EXTRA if dt in ("f32", "s32") else [])
    for srcs, _ in shapes:
        ins = cases.concat_inputs(dt, srcs, data)
        got = df.concat(ins, cases.DT[dt], relu)
        want = O.concat(cases.DT[dt], relu, ins)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (srcs, data)


def test_concat_golden_fixtures(df):
    g = np.load(os.path.join(GOLD, "concat_relu.npz"))
    for key in g.files:
        _, dt, ci, data = key.split("_")
        ins = cases.concat_inputs(dt, cases.CONCAT_BASIC[int(ci)][0], data)
        assert np.array_equal(df.concat(ins, cases.DT[dt], True).view(np.uint8), g[key].view(np.uint8)), key


def test_concat_cfg2_full_size_and_idempotence(df):
    srcs, _ = cases.CONCAT_CFG2
    ins = cases.concat_inputs("u8", srcs, "full")
    got = df.concat(ins, cases.DT["u8"], True)
    assert np.array_equal(got, O.concat(O.U8, True, ins))
    # ReLU(ReLU(x)) == ReLU(x): concatenating the already-clamped slices again changes nothing
    again = df.concat([np.ascontiguousarray(got[..., a:b]) for a, b in ((0, 64), (64, 192), (192, 224), (224, 256))],
                      cases.DT["u8"], True)
    assert np.array_equal(again, got)
    # without ReLU the op is a pure byte permutation: multiset of bytes preserved
    plain = df.concat(ins, cases.DT["u8"], False)
    assert np.array_equal(np.bincount(plain.reshape(-1), minlength=256),
                          sum(np.bincount(i.reshape(-1), minlength=256) for i in ins))


def test_concat_many_inputs_and_odd_widths(df):
    ics = [16, 48, 16, 80, 32, 16, 112, 16, 16, 64, 16, 16, 48, 16, 32, 16, 16, 96, 16]  # 19 inputs > one launch group
    ins = cases.concat_inputs("s8", [(3, c, 5, 7) for c in ics], "full")
    for relu in (False, True):
        assert np.array_equal(df.concat(ins, cases.DT["s8"], relu), O.concat(O.S8, relu, ins))
    ins32 = cases.concat_inputs("s32", [(2, c, 3, 3) for c in (4, 12, 20, 8)], "full")
    assert np.array_equal(df.concat(ins32, cases.DT["s32"], True), O.concat(O.S32, True, ins32))


# ------------------------------------------------------------------- alternative kernel paths
def _run_cfg3_in_subprocess(env_extra):
    """cfg3 (N=8) through the C-ABI in a fresh process with a kernel-selection hook set."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys
sys.path.insert(0, {os.path.join(root, 'tests')!r}); sys.path.insert(0, {os.path.join(root, 'deep-fusion_b200')!r})
import numpy as np, cases, dfb200 as df, oracle_lib as O
c = cases.ConvCase("alt", 8, 28, 28, 128, 128, 512, "u8", "s32", "s32")
src, w0, w1, b0, b1, s0, s1 = c.tensors()
wb, w1b = c.blocked(w0, w1)
op = df.Conv(c.n, c.h, c.w, c.ic, c.oc, c.oc1, df.U8, wb, w1b, b0, b1, s0, s1, df.S32, df.S32)
i = op.info()
got = op(src)
d = O.make_desc(c.n, c.h, c.w, c.ic, c.oc, c.oc1, O.U8, O.S32, O.S32, nscale0=c.oc, nscale1=c.oc1)
fn = O.replay_conv if O.replay_supported() else O.conv
want = fn(d, src, wb, b0, s0, w1b, b1, s1)
print("RES", i.w0_resident, i.w1_resident, "OK" if np.array_equal(got, want) else "MISMATCH")
"""
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_conv_cta_pair_kernel_is_the_default_for_cfg3(df):
    """cta_group::2 kernel (weight halves resident in a CTA pair): the default for the cfg3 shape;
    DF_PAIR=0 selects the single-CTA kernels (weights streamed), which must give the same bytes."""
    out = _run_cfg3_in_subprocess({})
    assert "RES 2 2 OK" in out, out
    out = _run_cfg3_in_subprocess({"DF_PAIR": "0"})
    assert "RES 0 1 OK" in out, out


def test_conv_store_and_launch_variants(df):
    """Direct (unstaged) stores of 1-byte output, per-channel offset-magic constants, launches without
    programmatic dependent launch: all bit-identical to the default path."""
    for env in ({"DF_FORCE_DYNAMIC_GEOMETRY": "1", "DF_NO_STAGED_STORE": "1"}, {"DF_NO_UNIFORM_K": "1"},
                {"DF_NO_UNIFORM_K": "1", "DF_NO_STAGED_STORE": "1"}, {"DF_NO_PDL": "1"}):
        assert "OK" in _run_cfg3_in_subprocess(env), env


def test_conv_back_to_back_launches_overlap_safely(df):
    """Programmatic dependent launch lets launch k+1 start its prologue under the tail of launch k.  Chain
    launches that reuse the same destination and feed on different sources: every result must be the one
    of its own source (no early read, no late write)."""
    c = cases.ConvCase("pdl", 16, 28, 28, 128, 128, 512, "u8", "s32", "s32")
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    op = df.Conv(c.n, c.h, c.w, c.ic, c.oc, c.oc1, cases.DT[c.dst], wb, w1b, b0, b1, s0, s1, cases.DT[c.b0], cases.DT[c.b1])
    srcs = [src, np.ascontiguousarray(src[::-1]), (255 - src)]
    want = [op(s) for s in srcs]                      # one launch at a time
    sb = [df.DeviceBuffer.from_numpy(s) for s in srcs]
    outs = [df.DeviceBuffer(want[0].size) for _ in range(2)]
    for rep in range(20):                             # many launches in flight, two destinations reused
        for k in range(3):
            op.run(sb[k], outs[(rep * 3 + k) % 2])
    last = (19 * 3 + 2) % 2
    df.sync()
    assert np.array_equal(outs[last].download(want[2].shape, np.uint8), want[2])
    assert np.array_equal(outs[1 - last].download(want[1].shape, np.uint8), want[1])
    op.close()


def test_conv_generic_geometry_and_i2f_paths(df):
    """The run-time-geometry kernel and the plain I2F conv1 epilogue on a BASELINE shape."""
    assert "OK" in _run_cfg3_in_subprocess({"DF_FORCE_DYNAMIC_GEOMETRY": "1"})
    assert "OK" in _run_cfg3_in_subprocess({"DF_NO_FAST_CONV1": "1"})
    assert "OK" in _run_cfg3_in_subprocess({"DF_FORCE_DYNAMIC_GEOMETRY": "1", "DF_NO_FAST_CONV1": "1"})


# ---------------------------------------------------------------- concat fused into the conv's A-operand load
# op_concat (src/op_concat.cc:22-72) feeding op_conv (src/op_conv.cc:140-260) as ONE kernel.  The
# checker is the oracle's concat followed by the oracle's conv on the concatenated tensor; full-range u8 inputs,
# so the literal ReLU (bytes >= 128 -> 0) is exercised.
CONCAT_CONV = [
    # name, n, h, w, input channels, oc, oc1, dst
    ("two_32", 2, 7, 9, (32, 32), 48, 80, "u8"),
    ("three_mixed", 2, 9, 7, (64, 32, 96), 64, 144, "s8"),
    ("one_input", 2, 6, 6, (64,), 32, 64, "u8"),
    ("four_128", 1, 8, 8, (128, 128, 128, 128), 64, 128, "s32"),
    ("cfg2_crop", 2, 28, 28, (64, 128, 32, 32), 64, 256, "u8"),
    ("cfg2_conv0_only", 2, 14, 28, (64, 128, 32, 32), 128, 0, "u8"),
    ("cfg2_f32", 1, 12, 28, (64, 128, 32, 32), 128, 256, "f32"),
]


@pytest.mark.parametrize("relu", [False, True], ids=["copy", "relu"])
@pytest.mark.parametrize("case", CONCAT_CONV, ids=lambda c: c[0])
def test_concat_fused_into_conv(df, case, relu):
    name, n, h, w, ics, oc, oc1, dst = case
    ic = sum(ics)
    srcs = [cases.synth.uniform_int(20 + i, (n, h, w, c), 0, 255, np.uint8) for i, c in enumerate(ics)]
    w0 = cases.synth.wei_s8(2, (oc, ic, 3, 3))
    wb = cases.layout.oihw_to_blocked(w0)
    b0 = cases.synth.bias(4, oc, "s32")
    k0 = {64: 12, 96: 12, 128: 13, 192: 13, 256: 14, 512: 15}.get(ic, 12) + (1 if not relu else 0)
    s0 = cases.synth.channel_scales(oc, k0)
    if oc1:
        w1 = cases.synth.wei_s8(3, (oc1, oc))
        w1b = cases.layout.oihw_to_blocked(w1.reshape(oc1, oc, 1, 1))
        b1 = cases.synth.bias(5, oc1, "s32")
        s1 = cases.synth.channel_scales(oc1, 12)
    else:
        w1b, b1, s1 = None, None, np.array([1.0], np.float32)
    cat = O.concat(O.U8, relu, srcs)
    d = O.make_desc(n, h, w, ic, oc, oc1, cases.DT[dst], O.S32, O.S32 if oc1 else O.UNDEF, nscale0=oc, nscale1=s1.size)
    fn = O.replay_conv if (O.replay_supported() and oc1) else O.conv
    want = fn(d, cat, wb, b0, s0, w1b, b1, s1)
    op = df.ConcatConv(n, h, w, ics, relu, oc, oc1, cases.DT[dst], wb, w1b, b0, b1, s0, s1, df.S32, df.S32 if oc1 else df.UNDEF)
    got = op(srcs)
    op.close()
    _assert_same(got, want, dst)
    assert want.any(), "degenerate case: the expected output is all zero"


def test_concat_fused_rejects_what_it_cannot_load(df):
    """Channel counts that are multiples of 16 but not of 32 are valid for concat (jit_concat_kernel.cc:157-176) but
    not for the fused load: DF_E_UNSUPPORTED, so that the caller runs the two ops instead -- never a wrong result."""
    wb = cases.layout.oihw_to_blocked(cases.synth.wei_s8(2, (32, 48, 3, 3)))
    with pytest.raises(df.DfError) as e:
        df.ConcatConv(1, 4, 4, (16, 32), True, 32, 0, df.U8, wb, None, None, None, np.ones(32, np.float32))
    assert e.value.code == -2


# ------------------------------------------------------------------------------ general conv0 windows
# jit_conv_kernel::init_conf (src/jit_conv_kernel.cc:586-661) accepts any kh x kw window; the
# B200 kernel runs every stride-1 window whose output is not larger than its input (2 p <= k - 1), for the
# conv-only operator and for the fused pair alike.  Checker: the scalar oracle (general in k / s / p).
WINDOWS = [
    # kh, kw, ph, pw
    (1, 1, 0, 0), (3, 3, 0, 0), (3, 3, 1, 0), (5, 5, 2, 2), (5, 5, 1, 2), (7, 7, 3, 3), (1, 3, 0, 1), (3, 1, 1, 0),
    (2, 2, 0, 0), (1, 7, 0, 3),
]


def _window_case(kh, kw, ic, oc, seed=2):
    w0 = cases.synth.wei_s8(seed, (oc, ic, kh, kw))
    return cases.layout.oihw_to_blocked(w0)


@pytest.mark.parametrize("dst", ["u8", "s32"])
@pytest.mark.parametrize("win", WINDOWS, ids=lambda w: "k%dx%d_p%dx%d" % w)
def test_conv0_only_general_window(df, win, dst):
    kh, kw, ph, pw = win
    n, h, w, ic, oc = 2, 11, 13, 32, 48
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = _window_case(kh, kw, ic, oc)
    b0 = cases.synth.bias(4, oc, "s32")
    s0 = cases.synth.channel_scales(oc, 10 + (kh * kw > 8))
    d = O.make_desc(n, h, w, ic, oc, 0, cases.DT[dst], O.S32, O.UNDEF, k=(kh, kw), pad=(ph, pw), relu0=1, nscale0=oc)
    want = O.conv(d, src, wb, b0, s0)
    op = df.Conv(n, h, w, ic, oc, 0, cases.DT[dst], wb, None, b0, None, s0, (1.0,), df.S32, df.UNDEF, relu0=True,
                 k=(kh, kw), pad=(ph, pw))
    got = op(src)
    op.close()
    assert got.shape == want.shape
    _assert_same(got, want, dst)
    assert want.any()


@pytest.mark.parametrize("win", [(1, 1, 0, 0), (5, 5, 2, 2), (3, 3, 0, 0), (1, 3, 0, 1), (7, 7, 3, 3)], ids=lambda w: "k%dx%d_p%dx%d" % w)
def test_fused_conv_general_window(df, win):
    """kh x kw conv + ReLU + 1x1 conv + ReLU: the reference's fused operator with another first-stage window."""
    kh, kw, ph, pw = win
    n, h, w, ic, oc, oc1 = 2, 12, 10, 64, 64, 144
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = _window_case(kh, kw, ic, oc)
    w1b = cases.layout.oihw_to_blocked(cases.synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1))
    b0, b1 = cases.synth.bias(4, oc, "s32"), cases.synth.bias(5, oc1, "s32")
    s0, s1 = cases.synth.channel_scales(oc, 10 + (kh * kw > 8) + (kh * kw > 24)), cases.synth.channel_scales(oc1, 12)
    d = O.make_desc(n, h, w, ic, oc, oc1, O.U8, O.S32, O.S32, k=(kh, kw), pad=(ph, pw), nscale0=oc, nscale1=oc1)
    want = O.conv(d, src, wb, b0, s0, w1b, b1, s1)
    op = df.Conv(n, h, w, ic, oc, oc1, df.U8, wb, w1b, b0, b1, s0, s1, df.S32, df.S32, k=(kh, kw), pad=(ph, pw))
    got = op(src)
    op.close()
    _assert_same(got, want, "u8")
    assert want.any()


def test_one_by_one_conv_full_width(df):
    """1x1 p0: no zero columns at all (Wp == W when W is aligned) -- every tile row is a real pixel."""
    n, h, w, ic, oc = 3, 8, 16, 128, 256
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = _window_case(1, 1, ic, oc)
    s0 = cases.synth.channel_scales(oc, 9)
    d = O.make_desc(n, h, w, ic, oc, 0, O.U8, O.UNDEF, O.UNDEF, k=1, pad=0, nscale0=oc)
    want = O.conv(d, src, wb, None, s0)
    op = df.Conv(n, h, w, ic, oc, 0, df.U8, wb, None, None, None, s0, (1.0,), k=1, pad=0)
    assert abs(op.info().mma_efficiency - 1.0) < 0.2
    got = op(src)
    op.close()
    _assert_same(got, want, "u8")


# ------------------------------------------------- strides, wide rows, more channels than one accumulator holds
def _conv0_only_case(df, n, h, w, ic, oc, k, stride, pad, dst="u8", k0=None, relu0=1):
    kk = (k, k) if isinstance(k, int) else k
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = cases.layout.oihw_to_blocked(cases.synth.wei_s8(2, (oc, ic) + tuple(kk)))
    b0 = cases.synth.bias(4, oc, "s32")
    if k0 is None:
        k0 = int(np.ceil(np.log2(ic * kk[0] * kk[1] * 64.0))) - 4
    s0 = cases.synth.channel_scales(oc, k0)
    d = O.make_desc(n, h, w, ic, oc, 0, cases.DT[dst], O.S32, O.UNDEF, k=k, stride=stride, pad=pad, relu0=relu0, nscale0=oc)
    want = O.conv(d, src, wb, b0, s0)
    op = df.Conv(n, h, w, ic, oc, 0, cases.DT[dst], wb, None, b0, None, s0, (1.0,), df.S32, df.UNDEF, relu0=bool(relu0),
                 k=k, stride=stride, pad=pad)
    got = op(src)
    op.close()
    assert got.shape == want.shape
    _assert_same(got, want, dst)
    assert want.any()


@pytest.mark.parametrize("dst", ["u8", "f32"])
@pytest.mark.parametrize("k,stride,pad", [(3, 2, 1), (1, 2, 0), (7, 2, 3), (3, (2, 1), 1), (2, 2, 0), (5, (1, 3), 2), (3, 2, 0)],
                         ids=lambda v: str(v).replace(" ", ""))
def test_conv0_only_strided(df, k, stride, pad, dst):
    """Strided windows (accepted by jit_conv_kernel::init_conf, pinned to the reference generators in
    tests/test_ref_pin.py): stride-1 arithmetic, strided store."""
    _conv0_only_case(df, 2, 13, 15, 32, 48, k, stride, pad, dst)


@pytest.mark.parametrize("w", [255, 256, 300, 520])
def test_conv0_only_wide_rows(df, w):
    """Rows wider than one TMA box (256 positions): several boxes per halo row."""
    _conv0_only_case(df, 1, 3, w, 16, 32, 3, 1, 1, "u8")


@pytest.mark.parametrize("dst", ["u8", "s8", "s32"])
@pytest.mark.parametrize("oc,k,pad", [(320, 3, 1), (512, 1, 0), (272, 3, 1), (784, 1, 0)])
def test_conv0_only_many_output_channels(df, oc, k, pad, dst):
    """oc > 256: groups of <= 256 channels, one launch each, into channel ranges of the same pixels."""
    _conv0_only_case(df, 2, 7, 9, 32, oc, k, 1, pad, dst, relu0=0 if dst != "u8" else 1)


def test_conv0_only_deep_input_single_halo_stage(df):
    """28x28 with 512 input channels: one halo stage is 119 KB, so the kernel runs with a single stage."""
    _conv0_only_case(df, 1, 28, 28, 512, 32, 3, 1, 1, "u8")


@pytest.mark.parametrize("case", [
    # n, h, w, ic, oc, oc1, k, stride, pad, dst
    (2, 7, 9, 32, 288, 272, 3, 1, 1, "u8"),       # first stage wider than one accumulator -> two chained launches
    (2, 7, 7, 64, 512, 1040, 3, 1, 1, "s32"),     # ... and a second stage in channel groups
    (2, 13, 15, 32, 64, 144, 3, 2, 1, "u8"),      # strided first stage inside the fused kernel
    (1, 9, 300, 16, 32, 48, 3, 1, 1, "u8"),       # wide rows inside the fused kernel
    (2, 14, 14, 48, 320, 96, 1, 1, 0, "f32"),
], ids=lambda c: "x".join(str(v) for v in c))
def test_fused_conv_beyond_one_accumulator(df, case):
    n, h, w, ic, oc, oc1, k, stride, pad, dst = case
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = cases.layout.oihw_to_blocked(cases.synth.wei_s8(2, (oc, ic, k, k)))
    w1b = cases.layout.oihw_to_blocked(cases.synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1))
    b0, b1 = cases.synth.bias(4, oc, "s32"), cases.synth.bias(5, oc1, "s32")
    s0 = cases.synth.channel_scales(oc, int(np.ceil(np.log2(ic * k * k * 64.0))) - 4)
    s1 = cases.synth.channel_scales(oc1, int(np.ceil(np.log2(oc * 64.0 * 64))) - 6)
    d = O.make_desc(n, h, w, ic, oc, oc1, cases.DT[dst], O.S32, O.S32, k=k, stride=stride, pad=pad, nscale0=oc, nscale1=oc1)
    want = O.conv(d, src, wb, b0, s0, w1b, b1, s1)
    op = df.Conv(n, h, w, ic, oc, oc1, cases.DT[dst], wb, w1b, b0, b1, s0, s1, df.S32, df.S32, k=k, stride=stride, pad=pad)
    got = op(src)
    op.close()
    assert got.shape == want.shape
    _assert_same(got, want, dst)
    assert want.any()


# --------------------------------------------------- the reference's planned operators (README.md:64-65)
# conv+relu+pooling and eltwise-sum+relu are listed, not implemented, in the reference; its test file only runs the
# MKL-DNN yardstick (test/test_conv_relu_pooling.cc).  Semantics are defined in include/dfcuda.h and restated by the
# oracle (dfo_pool / dfo_conv_sum, "parity unpinned" there); these tests hold the CUDA path to that restatement.
def _pool_input(dt, shape, seed=31):
    if dt == "f32":
        a = (cases.synth.uniform_int(seed, shape, -100000, 100000, np.int32).astype(np.float32) / np.float32(7))
        a.reshape(-1)[::97] = -0.0
        return a
    lo, hi = {"u8": (0, 255), "s8": (-128, 127), "s32": (-(2 ** 31), 2 ** 31 - 1)}[dt]
    return cases.synth.uniform_int(seed, shape, lo, hi, cases.NPDT[dt])


POOLS = [
    # n, h, w, c, k, stride, pad, out_hw
    (2, 4, 4, 16, 2, 2, 0, None),            # test_conv_relu_pooling.cc:314-315 after the conv: 2x2 -> 1x1 ...
    (2, 14, 14, 64, 2, 2, 0, None),
    (2, 7, 7, 128, 7, 7, 0, None),           # :341 global average
    (2, 9, 11, 32, 3, 2, 1, None),
    (2, 9, 11, 32, 3, 2, 1, (5, 6)),         # windows past the bottom / right edge (mkldnn padR)
    (1, 5, 5, 16, (2, 3), (1, 2), (1, 0), None),
]


@pytest.mark.parametrize("round_mode", [0, 1], ids=["rn", "rd"])
@pytest.mark.parametrize("kind", [0, 1, 2], ids=["max", "avg_incl", "avg_excl"])
@pytest.mark.parametrize("dt", ["u8", "s8", "s32", "f32"])
@pytest.mark.parametrize("case", POOLS, ids=lambda c: "x".join(str(v) for v in c[:7]).replace(" ", ""))
def test_pool(df, case, dt, kind, round_mode):
    n, h, w, c, k, stride, pad, out_hw = case
    if kind == 0 and round_mode == 1:
        pytest.skip("max pooling does not round")
    src = _pool_input(dt, (n, h, w, c))
    want = O.pool(src, kind, k, stride, pad, out_hw, round_mode)
    got = df.pool(src, cases.DT[dt], kind, k, stride, pad, out_hw, round_mode)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("dst", ["u8", "s8", "s32", "f32"])
@pytest.mark.parametrize("shape", [(2, 9, 7, 32, 48, 0, 3, 1), (2, 7, 7, 64, 320, 0, 1, 0), (2, 8, 8, 64, 64, 144, 3, 1),
                                   (1, 7, 7, 32, 288, 272, 3, 1)], ids=lambda s: "x".join(str(v) for v in s))
def test_conv_eltwise_sum(df, shape, dst):
    """conv (+1x1) + residual + ReLU: conv-only, conv-only in channel groups, fused, fused as chained launches."""
    n, h, w, ic, oc, oc1, k, pad = shape
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = cases.layout.oihw_to_blocked(cases.synth.wei_s8(2, (oc, ic, k, k)))
    b0 = cases.synth.bias(4, oc, "s32")
    s0 = cases.synth.channel_scales(oc, int(np.ceil(np.log2(ic * k * k * 64.0))) - 4)
    if oc1:
        w1b = cases.layout.oihw_to_blocked(cases.synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1))
        b1 = cases.synth.bias(5, oc1, "s32")
        s1 = cases.synth.channel_scales(oc1, int(np.ceil(np.log2(oc * 64.0 * 64))) - 6)
    else:
        w1b, b1, s1 = None, None, np.array([1.0], np.float32)
    oh, ow = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    res = _pool_input(dst, (n, oh, ow, oc1 or oc), seed=41)
    if dst == "s32":
        res = (res >> 20).astype(np.int32)  # comparable in size to the conv's output, so the add matters
    if dst == "f32":
        res = (res / np.float32(64)).astype(np.float32)
    d = O.make_desc(n, h, w, ic, oc, oc1, cases.DT[dst], O.S32, O.S32 if oc1 else O.UNDEF, k=k, pad=pad, relu0=1, relu1=1,
                    nscale0=oc, nscale1=s1.size)
    want = O.conv_sum(d, src, wb, b0, s0, res, w1b, b1, s1)
    plain = O.conv(d, src, wb, b0, s0, w1b, b1, s1)
    assert not np.array_equal(want, plain), "degenerate case: the residual does not change the result"
    op = df.Conv(n, h, w, ic, oc, oc1, cases.DT[dst], wb, w1b, b0, b1, s0, s1, df.S32, df.S32 if oc1 else df.UNDEF,
                 relu0=True, relu1=True, k=k, pad=pad, with_sum=True)
    got = op(src, res)
    op.close()
    _assert_same(got, want, dst)


# ------------------------------------------------------------------ geometry edges of the padded pixel space
# (advisor, round 1: widths at the TMA box limit, Wp values for which 128 % Wp hits the halo-row edge, tall images,
# many conv1 chunks)
EDGE_GEOMETRY = [
    # n, h, w, ic, oc, oc1
    (2, 5, 2, 16, 16, 32), (1, 4, 42, 64, 32, 48), (1, 3, 128, 128, 32, 32), (1, 2, 253, 16, 32, 32), (1, 2, 254, 64, 16, 32),
    (1, 2, 255, 128, 16, 16), (1, 3, 127, 16, 32, 32), (1, 3, 63, 32, 16, 16), (1, 70, 3, 16, 16, 16), (1, 5, 5, 32, 32, 1040),
    (3, 1, 129, 16, 16, 32), (1, 3, 31, 128, 32, 32),
]


@pytest.mark.parametrize("g", EDGE_GEOMETRY, ids=lambda g: "x".join(map(str, g)))
def test_fused_conv_geometry_edges(df, g):
    n, h, w, ic, oc, oc1 = g
    c = cases.ConvCase("edge", n, h, w, ic, oc, oc1, "u8", "s32", "s32")
    got = _gpu_conv(df, c)
    want = _oracle_conv(c, fast=False)
    _assert_same(got, want, "u8")
    assert want.any()


# ------------------------------------------- run-time geometry on CTA pairs (weights that have to stream)
PAIR_DYN = [
    # n, h, w, ic, oc, oc1, k, stride, pad, dst
    (3, 28, 28, 256, 128, 512, 3, 1, 1, "u8"),     # the conv behind BASELINE configs[1]'s concat
    (5, 14, 14, 256, 256, 0, 3, 1, 1, "u8"),       # conv-only, 576 KB of weights
    (2, 14, 14, 256, 256, 0, 3, 1, 1, "s32"),
    (3, 12, 12, 128, 128, 272, 5, 1, 2, "s8"),     # 5x5 window, ragged last conv1 chunk
    (2, 20, 20, 128, 64, 144, 7, 2, 3, "u8"),      # 7x7 stride 2
    (7, 7, 7, 256, 208, 528, 3, 1, 1, "f32"),      # odd tile count (one tile of the last pair is empty), oc not a multiple of 32
    (1, 9, 40, 192, 96, 0, 3, 1, 1, "u8"),
    # the interleaved issue / ring order (GEMM2 chunks and GEMM1 taps alternate) at its edges:
    (40, 14, 14, 256, 256, 1024, 1, 1, 0, "u8"),   # ONE tap, eight chunks: nothing to interleave after the leading tap; many tiles per pair
    (9, 13, 13, 256, 256, 1024, 2, 1, 0, "s32"),   # four taps over eight chunks (2x2 window)
    (33, 14, 14, 256, 256, 128, 3, 1, 1, "u8"),    # nine taps behind ONE chunk
]


@pytest.mark.parametrize("case", PAIR_DYN, ids=lambda c: "x".join(str(v) for v in c))
def test_dynamic_geometry_on_cta_pairs(df, case):
    """Shapes whose weights do not fit one CTA's shared memory run conv_pair_kernel<DynGeom> (weight halves streamed by
    a CTA pair); bit-exact like everything else."""
    n, h, w, ic, oc, oc1, k, stride, pad, dst = case
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = cases.layout.oihw_to_blocked(cases.synth.wei_s8(2, (oc, ic, k, k)))
    b0 = cases.synth.bias(4, oc, "s32")
    s0 = cases.synth.channel_scales(oc, int(np.ceil(np.log2(ic * k * k * 64.0))) - 4)
    if oc1:
        w1b = cases.layout.oihw_to_blocked(cases.synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1))
        b1 = cases.synth.bias(5, oc1, "s32")
        s1 = cases.synth.channel_scales(oc1, int(np.ceil(np.log2(oc * 64.0 * 64))) - 6)
    else:
        w1b, b1, s1 = None, None, np.array([1.0], np.float32)
    d = O.make_desc(n, h, w, ic, oc, oc1, cases.DT[dst], O.S32, O.S32 if oc1 else O.UNDEF, k=k, stride=stride, pad=pad, relu0=1,
                    nscale0=oc, nscale1=s1.size)
    want = O.conv(d, src, wb, b0, s0, w1b, b1, s1)
    op = df.Conv(n, h, w, ic, oc, oc1, cases.DT[dst], wb, w1b, b0, b1, s0, s1, df.S32, df.S32 if oc1 else df.UNDEF, relu0=True,
                 k=k, stride=stride, pad=pad)
    assert op.info().w0_resident == 3, "expected the CTA-pair kernel with streamed weight halves"
    got = op(src)
    got2 = op(src)
    op.close()
    _assert_same(got, want, dst)
    _assert_same(got2, want, dst)
    assert want.any()


# --------------------------------------------------------------- inputs too deep for one halo stage (K-sliced)
DEEP = [
    # n, h, w, ic, oc, oc1, k, pad, dst
    (2, 7, 7, 2048, 256, 0, 1, 0, "u8"),        # ResNet stage-4 1x1: a 7x7 x 2048 halo is 286 KB
    (1, 7, 7, 2048, 2048, 0, 1, 0, "s32"),      # test/test_conv_relu_pooling.cc:341 (1x1 2048 -> 2048), eight channel groups
    (2, 14, 14, 1024, 64, 0, 3, 1, "u8"),
    (2, 9, 9, 1536, 64, 144, 3, 1, "u8"),       # fused, three K-slices
]


@pytest.mark.parametrize("case", DEEP, ids=lambda c: "x".join(str(v) for v in c))
def test_deep_inputs_are_k_sliced(df, case):
    n, h, w, ic, oc, oc1, k, pad, dst = case
    src = cases.synth.src_u8(1, (n, h, w, ic))
    wb = cases.layout.oihw_to_blocked(cases.synth.wei_s8(2, (oc, ic, k, k)))
    b0 = cases.synth.bias(4, oc, "s32")
    s0 = cases.synth.channel_scales(oc, int(np.ceil(np.log2(ic * k * k * 64.0))) - 4)
    if oc1:
        w1b = cases.layout.oihw_to_blocked(cases.synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1))
        b1 = cases.synth.bias(5, oc1, "s32")
        s1 = cases.synth.channel_scales(oc1, 12)
    else:
        w1b, b1, s1 = None, None, np.array([1.0], np.float32)
    d = O.make_desc(n, h, w, ic, oc, oc1, cases.DT[dst], O.S32, O.S32 if oc1 else O.UNDEF, k=k, pad=pad, relu0=1, nscale0=oc,
                    nscale1=s1.size)
    want = O.conv(d, src, wb, b0, s0, w1b, b1, s1)
    op = df.Conv(n, h, w, ic, oc, oc1, cases.DT[dst], wb, w1b, b0, b1, s0, s1, df.S32, df.S32 if oc1 else df.UNDEF, relu0=True,
                 k=k, pad=pad)
    got = op(src)
    op.close()
    _assert_same(got, want, dst)
    assert want.any()
