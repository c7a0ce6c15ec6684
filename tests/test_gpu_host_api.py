"""GPU tests (-m gpu) through the reference-facing C++ API (memory / concat() / conv() / submit()),
written like the reference's own test/test_concat.cc: build memories, fill them, create the op,
submit(), compare the destination's HOST buffer with the oracle."""
import numpy as np
import pytest

import cases
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    import dfb200
    assert dfb200.device_count() >= 1
    from dfb200 import hostapi
    return hostapi


@pytest.mark.parametrize("dt", ["u8", "s8", "s32", "f32"])
def test_concat_submit_like_reference_test(H, dt):
    shapes = cases.CONCAT_BASIC + (cases.CONCAT_32BIT_EXTRA if dt in ("f32", "s32") else [])
    for srcs_dims, dst_dims in shapes:
        ins = cases.concat_inputs(dt, srcs_dims, "reference-range")
        srcs = [H.Memory(d, "nhwc", dt) for d in srcs_dims]
        for m, a in zip(srcs, ins):
            m.set(a)
        dst = H.Memory(dst_dims, "nhwc", dt)
        for post_relu in (True, False):  # same order as test/test_concat.cc:103-107
            c = H.concat(srcs, dst, post_relu)
            c.submit()
            want = O.concat(cases.DT[dt], post_relu, ins)
            assert np.array_equal(dst.array().view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("name", ["cfg1_crop", "cfg3_crop", "cfg4_f32", "s8_relu_down", "single_scale"])
def test_conv_submit(H, name):
    c = {x.name: x for x in cases.SMALL_CONV}[name]
    src_a, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    src = H.Memory((c.n, c.ic, c.h, c.w), "nhwc", "u8")
    src.set(src_a)
    wei = H.Memory((c.oc, c.ic, 3, 3), "OIhw4i16o4i", "s8")
    wei.array().reshape(-1)[...] = wb
    wei1 = H.Memory((c.oc1, c.oc, 1, 1), "OIhw4i16o4i", "s8")
    wei1.array().reshape(-1)[...] = w1b
    bia = bia1 = None
    if c.b0:
        bia = H.Memory((c.oc,), "x", c.b0, nchw=False)
        bia.set(b0)
    if c.b1:
        bia1 = H.Memory((c.oc1,), "x", c.b1, nchw=False)
        bia1.set(b1)
    dst = H.Memory((c.n, c.oc1, c.h, c.w), "nhwc", c.dst)
    op = H.conv(src, wei, bia, (1, 1), (1, 1), dst, wei1x1=wei1, bia1x1=bia1, conv0_relu=bool(c.relu0),
                conv0_scales=s0, conv0_round_mode=c.r0, conv1_relu=bool(c.relu1), conv1_scales=s1,
                conv1_round_mode=c.r1)
    op.submit()
    d = O.make_desc(c.n, c.h, c.w, c.ic, c.oc, c.oc1, cases.DT[c.dst], cases.DT[c.b0], cases.DT[c.b1], relu0=c.relu0,
                    relu1=c.relu1, round0=c.r0, round1=c.r1, nscale0=s0.size, nscale1=s1.size)
    want = O.conv(d, src_a, wb, b0, s0, w1b, b1, s1)
    assert np.array_equal(dst.array().view(np.uint8), want.view(np.uint8))
    # a second submit with new source data through the same op (handles are not re-bound)
    src.set(255 - src_a)
    op.submit()
    want2 = O.conv(d, 255 - src_a, wb, b0, s0, w1b, b1, s1)
    assert np.array_equal(dst.array().view(np.uint8), want2.view(np.uint8))
    assert op.launches() in (1, 4)


@pytest.mark.parametrize("dst_dt", ["u8", "f32"])
def test_conv0_only_submit(H, dst_dt):
    """The reference's 9-argument conv(): 3x3 stage only (include/deepfusion.h:121-129)."""
    from dfb200 import layout, synth
    n, h, w, ic, oc = 2, 9, 7, 32, 48
    src_a = synth.src_u8(1, (n, h, w, ic))
    w0 = synth.wei_s8(2, (oc, ic, 3, 3))
    b0 = synth.bias(4, oc, "s32")
    s0 = synth.channel_scales(oc, 10)
    wb = layout.oihw_to_blocked(w0)
    src = H.Memory((n, ic, h, w), "nhwc", "u8")
    src.set(src_a)
    wei = H.Memory((oc, ic, 3, 3), "OIhw4i16o4i", "s8")
    wei.array().reshape(-1)[...] = wb
    bia = H.Memory((oc,), "x", "s32", nchw=False)
    bia.set(b0)
    dst = H.Memory((n, oc, h, w), "nhwc", dst_dt)
    op = H.conv(src, wei, bia, (1, 1), (1, 1), dst, conv0_relu=True, conv0_scales=s0)
    op.submit()
    d = O.make_desc(n, h, w, ic, oc, 0, cases.DT[dst_dt], O.S32, 0, relu0=1, nscale0=oc)
    want = O.conv(d, src_a, wb, b0, s0)
    assert np.array_equal(dst.array().view(np.uint8), want.view(np.uint8))


def test_concat_feeds_conv_on_device(H):
    """concat+ReLU output consumed by the fused conv without leaving HBM (ext API)."""
    from dfb200 import synth
    n, h, w = 2, 14, 14
    ics = (64, 128, 32, 32)
    ins = [synth.src_u8(20 + i, (n, h, w, c), 0, 127) for i, c in enumerate(ics)]
    srcs = [H.Memory((n, c, h, w), "nhwc", "u8") for c in ics]
    for m, a in zip(srcs, ins):
        m.set(a)
        m.to_device()
    cat = H.Memory((n, 256, h, w), "nhwc", "u8")
    c = cases.ConvCase("chain", n, h, w, 256, 128, 256, "u8", "s32", "s32")
    _, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    wei = H.Memory((128, 256, 3, 3), "OIhw4i16o4i", "s8"); wei.array().reshape(-1)[...] = wb
    wei1 = H.Memory((256, 128, 1, 1), "OIhw4i16o4i", "s8"); wei1.array().reshape(-1)[...] = w1b
    bia = H.Memory((128,), "x", "s32", nchw=False); bia.set(b0)
    bia1 = H.Memory((256,), "x", "s32", nchw=False); bia1.set(b1)
    dst = H.Memory((n, 256, h, w), "nhwc", "u8")
    op_cat = H.concat(srcs, cat, True)
    op_conv = H.conv(cat, wei, bia, (1, 1), (1, 1), dst, wei1x1=wei1, bia1x1=bia1, conv0_scales=s0, conv1_scales=s1)
    op_cat.submit_device()
    op_conv.submit_device()
    H.sync()
    dst.to_host()
    H.sync()
    cat_ref = O.concat(O.U8, True, ins)
    d = O.make_desc(n, h, w, 256, 128, 256, O.U8, O.S32, O.S32, nscale0=128, nscale1=256)
    want = O.conv(d, cat_ref, wb, b0, s0, w1b, b1, s1)
    assert np.array_equal(dst.array(), want)


@pytest.mark.parametrize("ics,fused", [((32, 64), True), ((64, 128, 32, 32), False), ((16, 48, 64), False)],
                         ids=["fused", "two_kernels_pair_route", "two_kernels_channel_split"])
def test_concat_conv_submit(H, ics, fused):
    """ext::concat_conv through the C++ API: submit() == oracle concat followed by oracle conv.  The op fuses the concat
    into the conv's halo load where that is the faster route; where the plain conv runs on CTA pairs (weights that have
    to stream) or the channel split cannot be loaded fused (multiples of 16 that are not multiples of 32) it runs the
    two kernels back to back -- same result."""
    from dfb200 import synth
    n, h, w, ic = 2, 14, 14, sum(ics)
    ins = [synth.uniform_int(20 + i, (n, h, w, c), 0, 255, np.uint8) for i, c in enumerate(ics)]
    srcs = [H.Memory((n, c, h, w), "nhwc", "u8") for c in ics]
    for m, a in zip(srcs, ins):
        m.set(a)
    c = cases.ConvCase("cc", n, h, w, ic, 64, 128, "u8", "s32", "s32", k0=13)
    _, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    wei = H.Memory((64, ic, 3, 3), "OIhw4i16o4i", "s8"); wei.array().reshape(-1)[...] = wb
    wei1 = H.Memory((128, 64, 1, 1), "OIhw4i16o4i", "s8"); wei1.array().reshape(-1)[...] = w1b
    bia = H.Memory((64,), "x", "s32", nchw=False); bia.set(b0)
    bia1 = H.Memory((128,), "x", "s32", nchw=False); bia1.set(b1)
    dst = H.Memory((n, 128, h, w), "nhwc", "u8")
    op = H.concat_conv(srcs, True, wei, bia, (1, 1), (1, 1), dst, wei1x1=wei1, bia1x1=bia1, conv0_scales=s0, conv1_scales=s1)
    assert H.concat_conv_is_fused(op) == fused
    assert op.launches() == (1 if fused else 2)
    op.submit()
    d = O.make_desc(n, h, w, ic, 64, 128, O.U8, O.S32, O.S32, nscale0=64, nscale1=128)
    want = O.conv(d, O.concat(O.U8, True, ins), wb, b0, s0, w1b, b1, s1)
    assert want.any()
    assert np.array_equal(dst.array(), want)


@pytest.mark.parametrize("dst_dt", ["u8", "s8", "s32"])
@pytest.mark.parametrize("shape", [
    # src nchw, conv window / pad, conv dst nchw, pool window / stride, kind, dst nchw
    ((1, 16, 4, 4), 3, 0, (1, 16, 2, 2), 2, 2, 0, (1, 16, 1, 1)),          # test_conv_relu_pooling.cc:314-315 (first VGG entry)
    ((2, 64, 12, 12), 3, 1, (2, 128, 12, 12), 2, 2, 0, (2, 128, 6, 6)),    # :316-317 at a reduced image size
    ((2, 32, 7, 7), 1, 0, (2, 320, 7, 7), 7, 7, 2, (2, 320, 1, 1)),        # :341-342 style: 1x1 conv + 7x7 average, excl. padding
], ids=["vgg_first", "vgg_block", "global_avg"])
def test_conv_relu_pool_submit(H, shape, dst_dt):
    """ext::conv_pool through the C++ API on the reference's own shape list (reduced sizes): submit() == oracle conv
    followed by oracle pooling."""
    from dfb200 import synth, layout
    sd, k, pad, cd, pk, ps, kind, od = shape
    n, ic, h, w = sd
    oc = cd[1]
    src_a = synth.src_u8(1, (n, h, w, ic))
    wb = layout.oihw_to_blocked(synth.wei_s8(2, (oc, ic, k, k)))
    b0 = synth.bias(4, oc, "s32")
    s0 = synth.channel_scales(oc, int(np.ceil(np.log2(ic * k * k * 64.0))) - 4)
    src = H.Memory(sd, "nhwc", "u8"); src.set(src_a)
    wei = H.Memory((oc, ic, k, k), "OIhw4i16o4i", "s8"); wei.array().reshape(-1)[...] = wb
    bia = H.Memory((oc,), "x", "s32", nchw=False); bia.set(b0)
    conv_dst = H.Memory(cd, "nhwc", dst_dt)
    dst = H.Memory(od, "nhwc", dst_dt)
    op = H.conv_pool(src, wei, bia, (1, 1), (pad, pad), conv_dst, dst, kind, (pk, pk), (ps, ps), (0, 0), conv_relu=True, conv_scales=s0)
    assert op.launches() == 2
    op.submit()
    d = O.make_desc(n, h, w, ic, oc, 0, O.DT_OF[dst_dt], O.S32, O.UNDEF, k=k, pad=pad, relu0=1, nscale0=oc)
    want = O.pool(O.conv(d, src_a, wb, b0, s0), kind, pk, ps, 0)
    assert want.any()
    assert np.array_equal(dst.array().view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("fused", [False, True], ids=["conv_only", "fused"])
def test_conv_sum_submit(H, fused):
    """ext::conv_sum through the C++ API: conv (+1x1) + residual + ReLU == the oracle's dfo_conv_sum."""
    from dfb200 import synth, layout
    n, h, w, ic, oc, oc1 = 2, 7, 7, 64, 64, (256 if fused else 0)
    src_a = synth.src_u8(1, (n, h, w, ic))
    wb = layout.oihw_to_blocked(synth.wei_s8(2, (oc, ic, 3, 3)))
    b0 = synth.bias(4, oc, "s32")
    s0 = synth.channel_scales(oc, 12)
    out_c = oc1 or oc
    res_a = synth.uniform_int(41, (n, h, w, out_c), 0, 255, np.uint8)
    src = H.Memory((n, ic, h, w), "nhwc", "u8"); src.set(src_a)
    wei = H.Memory((oc, ic, 3, 3), "OIhw4i16o4i", "s8"); wei.array().reshape(-1)[...] = wb
    bia = H.Memory((oc,), "x", "s32", nchw=False); bia.set(b0)
    res = H.Memory((n, out_c, h, w), "nhwc", "u8"); res.set(res_a)
    dst = H.Memory((n, out_c, h, w), "nhwc", "u8")
    if fused:
        w1b = layout.oihw_to_blocked(synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1))
        b1 = synth.bias(5, oc1, "s32")
        s1 = synth.channel_scales(oc1, 12)
        wei1 = H.Memory((oc1, oc, 1, 1), "OIhw4i16o4i", "s8"); wei1.array().reshape(-1)[...] = w1b
        bia1 = H.Memory((oc1,), "x", "s32", nchw=False); bia1.set(b1)
        op = H.conv_sum(src, wei, bia, (1, 1), (1, 1), res, dst, wei1x1=wei1, bia1x1=bia1, conv0_scales=s0, conv1_scales=s1)
        d = O.make_desc(n, h, w, ic, oc, oc1, O.U8, O.S32, O.S32, relu0=1, relu1=1, nscale0=oc, nscale1=oc1)
        want = O.conv_sum(d, src_a, wb, b0, s0, res_a, w1b, b1, s1)
    else:
        op = H.conv_sum(src, wei, bia, (1, 1), (1, 1), res, dst, conv0_scales=s0)
        d = O.make_desc(n, h, w, ic, oc, 0, O.U8, O.S32, O.UNDEF, relu0=1, nscale0=oc)
        want = O.conv_sum(d, src_a, wb, b0, s0, res_a)
    op.submit()
    assert np.array_equal(dst.array(), want)


def test_resnet_entry_of_the_reference_list(H):
    """test/test_conv_relu_pooling.cc:341-342: 1x1 conv 2048 -> 2048 @7x7 + eltwise sum + ReLU + 7x7 average pooling
    (excluding padding), at batch 2: ext::conv_sum (K-sliced halo, eight channel groups) chained with ext::pool on the
    device == the oracle's dfo_conv_sum followed by dfo_pool."""
    from dfb200 import synth, layout
    n, c = 2, 2048
    src_a = synth.src_u8(1, (n, 7, 7, c))
    wb = layout.oihw_to_blocked(synth.wei_s8(2, (c, c, 1, 1)))
    b0 = synth.bias(4, c, "s32")
    s0 = synth.channel_scales(c, 13)
    res_a = synth.uniform_int(41, (n, 7, 7, c), 0, 255, np.uint8)
    src = H.Memory((n, c, 7, 7), "nhwc", "u8"); src.set(src_a); src.to_device()
    wei = H.Memory((c, c, 1, 1), "OIhw4i16o4i", "s8"); wei.array().reshape(-1)[...] = wb
    bia = H.Memory((c,), "x", "s32", nchw=False); bia.set(b0)
    res = H.Memory((n, c, 7, 7), "nhwc", "u8"); res.set(res_a); res.to_device()
    mid = H.Memory((n, c, 7, 7), "nhwc", "u8")
    dst = H.Memory((n, c, 1, 1), "nhwc", "u8")
    conv = H.conv_sum(src, wei, bia, (1, 1), (0, 0), res, mid, conv0_scales=s0)
    pool = H.pool(mid, dst, 2, (7, 7), (7, 7), (0, 0))
    conv.submit_device()
    pool.submit_device()
    H.sync()
    dst.to_host()
    H.sync()
    d = O.make_desc(n, 7, 7, c, c, 0, O.U8, O.S32, O.UNDEF, k=1, pad=0, relu0=1, nscale0=c)
    want = O.pool(O.conv_sum(d, src_a, wb, b0, s0, res_a), 2, 7, 7, 0)
    assert want.any() and (want < 255).any()
    assert np.array_equal(dst.array(), want)


def _conv_memories(H, c, n=None):
    n = n or c.n
    src_a, w0, w1, b0, b1, s0, s1 = cases.ConvCase(c.name, n, c.h, c.w, c.ic, c.oc, c.oc1, c.dst, c.b0, c.b1, c.r0, c.r1,
                                                   c.relu0, c.relu1, c.per_channel, c.k0, c.k1, c.data).tensors()
    wb, w1b = c.blocked(w0, w1)
    src = H.Memory((n, c.ic, c.h, c.w), "nhwc", "u8")
    src.set(src_a)
    wei = H.Memory((c.oc, c.ic, 3, 3), "OIhw4i16o4i", "s8")
    wei.array().reshape(-1)[...] = wb
    wei1 = H.Memory((c.oc1, c.oc, 1, 1), "OIhw4i16o4i", "s8")
    wei1.array().reshape(-1)[...] = w1b
    bia = H.Memory((c.oc,), "x", c.b0, nchw=False)
    bia.set(b0)
    bia1 = H.Memory((c.oc1,), "x", c.b1, nchw=False)
    bia1.set(b1)
    dst = H.Memory((n, c.oc1, c.h, c.w), "nhwc", c.dst)
    return src, wei, bia, wei1, bia1, dst, s0, s1


@pytest.mark.parametrize("n", [1, 3, 13, 64])
@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_sharded_equals_unsharded(H, n, n_dev):
    """ext::conv_sharded: contiguous batch slabs over the GPUs of one box, no collective.  The sharded
    result must be bit-identical to the single-device op -- also when the batch is smaller than the device count
    (idle devices) and when the slabs are ragged.  Devices beyond those present are skipped."""
    import dfb200 as df
    if df.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} devices")
    c = {x.name: x for x in cases.FULL_CONV}["cfg3"]
    src, wei, bia, wei1, bia1, dst, s0, s1 = _conv_memories(H, c, n)
    H.conv(src, wei, bia, (1, 1), (1, 1), dst, wei1x1=wei1, bia1x1=bia1, conv0_scales=s0, conv1_scales=s1).submit()
    want = dst.array().copy()
    dst.array()[...] = 0
    op = H.conv_sharded(list(range(n_dev)), src, wei, bia, (1, 1), (1, 1), dst, wei1x1=wei1, bia1x1=bia1, conv0_scales=s0,
                        conv1_scales=s1)
    op.submit()
    assert np.array_equal(dst.array(), want)
    # device-resident path: upload once, kernels only, download
    dst.array()[...] = 0
    op.upload()
    op.submit_device()
    op.sync()
    op.download()
    assert np.array_equal(dst.array(), want)
    assert df.lib().df_get_device is not None
