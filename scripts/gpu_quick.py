"""Quick GPU parity run (development aid): CUDA path vs oracle on a handful of shapes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import numpy as np
import dfb200 as df
from dfb200 import synth, layout
import oracle_lib as O

def conv_case(n, h, w, ic, oc, oc1, dst, b0, b1, r0=0, r1=0, relu1=0, k0=13, k1=12, ref="replay"):
    src = synth.src_u8(1, (n, h, w, ic)); w0 = synth.wei_s8(2, (oc, ic, 3, 3)); w1 = synth.wei_s8(3, (oc1, oc))
    bia0 = synth.bias(4, oc, b0) if b0 else None; bia1 = synth.bias(5, oc1, b1) if b1 else None
    s0 = synth.channel_scales(oc, k0); s1 = synth.channel_scales(oc1, k1)
    wb = layout.oihw_to_blocked(w0); w1b = layout.oihw_to_blocked(w1.reshape(oc1, oc, 1, 1))
    d = O.make_desc(n, h, w, ic, oc, oc1, O.DT_OF[dst], O.DT_OF[b0], O.DT_OF[b1], round0=r0, round1=r1, relu1=relu1, nscale0=oc, nscale1=oc1)
    fn = O.replay_conv if (ref == "replay" and O.replay_supported()) else O.conv
    want = fn(d, src, wb, bia0, s0, w1b, bia1, s1)
    op = df.Conv(n, h, w, ic, oc, oc1, df.DT_OF[dst], wb, w1b, bia0, bia1, s0, s1, df.DT_OF[b0], df.DT_OF[b1], relu1=bool(relu1), round0=r0, round1=r1)
    i = op.info()
    got = op(src)
    ok = np.array_equal(got.view(np.uint8), want.view(np.uint8))
    nbad = int((got.view(np.uint8) != want.view(np.uint8)).sum())
    print(f"conv n={n} {h}x{w} {ic}->{oc}->{oc1} dst={dst} b={b0},{b1} r={r0},{r1}: {'PASS' if ok else 'FAIL'} bad_bytes={nbad} "
          f"[tiles={i.tiles_per_launch} grid={i.grid} smem={i.smem_bytes} res={i.w0_resident}{i.w1_resident} SA={i.a_stages} SB={i.b_stages} Wp={i.padded_w} eff={i.mma_efficiency:.3f}]", flush=True)
    if not ok:
        bad = np.argwhere(got != want)
        print("  first bad idx", bad[:5].tolist(), "got", got[tuple(bad[0])], "want", want[tuple(bad[0])])
    return ok

def concat_case(dt, shapes, relu):
    npdt = df.NP_OF[df.DT_OF[dt]]
    srcs = []
    for i, s in enumerate(shapes):
        if dt == "f32":
            a = (synth.uniform_int(10 + i, s, -1000, 1000, np.int32).astype(np.float32) / 7).astype(np.float32)
        elif dt == "s32":
            a = synth.uniform_int(10 + i, s, -100000, 100000, np.int32)
        elif dt == "s8":
            a = synth.uniform_int(10 + i, s, -128, 127, np.int8)
        else:
            a = synth.uniform_int(10 + i, s, 0, 255, np.uint8)
        srcs.append(a)
    want = O.concat(O.DT_OF[dt], relu, srcs)
    got = df.concat(srcs, df.DT_OF[dt], relu)
    ok = np.array_equal(got.view(np.uint8), want.view(np.uint8))
    print(f"concat {dt} relu={relu} {shapes}: {'PASS' if ok else 'FAIL'}", flush=True)
    return ok

def bench_conv(n, h, w, ic, oc, oc1, dst="u8", iters=20):
    src = synth.src_u8(1, (n, h, w, ic)); w0 = synth.wei_s8(2, (oc, ic, 3, 3)); w1 = synth.wei_s8(3, (oc1, oc))
    bia0 = synth.bias(4, oc, "s32"); bia1 = synth.bias(5, oc1, "s32")
    s0 = synth.channel_scales(oc, 13); s1 = synth.channel_scales(oc1, 12)
    wb = layout.oihw_to_blocked(w0); w1b = layout.oihw_to_blocked(w1.reshape(oc1, oc, 1, 1))
    op = df.Conv(n, h, w, ic, oc, oc1, df.DT_OF[dst], wb, w1b, bia0, bia1, s0, s1, df.S32, df.S32)
    sbuf = df.DeviceBuffer.from_numpy(src)
    dbuf = df.DeviceBuffer(n * h * w * oc1 * (4 if dst in ("f32", "s32") else 1))
    for _ in range(3): op.run(sbuf, dbuf)
    df.sync()
    e0, e1 = df.Event(), df.Event()
    e0.record()
    for _ in range(iters): op.run(sbuf, dbuf)
    e1.record()
    ms = e0.elapsed_ms(e1) / iters
    i = op.info()
    tops = 2 * i.macs_per_image * n / (ms * 1e-3) / 1e12
    print(f"bench conv n={n} {h}x{w} {ic}->{oc}->{oc1} {dst}: {ms*1e3:.1f} us/launch  {tops:.1f} TOPS  {n/(ms*1e-3):.0f} img/s "
          f"[tiles={i.tiles_per_launch} res={i.w0_resident}{i.w1_resident} SA={i.a_stages} SB={i.b_stages} smem={i.smem_bytes}]", flush=True)

if __name__ == "__main__":
    print("devices", df.device_count(), "sms", df.sm_count())
    ok = True
    for dt in ("u8", "s8", "s32", "f32"):
        for relu in (False, True):
            ok &= concat_case(dt, [(2, 4, 4, 64), (2, 4, 4, 32)], relu)
    ok &= concat_case("u8", [(32, 28, 28, 64), (32, 28, 28, 128), (32, 28, 28, 32), (32, 28, 28, 32)], True)
    ok &= conv_case(1, 8, 8, 64, 64, 128, "u8", "s32", "s32")
    ok &= conv_case(1, 56, 56, 64, 64, 256, "u8", "s32", "s32")
    ok &= conv_case(2, 28, 28, 128, 128, 512, "u8", "s32", "s32")
    ok &= conv_case(3, 14, 14, 256, 256, 1024, "u8", "s32", "s32")
    ok &= conv_case(3, 14, 14, 256, 256, 1024, "f32", "f32", "u8", r0=1)
    ok &= conv_case(3, 14, 14, 256, 256, 1024, "s32", "s8", None, r1=1)
    ok &= conv_case(2, 9, 7, 32, 48, 80, "s8", "u8", "f32")
    ok &= conv_case(5, 5, 3, 16, 16, 16, "u8", None, None, k0=8, k1=8)
    ok &= conv_case(64, 28, 28, 128, 128, 512, "u8", "s32", "s32")
    print("ALL PASS" if ok else "SOME FAILED")
    bench_conv(1, 56, 56, 64, 64, 256)
    bench_conv(64, 56, 56, 64, 64, 256)
    bench_conv(64, 28, 28, 128, 128, 512)
    bench_conv(256, 14, 14, 256, 256, 1024)
    bench_conv(256, 14, 14, 256, 256, 1024, "s32")
