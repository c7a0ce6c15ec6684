"""Shapes beyond BASELINE.json through the same entry points: the conv-only operator on the reference's own shape list
(test/test_conv_relu_pooling.cc:313-391: VGG 3x3 layers, ResNet 1x1 with eltwise sum) and a few fused ones -- which kernel
they run on (1 / 0 = single CTA with resident / streamed weights, 3 = CTA pair with streamed halves, composite = several
launches) and what they reach.  Device-resident, CUDA-graph replay, buffers rotating over > 2x L2 where the batch allows.
Needs a GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import numpy as np
import dfb200 as df
from dfb200 import synth, layout

L2 = 126 << 20
ZOO = [
    # name, n, h, w, ic, oc, oc1, k, stride, pad
    ("VGG conv2_1  64->128 @112 (conv-only)", 16, 112, 112, 64, 128, 0, 3, 1, 1),
    ("VGG conv3_1 128->256 @56  (conv-only)", 32, 56, 56, 128, 256, 0, 3, 1, 1),
    ("VGG conv4_1 256->512 @28  (conv-only)", 64, 28, 28, 256, 512, 0, 3, 1, 1),
    ("VGG conv5_1 512->512 @14  (conv-only)", 128, 14, 14, 512, 512, 0, 3, 1, 1),
    ("ResNet 1x1 512->2048 @7   (conv-only)", 256, 7, 7, 512, 2048, 0, 1, 1, 0),
    ("ResNet 3x3 s2 128->128 @56 (conv-only)", 32, 56, 56, 128, 128, 0, 3, 2, 1),
    ("fused 256->128->512 @28", 64, 28, 28, 256, 128, 512, 3, 1, 1),
    ("fused 128->64->256 @56", 32, 56, 56, 128, 64, 256, 3, 1, 1),
    ("fused 512->512->2048 @7 (chained)", 256, 7, 7, 512, 512, 2048, 3, 1, 1),
    ("fused 5x5 128->128->512 @28", 32, 28, 28, 128, 128, 512, 5, 1, 2),
]
df.set_device(0)
st = df.Stream()
print(f"{'shape':44s} {'batch':>5s} {'kernel':>9s} {'us/launch':>10s} {'TOPS':>8s}")
for name, n, h, w, ic, oc, oc1, k, s, pd in ZOO:
    w0b = layout.oihw_to_blocked(synth.wei_s8(2, (oc, ic, k, k)))
    w1b = layout.oihw_to_blocked(synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1)) if oc1 else None
    b0 = synth.bias(4, oc, "s32")
    b1 = synth.bias(5, oc1, "s32") if oc1 else None
    s0 = synth.channel_scales(oc, int(np.ceil(np.log2(ic * k * k * 64.0))) - 4)
    s1 = synth.channel_scales(oc1, 12) if oc1 else (1.0,)
    op = df.Conv(n, h, w, ic, oc, oc1, df.U8, w0b, w1b, b0, b1, s0, s1, df.S32, df.S32 if oc1 else df.UNDEF, relu0=True, k=k, stride=s, pad=pd)
    i = op.info()
    oh, ow = op.oh, op.ow
    src_b, dst_b = n * h * w * ic, n * oh * ow * (oc1 or oc)
    n_sets = max(2, min(16, -(-2 * L2 // (src_b + dst_b))))
    base = synth.src_u8(1, (min(n, 8), h, w, ic))
    src = np.tile(base, (-(-n // base.shape[0]), 1, 1, 1))[:n]
    sets = [(df.DeviceBuffer.from_numpy(src), df.DeviceBuffer(dst_b)) for _ in range(n_sets)]
    steps = 20
    for j in range(3):
        op.run(*sets[j % n_sets], stream=st.ptr)
    with df.Graph(st) as g:
        for j in range(steps):
            op.run(*sets[j % n_sets], stream=st.ptr)
    g.launch(); st.sync()
    e0, e1 = df.Event(), df.Event()
    e0.record(st.ptr); g.launch(); e1.record(st.ptr); st.sync()
    us = e0.elapsed_ms(e1) / steps * 1e3
    tops = 2 * i.macs_per_image * n / us / 1e6
    kern = "composite" if (oc > 256) else str(i.w0_resident)
    print(f"{name:44s} {n:5d} {kern:>9s} {us:10.1f} {tops:8.1f}", flush=True)
    del g, e0, e1
    op.close(); del sets
