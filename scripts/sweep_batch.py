"""BASELINE configs[4]: the fused conv (cfg3 shape) over batch N = 1 .. 2048 on one GPU, each point timed as a
CUDA-graph replay of back-to-back launches over rotating buffers (> 2x L2 when the batch allows) -- needs a GPU.
usage: sweep_batch.py [cfg1|cfg3|cfg4]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import numpy as np
import dfb200 as df
from dfb200 import synth, layout

SHAPES = {"cfg1": (56, 56, 64, 64, 256), "cfg3": (28, 28, 128, 128, 512), "cfg4": (14, 14, 256, 256, 1024)}
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
h, w, ic, oc, oc1 = SHAPES[which]
df.set_device(0)
w0 = synth.wei_s8(2, (oc, ic, 3, 3)); w1 = synth.wei_s8(3, (oc1, oc))
args = (layout.oihw_to_blocked(w0), layout.oihw_to_blocked(w1.reshape(oc1, oc, 1, 1)), synth.bias(4, oc, "s32"),
        synth.bias(5, oc1, "s32"), synth.channel_scales(oc, {64: 12, 128: 13, 256: 14}[ic]), synth.channel_scales(oc1, 12))
st = df.Stream()
print(f"# {which}: {h}x{w} {ic}->{oc}->{oc1}, u8 out; peak = 3348 TOPS (2 x measured bf16)")
NS = [int(x) for x in os.environ["SWEEP_NS"].split(",")] if os.environ.get("SWEEP_NS") else (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048)
for n in NS:
    op = df.Conv(n, h, w, ic, oc, oc1, df.U8, *args, df.S32, df.S32)
    i = op.info()
    per_set = n * h * w * (ic + oc1)
    n_sets = max(2, min(64, -(-2 * 126 * 2 ** 20 // per_set)))
    if os.environ.get("SWEEP_SETS"): n_sets = int(os.environ["SWEEP_SETS"])  # development: 1 = the same buffers every launch (L2-resident when they fit)
    base = synth.src_u8(1, (n, h, w, ic))
    sets = [(df.DeviceBuffer.from_numpy(base), df.DeviceBuffer(n * h * w * oc1)) for _ in range(n_sets)]
    iters = 200 if n <= 256 else 40
    for k in range(3): op.run(*sets[k % n_sets], stream=st.ptr)
    with df.Graph(st) as g:
        for k in range(iters): op.run(*sets[k % n_sets], stream=st.ptr)
    g.launch(); st.sync()
    e0, e1 = df.Event(), df.Event()
    e0.record(st.ptr); g.launch(); e1.record(st.ptr); st.sync()
    us = e0.elapsed_ms(e1) / iters * 1e3
    tops = 2 * i.macs_per_image * n / us / 1e6
    print(f"N={n:5d} tiles={i.tiles_per_launch:6d} grid={i.grid:4d}  {us:9.2f} us/launch  {tops:8.1f} TOPS ({100 * tops / 3348.2:5.1f} %)  {n / us * 1e6:12.0f} images/s", flush=True)
    del g, e0, e1  # before the next capture begins (destroying a graph inside a capture invalidates it)
    op.close(); del sets
