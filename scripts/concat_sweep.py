"""concat+ReLU bandwidth vs footprint (development aid, needs a GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import numpy as np
import dfb200 as df
from dfb200 import synth

ics = (64, 128, 32, 32)
for n in (32, 128, 512, 2048):
    npix = n * 28 * 28
    ins = [df.DeviceBuffer(npix * c) for c in ics]
    for b in ins: b.fill(0x5A)
    out = df.DeviceBuffer(npix * 256)
    nbytes = 2 * npix * 256
    run = lambda: df.concat_run(df.U8, True, [b.ptr for b in ins], list(ics), out.ptr, npix)
    for _ in range(5): run()
    df.sync()
    iters = 50 if n <= 512 else 10
    e0, e1 = df.Event(), df.Event()
    e0.record()
    for _ in range(iters): run()
    e1.record()
    ms = e0.elapsed_ms(e1) / iters
    print(f"concat+relu u8 N={n:5d}: {nbytes/1e6:8.1f} MB moved, {ms*1e3:8.1f} us, {nbytes/ms/1e6:7.1f} GB/s ({nbytes/ms/1e6/6446.3*100:.1f}% of measured HBM copy)", flush=True)
