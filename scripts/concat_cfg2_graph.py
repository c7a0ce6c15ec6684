"""concat+ReLU cfg2 (12.8 MB) replayed from a CUDA graph over rotating buffers -- development aid, needs a GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import numpy as np
import dfb200 as df

ics = (64, 128, 32, 32)
npix = 32 * 28 * 28
sets = []
for i in range(21):
    ins = [df.DeviceBuffer(npix * c) for c in ics]
    for b in ins: b.fill(0x5A)
    sets.append((ins, df.DeviceBuffer(npix * 256)))
st = df.Stream()
calls = [df.ConcatCall(df.U8, True, [b.ptr for b in s[0]], list(ics), s[1].ptr, npix, stream=st.ptr) for s in sets]
for c in calls: c()
with df.Graph(st) as g:
    for i in range(200): calls[i % 21]()
g.launch(); st.sync()
e0, e1 = df.Event(), df.Event()
e0.record(st.ptr); g.launch(); e1.record(st.ptr); st.sync()
us = e0.elapsed_ms(e1) / 200 * 1e3
print(f"concat cfg2 graph replay: {us:.2f} us/launch  {2*npix*256/us/1e3:.0f} GB/s")
