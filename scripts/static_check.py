"""Parity of the three BASELINE shapes (static kernels) against the AVX-512 replay, a few batches each (needs a GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import numpy as np
import dfb200 as df, cases, oracle_lib as O
df.set_device(0)
ok = True
for (h, w, ic, oc, oc1), batches in (((56, 56, 64, 64, 256), (1, 5)), ((28, 28, 128, 128, 512), (3, 64)), ((14, 14, 256, 256, 1024), (2, 37))):
    for n in batches:
        for dst in ("u8", "s32"):
            c = cases.ConvCase("s", n, h, w, ic, oc, oc1, dst, "s32", "s32")
            src, w0, w1, b0, b1, s0, s1 = c.tensors()
            wb, w1b = c.blocked(w0, w1)
            op = df.Conv(n, h, w, ic, oc, oc1, cases.DT[dst], wb, w1b, b0, b1, s0, s1, df.S32, df.S32)
            i = op.info()
            got = op(src)
            got2 = op(src)  # a second launch on the same handle (barrier phases / seeds start over)
            d = O.make_desc(n, h, w, ic, oc, oc1, cases.DT[dst], O.S32, O.S32, nscale0=oc, nscale1=oc1)
            want = O.replay_conv(d, src, wb, b0, s0, w1b, b1, s1)
            same = np.array_equal(got, want) and np.array_equal(got2, want)
            ok &= same
            print(f"{h}x{w} {ic}->{oc}->{oc1} n={n} dst={dst} res={i.w0_resident}{i.w1_resident}: {'PASS' if same else 'FAIL'}", flush=True)
            op.close()
print("ALL PASS" if ok else "SOME FAILED")
