"""cfg4 on CTA pairs with streamed weight halves vs the single-CTA kernel: parity on a few batches, then timing (needs a GPU)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import numpy as np
import dfb200 as df, cases, oracle_lib as O
df.set_device(0)
ok = True
for n, dst in ((1, "u8"), (3, "u8"), (16, "s32"), (37, "f32"), (256, "u8")):
    c = cases.ConvCase("cfg4", n, 14, 14, 256, 256, 1024, dst, "s32", "s32")
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    op = df.Conv(n, 14, 14, 256, 256, 1024, cases.DT[dst], wb, w1b, b0, b1, s0, s1, df.S32, df.S32)
    i = op.info()
    got = op(src)
    d = O.make_desc(n, 14, 14, 256, 256, 1024, cases.DT[dst], O.S32, O.S32, nscale0=256, nscale1=1024)
    want = O.replay_conv(d, src, wb, b0, s0, w1b, b1, s1)
    same = np.array_equal(got.view(np.uint8), want.view(np.uint8))
    ok &= same
    print(f"cfg4 n={n} dst={dst} res={i.w0_resident}{i.w1_resident} SB={i.b_stages} smem={i.smem_bytes}: {'PASS' if same else 'FAIL'}", flush=True)
    op.close()
print("ALL PASS" if ok else "SOME FAILED")
