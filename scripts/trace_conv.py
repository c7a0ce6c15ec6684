"""Timeline of the fused conv kernel's warp roles for one CTA (development aid, needs a GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import numpy as np
import dfb200 as df
from dfb200 import synth, layout

TAGS = {1: "A.issue", 2: "A.landed", 3: "A.clamped", 8: "M.begin", 9: "M.start", 10: "M.g1_go", 11: "M.g1_issued", 12: "M.g2_go", 13: "M.g2_chunk", 14: "M.tap", 20: "B.iter",
        30: "E.acc0_ready", 31: "E.epi0_done", 32: "E.acc1_ready", 33: "E.chunk_done", 36: "E.unit_begin", 34: "E.e0_math_done", 35: "E.math_done", 37: "E.ld_issued", 38: "E.ld_landed", 39: "E.released",
        40: "S.full", 41: "S.issued", 42: "S.read_done"}

def run(n, h, w, ic, oc, oc1, dst="u8", cta=0, cap=512, max_lines=120, skip=0, ics=None):
    w0 = synth.wei_s8(2, (oc, ic, 3, 3)); w1 = synth.wei_s8(3, (max(oc1, 16), oc))
    args = (layout.oihw_to_blocked(w0), layout.oihw_to_blocked(w1.reshape(-1, oc, 1, 1)) if oc1 else None,
            synth.bias(4, oc, "s32"), synth.bias(5, oc1, "s32") if oc1 else None, synth.channel_scales(oc, 13),
            synth.channel_scales(oc1, 12) if oc1 else (1.0,), df.S32, df.S32 if oc1 else df.UNDEF)
    if ics:  # concat fused into the halo load
        op = df.ConcatConv(n, h, w, ics, True, oc, oc1, df.DT_OF[dst], *args)
        src = [df.DeviceBuffer.from_numpy(synth.src_u8(1 + k, (n, h, w, c))) for k, c in enumerate(ics)]
    else:
        op = df.Conv(n, h, w, ic, oc, oc1, df.DT_OF[dst], *args)
        src = df.DeviceBuffer.from_numpy(synth.src_u8(1, (n, h, w, ic)))
    i = op.info()
    out = df.DeviceBuffer(n * h * w * (oc1 or oc) * (4 if dst in ("s32", "f32") else 1))
    for _ in range(3): op.run(src, out)
    df.sync()
    tb = df.DeviceBuffer(i.grid * 4 * cap * 8); tb.fill(0)
    df.check(df.lib().df_conv_debug_trace(op.handle, tb.ptr, cap))
    op.run(src, out); df.sync()
    t = tb.download((i.grid, 4, cap), np.uint64)
    # kernel entry / exit wall-clock stamps (ns) of every CTA: last two words of role lane 2
    t_in = t[:, 2, cap - 1].astype(np.int64); t_out = t[:, 2, cap - 2].astype(np.int64)
    t[:, 2, cap - 2:] = 0
    if t_in.min() > 0:
        base = t_in.min()
        print(f"wall clock (ns, relative to the first CTA's entry): entry min/median/max = {0}/{int(np.median(t_in - base))}/{int((t_in - base).max())}  "
              f"exit min/median/max = {int((t_out - base).min())}/{int(np.median(t_out - base))}/{int((t_out - base).max())}  "
              f"this CTA: in {int(t_in[cta] - base)} out {int(t_out[cta] - base)}")
    ev = []
    for role in range(4):
        for x in t[cta, role]:
            if x == 0: continue
            ev.append((int(x & np.uint64(0xFFFFFFFFFFFF)), int(x >> np.uint64(48)), role))
    ev.sort()
    t0 = ev[0][0]
    print(f"--- {h}x{w} {ic}->{oc}->{oc1} n={n} dst={dst} tiles={i.tiles_per_launch} grid={i.grid} cta={cta} events={len(ev)} span={ev[-1][0]-t0} cycles")
    prev = t0
    for k, (c, tag, role) in enumerate(ev[skip:skip + max_lines]):
        print(f"{c - t0:8d} (+{c - prev:6d}) {'  ' * role}{TAGS.get(tag, tag)}")
        prev = c
    # per-tile steady state estimate from epilogue chunk-done events
    ends = [c for c, tag, _ in ev if tag == 31]
    if len(ends) > 2:
        d = np.diff(ends)
        print("cycles between successive epi0_done:", d.tolist()[:20])

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    if which == "cfg1": run(nb, 56, 56, 64, 64, 256)
    elif which == "cfg3": run(nb, 28, 28, 128, 128, 512, skip=skip)
    elif which == "cfg4": run(256, 14, 14, 256, 256, 1024)
    elif which == "vgg2": run(nb, 112, 112, 64, 128, 0, skip=skip, max_lines=200)   # conv-only, 64-byte K-blocks, one chunk per tile
    elif which == "catf": run(nb, 28, 28, 256, 128, 512, skip=skip, max_lines=260, ics=(64, 128, 32, 32))
    elif which == "cat": run(nb, 28, 28, 256, 128, 512, skip=skip, max_lines=260)  # the conv behind BASELINE configs[1]'s concat (run-time geometry)
