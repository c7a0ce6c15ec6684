"""Generates tests/golden/*.npz -- pinned outputs of the deep-fusion arithmetic on seeded inputs.

The reference ships no golden vectors for this path and cannot be built here, so the
fixtures are produced by THREE independent statements of the arithmetic that must agree before
anything is written: the scalar oracle (oracle/df_oracle.c), the AVX-512-VNNI replay of the emitted
x86 instructions (oracle/df_replay_avx512.c) and the numpy model (tests/np_model.py).  Inputs are
regenerated from seeds by tests/cases.py; only outputs are stored.

Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))

import cases  # noqa: E402
import np_model as M  # noqa: E402
import oracle_lib as O  # noqa: E402

GOLDEN_CONV = ["tiny16", "ragged", "one_pixel", "one_row", "one_col", "s8_norelu", "s8_relu_down", "single_scale",
               "refrange", "refrange_s32", "ties_rn", "ties_rd", "extreme", "ic96_oc80"]


def conv_outputs(c):
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    d = O.make_desc(c.n, c.h, c.w, c.ic, c.oc, c.oc1, cases.DT[c.dst], cases.DT[c.b0], cases.DT[c.b1], relu0=c.relu0,
                    relu1=c.relu1, round0=c.r0, round1=c.r1, nscale0=s0.size, nscale1=s1.size)
    a = O.conv(d, src, wb, b0, s0, w1b, b1, s1)
    m = M.conv_fused(src, w0, b0, s0, w1, b1, s1, cases.DT[c.dst], relu0=c.relu0, relu1=c.relu1, down0=c.r0, down1=c.r1)
    assert np.array_equal(a.view(np.uint8), m.view(np.uint8)), f"{c.name}: scalar oracle != numpy model"
    if O.replay_supported():
        r = O.replay_conv(d, src, wb, b0, s0, w1b, b1, s1)
        assert np.array_equal(a.view(np.uint8), r.view(np.uint8)), f"{c.name}: scalar oracle != x86 replay"
    return a


def main():
    out = {}
    for c in cases.SMALL_CONV:
        if c.name in GOLDEN_CONV:
            out["conv_" + c.name] = conv_outputs(c)
    np.savez_compressed(os.path.join(HERE, "conv_small.npz"), **out)
    cc = {}
    for dt in ("u8", "s8", "s32", "f32"):
        for ci, (srcs, _dst) in enumerate(cases.CONCAT_BASIC[:5]):
            for data in ("reference-range", "full"):
                ins = cases.concat_inputs(dt, srcs, data)
                a = O.concat(cases.DT[dt], 1, ins)
                m = M.concat(ins, cases.DT[dt], True)
                assert np.array_equal(a.view(np.uint8), m.view(np.uint8)), (dt, ci, data)
                if O.replay_supported():
                    r = O.replay_concat(cases.DT[dt], 1, ins)
                    assert np.array_equal(a.view(np.uint8), r.view(np.uint8)), (dt, ci, data)
                cc[f"concat_{dt}_{ci}_{data}"] = a
    np.savez_compressed(os.path.join(HERE, "concat_relu.npz"), **cc)
    for f in ("conv_small.npz", "concat_relu.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
