"""Shared, seeded test cases for the deep-fusion hot path (used by CPU and GPU tests)."""
import numpy as np

from dfb200 import layout, synth

F32, S32, S8, U8 = 1, 2, 3, 4
DT = {"f32": F32, "s32": S32, "s8": S8, "u8": U8, None: 0}
NPDT = {"f32": np.float32, "s32": np.int32, "s8": np.int8, "u8": np.uint8}


class ConvCase:
    def __init__(self, name, n, h, w, ic, oc, oc1, dst="u8", b0="s32", b1="s32", r0=0, r1=0, relu0=0, relu1=0,
                 per_channel=True, k0=None, k1=12, data="full"):
        self.name, self.n, self.h, self.w, self.ic, self.oc, self.oc1 = name, n, h, w, ic, oc, oc1
        self.dst, self.b0, self.b1, self.r0, self.r1, self.relu0, self.relu1 = dst, b0, b1, r0, r1, relu0, relu1
        self.per_channel, self.data = per_channel, data
        # scale exponents chosen so that roughly half of the u8 intermediate is non-zero
        self.k0 = k0 if k0 is not None else {16: 10, 32: 11, 48: 11, 64: 12, 128: 13, 256: 14}.get(ic, 12)
        self.k1 = k1

    def tensors(self):
        """src NHWC u8, w0 OIHW s8, w1 (O1,O) s8, biases, scales -- all from fixed seeds."""
        if self.data == "reference-range":  # test/test_utils.h:56-59
            src = synth.src_u8(1, (self.n, self.h, self.w, self.ic), 0, 16)
            w0 = synth.wei_s8(2, (self.oc, self.ic, 3, 3), -10, 10)
            w1 = synth.wei_s8(3, (self.oc1, self.oc), -10, 10)
        elif self.data == "extreme":  # all-255 x all-127: |acc| far above 2^24
            src = np.full((self.n, self.h, self.w, self.ic), 255, np.uint8)
            w0 = np.full((self.oc, self.ic, 3, 3), 127, np.int8)
            w0[1::2] = -127
            w1 = synth.wei_s8(3, (self.oc1, self.oc))
        else:
            src = synth.src_u8(1, (self.n, self.h, self.w, self.ic))
            w0 = synth.wei_s8(2, (self.oc, self.ic, 3, 3))
            w1 = synth.wei_s8(3, (self.oc1, self.oc))
        bia0 = synth.bias(4, self.oc, self.b0) if self.b0 else None
        bia1 = synth.bias(5, self.oc1, self.b1) if self.b1 else None
        if self.data == "reference-range":
            s0 = np.array([1.0], np.float32) if not self.per_channel else np.full(self.oc, 1.0, np.float32)
            s1 = np.array([1.0], np.float32) if not self.per_channel else np.full(self.oc1, 1.0, np.float32)
        elif self.data == "ties":  # scale 0.5 makes every odd accumulator an exact .5 tie
            s0 = np.array([0.5], np.float32) if not self.per_channel else np.full(self.oc, 0.5, np.float32)
            s1 = np.array([0.5], np.float32) if not self.per_channel else np.full(self.oc1, 0.5, np.float32)
        elif self.per_channel:
            s0, s1 = synth.channel_scales(self.oc, self.k0), synth.channel_scales(self.oc1, self.k1)
        else:
            s0 = np.array([2.0 ** -self.k0], np.float32)
            s1 = np.array([2.0 ** -self.k1], np.float32)
        return src, w0, w1, bia0, bia1, s0, s1

    def blocked(self, w0, w1):
        return layout.oihw_to_blocked(w0), layout.oihw_to_blocked(w1.reshape(self.oc1, self.oc, 1, 1))


# small enough for the scalar oracle to finish in well under a second each
SMALL_CONV = [
    ConvCase("tiny16", 5, 5, 3, 16, 16, 16, "u8", None, None, k0=8, k1=8),
    ConvCase("ragged", 2, 9, 7, 32, 48, 80, "s8", "u8", "f32"),
    ConvCase("one_pixel", 3, 1, 1, 64, 64, 64, "u8", "s8", "s8", k0=9),
    ConvCase("one_row", 2, 1, 20, 32, 32, 48, "s32", "s32", None),
    ConvCase("one_col", 2, 20, 1, 32, 32, 48, "f32", "f32", "f32", relu1=1),
    ConvCase("cfg1_crop", 1, 12, 56, 64, 64, 256, "u8", "s32", "s32"),
    ConvCase("cfg3_crop", 2, 6, 28, 128, 128, 512, "u8", "s32", "s32"),
    ConvCase("cfg4_img", 2, 14, 14, 256, 256, 1024, "u8", "s32", "s32"),
    ConvCase("cfg4_f32", 2, 14, 14, 256, 256, 1024, "f32", "f32", "u8", r0=1),
    ConvCase("cfg4_s32", 2, 14, 14, 256, 256, 1024, "s32", "s8", None, r1=1),
    ConvCase("s8_norelu", 2, 8, 8, 64, 64, 128, "s8", "s32", "s32", k1=10),
    ConvCase("s8_relu_down", 2, 8, 8, 64, 64, 128, "s8", "s32", "s32", r0=1, r1=1, relu1=1, k1=10),
    ConvCase("single_scale", 2, 8, 8, 64, 64, 128, "u8", "s32", "s32", per_channel=False),
    ConvCase("refrange", 2, 8, 8, 32, 32, 64, "u8", "s32", "s32", data="reference-range"),
    ConvCase("refrange_s32", 2, 8, 8, 32, 32, 64, "s32", None, None, data="reference-range"),
    ConvCase("ties_rn", 2, 6, 6, 16, 16, 32, "u8", None, None, data="ties"),
    ConvCase("ties_rd", 2, 6, 6, 16, 16, 32, "u8", None, None, r0=1, r1=1, data="ties"),
    ConvCase("extreme", 1, 6, 6, 256, 32, 32, "s32", "s32", "s32", k0=16, data="extreme"),
    ConvCase("ic96_oc80", 2, 7, 9, 96, 80, 144, "u8", "s32", "s32"),
    ConvCase("ic160", 1, 7, 9, 160, 64, 272, "u8", "u8", "u8"),
]

# BASELINE.json configs at full size (GPU parity uses the AVX-512 replay as the checker)
FULL_CONV = [
    ConvCase("cfg1", 1, 56, 56, 64, 64, 256, "u8", "s32", "s32"),
    ConvCase("cfg3", 64, 28, 28, 128, 128, 512, "u8", "s32", "s32"),
    ConvCase("cfg4_u8", 256, 14, 14, 256, 256, 1024, "u8", "s32", "s32"),
    ConvCase("cfg4_f32", 256, 14, 14, 256, 256, 1024, "f32", "s32", "s32"),
    ConvCase("cfg4_s32", 256, 14, 14, 256, 256, 1024, "s32", "s32", "s32"),
]

# the reference's own concat test list, NCHW dims as written in test/test_concat.cc:122-153
CONCAT_BASIC = [
    ([(2, 64, 1, 1), (2, 96, 1, 1)], (2, 160, 1, 1)),
    ([(2, 64, 4, 4), (2, 32, 4, 4)], (2, 96, 4, 4)),
    ([(2, 16, 8, 8), (2, 32, 8, 8)], (2, 48, 8, 8)),
    ([(2, 32, 9, 9), (2, 96, 9, 9)], (2, 128, 9, 9)),
    ([(2, 16, 3, 3), (2, 32, 3, 3), (2, 64, 3, 3)], (2, 112, 3, 3)),
    ([(2, 256, 16, 16), (2, 256, 16, 16)], (2, 512, 16, 16)),
    ([(4, 128, 14, 14), (4, 256, 14, 14)], (4, 384, 14, 14)),
]
CONCAT_32BIT_EXTRA = [
    ([(2, 4, 4, 4), (2, 8, 4, 4)], (2, 12, 4, 4)),
    ([(2, 16, 4, 4), (2, 8, 4, 4)], (2, 24, 4, 4)),
]
# BASELINE.json configs[1]: Inception-style 28x28, C = 64/128/32/32, batch 32
CONCAT_CFG2 = ([(32, 64, 28, 28), (32, 128, 28, 28), (32, 32, 28, 28), (32, 32, 28, 28)], (32, 256, 28, 28))


def concat_inputs(dt, src_dims_nchw, data="reference-range", seed0=10):
    """NHWC numpy inputs.  'reference-range' = test/test_utils.h:49-63; 'full' = whole dtype range."""
    out = []
    for i, (n, c, h, w) in enumerate(src_dims_nchw):
        shape = (n, h, w, c)
        if dt == "f32":
            if data == "reference-range":
                idx = np.arange(int(np.prod(shape)), dtype=np.float32)
                a = (np.float32(1) + np.float32(1e-2) * np.sin(np.mod(idx, 37).astype(np.float32))).astype(np.float32)
                a = a.reshape(shape)
            else:
                a = (synth.uniform_int(seed0 + i, shape, -100000, 100000, np.int32).astype(np.float32) / np.float32(7))
                flat = a.reshape(-1)
                flat[::97] = -0.0
                flat[5::193] = np.nan
                flat[7::211] = -np.inf
        elif dt == "s32":
            lo, hi = (-10, 10) if data == "reference-range" else (-(2 ** 31), 2 ** 31 - 1)
            a = synth.uniform_int(seed0 + i, shape, lo, hi, np.int32)
        elif dt == "s8":
            lo, hi = (-10, 10) if data == "reference-range" else (-128, 127)
            a = synth.uniform_int(seed0 + i, shape, lo, hi, np.int8)
        else:
            lo, hi = (0, 16) if data == "reference-range" else (0, 255)
            a = synth.uniform_int(seed0 + i, shape, lo, hi, np.uint8)
        out.append(np.ascontiguousarray(a))
    return out
