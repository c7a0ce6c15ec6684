"""CPU tests of the C++ host layer (libdeepfusion.so): `memory` semantics and the creation-time
validation that mirrors op_conv::init_conf / jit_concat_kernel::init_conf.  Creation failures exit
the process like the reference (util/log.h:38-42), so those run in a subprocess."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from dfb200 import hostapi as H


def test_memory_nchw_ctor_reorders_for_nhwc():
    m = H.Memory((2, 48, 5, 7), "nhwc", "u8")
    assert m.nbytes == 2 * 48 * 5 * 7 and m.shape == (2, 5, 7, 48)
    a = m.array()
    a[...] = 3
    assert int(m.array().sum()) == 3 * m.nbytes
    assert a.ctypes.data % 4096 == 0  # default alignment (include/deepfusion.h:77-80)


def test_memory_sizes_per_dtype():
    assert H.Memory((1, 16, 2, 2), "nhwc", "f32").nbytes == 256
    assert H.Memory((1, 16, 2, 2), "nhwc", "s32").nbytes == 256
    assert H.Memory((1, 16, 2, 2), "nhwc", "s8").nbytes == 64
    assert H.Memory((64, 64, 3, 3), "OIhw4i16o4i", "s8").nbytes == 64 * 64 * 9
    assert H.Memory((256,), "x", "s32", nchw=False).nbytes == 1024


def _run(code):
    env = dict(os.environ)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pre = f"import sys; sys.path.insert(0, {os.path.join(root, 'deep-fusion_b200')!r})\nfrom dfb200 import hostapi as H\n"
    return subprocess.run([sys.executable, "-c", pre + textwrap.dedent(code)], capture_output=True, text=True, env=env,
                          timeout=120)


def test_concat_creation_failure_exits_with_reference_message():
    r = _run("""
        a = H.Memory((2, 24, 4, 4), "nhwc", "u8"); b = H.Memory((2, 16, 4, 4), "nhwc", "u8")
        d = H.Memory((2, 40, 4, 4), "nhwc", "u8")
        H.concat([a, b], d)
        print("NOT REACHED")
    """)
    assert r.returncode != 0 and "Init Concat op failed!" in r.stderr and "NOT REACHED" not in r.stdout


def test_concat_mixed_dtype_rejected():
    r = _run("""
        a = H.Memory((2, 16, 4, 4), "nhwc", "u8"); b = H.Memory((2, 16, 4, 4), "nhwc", "s8")
        d = H.Memory((2, 32, 4, 4), "nhwc", "u8")
        H.concat([a, b], d)
    """)
    assert r.returncode != 0 and "Init Concat op failed!" in r.stderr


def _conv_script(**kw):
    p = dict(n=1, ic=64, oc=64, oc1=128, h=8, w=8, oh=8, ow=8, src_dt="u8", dst_c=128, wfmt="OIhw4i16o4i", k1=1,
             stride=1, pad=1, ns0=1)
    p.update(kw)
    return """
        src = H.Memory(({n}, {ic}, {h}, {w}), "nhwc", "{src_dt}")
        wei = H.Memory(({oc}, {ic}, 3, 3), "{wfmt}", "s8")
        w1 = H.Memory(({oc1}, {oc}, {k1}, {k1}), "OIhw4i16o4i", "s8")
        dst = H.Memory(({n}, {dst_c}, {oh}, {ow}), "nhwc", "u8")
        H.conv(src, wei, None, ({stride}, {stride}), ({pad}, {pad}), dst, wei1x1=w1, conv0_scales=[1.0] * {ns0})
        print("CREATED")
    """.format(**p)


def test_conv_creation_rules_exit_like_reference():
    for kw, msg in [
        (dict(oh=7), "Output image size do not match"),          # op_conv.cc:291-298
        (dict(dst_c=64), "Conv1x1 output channel do not match"),  # :330-333
        (dict(k1=3), "Fused conv must be 1x1 kernel"),            # :334-337
        (dict(src_dt="s8"), "u8 src"),                            # jit_conv_kernel.cc:531
        (dict(wfmt="nchw"), "formats"),                           # :552-564
        (dict(ns0=3), ""),                                        # op_conv.cc:342-345
    ]:
        r = _run(_conv_script(**kw))
        assert r.returncode != 0, kw
        assert "Init Conv op failed!" in r.stderr, (kw, r.stderr)
        assert msg in r.stdout + r.stderr, (kw, r.stdout, r.stderr)
        assert "CREATED" not in r.stdout


def test_conv_outside_b200_path_says_so():
    """What the reference accepts but the B200 path does not run (padding that makes the output larger than the input)
    exits with its own message -- never a CPU fallback."""
    r = _run(_conv_script(pad=2, oh=10, ow=10))
    assert r.returncode != 0 and "unsupported on B200 path" in r.stderr
