"""BASELINE configs[4]: the fused conv (cfg3 shape) over batch N = 1 .. 2048 split across G = 1 / 2 / 4 / 8 GPUs of one
box by ONE process through ext::conv_sharded (contiguous batch slabs, no collective), next to the reference's CPU path
(the AVX-512 port, all host threads) on the same batch.  Needs GPUs.

  kernel : slabs resident on the devices; host wall clock from the first launch to the last sync,
           median of `reps` repetitions
  e2e    : op->submit() with pinned host buffers: upload + kernels + download of every slab

usage: sweep_sharded.py [out.jsonl] [max_gpus]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np
import dfb200 as df
from dfb200 import hostapi as H, synth, layout

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "sweep_sharded.jsonl")
max_g = int(sys.argv[2]) if len(sys.argv) > 2 else 8
h, w, ic, oc, oc1 = 28, 28, 128, 128, 512
ops_img = 2.0 * h * w * (9.0 * ic * oc + oc * oc1)
w0b = layout.oihw_to_blocked(synth.wei_s8(2, (oc, ic, 3, 3)))
w1b = layout.oihw_to_blocked(synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1))
b0, b1 = synth.bias(4, oc, "s32"), synth.bias(5, oc1, "s32")
s0, s1 = synth.channel_scales(oc, 13), synth.channel_scales(oc1, 12)
n_dev = min(df.device_count(), max_g)
gs = [g for g in (1, 2, 4, 8) if g <= n_dev]

os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
import bench  # CpuArm
rows = []
with open(out_path, "w") as f:
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048):
        p = dict(n=n, h=h, w=w, ic=ic, oc=oc, oc1=oc1, dst="u8", w0b=w0b, w1b=w1b, b0=b0, b1=b1, s0=s0, s1=s1)
        arm = bench.CpuArm(p)
        arm.warm(0.3, 2)
        cpu_s = arm.timed(max(2, min(20, arm.budget_steps(1.0))))
        row = {"n": n, "cpu_images_per_s": n / cpu_s, "cpu_tops": n * ops_img / cpu_s / 1e12, "cpu_cores": arm.cores}
        src = H.Memory((n, ic, h, w), "nhwc", "u8"); src.set(arm.src)
        wei = H.Memory((oc, ic, 3, 3), "OIhw4i16o4i", "s8"); wei.array().reshape(-1)[...] = w0b
        wei1 = H.Memory((oc1, oc, 1, 1), "OIhw4i16o4i", "s8"); wei1.array().reshape(-1)[...] = w1b
        hb0 = H.Memory((oc,), "x", "s32", nchw=False); hb0.set(b0)
        hb1 = H.Memory((oc1,), "x", "s32", nchw=False); hb1.set(b1)
        dst = H.Memory((n, oc1, h, w), "nhwc", "u8")
        for g in gs:
            op = H.conv_sharded(list(range(g)), src, wei, hb0, (1, 1), (1, 1), dst, wei1x1=wei1, bia1x1=hb1, conv0_scales=s0, conv1_scales=s1)
            op.upload()
            reps, inner = 7, (50 if n <= 256 else 10)
            ts = []
            for _ in range(reps):
                op.sync(); t = time.perf_counter()
                for _ in range(inner): op.submit_device()
                op.sync(); ts.append((time.perf_counter() - t) / inner)
            k_s = float(np.median(ts))
            op.submit(); ts = []
            for _ in range(5):
                t = time.perf_counter(); op.submit(); ts.append(time.perf_counter() - t)
            e_s = float(np.median(ts))
            row[f"g{g}"] = {"kernel_us": k_s * 1e6, "kernel_tops": n * ops_img / k_s / 1e12, "kernel_images_per_s": n / k_s,
                            "e2e_us": e_s * 1e6, "e2e_tops": n * ops_img / e_s / 1e12, "e2e_images_per_s": n / e_s}
            del op
        f.write(json.dumps(row) + "\n"); f.flush()
        print(f"N={n:5d} cpu {row['cpu_tops']:6.2f} TOPS | " + " | ".join(f"G={g}: {row[f'g{g}']['kernel_tops']:7.1f} / e2e {row[f'g{g}']['e2e_tops']:6.1f}" for g in gs), flush=True)
