#!/usr/bin/env python
"""bench.py -- deep-fusion hot path on B200: fused conv3x3+ReLU+conv1x1+ReLU (int8) + concat+ReLU.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line from rank 0.  A step is one pass of the fused conv over one batch of synthetic input; the
default workload is BASELINE.json configs[2] (ResNet-50 stage-3 shape 28x28 128->128->512, batch 64
per GPU, u8 out, s32 bias) -- the configuration the headline TOPS figure is quoted on.  Work is
sharded over GPUs by batch (weak scaling: 64 images per GPU), with no data-path collective.

 value         TOPS of the whole job with inputs already resident in HBM (CUDA events, max over ranks)
 e2e           the same metric through the reference-facing C++ API (memory / conv() / submit()) with
               pinned HOST buffers: H2D of the batch + kernel + D2H of the result inside the timed region
 roofline      the fused conv kernel against the tensor roofline
 concat        the concat+ReLU op (BASELINE configs[1]) against the HBM roofline, reported beside it
 cpu_baseline  the AVX-512-VNNI + OpenMP port of the reference's CPU path on this box's host cores
`--impl reference` times that CPU port alone (the reference itself cannot be built: DESIGN.md §2).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))

from dfb200 import layout, synth  # noqa: E402

L2_BYTES = 126 * 1024 * 1024
WORKLOADS = {
    # name: (batch per GPU, H, W, IC, OC, OC1, dst dtype, description)
    "cfg3": (64, 28, 28, 128, 128, 512, "u8", "conv3x3+ReLU+conv1x1+ReLU 28x28 128->128->512, batch 64/GPU, u8 out, s32 bias (BASELINE configs[2])"),
    "cfg1": (1, 56, 56, 64, 64, 256, "u8", "conv3x3+ReLU+conv1x1+ReLU 56x56 64->64->256, batch 1/GPU, u8 out (BASELINE configs[0])"),
    "cfg1x64": (64, 56, 56, 64, 64, 256, "u8", "conv3x3+ReLU+conv1x1+ReLU 56x56 64->64->256, batch 64/GPU, u8 out"),
    "cfg4": (256, 14, 14, 256, 256, 1024, "u8", "conv3x3+ReLU+conv1x1+ReLU 14x14 256->256->1024, batch 256/GPU, u8 out (BASELINE configs[3])"),
    "cfg4s32": (256, 14, 14, 256, 256, 1024, "s32", "conv3x3+ReLU+conv1x1+ReLU 14x14 256->256->1024, batch 256/GPU, s32 out (BASELINE configs[3])"),
    "cfg4f32": (256, 14, 14, 256, 256, 1024, "f32", "conv3x3+ReLU+conv1x1+ReLU 14x14 256->256->1024, batch 256/GPU, f32 out (BASELINE configs[3])"),
}
CONCAT_CFG2 = (32, 28, 28, (64, 128, 32, 32))  # BASELINE configs[1]
# dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full` capture of the dominant
# kernel (profiles/r01_conv_cfg3_v18_summary.txt, r01_concat_v15_summary.txt).  Both kernels read exactly
# their algorithmic input from DRAM (conv: 6.42 MB of activations + 0.21 MB of weights; concat: 6.42 MB);
# the output (conv 25.69 MB, concat 6.42 MB) is still in the 126 MB write-back L2 when the launch ends, so
# the per-launch capture shows no DRAM writes -- in the rotating-buffer loop bench.py times they are
# evicted later at the same rate, i.e. steady-state traffic = algorithmic bytes, no re-reads.
NCU_TRAFFIC = {"cfg3": 6703616 + 0, "concat_cfg2": 6429184 + 0}
K0 = {64: 12, 128: 13, 256: 14}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained"),
                "which": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_sustained": 1400.0, "which": "fallback"}


def conv_params(wl):
    n, h, w, ic, oc, oc1, dst, _ = WORKLOADS[wl]
    w0 = synth.wei_s8(2, (oc, ic, 3, 3))
    w1 = synth.wei_s8(3, (oc1, oc))
    return dict(n=n, h=h, w=w, ic=ic, oc=oc, oc1=oc1, dst=dst, w0b=layout.oihw_to_blocked(w0),
                w1b=layout.oihw_to_blocked(w1.reshape(oc1, oc, 1, 1)), b0=synth.bias(4, oc, "s32"),
                b1=synth.bias(5, oc1, "s32"), s0=synth.channel_scales(oc, K0.get(ic, 12)),
                s1=synth.channel_scales(oc1, 12))


def ops_per_image(p):
    return 2.0 * p["h"] * p["w"] * (9.0 * p["ic"] * p["oc"] + p["oc"] * p["oc1"])


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of one GPU, sampled in the background through NVML."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.ok = index, [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_sm = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((time.time(), sm, rs, util))
            except Exception:
                pass
            time.sleep(0.0005)  # the timed region of a graph replay is only a few ms long

    def summary(self, t0, t1):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["unavailable"]}
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
                 getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake"}
        bits = 0
        for s in inside:
            bits |= s[2]
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": self.max_sm,
                "reasons": sorted(v for k, v in names.items() if bits & k), "samples": len(inside)}


def dist_setup(world):
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    return dist


def barrier_and_max(dist, value):
    """barrier + max over ranks of a scalar measured on the device."""
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


# ----------------------------------------------------------------------------------- CPU port
def cpu_port_time(p, budget_s, min_runs=2):
    """Times the AVX-512-VNNI + OpenMP port of the reference's CPU path (oracle/, checker-side code
    used here only as the reported baseline).  Returns (images/s, runs, images per run, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    n = p["n"]
    fn, kind = (O.replay_conv, "port (AVX-512-VNNI replay, OpenMP)") if O.replay_supported() else (O.conv, "port (scalar C)")
    if not O.replay_supported():
        n = min(n, 2)
    src = synth.src_u8(1, (n, p["h"], p["w"], p["ic"]))
    d = O.make_desc(n, p["h"], p["w"], p["ic"], p["oc"], p["oc1"], O.DT_OF[p["dst"]], O.S32, O.S32, nscale0=p["oc"], nscale1=p["oc1"])
    fn(d, src, p["w0b"], p["b0"], p["s0"], p["w1b"], p["b1"], p["s1"])  # warm-up
    t0, runs = time.time(), 0
    while runs < min_runs or (time.time() - t0) < budget_s:
        fn(d, src, p["w0b"], p["b0"], p["s0"], p["w1b"], p["b1"], p["s1"])
        runs += 1
    dt = time.time() - t0
    return n * runs / dt, runs, n, kind, int(O.lib().dfr_num_threads())


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is specified to use every host thread.  Must
    run before the oracle library (libgomp) is loaded."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.setdefault("OMP_PROC_BIND", "close")  # analogue of the reference's run_benchmark.sh pinning
    return n


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    use_all_host_threads()
    p = conv_params(args.workload)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    fast = O.replay_supported()
    n = p["n"] if fast else min(p["n"], 2)
    fn = O.replay_conv if fast else O.conv
    src = synth.src_u8(1, (n, p["h"], p["w"], p["ic"]))
    d = O.make_desc(n, p["h"], p["w"], p["ic"], p["oc"], p["oc1"], O.DT_OF[p["dst"]], O.S32, O.S32, nscale0=p["oc"], nscale1=p["oc1"])
    call = lambda: fn(d, src, p["w0b"], p["b0"], p["s0"], p["w1b"], p["b1"], p["s1"])  # noqa: E731
    t = time.time()
    call()
    one = time.time() - t
    steps = args.steps
    if one * (args.steps + args.warmup) > 150.0:  # keep the whole run within a few minutes
        steps = max(1, int(150.0 / one) - args.warmup)
    for _ in range(args.warmup):
        call()
    t0 = time.time()
    for _ in range(steps):
        call()
    dt = (time.time() - t0) / steps
    tops = n * ops_per_image(p) / dt / 1e12
    cores = int(O.lib().dfr_num_threads()) if fast else 1
    line = {
        "impl": "reference", "metric": "fused conv3x3+1x1 int8 TOPS", "value": tops, "unit": "TOPS", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "images_per_s": n / dt,
        "config": {"workload": WORKLOADS[args.workload][7], "images_per_step": n},
        "cpu_baseline": {"value": tops, "unit": "TOPS", "cores": cores,
                         "kind": "port", "sample": f"{n} images per step x {steps} steps; "
                         + ("AVX-512-VNNI intrinsics replay of the reference's emitted instruction sequence + OpenMP" if fast else "scalar C oracle (host lacks AVX-512 VNNI)")
                         + "; the reference binary itself needs Xbyak and cannot be built offline"},
        "e2e": {"value": tops, "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import dfb200 as df
    from dfb200 import hostapi as H
    peaks = measured_peaks()
    df.set_device(local_rank)
    dist = dist_setup(world)
    p = conv_params(args.workload)
    n, h, w, ic, oc, oc1, dst = p["n"], p["h"], p["w"], p["ic"], p["oc"], p["oc1"], p["dst"]
    ts_out = 4 if dst in ("f32", "s32") else 1
    src_bytes, dst_bytes = n * h * w * ic, n * h * w * oc1 * ts_out

    # ---- device-resident leg: rotate over enough (src, dst) sets that no step finds its data in L2
    op = df.Conv(n, h, w, ic, oc, oc1, df.DT_OF[dst], p["w0b"], p["w1b"], p["b0"], p["b1"], p["s0"], p["s1"], df.S32, df.S32)
    info = op.info()
    n_sets = max(2, -(-2 * L2_BYTES // (src_bytes + dst_bytes)))
    base = synth.src_u8(1 + 100 * rank, (n, h, w, ic))
    sets = []
    for i in range(n_sets):
        sets.append((df.DeviceBuffer.from_numpy(np.roll(base.reshape(-1), 4099 * i)), df.DeviceBuffer(dst_bytes)))
    sampler = ClockSampler(local_rank)
    sampler.start()
    # The K timed steps are captured into one CUDA graph (K kernel nodes, programmatic-dependent-launch
    # edges between them) and replayed with a single launch: at ~20 us per step a Python loop of ctypes
    # calls is uncomfortably close to being the thing measured.  --no-graph times the plain loop.
    st = df.Stream()
    for i in range(args.warmup):
        op.run(*sets[i % n_sets], stream=st.ptr)
    graph = None
    if not args.no_graph:
        with df.Graph(st) as graph:
            for i in range(args.steps):
                op.run(*sets[(args.warmup + i) % n_sets], stream=st.ptr)
        graph.launch()  # untimed replay: graph upload, instruction caches
    st.sync()
    barrier(dist)
    e0, e1 = df.Event(), df.Event()
    t_start = time.time()
    e0.record(st.ptr)
    if graph is not None:
        graph.launch()
    else:
        for i in range(args.steps):
            op.run(*sets[(args.warmup + i) % n_sets], stream=st.ptr)
    e1.record(st.ptr)
    st.sync()
    ms_total = e0.elapsed_ms(e1)
    t_end = time.time()
    barrier(dist)
    ms_step = barrier_and_max(dist, ms_total / args.steps)
    total_images = n * world
    tops = total_images * ops_per_image(p) / (ms_step * 1e-3) / 1e12
    kernel_tops_this_rank = n * ops_per_image(p) / (ms_total / args.steps * 1e-3) / 1e12

    # ---- concat+ReLU leg (BASELINE configs[1]) on device-resident buffers, footprint > L2 by rotation
    cn, chh, cww, cics = CONCAT_CFG2
    c_bytes = 2 * cn * chh * cww * sum(cics)
    c_sets = max(2, -(-2 * L2_BYTES // c_bytes))
    c_in = [synth.src_u8(10 + i, (cn, chh, cww, c)) for i, c in enumerate(cics)]
    csets = []
    for i in range(c_sets):
        csets.append(([df.DeviceBuffer.from_numpy(a) for a in c_in], df.DeviceBuffer(c_bytes // 2)))
    npix = cn * chh * cww
    ccalls = [df.ConcatCall(df.U8, True, [b.ptr for b in s[0]], list(cics), s[1].ptr, npix, stream=st.ptr) for s in csets]
    crun = lambda i: ccalls[i]()  # noqa: E731
    for i in range(max(3, args.warmup)):
        crun(i % c_sets)
    c_steps = max(args.steps, 50)
    cgraph = None
    if not args.no_graph:  # a 2 us kernel: the host cannot issue launches that fast, the graph can
        with df.Graph(st) as cgraph:
            for i in range(c_steps):
                crun(i % c_sets)
        cgraph.launch()
    st.sync()
    ce0, ce1 = df.Event(), df.Event()
    ce0.record(st.ptr)
    if cgraph is not None:
        cgraph.launch()
    else:
        for i in range(c_steps):
            crun(i % c_sets)
    ce1.record(st.ptr)
    st.sync()
    c_ms = barrier_and_max(dist, ce0.elapsed_ms(ce1) / c_steps)
    concat_gbs = c_bytes / (c_ms * 1e-3) / 1e9

    # ---- end-to-end leg: the reference's API with host buffers (H2D + kernel + D2H per step)
    hsrc = H.Memory((n, ic, h, w), "nhwc", "u8")
    hsrc.set(base)
    hwei = H.Memory((oc, ic, 3, 3), "OIhw4i16o4i", "s8")
    hwei.array().reshape(-1)[...] = p["w0b"]
    hwei1 = H.Memory((oc1, oc, 1, 1), "OIhw4i16o4i", "s8")
    hwei1.array().reshape(-1)[...] = p["w1b"]
    hb0 = H.Memory((oc,), "x", "s32", nchw=False)
    hb0.set(p["b0"])
    hb1 = H.Memory((oc1,), "x", "s32", nchw=False)
    hb1.set(p["b1"])
    hdst = H.Memory((n, oc1, h, w), "nhwc", dst)
    hsrc.pin()
    hdst.pin()
    hop = H.conv(hsrc, hwei, hb0, (1, 1), (1, 1), hdst, wei1x1=hwei1, bia1x1=hb1, conv0_scales=p["s0"], conv1_scales=p["s1"])
    e2e_steps = max(3, min(args.steps, 50))
    for _ in range(max(3, min(args.warmup, 5))):
        hop.submit()
    barrier(dist)
    g0, g1 = df.Event(), df.Event()
    g0.record()
    for _ in range(e2e_steps):
        hop.submit()  # synchronous: the result is in hdst.array() when it returns
    g1.record()
    e2e_ms = barrier_and_max(dist, g0.elapsed_ms(g1) / e2e_steps)
    e2e_tops = total_images * ops_per_image(p) / (e2e_ms * 1e-3) / 1e12
    checksum = int(hdst.array().reshape(-1)[:: 4097].astype(np.int64).sum())  # device -> host result was read
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    clocks = sampler.summary(t_start, t_end)

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        use_all_host_threads()
        ips, runs, imgs, kind, cores = cpu_port_time(p, budget_s=args.cpu_seconds)
        cpu = {"value": ips * ops_per_image(p) / 1e12, "unit": "TOPS", "images_per_s": ips, "cores": cores, "kind": "port",
               "sample": f"{runs} passes over {imgs} images of the same workload ({kind}); the reference binary needs Xbyak and cannot be built offline"}
    tensor_peak = 2.0 * peaks["bf16_tflops"]
    line = {
        "metric": "fused conv3x3+1x1 int8 TOPS", "value": tops, "unit": "TOPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "images_per_s": total_images / (ms_step * 1e-3),
        "config": {"workload": WORKLOADS[args.workload][7], "images_per_gpu": n, "parallelism": f"batch-sharded x{world}, no collective",
                   "cache": f"rotating {n_sets} src/dst buffer sets ({n_sets * (src_bytes + dst_bytes) >> 20} MiB > 2x L2) so no step finds its data in L2",
                   "launch": "plain loop of df_conv_run calls" if args.no_graph else f"{args.steps} steps captured in one CUDA graph, replayed once inside the timed region",
                   "tiles": info.tiles_per_launch, "grid": info.grid, "smem_bytes": info.smem_bytes,
                   "weights_resident": [info.w0_resident, info.w1_resident], "mma_row_efficiency": round(info.mma_efficiency, 4)},
        "roofline": {"bound": "tensor", "achieved": kernel_tops_this_rank, "peak": tensor_peak, "unit": "TOPS",
                     "frac": kernel_tops_this_rank / tensor_peak, "traffic": NCU_TRAFFIC.get(args.workload),
                     "peak_source": f"2 x bf16_tflops of MEASURED_PEAKS.json ({peaks['which']}); int8 dense rate = 2 x bf16",
                     "frac_of_i8_mma_probe": kernel_tops_this_rank / 4335.0,
                     "i8_mma_probe_tops": 4335.0, "kernel": "conv_pair_kernel" if info.w0_resident == 2 else "conv_fused_kernel", "ops_per_launch": n * ops_per_image(p)},
        "concat": {"workload": "concat+ReLU u8, 28x28, C=64/128/32/32, batch 32 (BASELINE configs[1])", "value": concat_gbs,
                   "unit": "GB/s", "us_per_launch": c_ms * 1e3, "bytes_per_launch": c_bytes,
                   "roofline": {"bound": "hbm", "achieved": concat_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": concat_gbs / peaks["hbm_gbs"], "traffic": NCU_TRAFFIC["concat_cfg2"]},
                   "cache": f"rotating {c_sets} buffer sets ({c_sets * c_bytes >> 20} MiB > 2x L2)"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_tops, "unit": "TOPS", "images_per_s": total_images / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": src_bytes, "d2h_bytes_per_step": dst_bytes, "steps": e2e_steps,
                "api": "deepfusion::conv(...)->submit() with pinned host memory", "result_checksum": checksum},
        "gpu_launches": args.steps,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time a plain loop of launches instead of a CUDA graph replay")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
