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
 other_shapes  the other BASELINE conv shapes (cfg1 batch 1 / 64, cfg4 u8 / s32 / f32) and a >= 1 s sustained run of
               the headline shape, each against its own roofline (HBM where the arithmetic intensity is below the ridge)
 cpu_baseline  the AVX-512-VNNI + OpenMP port of the reference's CPU path on this box's host cores (hot and
               cold cache, destination pre-allocated, warmed up by time)
`--impl reference` times that CPU port alone.  The port is bit-identical to the reference's own kernel
generators executed through oracle/_ref (tests/test_ref_pin.py); that interpreter is a checker, far slower
than the JIT code it stands for, so it is not what is timed.
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
K0 = {64: 12, 128: 13, 256: 14}


def ncu_constants():
    """Numbers that come from profiler captures / probes rather than from this run, read from
    profiles/ncu_constants.json (written by scripts/ncu_constants.py from the .ncu-rep files; every entry
    names its source).  Absent file or key -> None: nothing here is hard-coded."""
    p = os.path.join(ROOT, "profiles", "ncu_constants.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


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
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import atexit
    atexit.register(lambda: dist.is_initialized() and dist.destroy_process_group())
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
class CpuArm:
    """The AVX-512-VNNI + OpenMP port of the reference's CPU path (oracle/df_replay_avx512.c; checker-side code,
    used here only as the reported baseline) on one batch of the workload.  The destination is allocated
    ONCE (the reference's ops also own theirs), the arm is warmed up by time, and besides the hot-cache rate
    it measures the reference's cold-cache protocol: every thread scrubs 2 MB between iterations
    (test/test_utils.cc:23-49, benchmark/bench_concat.cc:86-110)."""

    def __init__(self, p):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        self.O, self.p = O, p
        self.fast = O.replay_supported()
        self.n = p["n"] if self.fast else min(p["n"], 2)
        self.fn = O.replay_conv if self.fast else O.conv
        self.cores = int(O.lib().dfr_num_threads()) if self.fast else 1
        self.src = synth.src_u8(1, (self.n, p["h"], p["w"], p["ic"]))
        self.d = O.make_desc(self.n, p["h"], p["w"], p["ic"], p["oc"], p["oc1"], O.DT_OF[p["dst"]], O.S32, O.S32,
                             nscale0=p["oc"], nscale1=p["oc1"])
        self.dst = np.zeros((self.n, p["h"], p["w"], p["oc1"]), dtype=O.NP_OF[O.DT_OF[p["dst"]]])
        self.scrub = np.zeros(max(1, self.cores) * (2 << 20), np.uint8)
        self.kind = ("AVX-512-VNNI intrinsics replay of the reference's emitted instruction sequence + OpenMP, bit-identical to the "
                     "reference generators run through oracle/_ref" if self.fast else "scalar C oracle (host lacks AVX-512 VNNI)")

    def call(self):
        p = self.p
        self.fn(self.d, self.src, p["w0b"], p["b0"], p["s0"], p["w1b"], p["b1"], p["s1"], out=self.dst)

    def warm(self, seconds, min_calls):
        t0, k = time.time(), 0
        while k < min_calls or time.time() - t0 < seconds:
            self.call()
            k += 1
        return k

    def timed(self, steps, cold=False):
        """seconds per step (mean over `steps`), each step timed on its own"""
        tot = 0.0
        for _ in range(steps):
            if cold:
                self.scrub += 1  # evicts what the previous iteration left in the caches
            t = time.perf_counter()
            self.call()
            tot += time.perf_counter() - t
        return tot / steps

    def budget_steps(self, seconds, lo=3):
        t = time.perf_counter()
        self.call()
        one = time.perf_counter() - t
        return max(lo, int(seconds / max(one, 1e-6)))


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is specified to use every host thread.  Must
    run before the oracle library (libgomp) is loaded."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.setdefault("OMP_PROC_BIND", "close")  # analogue of the reference's run_benchmark.sh pinning
    return n


def shared_config(wl, n, world):
    """`config` is the same object in both arms (the driver compares them)."""
    return {"workload": WORKLOADS[wl][7], "images_per_gpu": n, "parallelism": f"batch-sharded x{world}, no collective"}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    use_all_host_threads()
    p = conv_params(args.workload)
    arm = CpuArm(p)
    n = arm.n
    steps = min(args.steps, arm.budget_steps(120.0))  # keep the whole run within a few minutes
    arm.warm(2.0, args.warmup)                         # by time: clocks, page tables, OpenMP team
    dt = arm.timed(steps)
    dt_cold = arm.timed(max(3, min(steps, 20)), cold=True)
    tops = n * ops_per_image(p) / dt / 1e12
    line = {
        "impl": "reference", "metric": "fused conv3x3+1x1 int8 TOPS", "value": tops, "unit": "TOPS", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "images_per_s": n / dt,
        "config": shared_config(args.workload, p["n"], args.gpus),
        "cpu_baseline": {"value": tops, "unit": "TOPS", "cores": arm.cores, "kind": "port",
                         "sample": f"{n} images per step x {steps} steps after a 2 s warm-up, destination pre-allocated; {arm.kind}",
                         "cold_cache_value": n * ops_per_image(p) / dt_cold / 1e12,
                         "cold_cache_protocol": "2 MB per thread scrubbed between iterations (reference test/test_utils.cc:23-49)"},
        "e2e": {"value": tops, "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- GPU arm
def shape_roofline(p, peaks, tops):
    """Per-shape ceiling = min(tensor peak, arithmetic intensity x HBM bandwidth): shapes whose
    intensity is below the ridge are reported against the HBM roofline in GB/s."""
    ts_out = 4 if p["dst"] in ("f32", "s32") else 1
    bytes_img = p["h"] * p["w"] * (p["ic"] + p["oc1"] * ts_out)
    ai = ops_per_image(p) / bytes_img
    tensor_peak = 2.0 * peaks["bf16_tflops"]
    hbm_ceiling_tops = ai * peaks["hbm_gbs"] / 1e3
    if hbm_ceiling_tops < tensor_peak:
        gbs = tops * 1e3 / ai
        return {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                "ops_per_byte": ai, "ceiling_tops": hbm_ceiling_tops}
    return {"bound": "tensor", "achieved": tops, "peak": tensor_peak, "unit": "TOPS", "frac": tops / tensor_peak,
            "ops_per_byte": ai, "ceiling_tops": tensor_peak}


def time_shape(df, st, wl, steps, peaks, rank=0):
    """One more conv shape, device-resident, rotating buffers (> 2x L2 when the batch allows), CUDA-graph replay."""
    p = conv_params(wl)
    n, h, w, ic, oc, oc1, dst = p["n"], p["h"], p["w"], p["ic"], p["oc"], p["oc1"], p["dst"]
    ts_out = 4 if dst in ("f32", "s32") else 1
    src_bytes, dst_bytes = n * h * w * ic, n * h * w * oc1 * ts_out
    op = df.Conv(n, h, w, ic, oc, oc1, df.DT_OF[dst], p["w0b"], p["w1b"], p["b0"], p["b1"], p["s0"], p["s1"], df.S32, df.S32)
    n_sets = max(2, min(64, -(-2 * L2_BYTES // (src_bytes + dst_bytes))))
    base = synth.src_u8(1 + 100 * rank, (n, h, w, ic))
    sets = [(df.DeviceBuffer.from_numpy(base), df.DeviceBuffer(dst_bytes)) for _ in range(n_sets)]
    for i in range(3):
        op.run(*sets[i % n_sets], stream=st.ptr)
    with df.Graph(st) as g:
        for i in range(steps):
            op.run(*sets[i % n_sets], stream=st.ptr)
    g.launch()
    st.sync()
    e0, e1 = df.Event(), df.Event()
    g.launch()  # untimed, keeps the stream busy while the host queues the timed replay (see the main conv leg)
    e0.record(st.ptr)
    g.launch()
    e1.record(st.ptr)
    st.sync()
    us = e0.elapsed_ms(e1) / steps * 1e3
    tops = n * ops_per_image(p) / us / 1e6
    info = op.info()
    out = {"workload": WORKLOADS[wl][7], "us_per_launch": us, "tops": tops, "images_per_s": n / us * 1e6, "steps": steps,
           "footprint_mib": n_sets * (src_bytes + dst_bytes) >> 20, "weights_resident": [info.w0_resident, info.w1_resident],
           "roofline": shape_roofline(p, peaks, tops)}
    del g, e0, e1
    op.close()
    return out


def time_concat_conv(df, st, steps, rank=0, batch=None):
    """concat+ReLU (BASELINE configs[1] shape) feeding a fused conv, four ways over the same rotating
    buffer sets (> 2x L2), each a CUDA-graph replay: the concat kernel alone, the conv alone on the materialised
    tensor, the two chained, and the ONE kernel whose halo loads read the concat's inputs directly."""
    n, h, w, ics = CONCAT_CFG2
    n = batch or n
    ic, oc, oc1 = sum(ics), 128, 512
    w0b = layout.oihw_to_blocked(synth.wei_s8(2, (oc, ic, 3, 3)))
    w1b = layout.oihw_to_blocked(synth.wei_s8(3, (oc1, oc)).reshape(oc1, oc, 1, 1))
    b0, b1 = synth.bias(4, oc, "s32"), synth.bias(5, oc1, "s32")
    s0, s1 = synth.channel_scales(oc, 13), synth.channel_scales(oc1, 12)
    conv = df.Conv(n, h, w, ic, oc, oc1, df.U8, w0b, w1b, b0, b1, s0, s1, df.S32, df.S32)
    fused = df.ConcatConv(n, h, w, ics, True, oc, oc1, df.U8, w0b, w1b, b0, b1, s0, s1, df.S32, df.S32)
    px = n * h * w
    per_set = px * (2 * ic + oc1)
    n_sets = max(2, min(32, -(-2 * L2_BYTES // per_set)))
    sets = []
    for k in range(n_sets):
        ins = [df.DeviceBuffer.from_numpy(np.tile(synth.uniform_int(20 + i + 100 * rank, (min(n, 32), h, w, c), 0, 127, np.uint8), (-(-n // 32), 1, 1, 1))[:n])
               for i, c in enumerate(ics)]
        sets.append((ins, df.DeviceBuffer(px * ic), df.DeviceBuffer(px * oc1)))
    cats = [df.ConcatCall(df.U8, True, [b.ptr for b in ins], list(ics), cat.ptr, px, stream=st.ptr) for ins, cat, _ in sets]

    def timed(body):
        for k in range(3):
            body(k % n_sets)
        with df.Graph(st) as g:
            for k in range(steps):
                body(k % n_sets)
        g.launch()
        st.sync()
        e0, e1 = df.Event(), df.Event()
        g.launch()  # untimed, keeps the stream busy while the host queues the timed replay
        e0.record(st.ptr)
        g.launch()
        e1.record(st.ptr)
        st.sync()
        us = e0.elapsed_ms(e1) / steps * 1e3
        del g, e0, e1
        return us

    for k in range(n_sets):
        cats[k]()  # materialise every set's concatenated tensor once (the conv-alone leg reads them)
    st.sync()
    us_cat = timed(lambda k: cats[k]())
    us_conv = timed(lambda k: conv.run(sets[k][1], sets[k][2], stream=st.ptr))

    def chained(k):
        cats[k]()
        conv.run(sets[k][1], sets[k][2], stream=st.ptr)
    us_chain = timed(chained)
    us_fused = timed(lambda k: fused.run(sets[k][0], sets[k][2], stream=st.ptr))
    ops = 2.0 * px * (9 * ic * oc + oc * oc1)
    out = {"workload": f"concat+ReLU u8 {h}x{w} C={'/'.join(map(str, ics))} batch {n} (BASELINE configs[1]) -> conv3x3+ReLU+conv1x1+ReLU {ic}->{oc}->{oc1}, u8 out",
           "us_concat_alone": us_cat, "us_conv_alone": us_conv, "us_two_kernels": us_chain, "us_fused_one_kernel": us_fused,
           "fused_tops": ops / us_fused / 1e6, "two_kernel_tops": ops / us_chain / 1e6,
           "hbm_bytes_saved_per_launch": 2 * px * ic,
           "cache": f"rotating {n_sets} buffer sets ({n_sets * per_set >> 20} MiB > 2x L2)", "steps": steps,
           "concatenated_tensor_mib": px * ic / 2 ** 20}
    conv.close()
    fused.close()
    return out


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
    # The K timed steps are nodes of one CUDA graph together with `lead` untimed steps in front of them and one behind,
    # and the two timing events are nodes of the same graph on a forked stream: e0 fires when the last lead-in step
    # completes, e1 when the last timed step completes -- exactly K step periods of a stream of launches that overlap the way
    # back-to-back df_conv_run calls do (programmatic dependent launch: step k+1's prologue and weight loads run under step
    # k's tail).  Events outside the graph, or events as nodes IN the chain of launches, put the graph's start-up latency
    # (~18 us) or one non-overlapped fill + drain (~16 us) inside a 20-step region: 15.0 / 14.85 us per step at K = 20
    # against 14.03 at K = 200 (measured); on the forked stream K = 20 and K = 200 agree.
    graph = None
    lead = max(3, args.warmup)
    e0, e1 = df.Event(), df.Event()
    if not args.no_graph:
        side = df.Stream()
        fork0, fork1, join = df.Event(), df.Event(), df.Event()  # capture-internal ordering markers
        with df.Graph(st) as graph:
            for i in range(lead):
                op.run(*sets[i % n_sets], stream=st.ptr)
            fork0.record(st.ptr)
            df.check(df.lib().df_stream_wait_event(side.ptr, fork0.ptr))
            e0.record_node(side.ptr)
            for i in range(args.steps):
                op.run(*sets[(lead + i) % n_sets], stream=st.ptr)
            fork1.record(st.ptr)
            df.check(df.lib().df_stream_wait_event(side.ptr, fork1.ptr))
            e1.record_node(side.ptr)
            op.run(*sets[(lead + args.steps) % n_sets], stream=st.ptr)  # untimed: the last timed step ends the way the others do
            join.record(side.ptr)
            df.check(df.lib().df_stream_wait_event(st.ptr, join.ptr))
        graph.launch()  # untimed replay: graph upload, instruction caches
    st.sync()
    barrier(dist)
    t_start = time.time()
    if graph is not None:
        graph.launch()
    else:
        e0.record(st.ptr)
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
    if cgraph is not None:
        cgraph.launch()  # keeps the stream busy while the host queues the timed replay (see the conv leg)
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

    # ---- concat+ReLU at a footprint far beyond L2 (N = 1024: 205 MB per launch): the bandwidth the kernel
    #      sustains once a launch is longer than one DRAM round trip
    big = None
    if not args.quick:
        bn = 1024
        b_bytes = 2 * bn * chh * cww * sum(cics)
        b_in = [df.DeviceBuffer(bn * chh * cww * c) for c in cics]
        for b in b_in:
            b.fill(7)
        b_out = df.DeviceBuffer(b_bytes // 2)
        bcall = df.ConcatCall(df.U8, True, [b.ptr for b in b_in], list(cics), b_out.ptr, bn * chh * cww, stream=st.ptr)
        for _ in range(3):
            bcall()
        st.sync()
        be0, be1 = df.Event(), df.Event()
        be0.record(st.ptr)
        for _ in range(10):
            bcall()
        be1.record(st.ptr)
        st.sync()
        b_ms = be0.elapsed_ms(be1) / 10
        big = {"workload": "concat+ReLU u8, 28x28, C=64/128/32/32, batch 1024", "bytes_per_launch": b_bytes, "us_per_launch": b_ms * 1e3,
               "value": b_bytes / (b_ms * 1e-3) / 1e9, "unit": "GB/s", "frac_of_hbm": b_bytes / (b_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        del b_in, b_out

    # ---- the headline shape sustained for >= 1 s (clocks settle under load), against 2 x the SUSTAINED bf16 figure
    sustained = None
    if not args.quick and graph is not None:
        reps = max(1, int(1000.0 / max(ms_total, 1e-3)) + 1)
        s0, s1 = df.Event(), df.Event()
        t_s0 = time.time()
        s0.record(st.ptr)
        for _ in range(reps):
            graph.launch()
        s1.record(st.ptr)
        st.sync()
        t_s1 = time.time()
        s_ms = s0.elapsed_ms(s1) / (reps * (args.steps + lead + 1))  # every replay runs the lead-in and the trailing step too
        s_tops = n * ops_per_image(p) / (s_ms * 1e-3) / 1e12
        peak_s = 2.0 * (peaks["bf16_sustained"] or peaks["bf16_tflops"])
        sustained = {"seconds": s0.elapsed_ms(s1) / 1e3, "launches": reps * (args.steps + lead + 1), "us_per_launch": s_ms * 1e3, "tops": s_tops,
                     "peak": peak_s, "frac": s_tops / peak_s, "peak_source": "2 x bf16_tflops_sustained of MEASURED_PEAKS.json",
                     "clocks": sampler.summary(t_s0, t_s1)}

    # ---- the other BASELINE conv shapes, each against its own roofline
    others = None
    if not args.quick and world == 1:
        others = {wl: time_shape(df, st, wl, 50 if wl != "cfg1" else 200, peaks, rank) for wl in ("cfg1", "cfg1x64", "cfg4", "cfg4s32", "cfg4f32")}

    # ---- concat fused into the conv's A-operand load
    concat_conv = time_concat_conv(df, st, 50, rank) if (not args.quick and world == 1) else None

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
        arm = CpuArm(p)
        arm.warm(2.0, 3)
        steps_cpu = arm.budget_steps(args.cpu_seconds)
        dt = arm.timed(steps_cpu)
        dt_cold = arm.timed(max(3, min(steps_cpu, 20)), cold=True)
        cpu = {"value": arm.n * ops_per_image(p) / dt / 1e12, "unit": "TOPS", "images_per_s": arm.n / dt, "cores": arm.cores, "kind": "port",
               "sample": f"{steps_cpu} passes over {arm.n} images of the same workload after a 2 s warm-up, destination pre-allocated ({arm.kind})",
               "cold_cache_value": arm.n * ops_per_image(p) / dt_cold / 1e12,
               "cold_cache_protocol": "2 MB per thread scrubbed between iterations (reference test/test_utils.cc:23-49)"}
    tensor_peak = 2.0 * peaks["bf16_tflops"]
    consts = ncu_constants()
    probe = (consts.get("i8_mma_probe") or {}).get("tops")
    line = {
        "metric": "fused conv3x3+1x1 int8 TOPS", "value": tops, "unit": "TOPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "images_per_s": total_images / (ms_step * 1e-3),
        "config": shared_config(args.workload, n, world),
        "run": {"cache": f"rotating {n_sets} src/dst buffer sets ({n_sets * (src_bytes + dst_bytes) >> 20} MiB > 2x L2) so no step finds its data in L2",
                "launch": "plain loop of df_conv_run calls" if args.no_graph else f"one CUDA graph = {lead} untimed lead-in steps, the {args.steps} timed steps, one untimed trailing step; the two timing events are nodes of the same graph on a forked stream (first fires when the last lead-in step completes, second when the last timed step completes = exactly {args.steps} step periods of back-to-back launches); replayed once between the two synchronisations",
                "tiles": info.tiles_per_launch, "grid": info.grid, "smem_bytes": info.smem_bytes,
                "weights_resident": [info.w0_resident, info.w1_resident], "mma_row_efficiency": round(info.mma_efficiency, 4)},
        "roofline": {"bound": "tensor", "achieved": kernel_tops_this_rank, "peak": tensor_peak, "unit": "TOPS",
                     "frac": kernel_tops_this_rank / tensor_peak,
                     "traffic": (consts.get("conv_" + args.workload) or {}).get("dram_bytes_per_launch"),
                     "traffic_source": (consts.get("conv_" + args.workload) or {}).get("source"),
                     "peak_source": f"2 x bf16_tflops of MEASURED_PEAKS.json ({peaks['which']}); int8 dense rate = 2 x bf16",
                     "frac_of_i8_mma_probe": kernel_tops_this_rank / probe if probe else None,
                     "i8_mma_probe_tops": probe, "i8_mma_probe_source": (consts.get("i8_mma_probe") or {}).get("source"),
                     "kernel": "conv_pair_kernel" if info.w0_resident >= 2 else "conv_fused_kernel", "ops_per_launch": n * ops_per_image(p),
                     "sustained": sustained},
        "concat": {"workload": "concat+ReLU u8, 28x28, C=64/128/32/32, batch 32 (BASELINE configs[1])", "value": concat_gbs,
                   "unit": "GB/s", "us_per_launch": c_ms * 1e3, "bytes_per_launch": c_bytes,
                   "roofline": {"bound": "hbm", "achieved": concat_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": concat_gbs / peaks["hbm_gbs"],
                                "traffic": (consts.get("concat_cfg2") or {}).get("dram_bytes_per_launch"),
                                "traffic_source": (consts.get("concat_cfg2") or {}).get("source")},
                   "cache": f"rotating {c_sets} buffer sets ({c_sets * c_bytes >> 20} MiB > 2x L2)",
                   "large_footprint": big},
        "other_shapes": others,
        "concat_conv": concat_conv,
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
    ap.add_argument("--quick", action="store_true", help="headline legs only (skip the other shapes, the sustained run and the large concat)")
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
