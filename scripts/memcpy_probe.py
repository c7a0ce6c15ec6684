"""What the host can deliver: pinned-memory H2D / D2H bandwidth with 1, 2, 4, 8 GPUs copying AT THE SAME TIME, from
one process (one stream per device).  The end-to-end submit() path is bound by these copies (32 MB per cfg3 batch,
the kernel is ~3 % of it), so this is the ceiling of the e2e scaling curve.  Needs GPUs.
usage: memcpy_probe.py [out.json]"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import numpy as np
import dfb200 as df

L = df.lib()
n_dev = df.device_count()
MB = 64
res = {"buffer_mb": MB, "devices_present": n_dev, "rows": []}
for g in [x for x in (1, 2, 4, 8) if x <= n_dev]:
    host, dev, streams = [], [], []
    for d in range(g):
        df.set_device(d)
        a = np.zeros(MB << 20, np.uint8)
        df.check(L.df_host_register(a.ctypes.data, a.nbytes))
        host.append(a)
        dev.append(df.DeviceBuffer(a.nbytes))
        streams.append(df.Stream())
    def run(kind, reps=8):
        for d in range(g):
            df.set_device(d); df.check(L.df_stream_sync(streams[d].ptr))
        t = time.perf_counter()
        for _ in range(reps):
            for d in range(g):
                df.set_device(d)
                if kind in ("h2d", "both"):
                    df.check(L.df_h2d(dev[d].ptr, host[d].ctypes.data, host[d].nbytes, streams[d].ptr))
                if kind in ("d2h", "both"):
                    df.check(L.df_d2h(host[d].ctypes.data, dev[d].ptr, host[d].nbytes, streams[d].ptr))
        for d in range(g):
            df.set_device(d); df.check(L.df_stream_sync(streams[d].ptr))
        dt = time.perf_counter() - t
        moved = reps * g * (MB << 20) * (2 if kind == "both" else 1)
        return moved / dt / 1e9
    run("h2d", 2)
    row = {"gpus": g, "h2d_gbs_total": run("h2d"), "d2h_gbs_total": run("d2h"), "serial_both_gbs_total": run("both")}
    row["h2d_gbs_per_gpu"] = row["h2d_gbs_total"] / g
    row["d2h_gbs_per_gpu"] = row["d2h_gbs_total"] / g
    res["rows"].append(row)
    print(row, flush=True)
    for d in range(g):
        L.df_host_unregister(host[d].ctypes.data)
    del dev, streams, host
    df.set_device(0)
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "memcpy_probe.json")
json.dump(res, open(out, "w"), indent=1)
