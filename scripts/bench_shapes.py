"""bench.py's other_shapes section on its own: usage bench_shapes.py [workload ...] (needs a GPU)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import dfb200 as df
import bench
df.set_device(0)
peaks = bench.measured_peaks()
st = df.Stream()
for wl in (sys.argv[1:] or ["cfg1", "cfg1x64", "cfg4", "cfg4s32", "cfg4f32"]):
    r = bench.time_shape(df, st, wl, 50 if wl != "cfg1" else 200, peaks)
    print(f"{wl:8s} {r['us_per_launch']:8.2f} us  {r['tops']:7.1f} TOPS  {r['roofline']['bound']} frac {r['roofline']['frac']:.3f}", flush=True)
