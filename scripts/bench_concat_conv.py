"""concat+ReLU -> conv four ways (bench.py's concat_conv section on its own) -- needs a GPU."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import dfb200 as df
import bench
df.set_device(0)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else None  # default: BASELINE configs[1]'s 32
print(json.dumps(bench.time_concat_conv(df, df.Stream(), 50 if not batch or batch <= 64 else 10, batch=batch), indent=1))
