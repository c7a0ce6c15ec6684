"""concat+ReLU -> conv four ways (bench.py's concat_conv section on its own) -- needs a GPU."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-fusion_b200"))
import dfb200 as df
import bench
df.set_device(0)
print(json.dumps(bench.time_concat_conv(df, df.Stream(), 50), indent=1))
