"""Time one fused-conv shape: back-to-back launches (CUDA events) -- development aid, needs a GPU.
usage: bench_one.py [cfg1|cfg3|cfg4] [batch] [iters] [dst]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from gpu_quick import bench_conv
import dfb200 as df

SHAPES = {"cfg1": (56, 56, 64, 64, 256), "cfg3": (28, 28, 128, 128, 512), "cfg4": (14, 14, 256, 256, 1024), "cat": (28, 28, 256, 128, 512)}
if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    dst = sys.argv[4] if len(sys.argv) > 4 else "u8"
    df.set_device(0)
    bench_conv(n, *SHAPES[which], dst=dst, iters=iters)
