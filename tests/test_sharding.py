"""Multi-process test of the batch sharding used for N > 1 GPUs, on CPU:
world_size-2 gloo.  Each rank computes its slab with the oracle (as the checker standing in for
the device), results are gathered and must equal the unsharded result bit for bit -- i.e. the
partition has no data-path dependency and needs no collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dfb200 import shard


def test_slab_partition_is_exact():
    for n in (0, 1, 7, 64, 100, 2048):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                s, c = shard.slab(n, world, r)
                cover += list(range(s, s + c))
            assert cover == list(range(n))
    assert shard.slab(64, 8, 3) == (24, 8)
    assert shard.slab(1, 8, 0) == (0, 1) and shard.slab(1, 8, 5) == (1, 0)  # N < G: idle ranks
    assert shard.byte_range(64, 2, 1, 100352) == (32 * 100352, 32 * 100352)
    with pytest.raises(ValueError):
        shard.slab(4, 2, 2)


def _worker(rank, world, port, tmp):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests"))
    sys.path.insert(0, os.path.join(root, "deep-fusion_b200"))
    import cases
    import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = cases.ConvCase("shard", 5, 6, 6, 32, 32, 48, "u8", "s32", "s32")
    src, w0, w1, b0, b1, s0, s1 = c.tensors()
    wb, w1b = c.blocked(w0, w1)
    start, count = shard.slab(c.n, world, rank)
    out = np.zeros((c.n, 6, 6, 48), np.uint8)
    if count:
        d = O.make_desc(count, 6, 6, 32, 32, 48, O.U8, O.S32, O.S32, nscale0=32, nscale1=48)
        out[start:start + count] = O.conv(d, src[start:start + count], wb, b0, s0, w1b, b1, s1)
    t = torch.from_numpy(out.astype(np.int32))
    dist.all_reduce(t)  # slabs are disjoint, so a sum assembles the batch (test-only gather)
    elapsed = torch.tensor([float(rank + 1)])
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)  # the max-over-ranks timing rule of bench.py
    if rank == 0:
        d = O.make_desc(c.n, 6, 6, 32, 32, 48, O.U8, O.S32, O.S32, nscale0=32, nscale1=48)
        full = O.conv(d, src, wb, b0, s0, w1b, b1, s1)
        ok = np.array_equal(t.numpy().astype(np.uint8), full) and elapsed.item() == world
        open(tmp, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    flag = str(tmp_path / "result")
    mp.spawn(_worker, args=(2, port, flag), nprocs=2, join=True)
    assert open(flag).read() == "ok"
