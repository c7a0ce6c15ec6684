"""Batch sharding of the deep-fusion ops across the GPUs of one box.

Every image is independent in both ops (the reference itself splits work over n*oh rows /
n*h*w pixels, src/op_conv.cc:155, src/op_concat.cc:28), so rank r of `world` simply owns a
contiguous slab of the batch.  NHWC makes that slab one contiguous byte range of src and dst.
Weights, biases and scales are replicated.  There is no data-path collective.
"""
from __future__ import annotations


def slab(n_total: int, world: int, rank: int):
    """(first image, image count) of `rank`: contiguous slabs of ceil(N/G) images."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-n_total // world)
    start = min(rank * per, n_total)
    return start, max(0, min(per, n_total - start))


def byte_range(n_total: int, world: int, rank: int, bytes_per_image: int):
    s, c = slab(n_total, world, rank)
    return s * bytes_per_image, c * bytes_per_image
