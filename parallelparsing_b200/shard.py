"""Multi-GPU partitioning of DecompressAll.

Index chunks are independent: chunk k needs only index[k], index[k+1] and the file
bytes [Input_k - 1, Input_{k+1}) (Decompressor/LazyFileReader.cs:53-69).  So one process
per GPU takes a CONTIGUOUS range of chunks, balanced by compressed bytes (Input deltas),
and there is no data-path collective; global record ordinals are an exclusive prefix sum
of per-rank record counts, done on the host over `world` integers.
"""
import numpy as np

__all__ = ["partition_chunks", "record_bases"]


def partition_chunks(inputs, world: int):
    """Split chunks 0..len(inputs)-2 into `world` contiguous ranges of near-equal compressed
    size.  `inputs` = Point.Input of every index point.  Returns [(first_chunk, n_chunks)] per
    rank; ranges are disjoint, ordered, cover every chunk, and may be empty when there are
    fewer chunks than ranks."""
    inputs = np.asarray(inputs, np.int64)
    n = max(inputs.size - 1, 0)
    if world <= 0:
        raise ValueError("world must be positive")
    if n == 0:
        return [(0, 0)] * world
    total = int(inputs[-1] - inputs[0])
    cuts = [0]
    for r in range(1, world):
        target = inputs[0] + total * r // world
        # first point whose Input reaches the target; never move backwards
        k = int(np.searchsorted(inputs, target, side="left"))
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(world)]


def record_bases(counts):
    """Exclusive prefix sum of per-rank record counts -> first global record ordinal per rank."""
    c = np.asarray(counts, np.int64)
    return np.concatenate([[0], np.cumsum(c)[:-1]]) if c.size else c
