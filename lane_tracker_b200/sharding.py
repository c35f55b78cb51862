"""Stream -> GPU sharding: streams are independent, so they are split into contiguous
blocks, one block per rank; nothing is exchanged on the data path.  Only the small per-frame result records are
gathered to the host of rank 0."""
from __future__ import annotations

import numpy as np


def stream_range(total_streams, world_size, rank):
    """Contiguous block of stream ids owned by `rank` (block sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(total_streams), int(world_size))
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owner_of(stream_id, total_streams, world_size):
    for r in range(world_size):
        if stream_id in stream_range(total_streams, world_size, r):
            return r
    raise ValueError("stream %d outside [0, %d)" % (stream_id, total_streams))


def gather_results(local_results, total_streams, group=None):
    """Gather each rank's structured result array (one record per owned stream) on rank 0, in stream order.

    Uses torch.distributed object collectives on the host (gloo or nccl process group); returns the
    concatenated array on rank 0 and None elsewhere."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    payload = (rank, np.asarray(local_results).tobytes(), str(np.asarray(local_results).dtype.descr))
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(payload, bucket, dst=0, group=group)
    if rank != 0:
        return None
    dtype = np.asarray(local_results).dtype
    parts = []
    for r, (src_rank, raw, _) in enumerate(sorted(bucket, key=lambda p: p[0])):
        arr = np.frombuffer(raw, dtype=dtype)
        if len(arr) != len(stream_range(total_streams, world, src_rank)):
            raise RuntimeError("rank %d returned %d records" % (src_rank, len(arr)))
        parts.append(arr)
    return np.concatenate(parts)


def gather_records(local_records, total_streams, group=None, out=None):
    """Gather the ranks' result records on rank 0 as raw bytes, in stream order, with ONE tensor collective on the host
    process group (gloo): no pickling, no device collective.  ``local_records``: CPU uint8 tensor (pinned is fine) holding
    this rank's records back to back; every record has the same size.  Returns a uint8 tensor of all ``total_streams``
    records on rank 0 (``out`` is reused when given) and None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = len(stream_range(total_streams, world, rank))
    if mine == 0 or local_records.numel() % mine:
        raise ValueError("rank %d: %d bytes do not hold %d records" % (rank, local_records.numel(), mine))
    rec = local_records.numel() // mine
    largest = len(stream_range(total_streams, world, 0))             # block sizes differ by at most one; rank 0 has the largest
    send = local_records
    if mine != largest:
        send = torch.zeros(largest * rec, dtype=torch.uint8)
        send[:mine * rec] = local_records
    bucket = [torch.empty(largest * rec, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(send, bucket, dst=0, group=group)
    if rank != 0:
        return None
    if out is None:
        out = torch.empty(total_streams * rec, dtype=torch.uint8)
    pos = 0
    for r in range(world):
        k = len(stream_range(total_streams, world, r)) * rec
        out[pos:pos + k] = bucket[r][:k]
        pos += k
    return out
