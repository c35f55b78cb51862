"""Multi-rank host logic on CPU: world_size-2 gloo process group, streams sharded with no data-path collective,
per-frame result records gathered on rank 0 in stream order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lane_tracker_b200 import sharding
from lane_tracker_b200._lib import lt_result

RESULT_DTYPE = np.dtype(lt_result)


def test_stream_ranges_partition_everything():
    for total in (1, 7, 64, 511, 512):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                seen += list(sharding.stream_range(total, world, r))
            assert seen == list(range(total))
            sizes = [len(sharding.stream_range(total, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.owner_of(63, 512, 8) == 0 and sharding.owner_of(64, 512, 8) == 1
    with pytest.raises(ValueError):
        sharding.stream_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sharding.stream_range(total, world, rank)
    local = np.zeros(len(ids), dtype=RESULT_DTYPE)
    local["counter"] = 1
    local["n_left"] = np.array(list(ids)) * 10          # stands in for per-stream results
    local["left_fit"][:, 2] = np.array(list(ids)) + 0.5
    got = sharding.gather_results(local, total)
    raw = sharding.gather_records(torch.from_numpy(local.view(np.uint8).copy()), total)     # the tensor form bench.py uses
    if rank == 0:
        assert np.array_equal(raw.numpy().view(RESULT_DTYPE), got)
    else:
        assert raw is None
    # timing plumbing of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert float(t) == world
        np.save(out_path, got)
    else:
        assert got is None
    dist.destroy_process_group()


def test_world_size_2_gloo_gather(tmp_path):
    total, world = 7, 2
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    got = np.load(out)
    assert got.dtype == RESULT_DTYPE and len(got) == total
    assert list(got["n_left"]) == [10 * i for i in range(total)]
    assert list(got["left_fit"][:, 2]) == [i + 0.5 for i in range(total)]
