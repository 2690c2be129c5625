"""N > 1 host logic on CPU: two gloo ranks each process their contiguous block of pairs (here with the CPU
oracle standing in for the per-GPU engine), results are gathered and must equal the single-process batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from stvo_pl_b200 import sharding


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(total, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _worker(rank, world, port, total, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    from stvo_pl_b200 import synth, types as T
    lo, hi = sharding.shard_range(total, world, rank)
    prev, curr, _, cam = synth.make_batch("kitti", hi - lo, first_pair=lo, n_pt=150, n_ls=40)
    res = Oracle().track_batch(cam, T.kitti_config(), prev, curr)["results"]
    allres = sharding.gather_results(res, total)
    tmax = sharding.max_over_ranks(float(rank + 1))
    dist.barrier()
    if rank == 0:
        np.save(out_path, allres.view(np.uint8))
        assert tmax == float(world)
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    total, world = 5, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, port, total, out), nprocs=world, join=True)
    from oracle.oracle import Oracle
    from stvo_pl_b200 import synth, types as T
    prev, curr, _, cam = synth.make_batch("kitti", total, n_pt=150, n_ls=40)
    ref = Oracle().track_batch(cam, T.kitti_config(), prev, curr)["results"]
    got = np.load(out).view(T.POSE_RESULT_DTYPE)
    assert got.tobytes() == ref.tobytes()
