"""
CPU tests of the N > 1 path: world_size 2 over gloo.  Rows are sharded, each rank evaluates its
block with a local evaluator, one all-gather returns the full lnL vector on every rank.  (On the
GPU box the local evaluator is the device model and the backend is NCCL; the collective and the
sharding arithmetic are the same code.)
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from evidence_b200.multigpu import ShardedLikelihood, shard_bounds, split_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_eval(block):  # a stand-in likelihood: any row-wise function will do
    return -0.5 * (block ** 2).sum(dim=1) + block[:, 0]


def _worker(rank, world, port, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        theta = torch.from_numpy(np.random.default_rng(0).normal(size=(B, 5)))
        sh = ShardedLikelihood(_local_eval, ndim=5)
        full = sh(theta)
        blk = torch.from_numpy(np.random.default_rng(1).normal(size=(8, 5)))
        weak = sh.evaluate_local(blk + rank)  # weak-scaling call: own block, gathered
        ret[rank] = (full.numpy().copy(), weak.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [1, 7, 64, 101])
def test_sharded_likelihood_world2(B):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, ret), nprocs=world, join=True)
    theta = torch.from_numpy(np.random.default_rng(0).normal(size=(B, 5)))
    want = _local_eval(theta).numpy()
    for r in range(world):
        full, weak = ret[r]
        assert full.shape == (B,) and np.array_equal(full, want)
        assert weak.shape == (16,)
        blk = torch.from_numpy(np.random.default_rng(1).normal(size=(8, 5)))
        assert np.array_equal(weak[:8], _local_eval(blk).numpy())
        assert np.array_equal(weak[8:], _local_eval(blk + 1).numpy())


def test_shard_bounds_cover_every_row_once():
    for B in (0, 1, 5, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_bounds(B, world, r)
                assert 0 <= lo <= hi <= B and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(B))
    parts = split_rows(np.arange(10).reshape(5, 2), 2)
    assert [len(p) for p in parts] == [3, 2]
