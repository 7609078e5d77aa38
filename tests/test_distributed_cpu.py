"""World-size-2 gloo tests (CPU) of the host-side logic of the item-sharded predict_rank(k): shard bounds, the packed
all-gather layout and list order.  The merge itself is a CUDA kernel (trk_topk_merge, GPU-tested); here the gathered
lists are checked against the oracle's top-k of the concatenated shards with a numpy merge written in the test."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    from tensorrec_b200.distributed import shard_bounds
    for n_items in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            bounds = [shard_bounds(n_items, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n_items
            assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import oracle
    from tensorrec_b200.distributed import shard_bounds, all_gather_candidates
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                           # same scores on every rank
        n_users, n_items, k = 9, 101, 5
        scores = rng.integers(-3, 4, size=(n_users, n_items)).astype(np.float32)     # ties across shards
        lo, hi = shard_bounds(n_items, world, rank)
        ids, vals = oracle.top_k_from_scores(scores[:, lo:hi], k)                    # this shard's candidates
        top_s = torch.from_numpy(vals.copy())
        top_i = torch.from_numpy((ids + lo).astype(np.int32))
        all_s, all_i = all_gather_candidates(top_s, top_i)
        assert tuple(all_s.shape) == (n_users, world, k) and all_s.dtype == torch.float32
        assert tuple(all_i.shape) == (n_users, world, k) and all_i.dtype == torch.int32
        # list r of every user is rank r's candidate list, bit for bit
        assert torch.equal(all_s[:, rank], top_s) and torch.equal(all_i[:, rank], top_i)
        # merge by (score desc, global id asc) == the oracle's top-k over the whole item axis
        s = all_s.numpy().reshape(n_users, -1)
        i = all_i.numpy().reshape(n_users, -1)
        order = np.lexsort((i, -s), axis=1)[:, :k]
        merged_i = np.take_along_axis(i, order, axis=1)
        merged_s = np.take_along_axis(s, order, axis=1)
        exp_i, exp_s = oracle.top_k_from_scores(scores, k)
        assert np.array_equal(merged_i, exp_i) and np.array_equal(merged_s, exp_s)
        open(os.path.join(out_dir, 'ok_%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_all_gather_candidates_world_size_2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(str(tmp_path))) == ['ok_0', 'ok_1']


def test_all_gather_candidates_world_size_3_uneven_shards(tmp_path):
    """101 items over 3 ranks: shards of 34 / 34 / 33 items, cross-shard ties resolve to the lower global id."""
    world = 3
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(str(tmp_path))) == ['ok_0', 'ok_1', 'ok_2']
