"""World-size-2 / 3 gloo tests (CPU) of the host-side logic of the item-sharded predict_rank(k): shard bounds, the
all-to-all exchange layout (rank r receives the candidates of ITS user slice from every shard), the optional
all-gather of the merged slices and the collective re-run decision.  The merge itself is a CUDA kernel (trk_topk_merge, GPU-tested); here the gathered
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
    from tensorrec_b200 import kernels
    from tensorrec_b200.distributed import shard_bounds, exchange_rows, all_gather_rows, union_of_indices
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                           # same scores on every rank
        n_users, n_items, k = 11, 101, 5                         # 11 users over 2 / 3 ranks: uneven user slices too
        scores = rng.integers(-3, 4, size=(n_users, n_items)).astype(np.float32)     # ties across shards
        lo, hi = shard_bounds(n_items, world, rank)
        ids, vals = oracle.top_k_from_scores(scores[:, lo:hi], k)                    # this shard's candidates
        packed = torch.empty((n_users, 2 * k), dtype=torch.int32)
        packed[:, :k] = torch.from_numpy(vals.copy()).view(torch.int32)
        packed[:, k:] = torch.from_numpy((ids + lo).astype(np.int32))
        recv, (u0, u1) = exchange_rows(packed)
        assert (u0, u1) == shard_bounds(n_users, world, rank)
        assert tuple(recv.shape) == (world, u1 - u0, 2 * k) and recv.dtype == torch.int32
        # list r of this rank's users is what rank r found for them, bit for bit
        assert torch.equal(recv[rank], packed[u0:u1])
        # merge by (score desc, global id asc) == the oracle's top-k over the whole item axis, for this rank's users
        s = recv[:, :, :k].contiguous().view(torch.float32).numpy().transpose(1, 0, 2).reshape(u1 - u0, -1)
        i = recv[:, :, k:].numpy().transpose(1, 0, 2).reshape(u1 - u0, -1)
        merged_i = np.empty((u1 - u0, k), np.int32)
        merged_s = np.empty((u1 - u0, k), np.float32)
        for r in range(u1 - u0):
            order = np.lexsort((i[r], -s[r]))[:k]
            merged_i[r], merged_s[r] = i[r][order], s[r][order]
        exp_i, exp_s = oracle.top_k_from_scores(scores, k)
        assert np.array_equal(merged_i, exp_i[u0:u1]) and np.array_equal(merged_s, exp_s[u0:u1])
        # gather='all': every rank ends up with all users, in user order
        mine = torch.empty((u1 - u0, 2 * k), dtype=torch.int32)
        mine[:, :k] = torch.from_numpy(merged_s).view(torch.int32)
        mine[:, k:] = torch.from_numpy(merged_i)
        full = all_gather_rows(kernels.PackedTopK(u1 - u0, k, 'cpu', buf=mine), n_users)
        assert np.array_equal(full.items.numpy(), exp_i) and np.array_equal(full.scores.numpy(), exp_s)
        # every rank takes the same decision about the blocks to re-run
        assert union_of_indices([rank] if rank != 1 else [], 4, None, 'cpu') == [r for r in range(world) if r != 1]
        open(os.path.join(out_dir, 'ok_%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_exchange_world_size_2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(str(tmp_path))) == ['ok_0', 'ok_1']


def test_exchange_world_size_3_uneven_shards(tmp_path):
    """101 items over 3 ranks: shards of 34 / 34 / 33 items; 11 users: slices of 4 / 4 / 3; cross-shard ties resolve to
    the lower global id."""
    world = 3
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(str(tmp_path))) == ['ok_0', 'ok_1', 'ok_2']
