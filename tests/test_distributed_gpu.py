"""Multi-GPU test of the item-sharded predict_rank(k): one process per GPU, one NCCL all-to-all of the
per-shard top-k, every rank merges its user slice (gather='slice'), optionally followed by an all-gather of the merged
slices (gather='all') == the oracle answer.  Skipped on boxes with < 2 GPUs (tests/test_api_gpu.py runs the same shards
one after the other on one GPU through the same entry points)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    import tensorrec_b200 as T
    from tensorrec_b200.distributed import predict_top_k_sharded, shard_bounds
    from tests import helpers as H
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        for integer, k in ((True, 10), (False, 7)):
            U, I, d = 300, 5001, 64
            uf = H.tag_features(U, 200, 20, seed=1, integer=integer)
            itf = H.tag_features(I, 200, 20, seed=2, integer=integer)
            wu, wi = H.linear_weights(200, d, seed=3, integer=integer), H.linear_weights(200, d, seed=4, integer=integer)
            bu, bi = H.feature_biases(200, seed=5, integer=integer), H.feature_biases(200, seed=6, integer=integer)
            model = T.TensorRec(n_components=d)
            model.set_weights({'linear_weights_user_0': wu, 'linear_weights_item': wi,
                               'feature_biases_user': bu[:, None], 'feature_biases_item': bi[:, None]})
            scores = oracle.OracleModel([wu], wi, bu, bi).predict(uf, itf)
            exp_i, exp_s = oracle.top_k_from_scores(scores, k)
            for gather, batch in (('all', None), ('slice', None), ('slice', 128)):
                top = predict_top_k_sharded(model, uf, itf, k, gather=gather, user_batch_size=batch)
                rows_of = model.last_topk_info['user_rows']
                if gather == 'all':
                    assert np.array_equal(rows_of, np.arange(U))
                elif batch is None:
                    lo, hi = shard_bounds(U, world, rank)
                    assert np.array_equal(rows_of, np.arange(lo, hi))
                assert top.items.shape == (len(rows_of), k)
                if integer:
                    assert np.array_equal(top.items, exp_i[rows_of]) and np.array_equal(top.scores, exp_s[rows_of])
                else:
                    assert np.all(np.abs(top.scores - scores[rows_of[:, None], top.items]) <= 1e-5 * 40 + 2e-6)
                    assert (top.items != exp_i[rows_of]).mean() < 0.01
        open(os.path.join(out_dir, 'ok_%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_item_sharded_top_k_matches_oracle(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip('needs at least 2 GPUs')
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(str(tmp_path))) == ['ok_%d' % r for r in range(world)]
