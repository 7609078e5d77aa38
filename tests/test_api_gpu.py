"""GPU tests of the drop-in boundary: the TensorRec class driving the kernels through the C ABI, checked against the
oracle with injected weights (the reference never seeds its initialiser, so values after fit() are unpinned) and
against the contract tests of the reference's test/test_tensorrec.py, test_readme.py (cited per test)."""
import logging
import os
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope='module')
def T():
    import torch
    import tensorrec_b200
    from tensorrec_b200 import kernels
    kernels.require_cuda()
    torch.cuda.set_device(0)
    return tensorrec_b200


def build(T, uf, itf, d, n_tastes=1, user_repr='linear', item_repr='linear', prediction='dot', biased=True,
          attention=False, integer=False, seed=0):
    """A TensorRec with injected weights and the OracleModel holding the same weights."""
    R, P = T.representation_graphs, T.prediction_graphs
    reprs = {'linear': R.LinearRepresentationGraph, 'normalized_linear': R.NormalizedLinearRepresentationGraph}
    preds = {'dot': P.DotProductPredictionGraph, 'cosine': P.CosineSimilarityPredictionGraph,
             'euclidean': P.EuclideanSimilarityPredictionGraph}
    model = T.TensorRec(n_components=d, n_tastes=n_tastes, user_repr_graph=reprs[user_repr](),
                        item_repr_graph=reprs[item_repr](), prediction_graph=preds[prediction](), biased=biased,
                        attention_graph=R.LinearRepresentationGraph() if attention else None)
    fu, fi = uf.shape[1], itf.shape[1]
    wu = [H.linear_weights(fu, d, seed=seed + 10 + t, integer=integer) for t in range(n_tastes)]
    wi = H.linear_weights(fi, d, seed=seed + 3, integer=integer)
    wa = [H.linear_weights(fu, d, seed=seed + 20 + t, integer=integer) for t in range(n_tastes)] if attention else None
    bu = H.feature_biases(fu, seed=seed + 4, integer=integer) if biased else None
    bi = H.feature_biases(fi, seed=seed + 5, integer=integer) if biased else None
    weights = {'linear_weights_item': wi}
    for t in range(n_tastes):
        weights['linear_weights_user_%d' % t] = wu[t]
        if attention:
            weights['linear_weights_attn_%d' % t] = wa[t]
    if biased:
        weights['feature_biases_user'] = bu.reshape(-1, 1)
        weights['feature_biases_item'] = bi.reshape(-1, 1)
    model.set_weights(weights)
    om = oracle.OracleModel(wu, wi, bu, bi, attention_weights=wa, user_repr=user_repr, item_repr=item_repr,
                            prediction=prediction)
    return model, om


def score_tolerance(om, uf, itf, rel=1e-5):
    ur = om.user_representation(uf)
    ir = om.item_representation(itf)
    if om.prediction == 'cosine':
        ur = np.stack([oracle.l2_normalize(u) for u in ur])
        ir = oracle.l2_normalize(ir)
    return H.norm_tolerance(ur, ir, rel).max(axis=0) + 2e-6


# ---- BASELINE config #1: the README flow (test/test_readme.py:6-31) --------------------------------------------
def test_readme_flow_fit_predict_rank_recall(T, caplog):
    model = T.TensorRec()
    interactions, user_features, item_features = T.util.generate_dummy_data(num_users=100, num_items=150,
                                                                            interaction_density=.05, seed=0)
    with caplog.at_level(logging.INFO):
        model.fit(interactions, user_features, item_features, epochs=5, verbose=True)
    assert any('EPOCH 4 BATCH 0 loss' in r.getMessage() for r in caplog.records)
    predictions = model.predict(user_features=user_features, item_features=item_features)
    predicted_ranks = model.predict_rank(user_features=user_features, item_features=item_features)
    assert predictions.shape == (100, 150) and predictions.dtype == np.float32
    assert predicted_ranks.shape == (100, 150) and predicted_ranks.dtype == np.int32
    assert (predicted_ranks > 0).all()                                           # test/test_tensorrec.py:204-212
    assert np.array_equal(np.sort(predicted_ranks, axis=1), np.tile(np.arange(1, 151), (100, 1)))
    # ranks are exactly the reference ranking of the returned scores
    assert np.array_equal(predicted_ranks, oracle.rank_predictions(predictions))
    r_at_k = T.eval.recall_at_k(predicted_ranks, interactions, k=10)
    assert 0.0 <= np.mean(r_at_k) <= 1.0
    # the whole fitted model agrees with the oracle evaluated on the fitted weights
    w = model.get_weights()
    om = oracle.OracleModel([w['linear_weights_user_0']], w['linear_weights_item'], w['feature_biases_user'][:, 0],
                            w['feature_biases_item'][:, 0])
    assert np.all(np.abs(predictions - om.predict(user_features, item_features))
                  <= score_tolerance(om, user_features, item_features))


# ---- end-to-end parity with injected weights ---------------------------------------------------------------------
@pytest.mark.parametrize('path', ['auto', 'exact'])
@pytest.mark.parametrize('prediction,user_repr', [('dot', 'linear'), ('cosine', 'linear'), ('dot', 'normalized_linear'),
                                                  ('cosine', 'normalized_linear')])
def test_predict_matches_oracle_float(T, monkeypatch, path, prediction, user_repr):
    monkeypatch.setattr(T.tensorrec, 'SCORE_PATH', path)
    uf, itf = H.tag_features(100, 200, 20, seed=1), H.tag_features(150, 200, 20, seed=2)
    model, om = build(T, uf, itf, d=100, user_repr=user_repr, prediction=prediction)
    got = model.predict(uf, itf)
    expect = om.predict(uf, itf)
    assert got.shape == (100, 150) and got.dtype == np.float32
    assert np.all(np.abs(got - expect) <= score_tolerance(om, uf, itf))
    # ranks: exact w.r.t. the kernel's own scores; vs the oracle they may differ only inside the score tolerance
    ranks = model.predict_rank(uf, itf)
    assert np.array_equal(ranks, oracle.rank_predictions(got))
    expect_ranks = om.predict_rank(uf, itf)
    differing = ranks != expect_ranks
    assert differing.mean() < 0.01


@pytest.mark.parametrize('path', ['auto', 'exact', 'auto-exact3'])
@pytest.mark.parametrize('U,I,d', [(100, 150, 100), (256, 4096, 64), (77, 5000, 128), (50, 300, 10)])
def test_predict_and_rank_exact_on_integer_fixture(T, monkeypatch, path, U, I, d):
    """SURVEY 8d parity fixture: features in {0,1}, weights in {-2..2}, integer biases -> scores are exact in every
    arithmetic path, ties are everywhere, and the full int32 rank matrix must equal the reference's double sort."""
    if path == 'auto-exact3':      # tensor cores, but the 3-pass top-k kernel instead of filter + re-scoring
        path = 'auto'
        monkeypatch.setattr(T.tensorrec, 'TOPK_PATH', 'exact')
    monkeypatch.setattr(T.tensorrec, 'SCORE_PATH', path)
    uf = H.tag_features(U, 200, 20, seed=U, integer=True)
    itf = H.tag_features(I, 200, 20, seed=I, integer=True)
    model, om = build(T, uf, itf, d=d, integer=True)
    scores = om.predict(uf, itf)
    assert np.array_equal(model.predict(uf, itf), scores)
    assert np.array_equal(model.predict_rank(uf, itf), oracle.rank_predictions(scores))
    top = model.predict_rank(uf, itf, k=10)
    exp_i, exp_s = oracle.top_k_from_scores(scores, 10)
    assert np.array_equal(top.items, exp_i) and np.array_equal(top.scores, exp_s)


def test_tastes_normalized_cosine_and_attention(T):
    # test/test_tensorrec.py:278-339 shapes; values against the oracle
    uf, itf = H.tag_features(15, 200, 20, seed=1), H.tag_features(30, 200, 20, seed=2)
    model, om = build(T, uf, itf, d=10, n_tastes=3, user_repr='normalized_linear', prediction='cosine')
    got = model.predict(uf, itf)
    assert np.all(np.abs(got - om.predict(uf, itf)) <= 3e-6)
    assert model.predict_user_representation(uf).shape == (3, 15, 10)
    assert np.array_equal(model.predict_rank(uf, itf), oracle.rank_predictions(got))
    top = model.predict_rank(uf, itf, k=5)                       # n_tastes > 1: one fused sweep per taste + merge
    assert model.last_topk_info['path'] == 'filter'
    exp_i, exp_s = oracle.top_k_from_scores(got, 5)
    assert np.array_equal(top.items, exp_i) and np.all(np.abs(top.scores - exp_s) <= 3e-6)

    model, om = build(T, uf, itf, d=10, n_tastes=3, attention=True)
    got = model.predict(uf, itf)
    expect = om.predict(uf, itf)
    assert np.allclose(got, expect, rtol=2e-5, atol=2e-5)
    assert model.predict_user_attention_representation(uf).shape == (3, 15, 10)

    model, om = build(T, uf, itf, d=10, prediction='euclidean', n_tastes=2)
    assert np.allclose(model.predict(uf, itf), om.predict(uf, itf), rtol=1e-4, atol=1e-4)


def test_representations_and_biases(T):
    # test/test_tensorrec.py:226-275
    uf, itf = H.tag_features(15, 200, 20, seed=1), H.tag_features(30, 200, 20, seed=2)
    model, om = build(T, uf, itf, d=10)
    ur, ir = model.predict_user_representation(uf), model.predict_item_representation(itf)
    assert ur.shape == (15, 10) and ir.shape == (30, 10)
    assert np.allclose(ur, om.user_representation(uf)[0], atol=1e-6)
    assert np.allclose(ir, om.item_representation(itf), atol=1e-6)
    ub, ib = model.predict_user_bias(uf), model.predict_item_bias(itf)
    assert ub.shape == (15,) and ib.shape == (30,) and np.any(ub != 0) and np.any(ib != 0)
    assert np.allclose(ub, oracle.project_biases(oracle.coo_from_sparse(uf), om.user_bias), atol=1e-6)
    unbiased, _ = build(T, uf, itf, d=10, biased=False)
    with pytest.raises(T.errors.ModelNotBiasedException):
        unbiased.predict_user_bias(uf)
    assert unbiased.predict(uf, itf).shape == (15, 30)


def test_predict_similar_items(T):
    itf = H.tag_features(30, 200, 20, seed=2)
    uf = H.tag_features(15, 200, 20, seed=1)
    model, om = build(T, uf, itf, d=10, prediction='cosine')
    sims = model.predict_similar_items(itf, item_ids=[6, 12], n_similar=5)
    assert len(sims) == 2 and all(len(s) == 5 for s in sims)
    expect = oracle.predict_similar_items('cosine', om.item_representation(itf), [6, 12])
    for row, ids in zip(sims, [6, 12]):
        assert row[0][0] == ids and abs(row[0][1] - 1.0) < 1e-5          # an item is most similar to itself
        for item_id, score in row:
            assert abs(score - expect[[6, 12].index(ids), item_id]) < 1e-5


# ---- determinism and persistence (test/test_tensorrec.py:418-458) -----------------------------------------------
def test_repeat_and_reload_are_bit_identical(T):
    interactions, uf, itf = T.util.generate_dummy_data(num_users=15, num_items=30, interaction_density=.5, seed=3)
    model = T.TensorRec(n_components=10)
    model.fit(interactions, uf, itf, epochs=5)
    predictions, ranks = model.predict(uf, itf), model.predict_rank(uf, itf)
    top = model.predict_rank(uf, itf, k=7)
    with tempfile.TemporaryDirectory() as tmp:
        model.save_model(os.path.join(tmp, 'm'))
        assert np.array_equal(predictions, model.predict(uf, itf))
        assert np.array_equal(ranks, model.predict_rank(uf, itf))
        T.session_management.set_session(None)
        new_model = T.TensorRec.load_model(os.path.join(tmp, 'm'))
    assert np.array_equal(predictions, new_model.predict(uf, itf))
    assert np.array_equal(ranks, new_model.predict_rank(uf, itf))
    again = new_model.predict_rank(uf, itf, k=7)
    assert np.array_equal(top.items, again.items) and np.array_equal(top.scores, again.scores)
    # results come back in page-locked buffers that are recycled once collected: live results must never alias
    assert not np.shares_memory(top.items, again.items) and not np.shares_memory(predictions, new_model.predict(uf, itf))
    snapshot = top.items.copy()
    for _ in range(3):
        new_model.predict_rank(uf, itf, k=7)
    assert np.array_equal(top.items, snapshot)


# ---- the plugin surface: a user-defined representation graph (test/test_readme.py:33-68) --------------------------
def test_custom_representation_graph_runs_through_the_kernels(T):
    import torch

    class TanhRepresentationGraph(T.representation_graphs.AbstractRepresentationGraph):
        def connect_representation_graph(self, tf_features, n_components, n_features, node_name_ending):
            tf_tanh_weights = T.session_management.get_variable(
                'tanh_weights_%s' % node_name_ending,
                lambda: torch.randn(n_features, n_components, device=tf_features.device) * .5)
            tf_repr = torch.tanh(torch.sparse.mm(tf_features, tf_tanh_weights))
            return tf_repr, [tf_tanh_weights]

    interactions, uf, itf = T.util.generate_dummy_data(num_users=100, num_items=150, interaction_density=.05, seed=1)
    model = T.TensorRec(n_components=10, user_repr_graph=TanhRepresentationGraph(),
                        item_repr_graph=T.representation_graphs.NormalizedLinearRepresentationGraph())
    model.fit(interactions, uf, itf, epochs=5)
    w = model.get_weights()
    ur = np.tanh(oracle.sparse_dense_matmul(oracle.coo_from_sparse(uf), w['tanh_weights_user_0']))
    ir = oracle.normalized_linear_representation(oracle.coo_from_sparse(itf), w['linear_weights_item'])
    expect = oracle.bias_prediction_dense(
        oracle.dot_product_dense(ur, ir), oracle.project_biases(oracle.coo_from_sparse(uf), w['feature_biases_user']),
        oracle.project_biases(oracle.coo_from_sparse(itf), w['feature_biases_item']))
    got = model.predict(uf, itf)
    assert np.all(np.abs(got - expect) <= H.norm_tolerance(ur, ir, 2e-5) + 1e-5)
    assert np.array_equal(model.predict_rank(uf, itf), oracle.rank_predictions(got))
    top = model.predict_rank(uf, itf, k=10)
    assert np.array_equal(top.items, oracle.top_k_from_scores(got, 10)[0])


# ---- BASELINE config #3: MovieLens-1M-shaped, cosine, recall@10 ----------------------------------------------------
def test_movielens_shaped_cosine_full_rank_and_topk_agree(T):
    n_users, n_items = 6040, 3706
    rng = np.random.default_rng(0)
    uf = sp.identity(n_users, format='csr', dtype=F32)
    genres = sp.random(n_items, 18, density=1.65 / 18, format='csr', random_state=rng, dtype=np.float64)
    genres.data[:] = 1.0
    itf = sp.hstack([sp.identity(n_items, format='csr', dtype=F32), genres.astype(F32)]).tocsr()
    model, om = build(T, uf, itf, d=128, prediction='cosine', biased=True)
    n_int = 200000
    interactions = sp.csr_matrix((np.ones(n_int, dtype=F32), (rng.integers(0, n_users, n_int),
                                                              rng.integers(0, n_items, n_int))), shape=(n_users, n_items))
    ranks = model.predict_rank(uf, itf)
    assert ranks.shape == (n_users, n_items) and ranks.dtype == np.int32
    scores = model.predict(uf, itf)
    sample = rng.choice(n_users, 64, replace=False)
    assert np.array_equal(ranks[sample], oracle.rank_predictions(scores[sample]))
    expect = om.predict(uf[sample], itf)
    assert np.all(np.abs(scores[sample] - expect) <= score_tolerance(om, uf[sample], itf))
    top = model.predict_rank(uf, itf, k=10)
    full_recall = T.eval.recall_at_k(ranks, interactions, k=10)
    topk_recall = T.eval.recall_at_k(top, interactions, k=10)
    assert np.allclose(full_recall, topk_recall, atol=0.05) and abs(full_recall.mean() - topk_recall.mean()) < 1e-4
    # the top-k route (filter + exact fp32 re-scoring) names the rank <= 10 entries of the full-rank route (3-pass
    # split-product scores); the two arithmetics may order a pair differently only when it is tied to ~1e-7
    rows = np.arange(n_users)[:, None]
    agree = ranks[rows, top.items] == np.tile(np.arange(1, 11), (n_users, 1))
    assert agree.mean() > 0.9995
    assert np.all(np.abs(top.scores - scores[rows, top.items]) <= 4e-6)


def test_recommendation_graph_functions_evaluate_on_the_device(T):
    """The functional surface of tensorrec/recommendation_graphs.py with the reference's own known answers
    (test/test_recommendation_graphs.py), numpy in -> kernels -> device tensor out."""
    import json
    G = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_known_answers.json')))
    RG, PG = T.recommendation_graphs, T.prediction_graphs
    g = G['rank_predictions']
    assert np.array_equal(RG.rank_predictions(np.array(g['predictions'], dtype=F32)).cpu().numpy(), g['expected'])
    g = G['project_biases']
    got = RG.project_biases_with(sp.coo_matrix(np.array(g['features'], dtype=F32)), g['feature_biases'])
    assert np.array_equal(got.cpu().numpy(), np.array(g['expected'], dtype=F32))
    g = G['bias_prediction_dense']
    got = RG.bias_prediction_dense(np.array(g['predictions'], dtype=F32), np.array(g['user_biases'], dtype=F32),
                                   np.array(g['item_biases'], dtype=F32))
    assert np.array_equal(got.cpu().numpy(), np.array(g['expected'], dtype=F32))
    g = G['collapse_mixture_of_tastes']
    got = RG.collapse_mixture_of_tastes([np.array(p, dtype=F32) for p in g['predictions']], None)
    assert np.array_equal(got.cpu().numpy(), np.array(g['expected'], dtype=F32))
    g = G['collapse_mixture_of_tastes_with_attention']
    got = RG.collapse_mixture_of_tastes([np.array(p, dtype=F32) for p in g['predictions']],
                                        [np.array(a, dtype=F32) for a in g['attentions']]).cpu().numpy()
    assert np.all(np.abs(got - np.array(g['expected'], dtype=F32)) <= 4 * np.spacing(np.array(g['expected'], dtype=F32)))
    g = G['dot_product_dense']
    got = PG.DotProductPredictionGraph().connect_dense_prediction_graph(
        tf_user_representation=np.array(g['user_repr']), tf_item_representation=np.array(g['item_repr']))
    assert np.allclose(got.cpu().numpy(), g['expected'])
    g = G['cosine_dense']
    got = PG.CosineSimilarityPredictionGraph().connect_dense_prediction_graph(
        tf_user_representation=np.array(g['user_repr']), tf_item_representation=np.array(g['item_repr']))
    assert np.allclose(got.cpu().numpy(), g['expected'], atol=1e-6)
    g = G['predict_similar_items_cosine']
    got = RG.predict_similar_items(PG.CosineSimilarityPredictionGraph(), np.array(g['item_repr'], dtype=F32),
                                   g['item_ids'])
    assert np.array_equal(got.cpu().numpy(), np.array(g['expected'], dtype=F32))


# ---- the training step runs the representation through K1 (forward) and K1 on A^T (backward) ---------
def test_training_matmul_runs_on_k1_and_matches_torch_sparse(T):
    import torch
    from tensorrec_b200.input_utils import SparseInput
    from tensorrec_b200.sparse_ops import sparse_dense_matmul
    rng = np.random.default_rng(5)
    for rows, feats, d in ((300, 120, 16), (257, 90, 7), (40, 33, 1)):
        nnz = rows * 6
        r, c = rng.integers(0, rows, nnz), rng.integers(0, feats, nnz)          # duplicates on purpose
        a = sp.coo_matrix((rng.standard_normal(nnz).astype(np.float32), (r, c)), shape=(rows, feats))
        src = SparseInput(a)
        tf_a = src.torch_sparse(torch.device('cuda'))
        w1 = torch.randn(feats, d, device='cuda', requires_grad=True)
        w2 = w1.detach().clone().requires_grad_(True)
        g = torch.randn(rows, d, device='cuda')
        out1 = sparse_dense_matmul(tf_a, w1)
        out2 = torch.sparse.mm(tf_a, w2)
        assert type(out1.grad_fn).__name__.startswith('_CsrMatmul')
        (out1 * g).sum().backward()
        (out2 * g).sum().backward()
        scale = float(out2.detach().abs().max()) + 1e-6
        assert float((out1 - out2).detach().abs().max()) <= 1e-5 * scale
        assert float((w1.grad - w2.grad).abs().max()) <= 1e-5 * (float(w2.grad.abs().max()) + 1e-6)
        again = sparse_dense_matmul(tf_a, w1.detach().clone().requires_grad_(True))
        assert torch.equal(again, out1)                                           # deterministic forward
    # and fit() of the built-in linear graphs goes through it (loss decreases, results finite)
    interactions, uf, itf = T.util.generate_dummy_data(num_users=60, num_items=80, interaction_density=.2, seed=2)
    model = T.TensorRec(n_components=8)
    model.fit(interactions, uf, itf, epochs=30, learning_rate=.05)
    assert np.all(np.isfinite(model.predict(uf, itf)))


@pytest.mark.parametrize('integer', [True, False])
def test_item_shards_run_one_after_the_other_through_predict_top_k(T, integer):
    """The multi-GPU layout on ONE GPU (the driver's test box has one): every item shard goes through
    predict_top_k(item_id_offset=...) -- filter path, device-side fallback -- and the per-shard results are merged in the
    layout the all-to-all delivers ([shards, U_slice, 2k], trk_topk_merge with list / user strides).  Equals the oracle
    and the unsharded call."""
    import torch
    from tensorrec_b200 import kernels
    from tensorrec_b200.distributed import shard_bounds
    U, I, d, k, world = 517, 9001, 128, 10, 4
    uf = H.tag_features(U, 200, 20, seed=1, integer=integer) if integer else H.indicator_features(U, seed=1)
    itf = H.tag_features(I, 200, 20, seed=2, integer=integer) if integer else H.indicator_features(I, seed=2)
    wu = H.linear_weights(uf.shape[1], d, seed=3, integer=integer)
    wi = H.linear_weights(itf.shape[1], d, seed=4, integer=integer)
    bu = H.feature_biases(uf.shape[1], seed=5, integer=integer)
    bi = H.feature_biases(itf.shape[1], seed=6, integer=integer)
    model = T.TensorRec(n_components=d)
    model.set_weights({'linear_weights_user_0': wu, 'linear_weights_item': wi, 'feature_biases_user': bu[:, None],
                       'feature_biases_item': bi[:, None]})
    whole = model.predict_top_k(uf, itf, k)
    assert model.last_topk_info['path'] == 'filter'
    per_shard = []
    for r in range(world):
        lo, hi = shard_bounds(I, world, r)
        top = model.predict_top_k(uf, sp.csr_matrix(itf)[lo:hi], k, item_id_offset=lo, to_host=False)
        per_shard.append(torch.cat([top.scores.view(torch.int32), top.items], dim=1))      # PackedTopK rows [U, 2k]
    items_out, scores_out = [], []
    for r in range(world):                       # rank r's user slice, merged from every shard's rows of that slice
        u0, u1 = shard_bounds(U, world, r)
        recv = torch.stack([p[u0:u1] for p in per_shard]).contiguous()                     # [shards, U_slice, 2k]
        merged = kernels.topk_merge_received(recv, u1 - u0, world, k)
        items_out.append(merged.items.cpu().numpy())
        scores_out.append(merged.scores.cpu().numpy())
    got_i, got_s = np.concatenate(items_out), np.concatenate(scores_out)
    scores = oracle.OracleModel([wu], wi, bu, bi).predict(uf, itf)
    exp_i, exp_s = oracle.top_k_from_scores(scores, k)
    if integer:
        assert np.array_equal(got_i, exp_i) and np.array_equal(got_s, exp_s)
        assert np.array_equal(whole.items, exp_i) and np.array_equal(whole.scores, exp_s)
    else:
        assert np.array_equal(got_i, whole.items) and np.array_equal(got_s, whole.scores)   # sharding changes nothing
        assert (got_i != exp_i).mean() < 0.01


def test_predict_top_k_user_batches_equal_one_batch(T):
    U, I, d, k = 1000, 3000, 64, 10
    uf, itf = H.indicator_features(U, seed=1), H.indicator_features(I, seed=2)
    wu, wi = H.linear_weights(uf.shape[1], d, seed=3), H.linear_weights(itf.shape[1], d, seed=4)
    bu, bi = H.feature_biases(uf.shape[1], seed=5), H.feature_biases(itf.shape[1], seed=6)
    model = T.TensorRec(n_components=d)
    model.set_weights({'linear_weights_user_0': wu, 'linear_weights_item': wi, 'feature_biases_user': bu[:, None],
                       'feature_biases_item': bi[:, None]})
    one = model.predict_top_k(uf, itf, k)
    many = model.predict_top_k(uf, itf, k, user_batch_size=300)
    assert np.array_equal(one.items, many.items) and np.array_equal(one.scores, many.scores)
    assert np.array_equal(model.last_topk_info['user_rows'], np.arange(U))


def test_wide_representations_are_tiled_over_k1(T):
    """n_components beyond one K1 launch (> 512 columns): the component axis is tiled, values equal the oracle's."""
    U, I, d = 40, 60, 640
    uf, itf = H.tag_features(U, 50, 5, seed=1, integer=True), H.tag_features(I, 50, 5, seed=2, integer=True)
    wu, wi = H.linear_weights(50, d, seed=3, integer=True), H.linear_weights(50, d, seed=4, integer=True)
    model = T.TensorRec(n_components=d, biased=False)
    model.set_weights({'linear_weights_user_0': wu, 'linear_weights_item': wi})
    om = oracle.OracleModel([wu], wi)
    assert np.array_equal(model.predict_user_representation(uf), om.user_representation(uf)[0])
    assert np.array_equal(model.predict(uf, itf), om.predict(uf, itf))


@pytest.mark.parametrize('kind', ['tensor', 'tastes', 'euclidean'])
def test_streamed_predict_equals_one_shot(T, kind, tmp_path):
    """predict() in user blocks (the form BASELINE config #2 needs: the 400 GB result never exists at once) is bit-identical
    to the one-shot call, into a new array, a caller-provided array and a numpy.memmap; predict_batches covers every row
    exactly once, in order."""
    U, I, d = 700, 900, 64
    uf, itf = H.tag_features(U, 200, 20, seed=1), H.tag_features(I, 200, 20, seed=2)
    weights = {'linear_weights_item': H.linear_weights(200, d, seed=4),
               'feature_biases_user': H.feature_biases(200, seed=5)[:, None],
               'feature_biases_item': H.feature_biases(200, seed=6)[:, None]}
    if kind == 'tastes':
        model = T.TensorRec(n_components=d, n_tastes=3)
        for t in range(3):
            weights['linear_weights_user_%d' % t] = H.linear_weights(200, d, seed=10 + t)
    else:
        pg = T.prediction_graphs.EuclideanSimilarityPredictionGraph() if kind == 'euclidean' else \
            T.prediction_graphs.DotProductPredictionGraph()
        model = T.TensorRec(n_components=d, prediction_graph=pg)
        weights['linear_weights_user_0'] = H.linear_weights(200, d, seed=3)
    model.set_weights(weights)
    whole = model.predict(uf, itf)
    assert whole.shape == (U, I)
    blocks = model.predict(uf, itf, user_batch_size=128)
    assert np.array_equal(whole, blocks)
    out = np.full((U, I), np.nan, dtype=np.float32)
    assert model.predict(uf, itf, out=out, user_batch_size=300) is out and np.array_equal(out, whole)
    mm = np.lib.format.open_memmap(str(tmp_path / 'scores.npy'), mode='w+', dtype=np.float32, shape=(U, I))
    model.predict(uf, itf, out=mm, user_batch_size=256)
    assert np.array_equal(np.asarray(mm), whole)
    seen = []
    for u0, u1, block in model.predict_batches(uf, itf, user_batch_size=200):
        assert np.array_equal(block, whole[u0:u1])
        seen.append((u0, u1))
    assert seen == [(0, 200), (200, 400), (400, 600), (600, 700)]
    with pytest.raises(ValueError):
        model.predict(uf, itf, out=np.zeros((U, I + 1), np.float32))


@pytest.mark.parametrize('integer', [True, False])
def test_mixture_of_tastes_top_k_stays_on_the_fused_kernels(T, integer):
    """n_tastes > 1 without attention: prediction = max over tastes, so predict_rank(k) = one fused sweep per taste + a
    de-duplicating merge -- no [n_users, n_items] matrix (SURVEY a6 / VERDICT r1 'missing 3')."""
    U, I, d, k, n_tastes = 300, 4000, 64, 10, 3
    uf = H.tag_features(U, 200, 20, seed=1, integer=integer)
    itf = H.tag_features(I, 200, 20, seed=2, integer=integer)
    wus = [H.linear_weights(200, d, seed=10 + t, integer=integer) for t in range(n_tastes)]
    wi = H.linear_weights(200, d, seed=4, integer=integer)
    bu, bi = H.feature_biases(200, seed=5, integer=integer), H.feature_biases(200, seed=6, integer=integer)
    model = T.TensorRec(n_components=d, n_tastes=n_tastes)
    weights = {'linear_weights_item': wi, 'feature_biases_user': bu[:, None], 'feature_biases_item': bi[:, None]}
    for t in range(n_tastes):
        weights['linear_weights_user_%d' % t] = wus[t]
    model.set_weights(weights)
    top = model.predict_rank(uf, itf, k=k)
    assert model.last_topk_info['path'] == 'filter'
    scores = oracle.OracleModel(wus, wi, bu, bi).predict(uf, itf)
    exp_i, exp_s = oracle.top_k_from_scores(scores, k)
    if integer:
        assert np.array_equal(top.items, exp_i) and np.array_equal(top.scores, exp_s)
    else:
        rows = np.arange(U)[:, None]
        assert np.all(np.abs(top.scores - scores[rows, top.items]) <= 1e-5 * 40 + 2e-6)
        assert (top.items != exp_i).mean() < 0.01
        assert all(len(set(r)) == k for r in top.items)            # no item twice
    # the dense API of the same model agrees with the oracle as before
    if integer:
        assert np.array_equal(model.predict(uf, itf), scores)
