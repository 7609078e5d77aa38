"""CPU tests: the oracle (oracle/) against every known-answer vector the reference's own tests hold for the
predict / predict_rank path (tests/golden/reference_known_answers.json, each entry cites the reference test),
plus the oracle's internal consistency on the cases the reference never pins (ties, duplicates, empty rows)."""
import json
import math
import os

import numpy as np
import scipy.sparse as sp

import oracle
from oracle import reference_ops as R
from tests import helpers as H

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_known_answers.json')))
F32 = np.float32


def arr(x):
    return np.array(x, dtype=F32)


def test_dot_product_dense_golden():
    g = GOLDEN['dot_product_dense']
    assert np.allclose(oracle.dot_product_dense(arr(g['user_repr']), arr(g['item_repr'])), arr(g['expected']))


def test_cosine_dense_golden():
    g = GOLDEN['cosine_dense']
    assert np.allclose(oracle.cosine_dense(arr(g['user_repr']), arr(g['item_repr'])), arr(g['expected']), atol=1e-6)


def test_euclidean_dense_golden():
    g = GOLDEN['euclidean_dense']
    expect = -np.sqrt(np.array(g['expected_neg_sqrt_of']))
    assert np.allclose(oracle.euclidean_dense(arr(g['user_repr']), arr(g['item_repr'])), expect, atol=1e-6)


def test_project_biases_golden():
    g = GOLDEN['project_biases']
    coo = oracle.coo_from_sparse(sp.coo_matrix(arr(g['features'])))
    assert np.array_equal(oracle.project_biases(coo, arr(g['feature_biases'])), arr(g['expected']))


def test_bias_prediction_dense_golden():
    g = GOLDEN['bias_prediction_dense']
    got = oracle.bias_prediction_dense(arr(g['predictions']), arr(g['user_biases']), arr(g['item_biases']))
    assert np.array_equal(got, arr(g['expected']))


def test_rank_predictions_golden():
    g = GOLDEN['rank_predictions']
    got = oracle.rank_predictions(arr(g['predictions']))
    assert got.dtype == np.int32
    assert np.array_equal(got, np.array(g['expected']))
    assert np.array_equal(oracle.rank_predictions_closed_form(arr(g['predictions'])), np.array(g['expected']))


def test_collapse_mixture_of_tastes_golden():
    g = GOLDEN['collapse_mixture_of_tastes']
    got = oracle.collapse_mixture_of_tastes([arr(p) for p in g['predictions']], None)
    assert np.array_equal(got, arr(g['expected']))


def test_collapse_with_attention_golden():
    g = GOLDEN['collapse_mixture_of_tastes_with_attention']
    got = oracle.collapse_mixture_of_tastes([arr(p) for p in g['predictions']], [arr(a) for a in g['attentions']])
    expect = arr(g['expected'])
    assert np.all(np.abs(got - expect) <= 2 * np.spacing(expect))


def test_predict_similar_items_golden():
    g = GOLDEN['predict_similar_items_cosine']
    got = oracle.predict_similar_items('cosine', arr(g['item_repr']), g['item_ids'])
    assert np.array_equal(got, arr(g['expected']))


def test_batched_alpha_golden():
    g = GOLDEN['calculate_batched_alpha']
    from tensorrec_b200.util import calculate_batched_alpha
    got = calculate_batched_alpha(num_batches=g['num_batches'], alpha=g['alpha'])
    assert abs(got / g['alpha'] - g['expected_ratio']) < 1e-4
    assert abs(1.0 / (math.e * math.log(2)) - g['expected_ratio']) < 1e-4


# ---- consistency on what the reference leaves unpinned ------------------------------------------------
def test_double_sort_equals_closed_form_with_ties():
    rng = np.random.default_rng(0)
    s = rng.integers(-3, 4, size=(17, 53)).astype(F32)
    s[0, :10] = 0.0
    s[0, 3] = -0.0
    assert np.array_equal(oracle.rank_predictions(s), oracle.rank_predictions_closed_form(s))
    ranks = oracle.rank_predictions(s)
    assert np.array_equal(np.sort(ranks, axis=1), np.tile(np.arange(1, 54, dtype=np.int32), (17, 1)))


def test_top_k_is_the_rank_le_k_set():
    rng = np.random.default_rng(1)
    s = rng.integers(-2, 3, size=(9, 40)).astype(F32)
    ids, vals = oracle.top_k_from_scores(s, 7)
    ranks = oracle.rank_predictions(s)
    for u in range(9):
        assert np.array_equal(ranks[u, ids[u]], np.arange(1, 8))
        assert np.array_equal(vals[u], s[u, ids[u]])


def test_fast_top_k_equals_the_full_sort_top_k():
    """top_k_from_scores_fast (argpartition + tie-aware ordering) == top_k_from_scores (full stable argsort), with
    massive ties, ties straddling the k-th place, -0.0 / +0.0, infinities and k >= n_items."""
    rng = np.random.default_rng(2)
    for shape, k in (((9, 40), 7), ((5, 3000), 10), ((4, 12), 12), ((3, 6), 10), ((6, 500), 1)):
        s = rng.integers(-2, 3, size=shape).astype(F32)
        s[0, :5] = 0.0
        s[0, 2] = -0.0
        if shape[1] > 20:
            s[1, 7] = np.inf
            s[1, 9] = -np.inf
        a_i, a_v = oracle.top_k_from_scores(s, k)
        b_i, b_v = oracle.top_k_from_scores_fast(s, k)
        assert np.array_equal(a_i, b_i) and np.array_equal(a_v, b_v)
    f = rng.standard_normal((7, 2000)).astype(F32)
    a_i, a_v = oracle.top_k_from_scores(f, 10)
    b_i, b_v = oracle.top_k_from_scores_fast(f, 10)
    assert np.array_equal(a_i, b_i) and np.array_equal(a_v, b_v)


def test_spmm_duplicates_unsorted_empty_rows():
    m = H.messy_coo(37, 23, 300, seed=5)
    w = H.linear_weights(23, 12, seed=1)
    got = oracle.sparse_dense_matmul(oracle.coo_from_sparse(m), w)
    dense = np.zeros((37, 23), dtype=np.float64)
    np.add.at(dense, (m.row, m.col), m.data.astype(np.float64))
    assert np.allclose(got, dense @ w.astype(np.float64), atol=1e-4)
    assert np.all(got[np.bincount(m.row, minlength=37) == 0] == 0)
    # the timed scipy path computes the same contraction
    fast = R.sparse_dense_matmul_fast(sp.csr_matrix(m), w)
    assert np.allclose(fast, got, atol=1e-5)


def test_l2_normalize_zero_rows_stay_zero():
    x = np.zeros((3, 5), dtype=F32)
    x[1] = [3, 0, 4, 0, 0]
    n = oracle.l2_normalize(x)
    assert np.all(n[0] == 0) and np.all(n[2] == 0)
    assert np.allclose(n[1], [0.6, 0, 0.8, 0, 0])


def test_oracle_model_composition_order():
    uf = H.tag_features(12, 30, 5, seed=0)
    itf = H.tag_features(20, 30, 5, seed=1)
    wu = [H.linear_weights(30, 8, seed=s) for s in (2, 3)]
    wi = H.linear_weights(30, 8, seed=4)
    bu, bi = H.feature_biases(30, 5), H.feature_biases(30, 6)
    m = oracle.OracleModel(wu, wi, bu, bi, user_repr='normalized_linear', prediction='cosine')
    pred = m.predict(uf, itf)
    ucoo, icoo = oracle.coo_from_sparse(uf), oracle.coo_from_sparse(itf)
    ir = oracle.linear_representation(icoo, wi)
    per_taste = [oracle.cosine_dense(oracle.normalized_linear_representation(ucoo, w), ir) for w in wu]
    expect = oracle.bias_prediction_dense(np.max(np.stack(per_taste), axis=0), oracle.project_biases(ucoo, bu),
                                          oracle.project_biases(icoo, bi))
    assert np.array_equal(pred, expect)
    assert np.array_equal(m.predict_rank(uf, itf), oracle.rank_predictions(pred))


# ---- the serial forms used by the training step -------------------------------------------------------
def test_serial_predictions_golden():
    g = GOLDEN['dot_product_serial']
    got = oracle.dot_product_serial(arr(g['user_repr']), arr(g['item_repr']), g['x_user'], g['x_item'])
    assert got.dtype == F32 and np.allclose(got, arr(g['expected']))
    g = GOLDEN['cosine_serial']
    got = oracle.cosine_serial(arr(g['user_repr']), arr(g['item_repr']), g['x_user'], g['x_item'])
    assert np.allclose(got, arr(g['expected']), atol=1e-6)
    g = GOLDEN['euclidean_serial']
    got = oracle.euclidean_serial(arr(g['user_repr']), arr(g['item_repr']), g['x_user'], g['x_item'])
    expect = -np.sqrt(np.array(g['expected_neg_sqrt_of']))
    assert np.allclose(got, expect, atol=1e-6)


def test_serial_forms_equal_the_dense_forms_entry_by_entry():
    rng = np.random.default_rng(4)
    u, i = rng.standard_normal((7, 5)).astype(F32), rng.standard_normal((9, 5)).astype(F32)
    xu, xi = np.repeat(np.arange(7), 9), np.tile(np.arange(9), 7)
    assert np.allclose(oracle.dot_product_serial(u, i, xu, xi).reshape(7, 9), oracle.dot_product_dense(u, i), atol=1e-5)
    assert np.allclose(oracle.cosine_serial(u, i, xu, xi).reshape(7, 9), oracle.cosine_dense(u, i), atol=1e-5)
    assert np.allclose(oracle.euclidean_serial(u, i, xu, xi).reshape(7, 9), oracle.euclidean_dense(u, i), atol=1e-4)


def test_split_indices_bias_serial_and_densify_golden():
    g = GOLDEN['split_sparse_tensor_indices']
    x_user, x_item = oracle.split_sparse_tensor_indices(sp.coo_matrix(arr(g['interactions'])))
    assert np.array_equal(x_user, g['expected_user']) and np.array_equal(x_item, g['expected_item'])
    g = GOLDEN['bias_prediction_serial']
    got = oracle.bias_prediction_serial(arr(g['predictions']), arr(g['user_biases']), arr(g['item_biases']),
                                        g['x_user'], g['x_item'])
    assert np.array_equal(got, arr(g['expected']))
    g = GOLDEN['densify_sampled_item_predictions']
    got = oracle.densify_sampled_item_predictions(np.array(g['input']), g['n_sampled_items'], g['n_users'])
    assert np.array_equal(got, np.array(g['expected']))


# ---- loss graphs (parity unpinned by the reference: hand-computed cases) ------------------------------
def test_loss_oracle_hand_computed_cases():
    from oracle import loss_ops as L
    assert abs(float(L.rmse([1.0, 2.0, 4.0], [1.0, 0.0, 0.0])) - math.sqrt((0 + 4 + 16) / 3.0)) < 1e-6
    inter = oracle.coo_from_sparse(sp.coo_matrix(np.array([[1.0, 0.0], [0.0, 2.0]], dtype=F32)))
    pred = np.array([[0.5, 0.5], [1.0, 1.0]], dtype=F32)
    assert abs(float(L.rmse_dense(inter, pred)) - math.sqrt((0.25 + 0.25 + 1.0 + 1.0) / 4.0)) < 1e-6
    # one positive interaction (user 1, item 0, prediction 0.5), two sampled items with predictions 0.0 and 2.0:
    # margins max(0, 1 - 0.5 + 0) = 0.5 and max(0, 1 - 0.5 + 2) = 2.5 -> log(1 + 10 / 2 * 3.0)
    inter = oracle.coo_from_sparse(sp.coo_matrix(np.array([[0.0, -1.0], [3.0, 0.0]], dtype=F32)))
    pred_serial = np.array([9.0, 0.5], dtype=F32)                       # COO order: (0,1) = -1 first, then (1,0) = 3
    samples = np.array([[7.0, 7.0], [0.0, 2.0]], dtype=F32)
    got = L.wmrb(pred_serial, inter, samples, n_items=10, n_sampled_items=2)
    assert got.shape == (1,) and abs(float(got[0]) - math.log(1.0 + 5.0 * 3.0)) < 1e-6
    # balanced: x interaction value 3 / item 0's positive mass 3 -> unchanged here
    got_b = L.balanced_wmrb(pred_serial, inter, samples, n_items=10, n_sampled_items=2)
    assert abs(float(got_b[0]) - float(got[0])) < 1e-6
    # separation: positives {2, 4} (mean 3, var 1), negatives {0, 0} (mean 0, var 0): 1 - Phi((0 - (-3)) / 1)
    sep = L.separation([2.0, 4.0, 0.0, 0.0], [1.0, 5.0, 0.0, -2.0])
    assert abs(float(sep) - (1.0 - 0.5 * (1.0 + math.erf(3.0 / math.sqrt(2.0))))) < 1e-6
