"""CPU tests of the host-side mirror of the reference API (no kernel is launched here): constructor validation,
error conventions, input coercion order, the training step over every plugin combination, persistence, metrics.
Cases follow the reference's test/test_tensorrec.py, test_representation_graphs.py, test_loss_graphs.py, test_eval.py,
test_util.py (cited per test)."""
import os
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle
import tensorrec_b200 as tensorrec
from tensorrec_b200 import TensorRec
from tensorrec_b200.errors import (
    ModelNotBiasedException, ModelNotFitException, ModelWithoutAttentionException, BatchNonSparseInputException
)
from tensorrec_b200.eval import recall_at_k, precision_at_k, f1_score_at_k, ndcg_at_k, _setup_ndcg, _idcg, _dcg
from tensorrec_b200.input_utils import create_tensorrec_dataset_from_sparse_matrix, SparseInput
from tensorrec_b200.kernels import DeviceCSR
from tensorrec_b200.loss_graphs import (
    RMSELossGraph, RMSEDenseLossGraph, SeparationLossGraph, SeparationDenseLossGraph, WMRBLossGraph,
    BalancedWMRBLossGraph
)
from tensorrec_b200.prediction_graphs import (
    DotProductPredictionGraph, CosineSimilarityPredictionGraph, EuclideanSimilarityPredictionGraph
)
from tensorrec_b200.representation_graphs import (
    LinearRepresentationGraph, NormalizedLinearRepresentationGraph, FeaturePassThroughRepresentationGraph,
    WeightedFeaturePassThroughRepresentationGraph, ReLURepresentationGraph, AbstractRepresentationGraph
)
from tensorrec_b200.session_management import set_session, get_session
from tensorrec_b200.util import generate_dummy_data, generate_dummy_data_with_indicator, sample_items
from tests import helpers as H


@pytest.fixture(autouse=True)
def fresh_session():
    set_session(None)          # as the reference tests do (test/test_tensorrec.py:32)
    yield
    set_session(None)


@pytest.fixture(scope='module')
def data():
    return generate_dummy_data(num_users=15, num_items=30, interaction_density=.5, num_user_features=200,
                               num_item_features=200, n_features_per_user=20, n_features_per_item=20,
                               pos_int_ratio=.5, seed=0)


# ---- constructor validation: test/test_tensorrec.py:46-71 -------------------------------------------------
def test_init_argument_checks():
    assert TensorRec() is not None
    assert TensorRec(n_components=10) is not None
    with pytest.raises(ValueError):
        TensorRec(n_components=0)
    with pytest.raises(ValueError):
        TensorRec(n_tastes=0)
    with pytest.raises(ValueError):
        TensorRec(user_repr_graph=None)
    with pytest.raises(ValueError):
        TensorRec(item_repr_graph=None)
    with pytest.raises(ValueError):
        TensorRec(prediction_graph=None)
    with pytest.raises(ValueError):
        TensorRec(loss_graph=None)
    with pytest.raises(ValueError):
        TensorRec(user_repr_graph=np.mean)
    with pytest.raises(ValueError):
        TensorRec(item_repr_graph=np.mean)
    with pytest.raises(ValueError):
        TensorRec(prediction_graph=np.mean)
    with pytest.raises(ValueError):
        TensorRec(loss_graph=np.mean)
    with pytest.raises(ValueError):                       # attention_graph must be None if n_tastes == 1
        TensorRec(attention_graph=LinearRepresentationGraph())
    with pytest.raises(ValueError):
        TensorRec(n_tastes=2, attention_graph=np.mean)
    # abstract plugin classes are not enforced abstract (py2-style metaclass in the reference)
    assert AbstractRepresentationGraph() is not None


# ---- predict before fit: test/test_tensorrec.py:73-93 -----------------------------------------------------
def test_predict_before_fit_raises(data):
    _, uf, itf = data
    model = TensorRec()
    for call in (lambda: model.predict(uf, itf), lambda: model.predict_rank(uf, itf),
                 lambda: model.predict_rank(uf, itf, k=5), lambda: model.predict_user_representation(uf),
                 lambda: model.predict_item_representation(itf), lambda: model.predict_user_bias(uf),
                 lambda: model.predict_item_bias(itf), lambda: model.predict_similar_items(itf, [1], 2),
                 lambda: model.predict_user_attention_representation(uf), lambda: model.save_model('/tmp/x')):
        with pytest.raises(ModelNotFitException) as err:
            call()
        assert 'has been called before model fitting' in str(err.value)


def test_bias_and_attention_exceptions_precede_any_device_work(data):
    _, uf, itf = data
    model = TensorRec(biased=False, n_components=4)
    model.set_weights({'linear_weights_user_0': np.zeros((200, 4)), 'linear_weights_item': np.zeros((200, 4))})
    with pytest.raises(ModelNotBiasedException):
        model.predict_user_bias(uf)
    with pytest.raises(ModelNotBiasedException):
        model.predict_item_bias(itf)
    with pytest.raises(ModelWithoutAttentionException):
        model.predict_user_attention_representation(uf)


# ---- fit smoke over every plugin: test_representation_graphs.py:14-35, test_loss_graphs.py:17-47 -----------
@pytest.mark.parametrize('user_repr,item_repr,n_user_features,n_item_features,n_components', [
    (LinearRepresentationGraph, LinearRepresentationGraph, 50, 60, 20),
    (NormalizedLinearRepresentationGraph, NormalizedLinearRepresentationGraph, 50, 60, 20),
    (LinearRepresentationGraph, FeaturePassThroughRepresentationGraph, 50, 60, 60),
    (LinearRepresentationGraph, WeightedFeaturePassThroughRepresentationGraph, 50, 60, 60),
    (LinearRepresentationGraph, ReLURepresentationGraph, 50, 60, 20),
])
def test_fit_runs_for_every_representation_graph(user_repr, item_repr, n_user_features, n_item_features, n_components):
    interactions, uf, itf = generate_dummy_data(num_users=15, num_items=30, interaction_density=.5,
                                                num_user_features=n_user_features,
                                                num_item_features=n_item_features, n_features_per_user=20,
                                                n_features_per_item=20, pos_int_ratio=.5, seed=1)
    model = TensorRec(n_components=n_components, user_repr_graph=user_repr(), item_repr_graph=item_repr())
    model.fit(interactions, uf, itf, epochs=3)
    assert model.tf_prediction is not None
    assert all(np.isfinite(w).all() for w in model.get_weights().values())


def test_feature_pass_through_needs_matching_width(data):
    interactions, uf, itf = data
    model = TensorRec(n_components=5, item_repr_graph=FeaturePassThroughRepresentationGraph())
    with pytest.raises(ValueError):
        model.fit(interactions, uf, itf, epochs=1)


@pytest.mark.parametrize('loss,kwargs', [
    (RMSELossGraph, {}), (RMSEDenseLossGraph, {}), (SeparationLossGraph, {}), (SeparationDenseLossGraph, {}),
    (WMRBLossGraph, {'n_sampled_items': 10}), (BalancedWMRBLossGraph, {'n_sampled_items': 10}),
])
@pytest.mark.parametrize('biased', [True, False])
def test_fit_runs_and_learns_for_every_loss_graph(data, loss, kwargs, biased):
    interactions, uf, itf = data
    model = TensorRec(n_components=8, loss_graph=loss(), biased=biased)
    model.fit(interactions, uf, itf, epochs=1, **kwargs)
    before = model.get_weights()
    model.fit_partial(interactions, uf, itf, epochs=4, **kwargs)
    after = model.get_weights()
    assert any(not np.array_equal(before[k], after[k]) for k in before)
    assert all(np.isfinite(v).all() for v in after.values())


def test_rmse_fit_reduces_the_loss(data):
    interactions, uf, itf = data
    torch.manual_seed(0)
    model = TensorRec(n_components=8)
    dev = get_session().device

    def loss_now():
        from tensorrec_b200.session_management import variable_scope
        with variable_scope(model._variables):
            basic, _, _, _ = model._training_losses(SparseInput(interactions), SparseInput(uf), SparseInput(itf), None, dev)
        return float(basic.detach())

    model.fit(interactions, uf, itf, epochs=1, learning_rate=0.05)
    first = loss_now()
    model.fit_partial(interactions, uf, itf, epochs=40, learning_rate=0.05)
    assert loss_now() < first


def test_sampled_loss_requires_n_sampled_items(data):
    interactions, uf, itf = data
    model = TensorRec(loss_graph=WMRBLossGraph())
    with pytest.raises(ValueError):
        model.fit(interactions, uf, itf, epochs=1)
    with pytest.raises(ValueError):
        model.fit(interactions, uf, itf, epochs=1, n_sampled_items=0)


def test_fit_tastes_and_attention(data):
    # test/test_tensorrec.py:278-339
    interactions, uf, itf = data
    model = TensorRec(n_components=10, n_tastes=3, user_repr_graph=NormalizedLinearRepresentationGraph(),
                      attention_graph=LinearRepresentationGraph(), loss_graph=BalancedWMRBLossGraph())
    model.fit(interactions, uf, itf, epochs=2, n_sampled_items=5)
    names = set(model.get_weights())
    assert {'linear_weights_user_0', 'linear_weights_user_1', 'linear_weights_user_2', 'linear_weights_attn_0',
            'linear_weights_attn_2', 'linear_weights_item', 'feature_biases_user', 'feature_biases_item'} <= names


# ---- batching and input validation: test/test_tensorrec.py:95-172, util.py:57-58 ----------------------------
def test_fit_batched_and_lists(data):
    interactions, uf, itf = data
    model = TensorRec(n_components=10)
    model.fit(interactions, uf, itf, epochs=2, user_batch_size=2)
    model = TensorRec(n_components=10)
    model.fit([interactions.tocsr()[:7], interactions.tocsr()[7:]], [uf.tocsr()[:7], uf.tocsr()[7:]], itf, epochs=2)
    with pytest.raises(ValueError):
        model.fit([interactions, interactions], [uf], itf, epochs=1)
    with pytest.raises(ValueError):
        model.fit([interactions, interactions], [uf, uf], [itf, itf, itf], epochs=1)


def test_fit_rejects_non_sparse_input(data):
    interactions, uf, itf = data
    model = TensorRec(n_components=10)
    with pytest.raises(ValueError):
        model.fit(np.array([1, 2, 3, 4]), uf, itf, epochs=1)
    with pytest.raises(ValueError):
        model.fit(interactions, uf, 'not-a-matrix', epochs=1) if False else model.fit(interactions, uf, 7, epochs=1)
    with pytest.raises(BatchNonSparseInputException):
        model.fit(create_tensorrec_dataset_from_sparse_matrix(interactions), uf, itf, epochs=1, user_batch_size=2)
    with pytest.raises(FileNotFoundError):              # a str is a TFRecord path (tests/test_tfrecord.py)
        model.fit('/tmp/no-such-interactions.tfrecord', uf, itf, epochs=1)


def test_fit_accepts_dataset_tuples(data):
    # test/test_tensorrec.py:342-367 (tf.data.Dataset inputs become the 5-tuple itself)
    interactions, uf, itf = data
    model = TensorRec(n_components=10)
    model.fit(create_tensorrec_dataset_from_sparse_matrix(interactions), create_tensorrec_dataset_from_sparse_matrix(uf),
              create_tensorrec_dataset_from_sparse_matrix(itf), epochs=2)
    assert model.n_user_features == 200 and model.n_item_features == 200


def test_predict_without_cuda_fails_loudly(data):
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    interactions, uf, itf = data
    model = TensorRec(n_components=10)
    model.fit(interactions, uf, itf, epochs=1)
    for call in (lambda: model.predict(uf, itf), lambda: model.predict_rank(uf, itf),
                 lambda: model.predict_rank(uf, itf, k=3), lambda: model.predict_item_representation(itf),
                 lambda: model.predict_user_bias(uf)):
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            call()


# ---- input order: tensorrec/input_utils.py:22-40 -------------------------------------------------------------
@pytest.mark.parametrize('fmt', ['coo', 'csr', 'csc', 'lil', 'dok'])
def test_host_csr_keeps_the_reference_coo_order(fmt):
    m = H.messy_coo(40, 25, 300, seed=8)
    mm = m if fmt == 'coo' else getattr(m, 'to' + fmt)()
    row, col, val, d0, d1 = oracle.coo_from_sparse(mm)
    indptr, ccol, cval = DeviceCSR.host_arrays(mm)
    assert indptr.dtype == np.int32 and ccol.dtype == np.int32 and cval.dtype == np.float32
    assert indptr[0] == 0 and indptr[-1] == len(row) and len(indptr) == d0 + 1
    for r in range(d0):
        sel = row == r                                  # the row's entries, in the order the reference sees them
        assert np.array_equal(ccol[indptr[r]:indptr[r + 1]], col[sel])
        assert np.array_equal(cval[indptr[r]:indptr[r + 1]], val[sel])


def test_dataset_format_matches_reference_dtypes():
    ds = create_tensorrec_dataset_from_sparse_matrix(sp.random(5, 7, density=.5, format='csr', random_state=0))
    assert ds.row_index.dtype == np.int64 and ds.col_index.dtype == np.int64 and ds.values.dtype == np.float32
    assert (ds.d0, ds.d1) == (5, 7)


# ---- persistence: test/test_tensorrec.py:398-458 (value round trip; the predict half runs on the GPU) ---------
def test_save_and_load_round_trip(data):
    interactions, uf, itf = data
    model = TensorRec(n_components=10, n_tastes=2, prediction_graph=CosineSimilarityPredictionGraph())
    model.fit(interactions, uf, itf, epochs=2)
    with tempfile.TemporaryDirectory() as tmp:
        model.save_model(os.path.join(tmp, 'model'))
        assert model.tf_prediction is not None                      # still usable after saving
        assert sorted(os.listdir(os.path.join(tmp, 'model'))) == ['tensorrec.pkl', 'tensorrec_session.npz']
        loaded = TensorRec.load_model(os.path.join(tmp, 'model'))
    assert loaded.tf_prediction is not None and loaded.n_tastes == 2
    assert isinstance(loaded.prediction_graph_factory, CosineSimilarityPredictionGraph)
    for k, v in model.get_weights().items():
        assert np.array_equal(v, loaded.get_weights()[k])
    loaded.fit_partial(interactions, uf, itf, epochs=1)            # and trainable again


# ---- util: test/test_util.py, util.py:12-21 -----------------------------------------------------------------
def test_sample_items_shapes_and_uniqueness():
    pairs = sample_items(n_items=20, n_users=6, n_sampled_items=7, replace=False, rng=np.random.default_rng(0))
    assert pairs.shape == (42, 2) and pairs.dtype == np.int64
    assert np.array_equal(pairs[:, 0], np.repeat(np.arange(6), 7))
    for u in range(6):
        assert len(set(pairs[pairs[:, 0] == u, 1])) == 7
    assert sample_items(5, 3, 9, replace=True).shape == (27, 2)
    with pytest.raises(ValueError):
        sample_items(5, 3, 9, replace=False)


def test_dummy_data_generators_follow_the_reference_shapes():
    i, u, it = generate_dummy_data(num_users=100, num_items=150, interaction_density=.05, seed=0)
    assert i.shape == (100, 150) and u.shape == (100, 200) and it.shape == (150, 200)
    assert abs(u.nnz / 100.0 - 20) < 3
    i, u, it = generate_dummy_data_with_indicator(num_users=10, num_items=12, interaction_density=.5, seed=0)
    assert u.shape == (10, 12) and it.shape == (12, 14) and i.shape == (10, 12)
    assert np.all(u.toarray().diagonal() == 1)


# ---- metrics: test/test_eval.py:84-148 ------------------------------------------------------------------------
def test_idcg_known_value():
    hits = np.array([3, 3, 3, 2, 2, 2, 1, 0])
    rng = np.random.default_rng(0)
    shuffled = hits.copy()
    rng.shuffle(shuffled)
    assert abs(_idcg(shuffled) - 18.77105) < 1e-3
    assert _idcg(hits) == _idcg(shuffled)
    binary = np.array([1, 1, 1, 1, 0, 0])
    assert _idcg(binary) == np.sum([(2 ** e - 1) / np.log2(i + 2) for i, e in enumerate(binary)])


def test_ndcg_setup_and_dcg_known_values():
    rel, k_mask, ror, ror_at_k = _setup_ndcg(np.array([1, 2, 3, 4, 5, 6]), sp.lil_matrix(np.array([3, 2, 3, 0, 1, 2])))
    assert len(k_mask) == 5 and len(ror_at_k) == 5
    assert list(ror.data) == [1, 2, 3, 5, 6]
    rel, k_mask, ror, ror_at_k = _setup_ndcg(np.array([1, 2, 3, 4, 5]), sp.lil_matrix(np.array([3, 3, 1, 0, 2])))
    by_hand = np.sum((2 ** np.array([3, 3, 1, 2]) - 1) / np.log2(np.array([1, 2, 3, 5]) + 1))
    func_dcg = _dcg(rel, k_mask, ror_at_k, ror)
    assert by_hand == func_dcg
    assert abs((func_dcg / _idcg(np.array([3, 3, 1, 0, 2]))).item(0) - .979762) < 1e-3


def test_metrics_accept_full_ranks_and_top_k():
    rng = np.random.default_rng(3)
    scores = rng.standard_normal((12, 40)).astype(np.float32)
    ranks = oracle.rank_predictions(scores)
    interactions = sp.random(12, 40, density=.2, format='csr', random_state=rng)
    interactions.data[:] = rng.integers(1, 4, interactions.nnz)
    ids, vals = oracle.top_k_from_scores(scores, 10)
    topk = tensorrec.TopK(ids, vals)
    for fn in (recall_at_k, precision_at_k, ndcg_at_k):
        for preserve in (False, True):
            full = fn(ranks, interactions, k=10, preserve_rows=preserve)
            part = fn(topk, interactions, k=10, preserve_rows=preserve)
            assert np.allclose(full, part, equal_nan=True)
            assert np.allclose(fn(ranks, interactions, k=4, preserve_rows=preserve),
                               fn(topk, interactions, k=4, preserve_rows=preserve), equal_nan=True)
    assert abs(f1_score_at_k(ranks, interactions, k=10) - f1_score_at_k(topk, interactions, k=10)) < 1e-12
    n10 = np.mean(ndcg_at_k(ranks, interactions, k=40))
    assert np.mean(ndcg_at_k(ranks, interactions, k=5)) <= n10 < 1


def test_host_result_pool_never_aliases_live_results(monkeypatch):
    """kernels._HostResults hands out page-locked result buffers; one is reused only after the array returned for it
    (and every view of it) has been collected."""
    import gc
    import weakref
    from tensorrec_b200 import kernels
    real_empty = torch.empty
    monkeypatch.setattr(torch, 'empty', lambda *a, pin_memory=False, **k: real_empty(*a, **k))   # no CUDA here
    pool = kernels._HostResults()

    def hand_out(shape, dtype):
        buf = pool._take(shape, dtype)
        arr = buf.numpy()
        pool._idle.append((buf, weakref.ref(arr)))
        return arr

    first = hand_out((4, 3), torch.float32)
    first[:] = 7
    second = hand_out((4, 3), torch.float32)
    assert not np.shares_memory(first, second)
    view = first[:2]
    del first
    gc.collect()
    third = hand_out((4, 3), torch.float32)
    assert not np.shares_memory(view, third) and np.all(view == 7)      # a view keeps its buffer out of circulation
    addr = view.__array_interface__['data'][0]
    del view
    gc.collect()
    fourth = hand_out((4, 3), torch.float32)
    assert fourth.__array_interface__['data'][0] == addr                 # now it is recycled
    other = hand_out((4, 3), torch.int32)
    assert other.dtype == np.int32 and not np.shares_memory(other, fourth)


def test_transposed_csr_keeps_duplicates_and_row_order():
    """DeviceCSR.host_arrays_transposed: the backward operand of the K1 training step (dW = A^T . d_out)."""
    rows = np.array([2, 0, 2, 1, 2, 0], dtype=np.int64)
    cols = np.array([1, 3, 1, 0, 3, 3], dtype=np.int64)          # (2, 1) twice; unsorted COO
    vals = np.array([1.5, -2.0, 0.25, 4.0, 3.0, 7.0], dtype=np.float32)
    a = sp.coo_matrix((vals, (rows, cols)), shape=(3, 5))
    indptr, col, val = DeviceCSR.host_arrays_transposed(a)
    assert indptr.dtype == np.int32 and col.dtype == np.int32 and val.dtype == np.float32
    assert list(indptr) == [0, 1, 3, 3, 6, 6]                     # columns 2 and 4 are empty
    assert list(col) == [1, 2, 2, 0, 0, 2]                        # ascending row inside a column, duplicates kept
    assert list(val) == [4.0, 1.5, 0.25, -2.0, 7.0, 3.0]          # (2,1): storage order 1.5 then 0.25
    g = np.arange(12, dtype=np.float64).reshape(3, 4)
    dense_t = np.zeros((5, 3))
    for f in range(5):
        for p in range(indptr[f], indptr[f + 1]):
            dense_t[f, col[p]] += val[p]
    assert np.allclose(dense_t @ g, a.toarray().T.astype(np.float64) @ g)


def test_bias_processing_order_is_a_stable_descending_sort(monkeypatch):
    """kernels.bias_processing_order: the item order the filter kernel sweeps (highest bias first, ties by lower
    item index -> the order, and with it every result, is deterministic).  Here the library-sort form
    (TENSORREC_B200_BIAS_ORDER=torch); the default own-kernel form is checked against the same statement on the GPU
    (tests/test_kernels_gpu.py) and raises without a CUDA device like every other kernel call."""
    from tensorrec_b200 import kernels
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            kernels.bias_processing_order(torch.zeros(4))
    monkeypatch.setattr(kernels, 'BIAS_ORDER', 'torch')
    rng = np.random.default_rng(11)
    bias = rng.integers(-3, 4, size=5000).astype(np.float32) * 0.25          # many exact ties
    perm = kernels.bias_processing_order(torch.from_numpy(bias))
    assert perm.dtype == torch.int32
    assert np.array_equal(perm.numpy(), np.argsort(-bias, kind='stable'))
    assert kernels.bias_processing_order(None) is None


# ---- the host mirror's serial forms against the reference's known answers ---------------------------
def test_serial_prediction_graphs_match_the_reference_known_answers():
    import json
    golden = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_known_answers.json')))
    from tensorrec_b200.recommendation_graphs import (
        bias_prediction_serial, densify_sampled_item_predictions, split_sparse_tensor_indices)

    def t(x, dtype=torch.float32):
        return torch.tensor(x, dtype=dtype)

    for key, graph in (('dot_product_serial', DotProductPredictionGraph()),
                       ('cosine_serial', CosineSimilarityPredictionGraph()),
                       ('euclidean_serial', EuclideanSimilarityPredictionGraph())):
        g = golden[key]
        got = graph.connect_serial_prediction_graph(
            tf_user_representation=t(g['user_repr']), tf_item_representation=t(g['item_repr']),
            tf_x_user=t(g['x_user'], torch.long), tf_x_item=t(g['x_item'], torch.long)).numpy()
        expect = np.array(g['expected']) if 'expected' in g else -np.sqrt(np.array(g['expected_neg_sqrt_of']))
        assert np.allclose(got, expect, atol=1e-6), key

    g = golden['split_sparse_tensor_indices']
    interactions = SparseInput(sp.coo_matrix(np.array(g['interactions'], dtype=np.float32)))
    x_user, x_item = split_sparse_tensor_indices(tf_sparse_tensor=interactions.torch_sparse('cpu'), n_dimensions=2)
    assert x_user.tolist() == g['expected_user'] and x_item.tolist() == g['expected_item']

    g = golden['bias_prediction_serial']
    got = bias_prediction_serial(
        tf_prediction_serial=t(g['predictions']), tf_projected_user_biases=t(g['user_biases']),
        tf_projected_item_biases=t(g['item_biases']), tf_x_user=t(g['x_user'], torch.long),
        tf_x_item=t(g['x_item'], torch.long)).numpy()
    assert np.array_equal(got, np.array(g['expected'], dtype=np.float32))

    g = golden['densify_sampled_item_predictions']
    got = densify_sampled_item_predictions(tf_sample_predictions_serial=t(g['input'], torch.long),
                                           tf_n_sampled_items=g['n_sampled_items'], tf_n_users=g['n_users']).numpy()
    assert np.array_equal(got, np.array(g['expected']))


def test_loss_graphs_match_the_oracle_restatement():
    """Host mirror (torch) vs oracle/loss_ops.py (numpy, follows tensorrec/loss_graphs.py line by line) on random
    inputs with duplicates, non-positive interactions and empty rows."""
    from oracle import loss_ops as L
    rng = np.random.default_rng(21)
    n_users, n_items, n_sampled = 13, 17, 5
    nnz = 60
    row, col = rng.integers(0, n_users, nnz), rng.integers(0, n_items, nnz)
    val = rng.integers(-1, 4, nnz).astype(np.float32)
    inter = sp.coo_matrix((val, (row, col)), shape=(n_users, n_items))
    coo = oracle.coo_from_sparse(inter)
    tf_inter = SparseInput(inter).torch_sparse('cpu')
    pred = rng.standard_normal((n_users, n_items)).astype(np.float32)
    pred_serial = pred[coo[0], coo[1]]
    samples = rng.standard_normal((n_users, n_sampled)).astype(np.float32)
    t = torch.from_numpy
    kw = dict(tf_prediction_serial=t(pred_serial), tf_interactions_serial=t(coo[2]), tf_interactions=tf_inter,
              tf_n_users=n_users, tf_n_items=n_items, tf_prediction=t(pred), tf_rankings=None,
              tf_sample_predictions=t(samples), tf_n_sampled_items=n_sampled)
    pairs = [(RMSELossGraph(), L.rmse(pred_serial, coo[2])),
             (RMSEDenseLossGraph(), L.rmse_dense(coo, pred)),
             (SeparationLossGraph(), L.separation(pred_serial, coo[2])),
             (SeparationDenseLossGraph(), L.separation_dense(pred, coo)),
             (WMRBLossGraph(), L.wmrb(pred_serial, coo, samples, n_items, n_sampled)),
             (BalancedWMRBLossGraph(), L.balanced_wmrb(pred_serial, coo, samples, n_items, n_sampled))]
    for graph, expect in pairs:
        got = graph.connect_loss_graph(**kw).numpy()
        assert got.shape == np.shape(expect), type(graph).__name__
        assert np.allclose(got, expect, rtol=2e-5, atol=2e-6), type(graph).__name__
