"""CPU tests of the round-1 advisor findings: anonymous plugin variables keep their identity across steps,
a failed first fit leaves the model unbuilt, the negative sampler has bounded memory, tf_rankings is evaluated only for
loss graphs that name it, wide rows are tiled over K1."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compat'))

import tensorrec_b200 as T                      # noqa: E402
from tensorrec_b200 import util                 # noqa: E402
from tensorrec_b200.errors import ModelNotFitException      # noqa: E402


@pytest.fixture(autouse=True)
def cpu_session():
    from tensorrec_b200 import session_management as sm
    sm.set_session(sm.Session('cpu'))
    yield
    sm.set_session(None)


def test_unnamed_plugin_variables_are_created_once():
    """A plugin written against the reference may call tf.Variable(...) without a name (the reference's own
    project_biases does).  connect_* methods run on every training step here, so the variable must resolve to the same
    tensor every time: the model trains and _variables does not grow."""
    import tensorflow as tf          # the stand-in under compat/

    class Anonymous(T.representation_graphs.AbstractRepresentationGraph):
        def connect_representation_graph(self, tf_features, n_components, n_features, node_name_ending):
            w = tf.Variable(tf.random_normal([n_features, n_components], stddev=0.1))       # no name=
            b = tf.Variable(np.zeros((1, n_components), dtype=np.float32))                  # a second anonymous one
            return tf.sparse_tensor_dense_matmul(tf_features, w) + b, [w, b]

    interactions, uf, itf = util.generate_dummy_data(num_users=15, num_items=20, interaction_density=.3, seed=1)
    model = T.TensorRec(n_components=4, user_repr_graph=Anonymous(), item_repr_graph=Anonymous())
    model.fit(interactions, uf, itf, epochs=1)
    names = list(model._variables)
    first = {k: v.detach().clone() for k, v in model._variables.items()}
    model.fit_partial(interactions, uf, itf, epochs=4)
    assert list(model._variables) == names, 'training steps registered new variables: %r' % list(model._variables)
    assert sorted(n for n in names if n.startswith('Variable_')) == [
        'Variable_item_1', 'Variable_item_2', 'Variable_user_0_1', 'Variable_user_0_2']
    moved = [k for k in names if not np.array_equal(first[k].numpy(), model._variables[k].detach().numpy())]
    assert 'Variable_item_1' in moved and 'Variable_user_0_1' in moved, 'the anonymous weights were not trained'


def test_failed_first_fit_leaves_the_model_unbuilt():
    interactions, uf, itf = util.generate_dummy_data(num_users=10, num_items=12, interaction_density=.3, seed=2)
    model = T.TensorRec(n_components=3, loss_graph=T.loss_graphs.WMRBLossGraph())
    with pytest.raises(ValueError):
        model.fit(interactions, uf, itf, epochs=1, n_sampled_items=500)      # more samples than items, no replacement
    assert model.tf_prediction is None and model.n_user_features is None and not model._variables
    with pytest.raises(ModelNotFitException):
        model.predict(uf, itf)
    # a later fit may use other feature counts
    interactions2, uf2, itf2 = util.generate_dummy_data(num_users=10, num_items=12, interaction_density=.3,
                                                        n_features_per_user=7, n_features_per_item=9, seed=3)
    model.fit(interactions2, uf2, itf2, epochs=1, n_sampled_items=5)
    assert model.n_user_features == uf2.shape[1] and model.tf_prediction is not None


def test_sampler_without_replacement_is_a_uniform_subset_with_bounded_temporaries():
    rng = np.random.default_rng(0)
    n_items, n_users, n_s = 50, 4000, 7
    pairs = util.sample_items(n_items, n_users, n_s, replace=False, rng=rng)
    assert pairs.shape == (n_users * n_s, 2) and pairs.dtype == np.int64
    assert np.array_equal(pairs[:, 0], np.repeat(np.arange(n_users), n_s))
    items = pairs[:, 1].reshape(n_users, n_s)
    assert items.min() >= 0 and items.max() < n_items
    assert all(len(set(row)) == n_s for row in items)                      # no item twice for one user
    counts = np.bincount(items.reshape(-1), minlength=n_items)
    expect = n_users * n_s / n_items
    assert np.all(np.abs(counts - expect) < 6 * np.sqrt(expect))           # every item equally likely
    # user chunking: 70000 x 3000 would be a 1.7 GB float64 temporary in one piece
    big = util.sample_items(3000, 70000, 2, replace=False, rng=rng)
    assert big.shape == (140000, 2) and np.all(big[0::2, 1] != big[1::2, 1])
    full = util.sample_items(6, 5, 6, replace=False, rng=rng)[:, 1].reshape(5, 6)
    assert np.array_equal(np.sort(full, axis=1), np.tile(np.arange(6), (5, 1)))
    with pytest.raises(ValueError):
        util.sample_items(5, 3, 6, replace=False)


def test_rankings_are_evaluated_only_for_loss_graphs_that_name_them():
    from tensorrec_b200.tensorrec import _names_argument
    lg = T.loss_graphs
    assert not _names_argument(lg.RMSEDenseLossGraph().connect_loss_graph, 'tf_rankings')
    assert not _names_argument(lg.SeparationDenseLossGraph().connect_loss_graph, 'tf_rankings')

    class UsesRanks(lg.AbstractLossGraph):
        is_dense = True

        def connect_loss_graph(self, tf_prediction, tf_rankings, **kwargs):
            return tf_prediction.sum()

    assert _names_argument(UsesRanks().connect_loss_graph, 'tf_rankings')
