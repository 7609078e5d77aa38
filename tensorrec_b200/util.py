"""Host helpers mirroring tensorrec/util.py (dummy-data generators, negative sampler, batched alpha, input coercion).

The generators are the synthetic-input spec of the measurement harness.  Unlike the reference's
(util.py:61-117, unseeded), they accept a `seed` so that parity tests and benches are reproducible; seed=None keeps
the reference's behaviour."""
import math

import numpy as np
import scipy.sparse as sp


def sample_items(n_items, n_users, n_sampled_items, replace, rng=None):
    """util.py:12-21: for every user, n_sampled_items item ids; returns int64 [n_users * n_sampled_items, 2] of
    (user, item) pairs, user-major.  Vectorised (the reference loops in Python)."""
    rng = np.random.default_rng() if rng is None else rng
    if replace:
        items = rng.integers(0, n_items, size=(n_users, n_sampled_items))
    else:
        if n_sampled_items > n_items:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        # a uniformly random n_sampled_items-subset per user = the positions of the n_sampled_items smallest of n_items
        # iid uniforms (argpartition), in user chunks so that the temporary stays ~64 MB whatever n_users x n_items is
        items = np.empty((n_users, n_sampled_items), dtype=np.int64)
        chunk = max(1, (1 << 24) // max(n_items, 1))
        for u0 in range(0, n_users, chunk):
            u1 = min(n_users, u0 + chunk)
            draws = rng.random((u1 - u0, n_items), dtype=np.float32)
            if n_sampled_items < n_items:
                part = np.argpartition(draws, n_sampled_items - 1, axis=1)[:, :n_sampled_items]
            else:
                part = np.argsort(draws, axis=1)
            items[u0:u1] = part
    users = np.repeat(np.arange(n_users, dtype=np.int64), n_sampled_items)
    return np.stack([users, items.reshape(-1).astype(np.int64)], axis=1)


def calculate_batched_alpha(num_batches, alpha):
    """util.py:24-31."""
    if num_batches < 1:
        raise ValueError('num_batches must be >=1, num_batches={}'.format(num_batches))
    elif num_batches > 1:
        batched_alpha = alpha / (math.e * math.log(num_batches))
    else:
        batched_alpha = alpha
    return batched_alpha


def matrices_from_raw_input(raw_input):
    """util.py:34-58 (datasets_from_raw_input): a scipy sparse matrix or a list of them -> list of matrices.

    The reference also accepts tf.data.Dataset objects and TFRecord paths; TensorRec's own input handling
    (TensorRec._inputs_from_raw) takes the dataset 5-tuples and TFRecord paths (input_utils.py / tfrecord.py), this
    helper covers the in-memory matrices."""
    if sp.issparse(raw_input):
        return [raw_input]
    if isinstance(raw_input, list) and len(raw_input) > 0 and all(sp.issparse(v) for v in raw_input):
        return list(raw_input)
    raise ValueError('Input must be a scipy sparse matrix, an iterable of scipy sprase matrices, or a TensorFlow '
                     'Dataset')


datasets_from_raw_input = matrices_from_raw_input


def _rand(rows, cols, density, rng):
    return sp.random(rows, cols, density=density, format='coo', dtype=np.float64, random_state=rng)


def generate_dummy_data(num_users=15000, num_items=30000, interaction_density=.00045, num_user_features=200,
                        num_item_features=200, n_features_per_user=20, n_features_per_item=20, pos_int_ratio=.5,
                        return_datasets=False, seed=None):
    """util.py:61-85."""
    if pos_int_ratio <= 0.0:
        raise Exception("pos_int_ratio must be > 0")
    if return_datasets:
        raise ValueError('return_datasets=True needs tf.data, which this build does not provide')
    rng = np.random.default_rng(seed)
    interactions = _rand(num_users, num_items, interaction_density * pos_int_ratio, rng)
    if pos_int_ratio < 1.0:
        interactions = interactions + -1 * _rand(num_users, num_items, interaction_density * (1 - pos_int_ratio), rng)
    user_features = _rand(num_users, num_user_features, float(n_features_per_user) / num_user_features, rng)
    item_features = _rand(num_items, num_item_features, float(n_features_per_item) / num_item_features, rng)
    return interactions, user_features, item_features


def _indicator(rows, rng):
    n_features = int(rows * 1.2)
    n_tags = rows * 3
    r = np.concatenate([np.arange(rows), rng.integers(0, rows, n_tags)])
    c = np.concatenate([np.arange(rows), rng.integers(rows, max(n_features, rows + 1), n_tags)])
    m = sp.csr_matrix((np.ones(len(r)), (r, c)), shape=(rows, max(n_features, rows + 1)))
    m.data[:] = 1.0     # `lil[i, j] = 1` in the reference: repeated hits stay 1
    return m.tolil()


def generate_dummy_data_with_indicator(num_users=15000, num_items=30000, interaction_density=.00045, pos_int_ratio=.5,
                                       seed=None):
    """util.py:88-117 (vectorised)."""
    rng = np.random.default_rng(seed)
    user_features = _indicator(num_users, rng)
    item_features = _indicator(num_items, rng)
    n_interactions = (num_users * num_items) * interaction_density
    n_pos = int(n_interactions * pos_int_ratio)
    n_neg = int(n_interactions * (1 - pos_int_ratio))
    interactions = sp.lil_matrix((num_users, num_items))
    for count, value in ((n_pos, 1), (n_neg, -1)):
        if count:
            interactions[rng.integers(0, num_users, count), rng.integers(0, num_items, count)] = value
    return interactions, user_features, item_features


def append_to_string_at_point(string, value, point):
    for _ in range(0, (point - len(string))):
        string += " "
    string += "{}".format(value)
    return string
