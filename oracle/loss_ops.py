"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy restatement of the reference's loss graphs
(tensorrec/loss_graphs.py) -- the training step, SURVEY 8 row f1.  float32 like the reference's graph.

PARITY UNPINNED: the reference's tests only smoke-fit every loss graph (test/test_loss_graphs.py:17-47), they hold no
numbers; these functions follow the reference source line by line and are cross-checked against hand-computed cases
in tests/test_oracle.py.

Arguments mirror connect_loss_graph's kwargs (loss_graphs.py:17-50):
  prediction_serial [n_interactions], interactions_serial [n_interactions] -- one entry per stored interaction, COO order
  interactions_coo = (row, col, val, n_users, n_items)                       -- oracle.coo_from_sparse(matrix)
  prediction [n_users, n_items] (dense losses), sample_predictions [n_users, n_sampled_items] (sampled losses)
"""
import math

import numpy as np

F32 = np.float32


def rmse(prediction_serial, interactions_serial):
    """RMSELossGraph (loss_graphs.py:58-59): sqrt(mean((interactions - predictions)^2))."""
    p, y = np.asarray(prediction_serial, dtype=F32), np.asarray(interactions_serial, dtype=F32)
    return F32(np.sqrt(np.mean(np.square(y - p), dtype=F32)))


def rmse_dense(interactions_coo, prediction):
    """RMSEDenseLossGraph (loss_graphs.py:70-72): tf.sparse_add(interactions, -prediction), missing entries are 0
    (duplicates of the SparseTensor add up)."""
    row, col, val, n_users, n_items = interactions_coo
    dense = np.zeros((n_users, n_items), dtype=F32)
    np.add.at(dense, (row, col), val)
    error = dense + (F32(-1.0) * np.asarray(prediction, dtype=F32))
    return F32(np.sqrt(np.mean(np.square(error), dtype=F32)))


def _normal_overlap(positive_predictions, negative_predictions):
    """loss_graphs.py:91-98: moments of both groups (tf.nn.moments: biased variance), then
    1 - Normal(neg_mean - pos_mean, sqrt(neg_var + pos_var)).cdf(0)."""
    pos = np.asarray(positive_predictions, dtype=F32)
    neg = np.asarray(negative_predictions, dtype=F32)
    pos_mean, neg_mean = np.mean(pos, dtype=F32), np.mean(neg, dtype=F32)
    pos_var = np.mean(np.square(pos - pos_mean), dtype=F32)
    neg_var = np.mean(np.square(neg - neg_mean), dtype=F32)
    loc = float(neg_mean - pos_mean)
    scale = float(np.sqrt(neg_var + pos_var))
    cdf0 = 0.5 * (1.0 + math.erf((0.0 - loc) / (scale * math.sqrt(2.0))))
    return F32(1.0 - cdf0)


def separation(prediction_serial, interactions_serial):
    """SeparationLossGraph (loss_graphs.py:82-98): groups {interaction > 0} / {interaction <= 0}."""
    p, y = np.asarray(prediction_serial, dtype=F32), np.asarray(interactions_serial, dtype=F32)
    return _normal_overlap(p[y > 0.0], p[y <= 0.0])


def separation_dense(prediction, interactions_coo):
    """SeparationDenseLossGraph (loss_graphs.py:111-134): the same over the dense matrix, non-interacted = 0 = negative."""
    row, col, val, n_users, n_items = interactions_coo
    dense = np.zeros((n_users, n_items), dtype=F32)
    np.add.at(dense, (row, col), val)
    p = np.asarray(prediction, dtype=F32).reshape(-1)
    y = dense.reshape(-1)
    return _normal_overlap(p[y > 0.0], p[y <= 0.0])


def _sampled_margin_terms(prediction_serial, interactions_coo, sample_predictions):
    row, col, val, _, _ = interactions_coo
    mask = val > 0.0                                                    # loss_graphs.py:155 / 192
    positive_predictions = np.asarray(prediction_serial, dtype=F32)[mask]          # :160-161
    mapped = np.asarray(sample_predictions, dtype=F32)[row[mask]]       # :167-168 gather by the USER index
    summation_term = np.maximum(F32(1.0) - positive_predictions[:, None] + mapped, F32(0.0))   # :171-174
    return mask, np.sum(summation_term, axis=1, dtype=F32)


def wmrb(prediction_serial, interactions_coo, sample_predictions, n_items, n_sampled_items):
    """WMRBLossGraph.weighted_margin_rank_batch (loss_graphs.py:153-180): one value per POSITIVE interaction,
    log(1 + n_items / n_sampled * sum_s max(0, 1 - positive + sample_s))."""
    _, summed = _sampled_margin_terms(prediction_serial, interactions_coo, sample_predictions)
    sampled_margin_rank = (F32(n_items) / F32(n_sampled_items)) * summed
    return np.log(sampled_margin_rank + F32(1.0)).astype(F32)


def balanced_wmrb(prediction_serial, interactions_coo, sample_predictions, n_items, n_sampled_items):
    """BalancedWMRBLossGraph (loss_graphs.py:190-227): the margin rank is scaled by the interaction value over the
    item's total positive interaction mass (tf.sparse_reduce_sum over users)."""
    row, col, val, _, n_items_total = interactions_coo
    mask, summed = _sampled_margin_terms(prediction_serial, interactions_coo, sample_predictions)
    positive_values = val[mask].astype(F32)
    listening_sum_per_item = np.zeros(n_items_total, dtype=F32)
    np.add.at(listening_sum_per_item, col[mask], positive_values)
    gathered_sums = listening_sum_per_item[col[mask]]
    sampled_margin_rank = (F32(n_items) / F32(n_sampled_items)) * summed * positive_values / gathered_sums
    return np.log(sampled_margin_rank + F32(1.0)).astype(F32)


# ---------------------------------------------------------------------------------------------------
# The whole sampled-rank training step, forward AND backward, for the configuration BASELINE config
# #4 names: LinearRepresentationGraph x DotProductPredictionGraph x WMRBLossGraph (or BalancedWMRBLossGraph), n_tastes 1.
# Forward follows tensorrec/tensorrec.py:307-313, 339-346 (representations), prediction_graphs.py:52-55 (serial dot),
# recommendation_graphs.py:4-19, 44-57 (biases), 60-70 (densify), loss_graphs.py:153-180 / 190-227.  Backward is the
# analytic gradient of sum(loss) that tf.gradients builds: tf.maximum passes the gradient to its first argument where
# it is >= the second, tf.gather's gradient is a scatter-add, tf.sparse_tensor_dense_matmul's is A^T . d_out.
# PARITY UNPINNED (the reference holds no numbers for its training step); pinned here against torch autograd over the
# host mirror of the same graph functions (tests/test_train_step_cpu.py).
# ---------------------------------------------------------------------------------------------------
def wmrb_step_reference(user_features, item_features, interactions, w_user, w_item, b_user, b_item, samples,
                        balanced=False, round_repr=None):
    """Returns dict(loss [n_pos] in the COO order of the positive interactions, pred_serial [nnz] in COO order,
    d_w_user, d_w_item, d_b_user, d_b_item = gradients of sum(loss); d_b_* None when the model is unbiased).

    user_features / item_features / interactions: scipy sparse; samples: int [n_users, n_sampled] item ids;
    round_repr: optional function applied to both representations before they are used (e.g. rounding to bfloat16,
    BASELINE config #4) -- the gradient passes straight through it."""
    import scipy.sparse as sp
    uf, itf = sp.csr_matrix(user_features, dtype=F32), sp.csr_matrix(item_features, dtype=F32)
    coo = sp.coo_matrix(interactions)
    row, col, val = coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data.astype(F32)
    n_users, n_items = uf.shape[0], itf.shape[0]
    samples = np.asarray(samples, dtype=np.int64)
    n_sampled = samples.shape[1]
    biased = b_user is not None
    w_user, w_item = np.asarray(w_user, dtype=F32), np.asarray(w_item, dtype=F32)

    user_repr = np.asarray(uf @ w_user, dtype=F32)                      # representation_graphs.py:40
    item_repr = np.asarray(itf @ w_item, dtype=F32)
    if round_repr is not None:
        user_repr, item_repr = round_repr(user_repr), round_repr(item_repr)
    ub = np.asarray(uf @ np.asarray(b_user, dtype=F32), dtype=F32) if biased else np.zeros(n_users, F32)
    ib = np.asarray(itf @ np.asarray(b_item, dtype=F32), dtype=F32) if biased else np.zeros(n_items, F32)

    def serial(users, items):                                           # prediction_graphs.py:52-55 + biases :44-57
        dots = np.einsum('nk,nk->n', user_repr[users], item_repr[items]).astype(F32)
        return (dots + ub[users]) + ib[items] if biased else dots

    pred_serial = serial(row, col)
    su = np.repeat(np.arange(n_users), n_sampled)
    sample_pred = serial(su, samples.reshape(-1)).reshape(n_users, n_sampled)        # densify :60-70

    mask = val > 0.0
    prow, pcol, pval = row[mask], col[mask], val[mask]
    term = (F32(1.0) - pred_serial[mask][:, None]) + sample_pred[prow]               # loss_graphs.py:171-174
    summed = np.sum(np.maximum(term, F32(0.0)), axis=1, dtype=F32)
    scale = F32(n_items) / F32(n_sampled)
    weight = np.full(prow.shape[0], scale, dtype=F32)
    smr = scale * summed
    if balanced:
        item_sum = np.zeros(n_items, dtype=F32)
        np.add.at(item_sum, pcol, pval)
        smr = smr * pval / item_sum[pcol]
        weight = weight * pval / item_sum[pcol]
    loss = np.log(smr + F32(1.0)).astype(F32)

    # backward of sum(loss)
    dsum = (weight / (smr + F32(1.0))).astype(F32)                      # d loss_n / d summed_n
    active = term >= 0.0
    d_pos = -dsum * active.sum(axis=1).astype(F32)                      # d / d pred of the positive interaction
    d_samp = np.zeros((n_users, n_sampled), dtype=F32)
    np.add.at(d_samp, prow, dsum[:, None] * active.astype(F32))         # d / d sample_pred[u, j]

    d_user_repr = np.zeros_like(user_repr, dtype=F32)
    d_item_repr = np.zeros_like(item_repr, dtype=F32)
    d_ub, d_ib = np.zeros(n_users, F32), np.zeros(n_items, F32)
    pairs_u = np.concatenate([prow, su])
    pairs_i = np.concatenate([pcol, samples.reshape(-1)])
    pairs_g = np.concatenate([d_pos, d_samp.reshape(-1)]).astype(F32)
    np.add.at(d_user_repr, pairs_u, pairs_g[:, None] * item_repr[pairs_i])
    np.add.at(d_item_repr, pairs_i, pairs_g[:, None] * user_repr[pairs_u])
    np.add.at(d_ub, pairs_u, pairs_g)
    np.add.at(d_ib, pairs_i, pairs_g)
    return {
        'loss': loss, 'pred_serial': pred_serial, 'sample_pred': sample_pred,
        'd_user_repr': d_user_repr, 'd_item_repr': d_item_repr,
        'd_w_user': np.asarray(uf.T @ d_user_repr, dtype=F32), 'd_w_item': np.asarray(itf.T @ d_item_repr, dtype=F32),
        'd_b_user': np.asarray(uf.T @ d_ub, dtype=F32) if biased else None,
        'd_b_item': np.asarray(itf.T @ d_ib, dtype=F32) if biased else None,
        'positive_mask': mask,
    }


def adam_learning_rate(learning_rate, t, beta1=0.9, beta2=0.999):
    """lr sqrt(1 - beta2^t) / (1 - beta1^t) as TensorFlow forms it: float32 throughout, the powers are float32 variables
    multiplied by float32(beta) once per step (training_ops.cc ApplyAdam; adam.py _finish)."""
    b1p, b2p = F32(1.0), F32(1.0)
    for _ in range(int(t)):
        b1p, b2p = F32(b1p * F32(beta1)), F32(b2p * F32(beta2))
    return F32(F32(learning_rate) * np.sqrt(F32(1.0) - b2p) / (F32(1.0) - b1p))


def adam_reference(w, grad, m, v, t, learning_rate, l2=0.0, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """tf.train.AdamOptimizer (defaults as the reference uses it, tensorrec.py:489) on grad + l2 * w; t = 1, 2, ...
    The arithmetic of TensorFlow's ApplyAdam functor, float32:
        m += (g - m) (1 - beta1);  v += (g g - v) (1 - beta2);  w -= (m lr_t) / (sqrt(v) + epsilon).
    Returns (w, m, v) after the step."""
    w, grad, m, v = (np.asarray(a, dtype=F32) for a in (w, grad, m, v))
    g = grad + F32(l2) * w
    m = m + (g - m) * (F32(1.0) - F32(beta1))
    v = v + (g * g - v) * (F32(1.0) - F32(beta2))
    lr_t = adam_learning_rate(learning_rate, t, beta1, beta2)
    return (w - (m * lr_t) / (np.sqrt(v) + F32(epsilon))).astype(F32), m.astype(F32), v.astype(F32)


def round_to_bfloat16(x):
    """float32 -> nearest-even bfloat16, returned as float32 (numpy has no bfloat16)."""
    bits = np.ascontiguousarray(x, dtype=F32).view(np.uint32).astype(np.uint64)
    rounded = ((bits + 0x7FFF + ((bits >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(F32).reshape(np.shape(x))
