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
